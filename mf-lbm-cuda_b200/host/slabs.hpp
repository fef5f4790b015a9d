// slabs.hpp — x-slab arithmetic of the host driver, free of any GPU call (unit-tested on the CPU: tests/test_host_driver.py):
// the cuts, the window of a global reference-layout array a slab uploads (own + ghost columns) and the columns it owns.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/mflbm.h"

namespace mfhost {

// contiguous, balanced x ranges: the first nx % n slabs get one extra column (same rule as mflbm/slab.py partition())
inline std::vector<mflbm_slab> cut_slabs(long long nxg, int n) {
    std::vector<mflbm_slab> cut;
    if (n <= 1) { cut.push_back(mflbm_slab{1, (int64_t)nxg, 0, 0}); return cut; }
    const long long base = nxg / n, rem = nxg % n;
    for (int r = 0; r < n; r++)
        cut.push_back(mflbm_slab{(int64_t)(1 + r * base + std::min<long long>(r, rem)), (int64_t)(base + (r < rem ? 1 : 0)), r > 0, r < n - 1});
    return cut;
}

// arrays are [rows][x] with x fastest: rows = everything but x (outer x z x (ny + 2g)), global width nxg + 2g, local width
// nx_local + 2g; local column lx holds global column x0 - 1 + lx
template <typename T>
void slab_window(const T* global, long long rows, int g, long long nxg, const mflbm_slab& s, std::vector<T>& local) {
    const long long wg = nxg + 2 * g, wl = s.nx_local + 2 * g, off = s.x0 - 1;
    local.resize((size_t)(rows * wl));
    for (long long n = 0; n < rows; n++) std::memcpy(&local[(size_t)(n * wl)], global + n * wg + off, sizeof(T) * (size_t)wl);
}
// the columns a slab owns: its real columns, plus the lattice's own ghost columns on a side without a neighbour
template <typename T>
void slab_gather(T* global, const T* local, long long rows, int g, long long nxg, const mflbm_slab& s) {
    const long long wg = nxg + 2 * g, wl = s.nx_local + 2 * g, off = s.x0 - 1;
    const long long lo = s.has_left ? g : 0, hi = s.has_right ? g + s.nx_local : wl;   // local columns [lo, hi)
    for (long long n = 0; n < rows; n++) std::memcpy(global + n * wg + off + lo, local + n * wl + lo, sizeof(T) * (size_t)(hi - lo));
}

}  // namespace mfhost
