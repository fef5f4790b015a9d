// TEST INFRASTRUCTURE (compiled by tests/test_host_driver.py::test_slab_windows_and_owned_columns): host/slabs.hpp on the CPU.
// For several lattice widths, slab counts and ghost widths: the cuts tile [1..nx] without gaps; windowing a random global
// array and gathering every slab's owned columns reproduces it exactly; a window's ghost columns are the neighbour's columns.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mf-lbm-cuda_b200/host/slabs.hpp"

int main() {
    using namespace mfhost;
    int checks = 0;
    for (long long nx : {8LL, 17LL, 60LL, 61LL})
        for (int n : {1, 2, 3, 4})
            for (int g : {1, 2, 4}) {
                if (nx / n < 4) continue;
                const auto cut = cut_slabs(nx, n);
                long long next = 1;
                for (int r = 0; r < n; r++) {
                    if (cut[r].x0 != next || cut[r].has_left != (r > 0) || cut[r].has_right != (r < n - 1)) { printf("bad cut nx=%lld n=%d r=%d\n", nx, n, r); return 1; }
                    next += cut[r].nx_local;
                }
                if (next != nx + 1) { printf("cuts do not cover nx=%lld n=%d\n", nx, n); return 1; }
                const long long rows = 3 * 5, wg = nx + 2 * g;
                std::vector<double> a((size_t)(rows * wg)), b(a.size(), -1.0), w;
                for (auto& v : a) v = (double)rand();
                for (int r = 0; r < n; r++) {
                    slab_window(a.data(), rows, g, nx, cut[r], w);
                    const long long wl = cut[r].nx_local + 2 * g;
                    for (long long row = 0; row < rows; row++)
                        for (long long lx = 0; lx < wl; lx++)
                            if (w[(size_t)(row * wl + lx)] != a[(size_t)(row * wg + cut[r].x0 - 1 + lx)]) { printf("bad window\n"); return 1; }
                    slab_gather(b.data(), w.data(), rows, g, nx, cut[r]);
                }
                if (a != b) { printf("gather does not reproduce the array: nx=%lld n=%d g=%d\n", nx, n, g); return 1; }
                checks++;
            }
    printf("SLABS_OK %d\n", checks);
    return 0;
}
