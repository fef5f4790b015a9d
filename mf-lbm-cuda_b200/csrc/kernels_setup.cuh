// kernels_setup.cuh — once per geometry: the site permutation of the PDF slots, the wall-link ranks and every compact
// site list of the stepping kernels, built on the device from the node types on the U grid.
//
// The first version copied the types to the host and walked all PX*PY*PZ sites on one thread (5.3 s at 512^3 next to
// 0.5 s of device-side geometry preprocessing).  Every list is "the sites that satisfy a predicate, in z,y,x order", and
// the U index grows in exactly that order, so a list is a stream compaction over the U grid: flag kernel, exclusive scan
// (scan.cuh), fill kernel.  Nothing of the reference corresponds to this file (its kernels scan the whole volume each
// step instead); the predicates cite the loops they replace.
#pragma once
#include "core.cuh"

namespace mflbm {

struct SetupInfo {
    int has_left, has_right;   // slab neighbours
    int open_z, kper, jper;    // which boundary kernels copy phi (see P_COPIED)
};

enum SetupPred {
    P_FLUID_B = 0,   // real fluid nodes of the columns that face a neighbour slab (first segment of the fluid order)
    P_FLUID_I,       // all other real fluid nodes
    P_PASSIVE,       // other sites of the 1-ghost box: they keep PDF storage (solids realise bounce-back through it)
    P_PHI,           // solid-boundary sites of [-2 .. n+3]^3: extrapolate_phi_toSolid, src/main_iteration_GPU.cu:737
    P_CN,            // solid-boundary sites of [0 .. n+1]^3: extrapolateNormalToSolid, :885 (list chain)
    P_COPIED,        // P_PHI sites of planes whose phi a boundary kernel copies into ghost layers (kernels_chain.cuh, k_chain_pre)
    P_SHELL,         // non-solid sites outside the real box (k_chain_pre)
    P_NORMALS        // non-solid sites of [-1 .. n+2]^3: normalDirectionsOfInterfaces, :760-764 (list chain)
};

template <typename T>
__device__ __forceinline__ bool setup_pred(const Lattice<T>& L, const SetupInfo& A, const int kind, const int X, const int Y, const int Z) {
    const int x = X - 3, y = Y - 3, z = Z - 3;   // reference coordinates (1-based, ghosts <= 0)
    const int t = L.types[X + L.PX * (Y + L.PY * Z)];
    const int nx = L.nx, ny = L.ny, nz = L.nz;
    const bool in0 = x >= 1 && x <= nx && y >= 1 && y <= ny && z >= 1 && z <= nz;
    const bool in1 = x >= 0 && x <= nx + 1 && y >= 0 && y <= ny + 1 && z >= 0 && z <= nz + 1;
    const bool in2 = x >= -1 && x <= nx + 2 && y >= -1 && y <= ny + 2 && z >= -1 && z <= nz + 2;
    const bool in3 = x >= -2 && x <= nx + 3 && y >= -2 && y <= ny + 3 && z >= -2 && z <= nz + 3;
    const bool bcol = (x == 1 && A.has_left) || (x == nx && A.has_right);
    switch (kind) {
        case P_FLUID_B: return t <= 0 && in0 && bcol;
        case P_FLUID_I: return t <= 0 && in0 && !bcol;
        case P_PASSIVE: return in1 && !(t <= 0 && in0);
        case P_PHI: return t == 2 && in3;
        case P_CN: return t == 2 && in3 && in1;
        case P_COPIED: {
            if (!(t == 2 && in3)) return false;
            if (A.open_z && (z == 0 || z == nz || z == nz + 1)) return true;      // inlet_phi reads k = 0, the outlet kernels k = nz, nz + 1
            if (A.kper && ((z >= 1 && z <= 4) || (z >= nz - 3 && z <= nz))) return true;   // source layers of k_periodic_phi
            if (A.jper && ((y >= 1 && y <= 4) || (y >= ny - 3 && y <= ny))) return true;
            return false;
        }
        case P_SHELL: return t <= 0 && !in0;
        case P_NORMALS: return t <= 0 && in2;
    }
    return false;
}

// grid (PX / 128 rounded up, PY, PZ), block 128
template <typename T>
__global__ void __launch_bounds__(128) k_setup_flags(const Lattice<T> L, const SetupInfo A, const int kind, int* __restrict__ flags) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x), Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    if (X >= L.PX) return;
    flags[X + L.PX * (Y + L.PY * Z)] = setup_pred(L, A, kind, X, Y, Z) ? 1 : 0;
}

// list[rank] = U index of every site that satisfies the predicate (rank = its exclusive scan); mask != nullptr: the 18-bit mask
// of non-solid D3Q18 neighbours next to it (replaces 18 flag loads per site and step, :742-750)
template <typename T>
__global__ void __launch_bounds__(128) k_setup_fill(const Lattice<T> L, const SetupInfo A, const int kind, const int* __restrict__ scan,
                                                    int* __restrict__ list, int* __restrict__ mask) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x), Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    if (X >= L.PX || !setup_pred(L, A, kind, X, Y, Z)) return;
    const int u = X + L.PX * (Y + L.PY * Z), e = scan[u];
    list[e] = u;
    if (mask) {
        int m = 0;
#pragma unroll
        for (int q = 1; q < 19; q++) if (L.types[u + L.off(q)] <= 0) m |= 1 << (q - 1);
        mask[e] = m;
    }
}

// the site map: entry of every site of the 1-ghost box inside a PDF slot.  Fluid nodes: boundary segment, then interior
// segment; then the passive sites.  Solid-type passive sites are marked -(e + 2) (Lattice::f) - except in a ghost column that
// mirrors a neighbour slab: links that end there stay in slot storage, which is what the halo messages carry.
template <typename T>
__global__ void __launch_bounds__(128) k_setup_site_map(const Lattice<T> L, const SetupInfo A, const int* __restrict__ scan_b, const int* __restrict__ scan_i,
                                                        const int* __restrict__ scan_p, const int n_b, const int n_fluid, int* __restrict__ cmap,
                                                        int* __restrict__ fl_u) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x), Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    if (X >= L.PX) return;
    const int u = X + L.PX * (Y + L.PY * Z);
    int c = -1;
    if (setup_pred(L, A, P_FLUID_B, X, Y, Z)) { c = scan_b[u]; fl_u[c] = u; }
    else if (setup_pred(L, A, P_FLUID_I, X, Y, Z)) { c = n_b + scan_i[u]; fl_u[c] = u; }
    else if (setup_pred(L, A, P_PASSIVE, X, Y, Z)) {
        const int e = n_fluid + scan_p[u], x = X - 3;
        const bool halo_col = (x == 0 && A.has_left) || (x == L.nx + 1 && A.has_right);
        c = (L.types[u] > 0 && !halo_col) ? -(e + 2) : e;
    }
    cmap[u] = c;
}

// zstart[seg][k - 1] = first entry of slice k in segment seg (k = 1 .. nz + 2: the last two close the ranges), k_monitor
template <typename T>
__global__ void k_setup_zstart(const Lattice<T> L, const int* __restrict__ scan_b, const int* __restrict__ scan_i, const int n_b, int* __restrict__ zstart) {
    const int k = 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k > L.nz + 2) return;
    const int u0 = L.sz * (k + 3);   // first U index of plane k: every real fluid node below it has z < k
    zstart[k - 1] = scan_b[u0];
    zstart[L.nz + 2 + k - 1] = n_b + scan_i[u0];
}

// wall links (core.cuh): per group of 32 fluid entries and direction, how many entries have a solid-type neighbour there.
// One warp per group; count[(q - 1) * n_groups + g]
template <typename T>
__global__ void __launch_bounds__(128) k_setup_link_count(const Lattice<T> L, const int n_groups, int* __restrict__ count) {
    const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const int t = g * 32 + lane;
    const int u = t < L.n_fluid ? L.fl_u[t] : -1;
#pragma unroll
    for (int q = 1; q < 19; q++) {
        const unsigned v = __ballot_sync(0xffffffffu, u >= 0 && L.cmap[u + L.off(q)] < -1);
        if (lane == 0) count[(q - 1) * n_groups + g] = __popc(v);
    }
}
// wbase[g * 18 + q - 1] = rank of the first link of group g in direction q = scan over the groups of that direction
__global__ void k_setup_wbase(const int* __restrict__ scan, const int n_groups, const int n_rows, int* __restrict__ wbase) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_rows * 18) return;
    const int g = i / 18, q = i - 18 * g;
    wbase[i] = g < n_groups ? scan[q * n_groups + g] - scan[q * n_groups] : 0;
}

// counters of the reference's geometry report (src/Geometry_preprocessing.cpp:389-401): solid- and fluid-boundary nodes over
// the whole 4-ghost grid ("global") and over [-2 .. n+3]^3.  An x-slab counts the columns it owns (its real columns, plus
// the lattice's own ghost columns on a side without a neighbour), so that the slabs' counters add up to the lattice's.
template <typename T>
__global__ void __launch_bounds__(128) k_setup_counts(const Lattice<T> L, const SetupInfo A, unsigned long long* __restrict__ counts) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x), Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    int t = 0;
    bool in3 = false;
    const int xlo = A.has_left ? 1 : -3, xhi = A.has_right ? L.nx : L.nx + 4;
    if (X - 3 >= xlo && X - 3 <= xhi) {   // (the padding cells of a U row are not lattice sites)
        t = L.types[X + L.PX * (Y + L.PY * Z)];
        const int x = X - 3, y = Y - 3, z = Z - 3;
        in3 = x >= -2 && x <= L.nx + 3 && y >= -2 && y <= L.ny + 3 && z >= -2 && z <= L.nz + 3;
    }
    const unsigned s = __ballot_sync(0xffffffffu, t == 2), f = __ballot_sync(0xffffffffu, t == -1);
    const unsigned s3 = __ballot_sync(0xffffffffu, t == 2 && in3), f3 = __ballot_sync(0xffffffffu, t == -1 && in3);
    if ((threadIdx.x & 31) == 0) {
        if (s) atomicAdd(&counts[0], (unsigned long long)__popc(s));
        if (f) atomicAdd(&counts[1], (unsigned long long)__popc(f));
        if (s3) atomicAdd(&counts[2], (unsigned long long)__popc(s3));
        if (f3) atomicAdd(&counts[3], (unsigned long long)__popc(f3));
    }
}

}  // namespace mflbm
