// kernels_activity.cuh — interface-activity map for the colour-gradient chain (opt-in: MFLBM_ACTIVITY=1).
//
// Away from interfaces the order parameter of every non-solid site is +1 or -1 to within rounding, and
// normalDirectionsOfInterfaces (:757-807) zeroes every gradient shorter than 1e-6 (:795-800): the chain's results there
// are known without evaluating a single stencil.  Let every non-solid phi within Chebyshev distance 3 of a site lie in
// [s - eps, s + eps], s = +1 or -1.  Then
//   extrapolate_phi_toSolid (:732-755)        a weighted mean of such values: within eps + d of s, d = the rounding of 18
//                                             accumulations (1e-15 in double; for eps = 0 numerator and denominator are the
//                                             same sums of the same addends and the mean is exactly s)
//   normalDirectionsOfInterfaces (:757-807)   every difference of two stencil points is <= 2(eps + d); a component of the
//                                             gradient is 1/6 of one difference + 1/12 of four, <= eps + d, so
//                                             |grad| <= sqrt(3)(eps + d) -> the reference stores cn = 0, c_norm = 0
//   alter_color_gradient_solid_surface (:809) c_norm <= 1e-6 -> untouched
//   extrapolateNormalToSolid (:880-906)       mean of zeros -> 0
// eps: 1e-7 in double precision (|grad| <= 1.8e-7, a fifth of the cut-off).  In single precision d can approach the
// cut-off itself (18 roundings of 6e-8), so the test is exact there, eps = 0 - which costs little: one float ulp IS ~1e-7.
// An exact test would be useless in double: the minority density leaks diffusively at the 1e-16 level (130 sites in 2000
// steps of the benchmark workload, the whole lattice eventually), while the 1e-7 contour stays within ~40 sites of the
// interface.
// The U grid is cut into bricks of 8 x 4 x 4 sites.  Per chain:
//   k_act_scan     one pass over phi and the node types: brick flag P = "holds a non-solid site with |phi - 1| > eps",
//                  M = "... with |phi + 1| > eps" (by VALUE, so ghost planes written by boundary kernels, periodic copies
//                  and slab halos need no special case)
//   k_act_dilate   quiet[b] over the 27 bricks around b:  0 = both kinds present (an interface may be near: evaluate),
//                  1 = every non-solid site is within eps of +1, 2 = of -1, 3 = no non-solid site at all
// The 27-brick neighbourhood reaches >= 4 sites in every direction, so k_normals_act / k_alter_act / k_extrap_cn_act below
// store exactly what the full kernels would store at the sites of a quiet brick, and skip their gathers.  k_extrap_phi
// always runs in full (its result there is s only to within rounding).
#pragma once
#include "core.cuh"
#include "kernels_step.cuh"

namespace mflbm {

constexpr int ACT_BX = 8, ACT_BY = 4, ACT_BZ = 4;   // brick extents in U coordinates (PX is a multiple of 16)
template <typename T> __device__ __forceinline__ T act_eps();              // see the bound above
template <> __device__ __forceinline__ double act_eps<double>() { return 1e-7; }
template <> __device__ __forceinline__ float act_eps<float>() { return 0.0f; }

struct ActGrid {
    int nbx, nby, nbz;   // bricks per axis
    __host__ __device__ __forceinline__ int brick(int X, int Y, int Z) const { return (X / ACT_BX) + nbx * ((Y / ACT_BY) + nby * (Z / ACT_BZ)); }
    __host__ __device__ __forceinline__ int count() const { return nbx * nby * nbz; }
};

// grid (PX / 128 rounded up, PY, PZ), block 128: one thread per U site, a warp = 32 consecutive sites of one row = 4 bricks
template <typename T>
__global__ void __launch_bounds__(128) k_act_scan(const Lattice<T> L, const ActGrid G, unsigned char* __restrict__ P, unsigned char* __restrict__ M) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    bool np = false, nm = false;
    if (X < L.PX) {
        const int u = X + L.PX * (Y + L.PY * Z);
        if (L.types[u] <= 0) {
            const T v = L.phi[u];
            np = !(fabs(v - T(1)) <= act_eps<T>());     // NaN counts as both
            nm = !(fabs(v + T(1)) <= act_eps<T>());
        }
    }
    const unsigned bp = __ballot_sync(0xffffffffu, np), bm = __ballot_sync(0xffffffffu, nm);
    const int lane = threadIdx.x & 31;
    if ((lane & (ACT_BX - 1)) == 0 && X < L.PX) {
        const int b = G.brick(X, Y, Z);
        if ((bp >> lane) & 0xffu) P[b] = 1;
        if ((bm >> lane) & 0xffu) M[b] = 1;
    }
}

// one thread per brick
__global__ void __launch_bounds__(128) k_act_dilate(const ActGrid G, const unsigned char* __restrict__ P, const unsigned char* __restrict__ M,
                                                    unsigned char* __restrict__ quiet) {
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= G.count()) return;
    const int bx = b % G.nbx, by = (b / G.nbx) % G.nby, bz = b / (G.nbx * G.nby);
    unsigned p = 0, m = 0;
    for (int dz = -1; dz <= 1; dz++) {
        const int z = bz + dz;
        if (z < 0 || z >= G.nbz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = by + dy;
            if (y < 0 || y >= G.nby) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const int x = bx + dx;
                if (x < 0 || x >= G.nbx) continue;
                const int n = x + G.nbx * (y + G.nby * z);
                p |= P[n]; m |= M[n];
            }
        }
    }
    quiet[b] = (unsigned char)((p && m) ? 0 : (m ? 1 : (p ? 2 : 3)));   // only M raised: every site is near +1; only P raised: near -1
}

// k_normals with the quiet shortcut (same live / near protocol)
template <typename T>
__global__ void __launch_bounds__(128, 16) k_normals_act(const Lattice<T> L, const int* __restrict__ list, const int* __restrict__ brick,
                                                         const unsigned char* __restrict__ quiet, unsigned char* __restrict__ live,
                                                         unsigned char* __restrict__ near, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    if (quiet[brick[t]] != 0) {   // |grad phi| < 1e-6: the reference stores zeros
        if (!live[t]) return;
        live[t] = 0;
        const int u = list[t];
        L.cn_x[u] = T(0); L.cn_y[u] = T(0); L.cn_z[u] = T(0); L.c_norm[u] = T(0);
        return;
    }
    const int u = list[t];
    T gx = iso4<T, 0>(L.phi, u, L.sy, L.sz);
    T gy = iso4<T, 1>(L.phi, u, L.sy, L.sz);
    T gz = iso4<T, 2>(L.phi, u, L.sy, L.sz);
    T nrm = sqrt(gx * gx + gy * gy + gz * gz);
    if (nrm < lit<T>(1e-6)) {
        if (!live[t]) return;
        live[t] = 0;
        gx = T(0); gy = T(0); gz = T(0); nrm = T(0);
    } else {
        gx = gx / nrm; gy = gy / nrm; gz = gz / nrm;
        live[t] = 1;
        raise_near(L, near, u);
    }
    L.cn_x[u] = gx; L.cn_y[u] = gy; L.cn_z[u] = gz; L.c_norm[u] = nrm;
}

// k_alter with the quiet shortcut: c_norm is zero there (k_normals_act), the wetting rotation leaves such sites alone (:816)
template <typename T>
__global__ void k_alter_act(const Lattice<T> L, const int* __restrict__ list, const int* __restrict__ brick, const unsigned char* __restrict__ quiet,
                            const T* __restrict__ snx, const T* __restrict__ sny, const T* __restrict__ snz, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    if (quiet[brick[t]] != 0) return;
    alter_site(L, list, snx, sny, snz, t);
}

// k_extrap_cn with the quiet shortcut: every neighbour's normal is zero, near[c2] cannot have been raised in this chain
template <typename T>
__global__ void k_extrap_cn_act(const Lattice<T> L, const int* __restrict__ list, const int* __restrict__ mask, const int* __restrict__ brick,
                                const unsigned char* __restrict__ quiet, unsigned char* __restrict__ live, unsigned char* __restrict__ near,
                                const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    bool any = false;
    int c2 = -1;
    if (quiet[brick[t]] == 0) {
        c2 = list[t];
        any = near[c2] != 0;
        if (any) near[c2] = 0;
    }
    if (!any) {
        if (!live[t]) return;
        live[t] = 0;
        if (c2 < 0) c2 = list[t];
        L.cn_x[c2] = T(0); L.cn_y[c2] = T(0); L.cn_z[c2] = T(0);
        return;
    }
    const int m = mask[t];
    T sx = T(0), sy = T(0), sz = T(0), wsum = T(0);
#pragma unroll
    for (int q = 1; q < 19; q++) {
        if (m & (1 << (q - 1))) {
            const int nb = c2 + L.off(q);
            sx += L.cn_x[nb] * w_equ<T>(q); sy += L.cn_y[nb] * w_equ<T>(q); sz += L.cn_z[nb] * w_equ<T>(q); wsum += w_equ<T>(q);
        }
    }
    live[t] = 1;
    L.cn_x[c2] = sx / wsum; L.cn_y[c2] = sy / wsum; L.cn_z[c2] = sz / wsum;
}

}  // namespace mflbm
