// kernels_chain.cuh — the colour-gradient chain evaluated brick by brick, only where an interface can be.
//
// Replaces extrapolate_phi_toSolid, normalDirectionsOfInterfaces, alter_color_gradient_solid_surface_GPU and
// extrapolateNormalToSolid (/root/reference/src/main_iteration_GPU.cu:732-906) on the stepping path.  The reference scans
// the whole volume four times per step; the first version of this library ran four compact-list kernels (kernels_step.cuh,
// still the `list` chain: MFLBM_CHAIN=list) that cost 182 us of a 1 060 us step at 256^3 - 15 times their algorithmic bytes -
// most of it spent evaluating 18-point stencils whose result is a stored zero.
//
// Criterion (proved in the comment of k_act_verdict): away from interfaces the order parameter of every non-solid site is
// +1 or -1 to within eps, normalDirectionsOfInterfaces zeroes every gradient shorter than 1e-6 (:795-800), and then every
// result of the chain is known without evaluating a stencil.  The U grid is cut into bricks of 8 x 4 x 4 sites:
//
//   raise      P[b] = "brick b holds a non-solid site with |phi - 1| > eps", M[b] = "... |phi + 1| > eps".  Raised where phi is
//              PRODUCED: by the collide kernels for the real fluid nodes (two compares and, per run of lanes in one brick, two
//              predicated byte stores: raise_activity below), by k_act_shell for the non-solid sites outside the real box
//              (ghost planes written by the inlet / outlet / periodic kernels, ghost columns filled by slab halos; a compact
//              list, by VALUE, so none of those writers needs a special case), and by k_act_scan (one pass over phi) when the
//              chain is evaluated without a preceding collide (initial state, restart).
//   verdict    k_act_verdict: a brick is quiet when its 27-brick neighbourhood does not hold both kinds.  Bricks that are
//              not quiet now, or were not quiet in the previous chain (their non-zero normals have to be zeroed once), are
//              appended to the list of bricks to process.
//   process    k_chain_normals_csr (first version: k_chain_normals, MFLBM_CHAIN=brick), persistent CTAs, one listed brick per
//              iteration: the 16 x 8 x 8 tile of phi around the brick is staged into shared memory by ONE tensor-map TMA request
//              (mbarrier transaction count); then
//                 phase 0  phi at the solid-boundary sites of the 10 x 6 x 6 box (:732-755) from the tile, stored to the tile
//                          and to global memory (neighbouring bricks recompute the same values: identical stores); the sites
//                          and the masks of their non-solid neighbours come from a per-brick list built once per geometry
//                          (first version: collected every step from a staged tile of node types)
//                 phase 1  interface normal of the thread's own site from the tile (18 LDS, :757-807) and, on fluid-boundary
//                          sites, the wetting rotation (:809-878) on the values still in registers
//              k_chain_extrap_cn_flat (first version: k_chain_extrap_cn, one warp per brick), same bricks: cn at the
//              solid-boundary sites from the fluid neighbours' cn (:880-906), one thread per site.
//
// Results are bit-identical to the list chain (tests/test_gpu_parity.py compares every array after every kind of step),
// with one documented exception that no kernel of the stepping path can observe: phi at the solid-boundary sites of a brick
// that has been quiet for more than one chain is the value of its last evaluation.  Only stencils of sites within one site
// of that brick could read it, all of them quiet; Solver::download_state refreshes those sites (k_extrap_phi over the full
// list) before the array is handed out.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "core.cuh"
#include "kernels_step.cuh"
#include "kernels_collide.cuh"

namespace mflbm {

constexpr int BR_X = 8, BR_Y = 4, BR_Z = 4;       // brick extents in U coordinates (PX is a multiple of 16)
constexpr int TL_X = 16, TL_Y = 8, TL_Z = 8;      // staged tile: x from -4 to +11, y and z from -2 to +5 relative to the brick
constexpr int TL_OX = 4, TL_OY = 2, TL_OZ = 2;    // offset of the brick inside the tile
constexpr int CHAIN_THREADS = BR_X * BR_Y * BR_Z;

// |phi -+ 1| <= eps counts as "pure".  Double: 1e-7 gives |grad phi| <= sqrt(3)(eps + 18 roundings) = 1.8e-7, a fifth of the
// cut-off.  Single: the roundings alone approach the cut-off, so the test is exact, which costs little: one float ulp IS 6e-8
// and the minority density leaks at the 1e-16 level, far below it.
template <typename T> __device__ __forceinline__ T act_eps();
template <> __device__ __forceinline__ double act_eps<double>() { return 1e-7; }
template <> __device__ __forceinline__ float act_eps<float>() { return 0.0f; }

// The weighted means of :732-755 and :880-906 divide by the sum of the weights of the contributing neighbours, accumulated in the
// order q = 1 .. 18: n6 times 1/18, then n12 times 1/36.  Identical addends make that sum a function of (n6, n12) alone: 91
// values, tabulated by the same additions, instead of 18 predicated additions per site.
constexpr int WSUM_N = 7 * 13;
template <typename T>
__device__ __forceinline__ void wsum_fill(T* __restrict__ table, const int tid, const int nthreads) {
    for (int e = tid; e < WSUM_N; e += nthreads) {
        const int n6 = e / 13, n12 = e - 13 * n6;
        T w = T(0);
        for (int n = 0; n < n6; n++) w += w_equ<T>(1);
        for (int n = 0; n < n12; n++) w += w_equ<T>(7);
        table[e] = w;
    }
}
__device__ __forceinline__ int wsum_index(const int m) { return 13 * __popc(m & 0x3f) + __popc(m >> 6); }

// Called by every lane named in `mask` (the live lanes of a collide warp: a prefix of the warp) once phi of fluid entry t is
// known.  A warp = one group of 32 consecutive fluid entries; it raises two flags for the GROUP (two compares, two ballots,
// two predicated byte stores by one lane).  Which bricks a group touches is geometry: k_chain_pre spreads the group flags
// over them (a group that holds both kinds marks all its bricks with both - more than a per-site map would, never less).
// A first version decoded the brick of every site inside the collide kernels: +2.5 % on the odd kernel.
template <typename T>
__device__ __forceinline__ void raise_activity(const Lattice<T>& L, const int t, const T phi, const unsigned mask) {
    const bool np = !(fabs(phi - T(1)) <= act_eps<T>());   // NaN counts as both
    const bool nm = !(fabs(phi + T(1)) <= act_eps<T>());
    const unsigned bp = __ballot_sync(mask, np), bm = __ballot_sync(mask, nm);
    if ((threadIdx.x & 31u) == 0u) {
        if (bp) L.grp_p[t >> 5] = 1;
        if (bm) L.grp_m[t >> 5] = 1;
    }
}

struct BrickGrid {
    int nbx, nby, nbz;
    __host__ __device__ __forceinline__ int count() const { return nbx * nby * nbz; }
};

// the same flags from one pass over phi and the node types (chain without a preceding collide; cross-check of the above).
// grid (PX / 128 rounded up, PY, PZ), block 128: a warp = 32 consecutive sites of one row = 4 bricks
template <typename T>
__global__ void __launch_bounds__(128) k_act_scan(const Lattice<T> L, unsigned char* __restrict__ P, unsigned char* __restrict__ M, int* __restrict__ counter) {
    const int X = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int Y = (int)blockIdx.y, Z = (int)blockIdx.z;
    if (X == 0 && Y == 0 && Z == 0) { counter[0] = 0; counter[1] = 0; }   // bricks to process, their cn-extrapolation entries (k_act_verdict)
    bool np = false, nm = false;
    if (X < L.PX) {
        const int u = X + L.PX * (Y + L.PY * Z);
        if (L.types[u] <= 0) {
            const T v = L.phi[u];
            np = !(fabs(v - T(1)) <= act_eps<T>());
            nm = !(fabs(v + T(1)) <= act_eps<T>());
        }
    }
    const unsigned bp = __ballot_sync(0xffffffffu, np), bm = __ballot_sync(0xffffffffu, nm);
    const int lane = threadIdx.x & 31;
    if ((lane & (BR_X - 1)) == 0 && X < L.PX) {
        const int b = (X / BR_X) + L.nbx * ((Y / BR_Y) + L.nby * (Z / BR_Z));
        if ((bp >> lane) & 0xffu) P[b] = 1;
        if ((bm >> lane) & 0xffu) M[b] = 1;
    }
}

// First kernel of a chain, three thread ranges:
//   [0, n_shell)            brick flags of the non-solid sites outside the real box (compact list built once per geometry), by value
//   [.., + n_bc)            phi at the solid-boundary sites of the planes whose phi the boundary kernels COPY (inlet k = 0, outlet
//                           k = nz / nz+1, the source layers of periodic copies; :732-755 on a sub-list).  Those copies land in
//                           ghost layers beyond the chain's own range and would otherwise carry the stale value of a long-quiet
//                           brick into the arrays download_state hands out (header of this file).
//   [.., + n_groups)        flags of one group of 32 fluid entries (raised by the collide kernel of this step) spread over the
//                           bricks the group touches (CSR built once per geometry), then cleared for the next step
template <typename T>
__global__ void __launch_bounds__(128) k_chain_pre(const Lattice<T> L, const int* __restrict__ shell, const int n_shell, const int* __restrict__ bc_list,
                                                   const int* __restrict__ bc_mask, const int n_bc, const int* __restrict__ grp_start,
                                                   const int* __restrict__ grp_bricks, const int n_groups, int* __restrict__ counter) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0 && counter) { counter[0] = 0; counter[1] = 0; }   // the list of bricks to process is rebuilt by the verdict kernel that follows
    if (t < n_shell) {
        const int u = shell[t];
        const T v = L.phi[u];
        const unsigned z = fastdiv((unsigned)u, L.dv_sz), r = (unsigned)u - z * (unsigned)L.sz;
        const unsigned y = fastdiv(r, L.dv_px), x = r - y * (unsigned)L.PX;
        const int b = (int)((x >> 3) + (unsigned)L.nbx * ((y >> 2) + (unsigned)L.nby * (z >> 2)));
        if (!(fabs(v - T(1)) <= act_eps<T>())) L.act_p[b] = 1;
        if (!(fabs(v + T(1)) <= act_eps<T>())) L.act_m[b] = 1;
        return;
    }
    t -= n_shell;
    if (t < n_bc) {
        const int n = bc_list[t], m = bc_mask[t];
        T phi_sum = T(0), weight_sum = T(0);
#pragma unroll
        for (int q = 1; q < 19; q++)
            if (m & (1 << (q - 1))) { phi_sum += L.phi[n + L.off(q)] * w_equ<T>(q); weight_sum += w_equ<T>(q); }
        L.phi[n] = phi_sum / weight_sum;
        return;
    }
    t -= n_bc;
    if (t < n_groups) {
        const unsigned char p = L.grp_p[t], m = L.grp_m[t];
        if (p) L.grp_p[t] = 0;
        if (m) L.grp_m[t] = 0;
        const int e0 = grp_start[t], e1 = grp_start[t + 1];
        for (int e = e0; e < e1; e++) {
            const int b = grp_bricks[e];
            if (p) L.act_p[b] = 1;
            if (m) L.act_m[b] = 1;
        }
    }
}

// bricks touched by every group of 32 consecutive fluid entries (runs of equal brick id; a brick can appear twice when a group
// leaves and re-enters it, harmless).  One warp per group.  MODE 0: count[g]; MODE 1: bricks at start[g] + rank
template <typename T, int MODE>
__global__ void __launch_bounds__(128) k_group_bricks(const Lattice<T> L, const int n_groups, int* __restrict__ count_or_start, int* __restrict__ out) {
    const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const int t = g * 32 + lane;
    int b = -1;
    if (t < L.n_fluid) {
        const unsigned u = (unsigned)L.fl_u[t];
        const unsigned z = fastdiv(u, L.dv_sz), r = u - z * (unsigned)L.sz;
        const unsigned y = fastdiv(r, L.dv_px), x = r - y * (unsigned)L.PX;
        b = (int)((x >> 3) + (unsigned)L.nbx * ((y >> 2) + (unsigned)L.nby * (z >> 2)));
    }
    const int prev = __shfl_up_sync(0xffffffffu, b, 1);
    const bool lead = b >= 0 && (lane == 0 || prev != b);
    const unsigned leaders = __ballot_sync(0xffffffffu, lead);
    if (MODE == 0) { if (lane == 0) count_or_start[g] = __popc(leaders); }
    else if (lead) out[count_or_start[g] + __popc(leaders & ((1u << lane) - 1u))] = b;
}

// One thread per brick.  quiet = the 27-brick neighbourhood of b (it reaches >= 4 sites in every direction) does not hold
// both kinds, i.e. every non-solid phi within Chebyshev distance 3 of a site of b lies in [s - eps, s + eps], s = +1 or -1:
//   extrapolate_phi_toSolid (:732-755)        a weighted mean of such values: within eps + d of s (d = the rounding of 18
//                                             accumulations; for eps = 0 numerator and denominator are the same sums of the
//                                             same addends and the mean is exactly s)
//   normalDirectionsOfInterfaces (:757-807)   every difference of two stencil points is <= 2(eps + d); a component of the
//                                             gradient is 1/6 of one difference + 1/12 of four, <= eps + d, so
//                                             |grad| <= sqrt(3)(eps + d) < 1e-6 -> the reference stores cn = 0, c_norm = 0
//   alter_color_gradient_solid_surface (:809) c_norm <= 1e-6 -> untouched
//   extrapolateNormalToSolid (:880-906)       mean of zeros -> 0
// (checked against the oracle on the CPU: tests/test_activity_criterion.py).  A brick is processed when it is not quiet or
// was not quiet in the previous chain: its sites then hold non-zero normals that this chain has to zero, and its
// solid-boundary phi gets one evaluation from pure neighbours.  `quiet` carries the verdict to the next chain.
// clear_p / clear_m: the flag set the NEXT step's collide raises into.
__global__ void __launch_bounds__(128) k_act_verdict(const BrickGrid G, const unsigned char* __restrict__ P, const unsigned char* __restrict__ M,
                                                     unsigned char* __restrict__ quiet, int* __restrict__ active, int* __restrict__ counter,
                                                     unsigned char* __restrict__ clear_p, unsigned char* __restrict__ clear_m,
                                                     const int* __restrict__ sb_start, int* __restrict__ ent_off) {
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    bool proc = false;
    if (b < G.count()) {
        const int bx = b % G.nbx, by = (b / G.nbx) % G.nby, bz = b / (G.nbx * G.nby);
        // all 54 flag loads are independent (neighbours outside the grid are clamped onto it: a brick read twice changes
        // nothing in an OR); with branches around them the loop was nine dependent rounds of L2 latency
        unsigned p = 0, m = 0;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++) {
            const int z = min(max(bz + dz, 0), G.nbz - 1);
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int y = min(max(by + dy, 0), G.nby - 1);
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    const int x = min(max(bx + dx, 0), G.nbx - 1);
                    const int n = x + G.nbx * (y + G.nby * z);
                    p |= P[n]; m |= M[n];
                }
            }
        }
        const unsigned char q = (p && m) ? 0 : 1;
        proc = q == 0 || quiet[b] == 0;
        quiet[b] = q;
        if (clear_p) { clear_p[b] = 0; clear_m[b] = 0; }
    }
    // Compaction.  counter = {bricks, cn-extrapolation entries of those bricks} advanced by ONE 64-bit atomic per warp, so that
    // slot and entry offset of a brick are handed out together: ent_off is non-decreasing in the slot, whatever order the warps
    // arrive in, and k_chain_extrap_cn_flat finds the brick of a global entry number by bisection.
    const unsigned vote = __ballot_sync(0xffffffffu, proc);
    const int lane = threadIdx.x & 31;
    const int ne = (proc && ent_off) ? sb_start[b + 1] - sb_start[b] : 0;
    int incl = ne;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0, ebase = 0;
    if (lane == 0 && vote) {
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(counter), ((unsigned long long)(unsigned)tot << 32) | (unsigned)__popc(vote));
        base = (int)(unsigned)(old & 0xffffffffull); ebase = (int)(unsigned)(old >> 32);
    }
    base = __shfl_sync(0xffffffffu, base, 0); ebase = __shfl_sync(0xffffffffu, ebase, 0);
    if (proc) {
        const int slot = base + __popc(vote & ((1u << lane) - 1u));
        active[slot] = b;
        if (ent_off) ent_off[slot] = ebase + incl - ne;
    }
}

// solid-surface normals of the fluid-boundary sites, compacted brick by brick (sites of a brick in z,y,x order = thread order)
template <typename T>
struct BrickNormals {
    const int* start;   // [bricks + 1] first entry of every brick
    const T* nx; const T* ny; const T* nz;
};

// persistent CTAs over the bricks listed by k_act_verdict; see the header of this file.  Two tile buffers: while a brick is
// evaluated the tile of the CTA's next brick is in flight.  The phi tile is ONE tensor-map TMA request
// (cp.async.bulk.tensor.3d, box 16 x 8 x 8 of the U grid; coordinates outside the grid are zero-filled by the engine and
// never read): a first version issued 64 row copies per tile with cp.async.bulk and was bound by the request rate of the
// engine - 99 us for 17 k bricks, the same in both precisions.
template <typename T>
struct ChainTile {
    T phi[TL_Z * TL_Y * TL_X];
    signed char ty[TL_Z * TL_Y * TL_X];
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(pipe::smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(pipe::smem_u32(bar))
                 : "memory");
}

template <typename T>
__global__ void __launch_bounds__(CHAIN_THREADS, 8) k_chain_normals(const Lattice<T> L, const __grid_constant__ CUtensorMap tm_phi, const int* __restrict__ active,
                                                                    const int* __restrict__ n_active_ptr, const BrickNormals<T> SN) {
    __shared__ __align__(128) ChainTile<T> tile[2];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int fb_count[CHAIN_THREADS / 32];
    __shared__ short sb_idx[10 * 6 * 6];
    __shared__ int sb_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { pipe::mbar_init(&bar[0], 1); pipe::mbar_init(&bar[1], 1); pipe::fence_mbar_init(); }
    __syncthreads();
    const int n_active = *n_active_ptr;
    const int nbxy = L.nbx * L.nby;
    auto stage = [&](const int b, const int k) {
        const int bz = (int)fastdiv((unsigned)b, L.dv_nbxy), r = b - bz * nbxy, by = (int)fastdiv((unsigned)r, L.dv_nbx), bx = r - by * L.nbx;
        const int X0 = bx * BR_X - TL_OX, Y0 = by * BR_Y - TL_OY, Z0 = bz * BR_Z - TL_OZ;   // tile origin, U coordinates
        if (tid == 0) {
            pipe::mbar_expect_tx(&bar[k], (uint32_t)sizeof(tile[k].phi));
            tma_load_3d(tile[k].phi, &tm_phi, X0, Y0, Z0, &bar[k]);
        }
        // node types: 64 rows of 16 bytes as 4-byte cp.async (a uint8 tensor map of this box is rejected by the hardware as an
        // illegal instruction); x over-runs of up to 4 bytes land in the neighbouring row or in the guard bytes around the array
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = tid + CHAIN_THREADS * j, row = w >> 2, part = w & 3;
            const int Y = Y0 + (row & 7), Z = Z0 + (row >> 3);
            int* dst = reinterpret_cast<int*>(tile[k].ty) + w;
            if (Y >= 0 && Y < L.PY && Z >= 0 && Z < L.PZ) pipe::cp_async<4>(dst, L.types + (X0 + 4 * part + (long long)L.PX * (Y + (long long)L.PY * Z)));
            else *dst = 0x01010101;   // outside the grid: solid
        }
        pipe::cp_async_commit();
    };
    uint32_t phase[2] = {0u, 0u};
    int i = blockIdx.x, k = 0;
    if (i < n_active) stage(active[i], 0);
    for (; i < n_active; i += gridDim.x, k ^= 1) {
        const int b = active[i];
        const int inext = i + gridDim.x;
        if (tid == 0) sb_n = 0;
        if (inext < n_active) { stage(active[inext], k ^ 1); pipe::cp_async_wait<1>(); } else pipe::cp_async_wait<0>();
        pipe::mbar_wait(&bar[k], phase[k]);
        phase[k] ^= 1u;
        __syncthreads();
        T* const phiS = tile[k].phi;
        const signed char* const tyS = tile[k].ty;
        const int bz = (int)fastdiv((unsigned)b, L.dv_nbxy), rr = b - bz * nbxy, by = (int)fastdiv((unsigned)rr, L.dv_nbx), bx = rr - by * L.nbx;
        const int X0 = bx * BR_X - TL_OX, Y0 = by * BR_Y - TL_OY, Z0 = bz * BR_Z - TL_OZ;
        // ---- phase 0: phi at the solid-boundary sites of the 10 x 6 x 6 box around the brick (:732-755), sites of [-2 .. n+3]^3.
        //      The sites are collected first (a warp that walked the 18 neighbours for one hit among 32 lanes cost half of the
        //      kernel's instructions), then evaluated one per thread.
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int s = tid + CHAIN_THREADS * j;
            bool hit = false;
            int c = 0;
            if (s < 10 * 6 * 6) {
                const int tz = s / 60, r2 = s - 60 * tz, ty = r2 / 10, tx = r2 - 10 * ty;
                c = (tx + TL_OX - 1) + TL_X * ((ty + TL_OY - 1) + TL_Y * (tz + TL_OZ - 1));
                const int X = X0 + tx + TL_OX - 1, Y = Y0 + ty + TL_OY - 1, Z = Z0 + tz + TL_OZ - 1;
                hit = tyS[c] == 2 && X >= 1 && X <= L.nx + 6 && Y >= 1 && Y <= L.ny + 6 && Z >= 1 && Z <= L.nz + 6;
            }
            const unsigned vote = __ballot_sync(0xffffffffu, hit);
            if (vote) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sb_n, __popc(vote));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) sb_idx[base + __popc(vote & ((1u << lane) - 1u))] = (short)c;
            }
        }
        // rank of a fluid-boundary site inside its brick (sites of the whole U grid are listed, whatever the kernel's range)
        const int tx = TL_OX + (tid & 7), ty = TL_OY + ((tid >> 3) & 3), tz = TL_OZ + (tid >> 5);
        const int c = tx + TL_X * (ty + TL_Y * tz);
        const int t = tyS[c];
        const unsigned fb = __ballot_sync(0xffffffffu, t == -1);
        if (lane == 0) fb_count[warp] = __popc(fb);
        __syncthreads();
        for (int e = tid; e < sb_n; e += CHAIN_THREADS) {
            const int cs = sb_idx[e];
            T phi_sum = T(0), weight_sum = T(0);
#pragma unroll
            for (int q = 1; q < 19; q++) {
                const int o = ex(q) + TL_X * (ey(q) + TL_Y * ez(q));
                if (tyS[cs + o] <= 0) { phi_sum += phiS[cs + o] * w_equ<T>(q); weight_sum += w_equ<T>(q); }
            }
            const T v = phi_sum / weight_sum;
            phiS[cs] = v;   // only non-solid sites are read in this phase: no hazard
            L.phi[(X0 + (cs & (TL_X - 1))) + L.PX * ((Y0 + ((cs >> 4) & (TL_Y - 1))) + L.PY * (Z0 + (cs >> 7)))] = v;
        }
        __syncthreads();
        // ---- phase 1: normal of my own site (:757-807) and, on fluid-boundary sites, the wetting rotation (:809-878)
        const int X = X0 + tx, Y = Y0 + ty, Z = Z0 + tz;
        const bool in_range = X >= 2 && X <= L.nx + 5 && Y >= 2 && Y <= L.ny + 5 && Z >= 2 && Z <= L.nz + 5;   // [-1 .. n+2]^3
        if (t <= 0 && in_range) {
            T gx = iso4<T, 0>(phiS, c, TL_X, TL_X * TL_Y);
            T gy = iso4<T, 1>(phiS, c, TL_X, TL_X * TL_Y);
            T gz = iso4<T, 2>(phiS, c, TL_X, TL_X * TL_Y);
            const T s2 = gx * gx + gy * gy + gz * gz;
            T nrm = T(0);
            // s2 < 2.5e-13 means sqrt(s2) <= 5e-7 < 1e-6, the reference's zero branch (:795): most sites of a processed brick
            // take it, without the square root and the three divisions
            if (s2 < lit<T>(2.5e-13)) { gx = T(0); gy = T(0); gz = T(0); }
            else {
                nrm = sqrt(s2);
                if (nrm < lit<T>(1e-6)) { gx = T(0); gy = T(0); gz = T(0); nrm = T(0); }
                else { gx = gx / nrm; gy = gy / nrm; gz = gz / nrm; }
                if (t == -1 && nrm > lit<T>(1e-6)) {
                    int e = SN.start[b] + __popc(fb & ((1u << lane) - 1u));
                    for (int w = 0; w < warp; w++) e += fb_count[w];
                    alter_values<T>(L.cos_theta, SN.nx[e], SN.ny[e], SN.nz[e], gx, gy, gz);
                }
            }
            const int u = X + L.PX * (Y + L.PY * Z);
            L.cn_x[u] = gx; L.cn_y[u] = gy; L.cn_z[u] = gz; L.c_norm[u] = nrm;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic accesses to this buffer before its next TMA fill
        __syncthreads();
    }
}

// The same chain stage with everything that is geometry taken out of the step (the default; MFLBM_CHAIN=brick selects the kernel above).  k_chain_normals spends a
// third of its instructions finding the solid-boundary sites of the 10 x 6 x 6 box in a staged tile of node types and probing
// the types of their 18 neighbours; both are fixed per geometry.  Here a per-brick CSR (k_brick_box_sites) lists those sites
// as tile index + 18-bit mask of non-solid neighbours in one int: no node-type tile, no collection pass, no shared counters.
//   * one __syncthreads per brick instead of three: the tile of the CTA's next brick is requested AFTER the barrier that
//     separates phase 0 from phase 1 of the current one, into the buffer of the PREVIOUS brick - every thread that has passed
//     that barrier has finished phase 1 of the previous brick (each thread's fence.proxy.async follows its last read of a
//     buffer); the per-warp counts of fluid-boundary sites are double-buffered for the same reason.  (A third buffer, requested a
//     whole iteration ahead, was measured: no faster, profiles/r03c_chain_variants.json.)
//   * the next brick's CSR entries (at most 3 per thread: 360 box sites / 128 threads) and the type of the thread's own site
//     are plain loads into registers issued one brick ahead, the entry range two bricks ahead, the brick id three ahead, so
//     that no load of an iteration depends on another load of the same iteration.
// Arithmetic, evaluation order and every store are those of k_chain_normals (bit-identical: tests/test_gpu_parity.py).
template <typename T, int CTAS>
__global__ void __launch_bounds__(CHAIN_THREADS, CTAS) k_chain_normals_csr(const Lattice<T> L, const __grid_constant__ CUtensorMap tm_phi, const int* __restrict__ active,
                                                                         const int* __restrict__ n_active_ptr, const BrickNormals<T> SN,
                                                                         const int* __restrict__ bx_start, const int* __restrict__ bx_ent, unsigned* __restrict__ live) {
    __shared__ __align__(128) T tile[2][TL_Z * TL_Y * TL_X];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int fb_count[2][CHAIN_THREADS / 32];
    __shared__ T wsum[WSUM_N];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { pipe::mbar_init(&bar[0], 1); pipe::mbar_init(&bar[1], 1); pipe::fence_mbar_init(); }
    wsum_fill<T>(wsum, tid, CHAIN_THREADS);
    __syncthreads();
    const int n_active = *n_active_ptr;
    const int nbxy = L.nbx * L.nby;
    const int stride = gridDim.x;
    const int tx = TL_OX + (tid & 7), ty = TL_OY + ((tid >> 3) & 3), tz = TL_OZ + (tid >> 5);
    const int c = tx + TL_X * (ty + TL_Y * tz);   // my own site inside the tile
    auto origin = [&](const int b, int& X0, int& Y0, int& Z0) {
        const int bz = (int)fastdiv((unsigned)b, L.dv_nbxy), r = b - bz * nbxy, by = (int)fastdiv((unsigned)r, L.dv_nbx), bx = r - by * L.nbx;
        X0 = bx * BR_X - TL_OX; Y0 = by * BR_Y - TL_OY; Z0 = bz * BR_Z - TL_OZ;   // tile origin, U coordinates
    };
    auto request = [&](const int b, const int k) {   // thread 0 only
        int X0, Y0, Z0;
        origin(b, X0, Y0, Z0);
        pipe::mbar_expect_tx(&bar[k], (uint32_t)sizeof(tile[k]));
        tma_load_3d(tile[k], &tm_phi, X0, Y0, Z0, &bar[k]);
    };
    auto brick_at = [&](const int i) -> int { return i < n_active ? __ldg(active + i) : -1; };
    auto own_type = [&](const int b) -> int {   // node type of my site of brick b; bricks may stick out of the grid in y and z
        int X0, Y0, Z0;
        origin(b, X0, Y0, Z0);
        const int Y = Y0 + ty, Z = Z0 + tz;
        return (Y < L.PY && Z < L.PZ) ? (int)L.types[(X0 + tx) + L.PX * (Y + L.PY * Z)] : 1;
    };
    int i = blockIdx.x;
    if (i >= n_active) return;
    // ---- prologue: brick i complete, range of i + stride, id of i + 2 stride
    int b = brick_at(i), bN = brick_at(i + stride), bN2 = brick_at(i + 2 * stride);
    if (tid == 0) request(b, 0);
    int s0 = __ldg(bx_start + b), s1 = __ldg(bx_start + b + 1);
    int s0N = 0, s1N = 0;
    if (bN >= 0) { s0N = __ldg(bx_start + bN); s1N = __ldg(bx_start + bN + 1); }
    int ent[3];
#pragma unroll
    for (int j = 0; j < 3; j++) { const int e = s0 + tid + CHAIN_THREADS * j; ent[j] = e < s1 ? __ldg(bx_ent + e) : -1; }
    int t = own_type(b);
    unsigned lv = live[4 * b + warp];   // bit l: the four outputs of lane l's site may be non-zero in memory
    uint32_t phases = 0u;   // bit k: parity of the next completion of bar[k]
    for (int k = 0; b >= 0; i += stride, k ^= 1) {
        // ---- loads for the bricks ahead; none depends on a load issued in this iteration
        int entN[3] = {-1, -1, -1}, tN = 1, s0N2 = 0, s1N2 = 0;
        unsigned lvN = 0u;
        if (bN >= 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) { const int e = s0N + tid + CHAIN_THREADS * j; if (e < s1N) entN[j] = __ldg(bx_ent + e); }
            tN = own_type(bN);
            lvN = live[4 * bN + warp];
        }
        if (bN2 >= 0) { s0N2 = __ldg(bx_start + bN2); s1N2 = __ldg(bx_start + bN2 + 1); }
        const int bN3 = brick_at(i + 3 * stride);
        int X0, Y0, Z0;
        origin(b, X0, Y0, Z0);
        pipe::mbar_wait(&bar[k], (phases >> k) & 1u);
        phases ^= 1u << k;
        T* const phiS = tile[k];
        // ---- phase 0: phi at the solid-boundary sites of the 10 x 6 x 6 box around the brick (:732-755), sites of [-2 .. n+3]^3
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (ent[j] < 0) continue;
            const int cs = ent[j] & 1023, m = ent[j] >> 10;
            T phi_sum = T(0);
#pragma unroll
            for (int q = 1; q < 19; q++) {
                const int o = ex(q) + TL_X * (ey(q) + TL_Y * ez(q));
                if (m & (1 << (q - 1))) phi_sum += phiS[cs + o] * w_equ<T>(q);
            }
            const T v = phi_sum / wsum[wsum_index(m)];
            phiS[cs] = v;   // only non-solid sites are read in this phase: no hazard
            L.phi[(X0 + (cs & (TL_X - 1))) + L.PX * ((Y0 + ((cs >> 4) & (TL_Y - 1))) + L.PY * (Z0 + (cs >> 7)))] = v;
        }
        const unsigned fb = __ballot_sync(0xffffffffu, t == -1);
        if (lane == 0) fb_count[k][warp] = __popc(fb);
        __syncthreads();
        if (tid == 0 && bN >= 0) request(bN, k ^ 1);   // the buffer of the PREVIOUS brick: its readers are all past the barrier above
        // ---- phase 1: normal of my own site (:757-807) and, on fluid-boundary sites, the wetting rotation (:809-878)
        const int X = X0 + tx, Y = Y0 + ty, Z = Z0 + tz;
        const bool in_range = X >= 2 && X <= L.nx + 5 && Y >= 2 && Y <= L.ny + 5 && Z >= 2 && Z <= L.nz + 5;   // [-1 .. n+2]^3
        // Most sites of a processed brick get zeros that are already there (the reference rewrites them every step, 4 stores per
        // site): `lv` remembers per site whether memory may hold anything else, and a site that was zero and stays zero stores nothing.
        const bool mine = t <= 0 && in_range;
        T gx = T(0), gy = T(0), gz = T(0), nrm = T(0);
        if (mine) {
            gx = iso4<T, 0>(phiS, c, TL_X, TL_X * TL_Y);
            gy = iso4<T, 1>(phiS, c, TL_X, TL_X * TL_Y);
            gz = iso4<T, 2>(phiS, c, TL_X, TL_X * TL_Y);
            const T s2 = gx * gx + gy * gy + gz * gz;
            if (s2 < lit<T>(2.5e-13)) { gx = T(0); gy = T(0); gz = T(0); }   // see k_chain_normals
            else {
                nrm = sqrt(s2);
                if (nrm < lit<T>(1e-6)) { gx = T(0); gy = T(0); gz = T(0); nrm = T(0); }
                else { gx = gx / nrm; gy = gy / nrm; gz = gz / nrm; }
                if (t == -1 && nrm > lit<T>(1e-6)) {
                    int e = SN.start[b] + __popc(fb & ((1u << lane) - 1u));
                    for (int w = 0; w < warp; w++) e += fb_count[k][w];
                    alter_values<T>(L.cos_theta, SN.nx[e], SN.ny[e], SN.nz[e], gx, gy, gz);
                }
            }
        }
        const unsigned now = __ballot_sync(0xffffffffu, nrm != T(0));   // nrm == 0 <=> all four outputs are +0
        if (mine && ((now | lv) >> lane) & 1u) {
            const int u = X + L.PX * (Y + L.PY * Z);
            L.cn_x[u] = gx; L.cn_y[u] = gy; L.cn_z[u] = gz; L.c_norm[u] = nrm;
        }
        if (lane == 0 && now != lv) live[4 * b + warp] = now;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic accesses to this buffer before its next TMA fill
        b = bN; bN = bN2; bN2 = bN3;
        s0N = s0N2; s1N = s1N2;
#pragma unroll
        for (int j = 0; j < 3; j++) ent[j] = entN[j];
        t = tN; lv = lvN;
    }
}

// cn at the solid-boundary sites of [0 .. n+1]^3 of the listed bricks <- weighted mean of the fluid neighbours' cn (:880-906).
// One warp per brick over the brick's entries of a compact list (sites + 18-bit fluid-neighbour masks, built once per
// geometry): most bricks of an open region hold none, a brick inside the pack a few dozen.  (Four bricks per warp, eight lanes
// each, was measured: 10 us slower, profiles/r03a_chain_variants.json.)
template <typename T>
__global__ void __launch_bounds__(CHAIN_THREADS, 12) k_chain_extrap_cn(const Lattice<T> L, const int* __restrict__ active, const int* __restrict__ n_active_ptr,
                                                                      const int* __restrict__ sb_start, const int* __restrict__ sb_list, const int* __restrict__ sb_mask) {
    const int lane = threadIdx.x & 31;
    const int n_active = *n_active_ptr;
    const int nwarps = gridDim.x * (CHAIN_THREADS / 32);
    int i = blockIdx.x * (CHAIN_THREADS / 32) + (threadIdx.x >> 5);
    if (i >= n_active) return;
    // the entry range of the warp's next brick is fetched while the current one is evaluated (the kernel is a chain of
    // dependent loads: brick -> range -> entry -> 54 normals)
    int b = active[i];
    int s0 = sb_start[b], s1 = sb_start[b + 1];
    while (true) {
        const int inext = i + nwarps;
        int s0n = 0, s1n = 0;
        if (inext < n_active) { const int bn = active[inext]; s0n = sb_start[bn]; s1n = sb_start[bn + 1]; }
        for (int e = s0 + lane; e < s1; e += 32) {
            const int c2 = sb_list[e], m = sb_mask[e];
            T sx = T(0), sy = T(0), sz = T(0), wsum = T(0);
#pragma unroll
            for (int q = 1; q < 19; q++) {
                if (m & (1 << (q - 1))) {
                    const int nb = c2 + L.off(q);
                    sx += L.cn_x[nb] * w_equ<T>(q); sy += L.cn_y[nb] * w_equ<T>(q); sz += L.cn_z[nb] * w_equ<T>(q); wsum += w_equ<T>(q);
                }
            }
            L.cn_x[c2] = sx / wsum; L.cn_y[c2] = sy / wsum; L.cn_z[c2] = sz / wsum;
        }
        if (inext >= n_active) break;
        i = inext; s0 = s0n; s1 = s1n;
    }
}

// The same over a flat numbering of the entries of all listed bricks (k_act_verdict hands every listed brick its entry offset):
// one thread per entry, every lane busy, no warp waits for the one brick of the pack that holds 60 entries while its
// neighbours hold none.  counter = {listed bricks, their entries}.  The brick of entry g = the last slot whose offset is <= g
// (empty bricks share their successor's offset and are never selected): 15 rounds of bisection over an array that stays in L1.
template <typename T>
__global__ void __launch_bounds__(CHAIN_THREADS, 12) k_chain_extrap_cn_flat(const Lattice<T> L, const int* __restrict__ active, const int* __restrict__ counter,
                                                                           const int* __restrict__ ent_off, const int* __restrict__ sb_start,
                                                                           const int* __restrict__ sb_list, const int* __restrict__ sb_mask) {
    __shared__ T wsum_t[WSUM_N];
    wsum_fill<T>(wsum_t, threadIdx.x, CHAIN_THREADS);
    __syncthreads();
    const int n_active = counter[0];
    const unsigned total = (unsigned)counter[1];
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_active - 1;   // ent_off[lo] <= g throughout (ent_off[0] == 0)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((unsigned)__ldg(ent_off + mid) <= g) lo = mid; else hi = mid - 1;
        }
        const int e = __ldg(sb_start + __ldg(active + lo)) + (int)(g - (unsigned)__ldg(ent_off + lo));
        const int c2 = sb_list[e], m = sb_mask[e];
        T sx = T(0), sy = T(0), sz = T(0);
#pragma unroll
        for (int q = 1; q < 19; q++) {
            if (m & (1 << (q - 1))) {
                const int nb = c2 + L.off(q);
                sx += L.cn_x[nb] * w_equ<T>(q); sy += L.cn_y[nb] * w_equ<T>(q); sz += L.cn_z[nb] * w_equ<T>(q);
            }
        }
        const T wsum = wsum_t[wsum_index(m)];
        L.cn_x[c2] = sx / wsum; L.cn_y[c2] = sy / wsum; L.cn_z[c2] = sz / wsum;
    }
}

// ---------------------------------------------------------------------------------------------------------
// geometry-time helpers: sites of one kind compacted brick by brick (count, scan on the host side of Solver, fill).
// KIND 0: fluid-boundary sites (type -1) of the whole U grid; KIND 1: solid-boundary sites (type 2) of [0 .. n+1]^3 with the
// mask of their non-solid D3Q18 neighbours.  One CTA per brick; sites of a brick in z,y,x order = thread order.
// MODE 0 count; MODE 1 write U indices (and masks) at start[b] + rank
// ---------------------------------------------------------------------------------------------------------
template <typename T, int KIND, int MODE>
__global__ void __launch_bounds__(CHAIN_THREADS) k_brick_sites(const Lattice<T> L, int* __restrict__ count_or_start, int* __restrict__ list, int* __restrict__ mask) {
    __shared__ int wc[CHAIN_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx = b % L.nbx, by = (b / L.nbx) % L.nby, bz = b / (L.nbx * L.nby);
    const int X = bx * BR_X + (tid & 7), Y = by * BR_Y + ((tid >> 3) & 3), Z = bz * BR_Z + (tid >> 5);
    const bool in = X < L.PX && Y < L.PY && Z < L.PZ;
    const int u = X + L.PX * (Y + L.PY * Z);
    bool hit;
    if (KIND == 0) hit = in && L.types[u] == -1;
    else hit = in && X >= 3 && X <= L.nx + 4 && Y >= 3 && Y <= L.ny + 4 && Z >= 3 && Z <= L.nz + 4 && L.types[u] == 2;
    const unsigned vote = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wc[warp] = __popc(vote);
    __syncthreads();
    if (MODE == 0) {
        if (tid == 0) count_or_start[b] = wc[0] + wc[1] + wc[2] + wc[3];
    } else if (hit) {
        int e = count_or_start[b] + __popc(vote & ((1u << lane) - 1u));
        for (int w = 0; w < warp; w++) e += wc[w];
        list[e] = u;
        if (KIND == 1) {
            int m = 0;
#pragma unroll
            for (int q = 1; q < 19; q++) if (L.types[u + L.off(q)] <= 0) m |= 1 << (q - 1);
            mask[e] = m;
        }
    }
}

// The solid-boundary sites of the 10 x 6 x 6 box around every brick that phase 0 of k_chain_normals_csr evaluates (sites of
// [-2 .. n+3]^3, :732-755), each packed as  tile index (10 bits) | mask of the non-solid D3Q18 neighbours << 10.  One CTA per
// brick, box sites in z,y,x order.  MODE 0 count; MODE 1 write at start[b] + rank.  (A site appears in the lists of up to
// 2.8 bricks on average - every brick evaluates its whole box, as k_chain_normals does, so that no stencil of a processed brick
// ever reads the stale value a quiet neighbour brick may hold.)
template <typename T, int MODE>
__global__ void __launch_bounds__(CHAIN_THREADS) k_brick_box_sites(const Lattice<T> L, int* __restrict__ count_or_start, int* __restrict__ ent, int* __restrict__) {
    __shared__ int wc[3][CHAIN_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx = b % L.nbx, by = (b / L.nbx) % L.nby, bz = b / (L.nbx * L.nby);
    const int X0 = bx * BR_X - TL_OX, Y0 = by * BR_Y - TL_OY, Z0 = bz * BR_Z - TL_OZ;
    unsigned vote[3];
    bool hit[3];
    int packed[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int s = tid + CHAIN_THREADS * j;
        hit[j] = false; packed[j] = 0;
        if (s < 10 * 6 * 6) {
            const int tz = s / 60, r2 = s - 60 * tz, ty = r2 / 10, tx = r2 - 10 * ty;
            const int c = (tx + TL_OX - 1) + TL_X * ((ty + TL_OY - 1) + TL_Y * (tz + TL_OZ - 1));
            const int X = X0 + tx + TL_OX - 1, Y = Y0 + ty + TL_OY - 1, Z = Z0 + tz + TL_OZ - 1;
            // the range lies inside the grid with one site to spare on every side (PX >= nx + 8, PY = ny + 8, PZ = nz + 8)
            if (X >= 1 && X <= L.nx + 6 && Y >= 1 && Y <= L.ny + 6 && Z >= 1 && Z <= L.nz + 6) {
                const int u = X + L.PX * (Y + L.PY * Z);
                hit[j] = L.types[u] == 2;
                if (MODE == 1 && hit[j]) {
                    int m = 0;
#pragma unroll
                    for (int q = 1; q < 19; q++) if (L.types[u + L.off(q)] <= 0) m |= 1 << (q - 1);
                    packed[j] = c | (m << 10);
                }
            }
        }
        vote[j] = __ballot_sync(0xffffffffu, hit[j]);
        if (lane == 0) wc[j][warp] = __popc(vote[j]);
    }
    __syncthreads();
    if (MODE == 0) {
        if (tid == 0) {
            int n = 0;
            for (int j = 0; j < 3; j++) for (int w = 0; w < CHAIN_THREADS / 32; w++) n += wc[j][w];
            count_or_start[b] = n;
        }
    } else {
        int base = count_or_start[b];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            int e = base + __popc(vote[j] & ((1u << lane) - 1u));
            for (int w = 0; w < CHAIN_THREADS / 32; w++) { if (w < warp) e += wc[j][w]; base += wc[j][w]; }
            if (hit[j]) ent[e] = packed[j];
        }
    }
}

}  // namespace mflbm
