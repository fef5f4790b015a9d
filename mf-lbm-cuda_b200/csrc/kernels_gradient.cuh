// kernels_gradient.cuh — interface normals from the phase-field gradient with phi staged in shared memory by TMA.
//
// Replaces normalDirectionsOfInterfaces (/root/reference/src/main_iteration_GPU.cu:757-807).  The list-driven k_normals
// (kernels_step.cuh) gathers its 18 stencil points from global memory per site: 18 LDG per warp, each touching 2-3
// cache lines, and the L1 wavefront rate bounds it (118 us at 256^3 for 150 MB of data).  Here a CTA owns GRAD_TY whole
// x-rows of the dense U grid and marches along z over a chunk of planes.  A plane = the GRAD_TY + 2 rows the stencil
// needs; planes live in a ring of four shared-memory slots, each filled by cp.async.bulk (one copy per row: rows are
// contiguous and 16-byte aligned by construction of the U grid, core.cuh) and guarded by an mbarrier with a transaction
// count.  While plane z is evaluated from slots z-1, z, z+1 the copy of plane z+2 is in flight; every phi value is
// fetched from L2/HBM (GRAD_TY+2)/GRAD_TY times and every stencil point is an LDS.
// Same arithmetic (iso4) and the same live-flag scheme as k_normals: results are bit-identical.
#pragma once
#include "core.cuh"
#include "kernels_collide.cuh"
#include "kernels_step.cuh"

namespace mflbm {

constexpr int GRAD_TY = 4, GRAD_RING = 4, GRAD_THREADS = 256;

template <typename T> __host__ __device__ constexpr int grad_align() { return 16 / (int)sizeof(T); }
// row width in elements: TX outputs, one halo element each side, rounded so both ends of the copy are 16-byte aligned
template <typename T> __host__ __device__ inline int grad_row_elems(int TX) { return TX + 2 * grad_align<T>(); }
template <typename T> inline size_t normals_tile_smem(int TX) { return sizeof(T) * (size_t)grad_row_elems<T>(TX) * (GRAD_TY + 2) * GRAD_RING + 8 * GRAD_RING; }

// live_u[u] != 0: the four outputs at U index u may be non-zero in memory (see k_normals).  near: see raise_near.
// grid = (x blocks, y blocks, z chunks of ZC planes)
template <typename T>
__global__ void __launch_bounds__(GRAD_THREADS) k_normals_tile(const Lattice<T> L, unsigned char* __restrict__ live_u, unsigned char* __restrict__ near,
                                                               const int TX, const int ZC) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int A = grad_align<T>();
    constexpr int NR = GRAD_TY + 2;
    const int RW = grad_row_elems<T>(TX);
    const int plane_elems = RW * NR;
    T* ring = reinterpret_cast<T*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(T) * (size_t)plane_elems * GRAD_RING);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // outputs: U coordinates ux in [ux0, ux0 + TX), uy in [uy0, uy0 + GRAD_TY), uz in [uz_lo, uz_hi)   (ux = x + 3; sites of
    // [-1 .. n+2]^3 have U coordinates 2 .. n+5, :760-764)
    const int ux0 = 2 + blockIdx.x * TX, uy0 = 2 + blockIdx.y * GRAD_TY;
    const int uz_lo = 2 + blockIdx.z * ZC, uz_hi = min(uz_lo + ZC, L.nz + 6);
    const int uxa = ((ux0 - 1) / A) * A;                       // first element copied, 16-byte aligned
    const int ncopy = min(RW, L.PX - uxa);                     // PX is a multiple of 16 elements: stays a multiple of A
    int nrows = 0;
    for (int r = 0; r < NR; r++) nrows += (uy0 - 1 + r < L.PY) ? 1 : 0;
    if (tid == 0) {
        for (int s = 0; s < GRAD_RING; s++) pipe::mbar_init(&bar[s], 1);
        pipe::fence_mbar_init();
    }
    __syncthreads();
    // plane uz -> ring slot uz & 3; issued by the first NR lanes of warp 0
    auto issue_plane = [&](const int uz) {
        if (uz >= L.PZ) return;
        const int s = uz & (GRAD_RING - 1);
        if (tid == 0) pipe::mbar_expect_tx(&bar[s], (uint32_t)(nrows * ncopy * (int)sizeof(T)));
        if (tid < NR && uy0 - 1 + tid < L.PY)
            pipe::bulk_g2s(ring + (size_t)s * plane_elems + (size_t)tid * RW, L.phi + (uxa + (long long)L.sy * (uy0 - 1 + tid) + (long long)L.sz * uz),
                           (uint32_t)(ncopy * (int)sizeof(T)), &bar[s]);
    };
    if (uz_lo >= uz_hi) return;
    if (warp == 0) { issue_plane(uz_lo - 1); issue_plane(uz_lo); issue_plane(uz_lo + 1); }
    // thread -> (row, x): warps 0..3 take rows 0..3 over x = lane, lane + 64, ...; warps 4..7 the same rows at x + 32
    const int ly = warp & (GRAD_TY - 1), xs = (warp >> 2) * 32 + lane;
    const int uy = uy0 + ly;
    const int ssy = RW;
    uint32_t phase_bits = 0;   // bit s = parity of the next completion of slot s this thread waits for
    auto wait_plane = [&](const int uz) {
        const int s = uz & (GRAD_RING - 1);
        pipe::mbar_wait(&bar[s], (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
    };
    wait_plane(uz_lo - 1);
    wait_plane(uz_lo);
    for (int uz = uz_lo; uz < uz_hi; uz++) {
        // the slot of plane uz + 2 held plane uz - 2: every thread passed the barrier below after using it.  Nothing is
        // issued that is not waited for (a CTA must not retire with a bulk copy into its shared memory in flight).
        if (warp == 0 && uz + 2 <= uz_hi) issue_plane(uz + 2);
        wait_plane(uz + 1);
        const T* pm = ring + (size_t)((uz - 1) & (GRAD_RING - 1)) * plane_elems;
        const T* p0 = ring + (size_t)(uz & (GRAD_RING - 1)) * plane_elems;
        const T* pp = ring + (size_t)((uz + 1) & (GRAD_RING - 1)) * plane_elems;
        if (uy <= L.ny + 5) {
            for (int lx = xs; lx < TX; lx += 64) {
                const int ux = ux0 + lx;
                if (ux > L.nx + 5) break;
                const int u = ux + L.sy * uy + L.sz * uz;
                if (L.types[u] > 0) continue;
                const int c = (ux - uxa) + ssy * (ly + 1);
                // the 18-point isotropic gradient of :765-791, operand order of iso4 (kernels_step.cuh)
                constexpr T W0 = T(1) / T(6), W1 = T(1) / T(12);
                T s, ax;
                ax = p0[c + 1] - p0[c - 1];
                s = p0[c + 1 + ssy] - p0[c - 1 - ssy];
                s = s + p0[c + 1 - ssy] - p0[c - 1 + ssy];
                s = s + pp[c + 1] - pm[c - 1];
                s = s + pm[c + 1] - pp[c - 1];
                T gx = W0 * ax + W1 * s;
                ax = p0[c + ssy] - p0[c - ssy];
                s = p0[c + 1 + ssy] - p0[c - 1 - ssy];
                s = s + p0[c - 1 + ssy] - p0[c + 1 - ssy];
                s = s + pp[c + ssy] - pm[c - ssy];
                s = s + pm[c + ssy] - pp[c - ssy];
                T gy = W0 * ax + W1 * s;
                ax = pp[c] - pm[c];
                s = pp[c + 1] - pm[c - 1];
                s = s + pp[c - 1] - pm[c + 1];
                s = s + pp[c + ssy] - pm[c - ssy];
                s = s + pp[c - ssy] - pm[c + ssy];
                T gz = W0 * ax + W1 * s;
                const T n2 = gx * gx + gy * gy + gz * gz;
                T nrm = T(0);
                // n2 < 0.98e-12 implies sqrt(n2) < 1e-6: the reference's zero branch, without the square root
                bool zero = n2 < lit<T>(0.98e-12);
                if (!zero) { nrm = sqrt(n2); zero = nrm < lit<T>(1e-6); }
                if (zero) {
                    if (!live_u[u]) continue;
                    live_u[u] = 0;
                    gx = T(0); gy = T(0); gz = T(0); nrm = T(0);
                } else {
                    gx = gx / nrm; gy = gy / nrm; gz = gz / nrm;
                    live_u[u] = 1;
                    raise_near(L, near, u);
                }
                L.cn_x[u] = gx; L.cn_y[u] = gy; L.cn_z[u] = gz; L.c_norm[u] = nrm;
            }
        }
        __syncthreads();   // plane uz - 1 is free for the copy of plane uz + 3
    }
}

}  // namespace mflbm
