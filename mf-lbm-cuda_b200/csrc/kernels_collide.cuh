// kernels_collide.cuh — software-pipelined AA collide+stream kernels for sm_100a.
//
// Same arithmetic and data flow as k_collide (kernels_step.cuh; reference src/main_iteration_GPU.cu:56-726), different
// memory engine.  The plain kernel is latency-bound: 140-254 registers per thread leave 8-12 warps per SM, and every
// warp serialises   site id -> neighbour map -> 38 PDF rows -> ~500 FP64 instructions -> 38 stores   (ncu: 65 % of the
// samples sit on three long-scoreboard waits).  Here the PDF rows of the NEXT tiles are in flight while a tile is
// being collided, independent of occupancy and of the register file:
//
//   EVEN (local read/write): a persistent CTA walks 128-entry tiles of the permuted fluid order.  One thread issues
//        38 TMA bulk copies (cp.async.bulk, 128*sizeof(T) bytes each, one per slot row) per tile into a shared-memory
//        ring; an mbarrier with a transaction count signals arrival.  Threads read their column, the stage is handed
//        back to the TMA, results go straight from registers to global memory (rows are contiguous and aligned).
//   ODD  (pull from x-e, push to x+e): neighbours are found through the site map, so rows are gathered per thread
//        with cp.async (LDGSTS, 4/8 bytes) into the thread's own shared-memory column D tiles ahead; the 18 map
//        look-ups of the tile after that are ordinary loads that land during the collision.  A thread only ever
//        reads shared memory it wrote itself, so the odd kernel needs no CTA barrier at all.
//
// Bulk skip: where c_norm == 0 the reference's surface-tension force 0.5*gamma*curv*c_norm*cn and the recolouring
// term are exactly zero (normalDirectionsOfInterfaces zeroes cn together with c_norm, :795-800), so cn_* and the
// 57-point curvature stencil are only read at nodes that carry an interface (`bulk_skip`, see Solver::cn_consistent).
#pragma once
#include "core.cuh"
#include "kernels_step.cuh"

namespace mflbm {
namespace pipe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace pipe

constexpr int COLLIDE_TILE = 128;

// surface-tension inputs of one node (:143-150): cn and 0.5*gamma*curv*c_norm
template <typename T>
__device__ __forceinline__ void node_force(const Lattice<T>& L, const int u, const T cnorm, const bool bulk_skip, T& cnx, T& cny, T& cnz, T& tmp) {
    if (bulk_skip && cnorm == T(0)) { cnx = T(0); cny = T(0); cnz = T(0); tmp = T(0); return; }
    cnx = L.cn_x[u]; cny = L.cn_y[u]; cnz = L.cn_z[u];
    tmp = lit<T>(0.5) * L.lbm_gamma * curvature_at(L, u) * cnorm;   // :147
}

template <typename T, int NST>
constexpr size_t collide_even_smem() { return sizeof(T) * NST * 38 * COLLIDE_TILE + 8 * NST; }
template <typename T, int D>
constexpr size_t collide_odd_smem() { return (sizeof(T) * 38 + sizeof(int) * 18) * (D + 1) * COLLIDE_TILE; }

// ---------------------------------------------------------------------------------------------------------
// EVEN step: f_q = local slot opc(q); collide; local slot q = f_q*     (:395-726)
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MRT, int NST, int CTAS>
__global__ void __launch_bounds__(COLLIDE_TILE, CTAS) k_collide_even_tma(const Lattice<T> L, const int ntiles, const int bulk_skip) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef T Stage[38][COLLIDE_TILE];
    Stage* buf = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * NST);
    const int tid = threadIdx.x;
    const long long NC = L.NC;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) pipe::mbar_init(&full[s], 1);
        pipe::fence_mbar_init();
    }
    __syncthreads();
    // threads 0..37 each copy one slot row of the tile (one UBLKCP per thread instead of 38 serial ones in one thread);
    // thread 0 posts the byte count.  The phase cannot complete before that arrival, whatever the order.
    auto issue = [&](const int tile, const int s) {
        if (tid == 0) pipe::mbar_expect_tx(&full[s], (uint32_t)sizeof(Stage));
        if (tid < 38) pipe::bulk_g2s(&buf[s][tid][0], L.pdf + (long long)tid * NC + (long long)tile * COLLIDE_TILE, (uint32_t)(sizeof(T) * COLLIDE_TILE), &full[s]);
    };
    int tile = blockIdx.x;
#pragma unroll
    for (int s = 0; s < NST; s++) {
        const int tl = tile + s * (int)gridDim.x;
        if (tl < ntiles) issue(tl, s);
    }
    // site id and c_norm of my entry, fetched one / two tiles ahead (a dependent pair of loads: exposed, they were
    // half of all stall samples)
    auto site = [&](const int tl) -> int {
        const int t = tl * COLLIDE_TILE + tid;
        return (tl < ntiles && t < L.n_fluid) ? L.fl_u[t] : -1;
    };
    int u = site(tile), uN = site(tile + (int)gridDim.x);
    T cnorm = u >= 0 ? L.c_norm[u] : T(0);
    int s = 0;
    uint32_t phase = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        const int t = tile * COLLIDE_TILE + tid;
        const bool live = u >= 0;
        const T cnormN = uN >= 0 ? L.c_norm[uN] : T(0);
        const int uNN = site(tile + 2 * (int)gridDim.x);
        T cnx = T(0), cny = T(0), cnz = T(0), tmp = T(0);
        if (live) node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);   // interface nodes only: cn + curvature stencil
        pipe::mbar_wait(&full[s], phase);
        T g1[19], g2[19];
#pragma unroll
        for (int q = 0; q < 19; q++) { g1[q] = buf[s][opc(q)][tid]; g2[q] = buf[s][opc(q) + 19][tid]; }
        __syncthreads();   // every thread has its column: the stage goes back to the TMA
        {
            const int tn = tile + NST * (int)gridDim.x;
            if (tn < ntiles) issue(tn, s);
        }
        if (live) {
            const T phi_loc = collide_node<T, MRT>(L, g1, g2, cnx, cny, cnz, tmp);
            L.phi[u] = phi_loc;
            T* __restrict__ po = L.pdf + t;
#pragma unroll
            for (int q = 0; q < 19; q++) { po[(long long)q * NC] = g1[q]; po[(long long)(q + 19) * NC] = g2[q]; }
        }
        u = uN; uN = uNN; cnorm = cnormN;
        if (++s == NST) { s = 0; phase ^= 1; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// ODD step: f_q pulled from x - e_q (slot q); collide; f_q* pushed to x + e_q (slot opc(q))     (:56-388)
// D = how many tiles ahead the PDF gathers are issued (D + 1 shared-memory stages).
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MRT, int D, int CTAS>
__global__ void __launch_bounds__(COLLIDE_TILE, CTAS) k_collide_odd_pipe(const Lattice<T> L, const int ntiles, const int bulk_skip) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NST = D + 1;
    typedef T Stage[38][COLLIDE_TILE];
    typedef int NbStage[18][COLLIDE_TILE];
    Stage* vals = reinterpret_cast<Stage*>(smem_raw);
    NbStage* nbS = reinterpret_cast<NbStage*>(smem_raw + sizeof(Stage) * NST);
    const int tid = threadIdx.x;
    const long long NC = L.NC;
    const int stride = gridDim.x;
    const T* __restrict__ p0 = L.pdf;

    // site id of my entry in a tile (-1: no such entry)
    auto site = [&](const int tile) -> int {
        const int t = tile * COLLIDE_TILE + tid;
        return (tile < ntiles && t < L.n_fluid) ? L.fl_u[t] : -1;
    };
    // The 18 neighbour cells of my entry.  load_index fetches the raw map entries (ordinary loads, consumed one
    // iteration later); resolve_index turns them into slot entries: the map entry itself for a non-solid neighbour, the
    // mailbox entry mb0 + rank for a wall link (core.cuh).  Both run in whole warps, `tile` is warp-uniform.
    const unsigned lanes_below = (1u << (tid & 31)) - 1u;
    auto load_index = [&](const int tile, const int u, int (&nb)[18], int& wb) {
        if (tile >= ntiles) return;
        const int lane = tid & 31;
        wb = lane < 18 ? L.wbase[(tile * (COLLIDE_TILE / 32) + (tid >> 5)) * 18 + lane] : 0;
#pragma unroll
        for (int q = 1; q < 19; q++) nb[q - 1] = u >= 0 ? L.cmap[u + L.off(q)] : 0;
    };
    auto resolve_index = [&](const int tile, int (&nb)[18], const int wb) {
        if (tile >= ntiles) return;
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int c = nb[q - 1];
            const unsigned walls = __ballot_sync(0xffffffffu, c < 0);
            const int base = __shfl_sync(0xffffffffu, wb, q - 1);
            if (c < 0) nb[q - 1] = L.mb0 + base + __popc(walls & lanes_below);
        }
    };
    // park the entries for the scatter and start the 38 gathers of that tile; always commits one group
    auto issue_gather = [&](const int tile, const int u, const int (&nb)[18], const int st) {
        if (u >= 0) {
            const int t = tile * COLLIDE_TILE + tid;
#pragma unroll
            for (int q = 1; q < 19; q++) nbS[st][q - 1][tid] = nb[q - 1];
#pragma unroll
            for (int q = 0; q < 19; q++) {
                const int src = (q == 0) ? t : nb[opc(q) - 1];   // x - e_q = x + e_opc(q), slot q there
                pipe::cp_async<sizeof(T)>(&vals[st][q][tid], p0 + (long long)q * NC + src);
                pipe::cp_async<sizeof(T)>(&vals[st][q + 19][tid], p0 + (long long)(q + 19) * NC + src);
            }
        }
        pipe::cp_async_commit();
    };

    int tile = blockIdx.x;
    // ring of my site ids: uR[j] belongs to tile + j*stride, j = 0 .. D+2
    int uR[D + 3];
#pragma unroll
    for (int j = 0; j < D + 3; j++) uR[j] = site(tile + j * stride);
    // prologue: gathers of my first D tiles, map entries of tile D
    int nbN[18], wbN = 0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        load_index(tile + j * stride, uR[j], nbN, wbN);
        resolve_index(tile + j * stride, nbN, wbN);
        issue_gather(tile + j * stride, uR[j], nbN, j % NST);
    }
    load_index(tile + D * stride, uR[D], nbN, wbN);
    T cnorm = uR[0] >= 0 ? L.c_norm[uR[0]] : T(0);

    int st = 0;   // stage of the current tile
    for (; tile < ntiles; tile += stride) {
        // 1. gathers of tile + D (its map entries arrived during the previous collision)
        int stD = st + D; if (stD >= NST) stD -= NST;
        resolve_index(tile + D * stride, nbN, wbN);
        issue_gather(tile + D * stride, uR[D], nbN, stD);
        // 2. map entries of tile + D + 1, c_norm of tile + 1, site id of tile + D + 3: land during this collision
        load_index(tile + (D + 1) * stride, uR[D + 1], nbN, wbN);
        const T cnormN = uR[1] >= 0 ? L.c_norm[uR[1]] : T(0);
        const int uNew = site(tile + (D + 3) * stride);
        // 3. this tile
        const int t = tile * COLLIDE_TILE + tid;
        const int u = uR[0];
        const bool live = u >= 0;
        T cnx = T(0), cny = T(0), cnz = T(0), tmp = T(0);
        if (live) node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);
        pipe::cp_async_wait<D>();   // all but the D most recent groups: the gathers of this tile have landed
        if (live) {
            T g1[19], g2[19];
#pragma unroll
            for (int q = 0; q < 19; q++) { g1[q] = vals[st][q][tid]; g2[q] = vals[st][q + 19][tid]; }
            const T phi_loc = collide_node<T, MRT>(L, g1, g2, cnx, cny, cnz, tmp);
            L.phi[u] = phi_loc;
            T* __restrict__ po = L.pdf;
            po[t] = g1[0];
            po[19 * NC + t] = g2[0];
#pragma unroll
            for (int q = 1; q < 19; q++) {
                const int dst = nbS[st][q - 1][tid];
                po[(long long)opc(q) * NC + dst] = g1[q];
                po[(long long)(opc(q) + 19) * NC + dst] = g2[q];
            }
        }
#pragma unroll
        for (int j = 0; j < D + 2; j++) uR[j] = uR[j + 1];
        uR[D + 2] = uNew;
        cnorm = cnormN;
        if (++st == NST) st = 0;
    }
    pipe::cp_async_wait<0>();
}

}  // namespace mflbm
