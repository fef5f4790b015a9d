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
constexpr size_t collide_odd_smem() { return (sizeof(T) * 38 * (D + 1) + sizeof(int) * 19 * (D + 2)) * COLLIDE_TILE; }

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
        T cnx = T(0), cny = T(0), cnz = T(0), tmp = T(0);
        if (live) node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);   // interface nodes only: cn + curvature stencil
        pipe::mbar_wait(&full[s], phase);
        // c_norm / site id of my next tiles, issued after the uses above (a load issued before them would share their
        // scoreboard and expose its full latency there); they land during the collision
        asm volatile("" ::: "memory");
        const T cnormN = uN >= 0 ? L.c_norm[uN] : T(0);
        const int uNN = site(tile + 2 * (int)gridDim.x);
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
// D = how many tiles ahead the PDF gathers are issued (D + 1 value stages, D + 2 index stages).
//
// Per iteration k of a CTA (tiles k, k+1, ... are the CTA's own tiles, `stride` apart):
//   wait     index(k+D) has landed                                   (cp.async group B of iteration k-1)
//   resolve  map entries of tile k+D -> slot entries (wall links -> mailbox entries), parked for the scatter
//   issue    index(k+D+1): 18 map entries per thread, cp.async 4 B   (group B_k)
//   issue    gathers(k+D): 38 PDFs per thread, cp.async 4/8 B        (group A_k)
//   wait     gathers(k) have landed;  collide tile k;  scatter through the parked entries of tile k
// Nothing a later tile needs is held in registers across the collision (the plain-load version kept 18 map entries
// live and ptxas tied constant-bank reloads to the scoreboard of the outstanding loads: 28 % of all stall samples).
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MRT, int D, int CTAS>
__global__ void __launch_bounds__(COLLIDE_TILE, CTAS) k_collide_odd_pipe(const Lattice<T> L, const int ntiles, const int bulk_skip) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NST = D + 1, NNB = D + 2;
    typedef T Stage[38][COLLIDE_TILE];
    typedef int NbStage[18][COLLIDE_TILE];
    Stage* vals = reinterpret_cast<Stage*>(smem_raw);
    NbStage* nbS = reinterpret_cast<NbStage*>(smem_raw + sizeof(Stage) * NST);
    typedef int WbStage[COLLIDE_TILE / 32][32];
    WbStage* wbS = reinterpret_cast<WbStage*>(smem_raw + sizeof(Stage) * NST + sizeof(NbStage) * NNB);
    const int tid = threadIdx.x;
    const long long NC = L.NC;
    const int stride = gridDim.x;
    const T* __restrict__ p0 = L.pdf;
    const int* __restrict__ cmap = L.cmap;
    const unsigned lanes_below = (1u << (tid & 31)) - 1u;

    // site id of my entry in a tile (-1: no such entry)
    auto site = [&](const int tile) -> int {
        const int t = tile * COLLIDE_TILE + tid;
        return (tile < ntiles && t < L.n_fluid) ? L.fl_u[t] : -1;
    };
    // raw map entries of my 18 neighbours -> my column of index stage sb, link ranks of my warp's 32-entry group ->
    // lanes 0..17; always commits one group
    auto issue_index = [&](const int tile, const int u, const int sb) {
        if (tile < ntiles && (tid & 31) < 18)
            pipe::cp_async<4>(&wbS[sb][tid >> 5][tid & 31], L.wbase + ((tile * (COLLIDE_TILE / 32) + (tid >> 5)) * 18 + (tid & 31)));
        if (u >= 0) {
#pragma unroll
            for (int q = 1; q < 19; q++) pipe::cp_async<4>(&nbS[sb][q - 1][tid], cmap + (u + L.off(q)));
        }
        pipe::cp_async_commit();
    };
    // slot entries of the 18 neighbour cells: the map entry itself for a non-solid neighbour, the mailbox entry
    // mb0 + rank for a wall link (core.cuh).  Whole warps (`tile` is warp-uniform); parked in place for the scatter.
    auto resolve_index = [&](const int tile, const int u, const int sb, int (&nb)[18]) {
        if (tile >= ntiles) return;
        const int wb = (tid & 31) < 18 ? wbS[sb][tid >> 5][tid & 31] : 0;
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int c = u >= 0 ? nbS[sb][q - 1][tid] : 0;
            const unsigned walls = __ballot_sync(0xffffffffu, c < 0);
            const int base = __shfl_sync(0xffffffffu, wb, q - 1);
            nb[q - 1] = c >= 0 ? c : L.mb0 + base + __popc(walls & lanes_below);
            if (c < 0) nbS[sb][q - 1][tid] = nb[q - 1];
        }
    };
    // the 38 gathers of a tile; always commits one group
    auto issue_gather = [&](const int tile, const int u, const int (&nb)[18], const int st) {
        if (u >= 0) {
            const int t = tile * COLLIDE_TILE + tid;
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
    // prologue: index + gathers of my first D tiles (groups B, A per tile, as in the loop), index of tile D
    int nb[18];
#pragma unroll
    for (int j = 0; j < D; j++) {
        issue_index(tile + j * stride, uR[j], j % NNB);
        pipe::cp_async_wait<0>();
        resolve_index(tile + j * stride, uR[j], j % NNB, nb);
        issue_gather(tile + j * stride, uR[j], nb, j % NST);
    }
    issue_index(tile + D * stride, uR[D], D % NNB);
    pipe::cp_async_commit();   // keeps the group pattern of the loop: (B, A) per iteration
    T cnorm = uR[0] >= 0 ? L.c_norm[uR[0]] : T(0);

    int st = 0, sb = 0;   // value / index stage of the current tile
    for (; tile < ntiles; tile += stride) {
        int stD = st + D; if (stD >= NST) stD -= NST;
        int sbD = sb + D; if (sbD >= NNB) sbD -= NNB;
        int sbD1 = sbD + 1; if (sbD1 >= NNB) sbD1 -= NNB;
        pipe::cp_async_wait<1>();   // index(tile + D) has landed (all but the newest group, the gathers of tile + D - 1)
        resolve_index(tile + D * stride, uR[D], sbD, nb);
        issue_index(tile + (D + 1) * stride, uR[D + 1], sbD1);
        issue_gather(tile + D * stride, uR[D], nb, stD);
        const int t = tile * COLLIDE_TILE + tid;
        const int u = uR[0];
        const bool live = u >= 0;
        T cnx = T(0), cny = T(0), cnz = T(0), tmp = T(0);
        if (live) node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);
        pipe::cp_async_wait<2 * D>();   // all but the 2D newest groups: the gathers of this tile have landed
        // c_norm of tile + 1 and site id of tile + D + 3: ordinary loads that land during this collision.  Issued only
        // now: a load issued before the uses above would share their scoreboard and expose its full latency there.
        asm volatile("" ::: "memory");
        const T cnormN = uR[1] >= 0 ? L.c_norm[uR[1]] : T(0);
        const int uNew = site(tile + (D + 3) * stride);
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
                const int dst = nbS[sb][q - 1][tid];
                po[(long long)opc(q) * NC + dst] = g1[q];
                po[(long long)(opc(q) + 19) * NC + dst] = g2[q];
            }
        }
#pragma unroll
        for (int j = 0; j < D + 2; j++) uR[j] = uR[j + 1];
        uR[D + 2] = uNew;
        cnorm = cnormN;
        if (++st == NST) st = 0;
        if (++sb == NNB) sb = 0;
    }
    pipe::cp_async_wait<0>();
}

}  // namespace mflbm
