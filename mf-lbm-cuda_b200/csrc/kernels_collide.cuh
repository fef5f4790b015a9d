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

// kernels_chain.cuh: flags of the interface-activity map, raised where phi is produced
template <typename T> __device__ __forceinline__ void raise_activity(const Lattice<T>& L, const int t, const T phi, const unsigned mask);

// surface-tension inputs of one node (:143-150): cn and 0.5*gamma*curv*c_norm
template <typename T>
__device__ __forceinline__ void node_force(const Lattice<T>& L, const int u, const T cnorm, const bool bulk_skip, T& cnx, T& cny, T& cnz, T& tmp) {
    if (bulk_skip && cnorm == T(0)) { cnx = T(0); cny = T(0); cnz = T(0); tmp = T(0); return; }
    cnx = L.cn_x[u]; cny = L.cn_y[u]; cnz = L.cn_z[u];
    tmp = lit<T>(0.5) * L.lbm_gamma * curvature_at(L, u) * cnorm;   // :147
}


// ---------------------------------------------------------------------------------------------------------
// EVEN step: f_q = local slot opc(q); collide; local slot q = f_q*     (:395-726)
// ---------------------------------------------------------------------------------------------------------
// Block = 4 consumer warps (one thread per entry of a 128-entry tile) + 1 producer warp.  Per stage two mbarriers:
//   full[s]   1 arrival (the producer's expect_tx) + 38 rows * 128 * sizeof(T) bytes of TMA traffic
//   empty[s]  4 arrivals, one per consumer warp once its lanes hold their columns in registers
// so a stage is refilled as soon as its last reader is done and no CTA-wide barrier sits between the tiles (the
// __syncthreads of the first version took 19 % / 26 % of all stall samples, profiles/r01d_collide_*_stalls.txt).
constexpr int COLLIDE_EVEN_THREADS = COLLIDE_TILE + 32;
template <typename T, int NST>
constexpr size_t collide_even_smem() { return sizeof(T) * NST * 38 * COLLIDE_TILE + 16 * NST + 2 * (sizeof(T) + sizeof(int)) * COLLIDE_TILE; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pipe::smem_u32(bar)) : "memory");
}

template <typename T, int MRT, int NST, int CTAS>
__global__ void __launch_bounds__(COLLIDE_EVEN_THREADS, CTAS) k_collide_even_tma(const Lattice<T> L, const int tile0, const int ntiles, const int bulk_skip) {   // tiles [tile0, ntiles)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef Pair<T> Stage[19][COLLIDE_TILE];
    Stage* buf = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * NST);
    uint64_t* empty = full + NST;
    T* cS = reinterpret_cast<T*>(empty + NST);                 // [2][tile] c_norm of my entry in the next tile
    int* uS = reinterpret_cast<int*>(cS + 2 * COLLIDE_TILE);   // [2][tile] site id of my entry two tiles ahead
    const int tid = threadIdx.x;
    const int stride = gridDim.x;
    const long long NC = L.NC;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) { pipe::mbar_init(&full[s], 1); pipe::mbar_init(&empty[s], COLLIDE_TILE / 32); }
        pipe::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= COLLIDE_TILE) {
        // ---- producer warp: lane l < 19 copies the row of direction l (both components, 128 pairs) of every tile of this CTA ----
        const int lane = tid - COLLIDE_TILE;
        int s = 0;
        uint32_t phase = 0;   // parity of the empty-phase a refill of stage s has to see completed
        int it = 0;
        for (int tile = tile0 + blockIdx.x; tile < ntiles; tile += stride, it++) {
            if (it >= NST) pipe::mbar_wait(&empty[s], phase);
            if (lane == 0) pipe::mbar_expect_tx(&full[s], (uint32_t)sizeof(Stage));
            if (lane < 19) pipe::bulk_g2s(&buf[s][lane][0], L.pairs(lane) + (long long)tile * COLLIDE_TILE, (uint32_t)(sizeof(Pair<T>) * COLLIDE_TILE), &full[s]);
            if (++s == NST) { s = 0; if (it >= NST) phase ^= 1; }
        }
        return;
    }

    // ---- consumer warps ----
    auto site = [&](const int tl) -> int {
        const int t = tl * COLLIDE_TILE + tid;
        return (tl < ntiles && t < L.n_fluid) ? L.fl_u[t] : -1;
    };
    // site id two tiles ahead and c_norm of the next tile's site: cp.async, so that no scoreboard is outstanding when the
    // collision calls the division subroutine (a CALL waits for all of them; 11.5 % of the stall samples)
    auto prefetch = [&](const int tile_site, const int u_cn, const int slot) {
        if (tile_site < ntiles && tile_site * COLLIDE_TILE + tid < L.n_fluid) pipe::cp_async<4>(&uS[slot * COLLIDE_TILE + tid], L.fl_u + (tile_site * COLLIDE_TILE + tid));
        if (u_cn >= 0) pipe::cp_async<sizeof(T)>(&cS[slot * COLLIDE_TILE + tid], L.c_norm + u_cn);
        pipe::cp_async_commit();
    };
    int tile = tile0 + blockIdx.x;
    int u = site(tile), uN = site(tile + stride);
    prefetch(tile + 2 * stride, u, 0);   // c_norm of the first tile, site of the third
    int s = 0, slot = 0;                 // slot: the pair (cS, uS) half this iteration reads; the other half is being filled
    uint32_t phase = 0;
    for (; tile < ntiles; tile += stride) {
        const int t = tile * COLLIDE_TILE + tid;
        const bool live = u >= 0;
        pipe::cp_async_wait<0>();
        const T cnorm = live ? cS[slot * COLLIDE_TILE + tid] : T(0);
        int uNN = -1;
        { const int tl = tile + 2 * stride; if (tl < ntiles && tl * COLLIDE_TILE + tid < L.n_fluid) uNN = uS[slot * COLLIDE_TILE + tid]; }
        slot ^= 1;
        prefetch(tile + 3 * stride, uN, slot);   // lands during this collision
        T cnx = T(0), cny = T(0), cnz = T(0), tmp = T(0);
        if (live) node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);   // interface nodes only: cn + curvature stencil
        pipe::mbar_wait(&full[s], phase);
        T g1[19], g2[19];
#pragma unroll
        for (int q = 0; q < 19; q++) { const Pair<T> v = buf[s][opc(q)][tid]; g1[q] = v.a; g2[q] = v.b; }
        // The columns must BE in registers before the stage is released: an LDS that is merely issued can still be queued
        // in the shared-memory pipe when the arrive below becomes visible, and the refill then overwrites what it was
        // about to read (seen once per ~10^5 warp-tiles at 256^3).  An empty asm that consumes the values makes the
        // compiler wait for their scoreboard here.
        // Handing the stage back.  An LDS that is merely ISSUED can still be pending when a later mbarrier arrive becomes
        // visible: releasing right after the loads let the refill overwrite columns that had not been read yet (about
        // one warp-tile in 10^5 at 256^3; compute-sanitizer racecheck reports the same pair).  The release therefore
        // follows the store of phi, the first instruction that needs all 38 values: STG(phi) cannot issue before the loads
        // have returned, and the arrive (release semantics) cannot pass the store.
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        if (live) {
            uint64_t* const bar = &empty[s];
            const int leader = __ffs(live_mask) - 1;
            const T phi_loc = collide_node<T, MRT>(L, g1, g2, cnx, cny, cnz, tmp, [&](const T phi_now) {
                L.phi[u] = phi_now;
                __syncwarp(live_mask);
                if ((tid & 31) == leader) mbar_arrive(bar);
            });
            Pair<T>* __restrict__ po = L.pairs(0) + t;
#pragma unroll
            for (int q = 0; q < 19; q++) po[(long long)q * NC] = Pair<T>{g1[q], g2[q]};
            if (L.grp_p) raise_activity(L, t, phi_loc, live_mask);   // after the stage has gone back and the stores are on their way
        } else if (live_mask == 0u && (tid & 31) == 0) {
            mbar_arrive(&empty[s]);   // a warp past the last fluid entry
        }
        u = uN; uN = uNN;
        if (++s == NST) { s = 0; phase ^= 1; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// ODD step: f_q pulled from x - e_q (direction row q); collide; f_q* pushed to x + e_q (row opc(q))     (:56-388)
// Neighbours are found through the site map, so rows are gathered per thread (no bulk copies).  Warp-specialised: a first
// version had one warp per scheduler (255 registers, one CTA per SM) issue everything - map look-ups, link ranks, gather
// addresses, the collision, scatter addresses - and ncu showed 0.27 eligible warps per scheduler with no dominant stall.
// Here a CTA is 4 consumer warps + 4 producer warps over the same 128-entry tiles:
//   producer thread p, tile i :  raw map entries (plain loads, one tile ahead) -> slot entries of the 18 neighbour cells
//                                (wall links -> mailbox entries) -> stage.nb / stage.u ;  38 PDF gathers + c_norm with
//                                cp.async into stage.val / stage.cn ;  completion is reported to full[stage] by
//                                cp.async.mbarrier.arrive (data) and a plain arrive (the st.shared of nb / u)
//   consumer thread p, tile i :  wait full[stage] -> collide -> scatter through stage.nb -> arrive empty[stage]
// The consumer's instruction stream shrinks by a third and the producer warp fills its issue slots.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct OddStage {
    Pair<T> val[19][COLLIDE_TILE];
    int nb[18][COLLIDE_TILE];
    T cn[COLLIDE_TILE];
    int u[COLLIDE_TILE];
};
template <typename T, int NST>
constexpr size_t collide_odd_ws_smem() { return sizeof(OddStage<T>) * NST + 16 * NST; }

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(pipe::smem_u32(bar)) : "memory");
}

// NCONS consumer warpgroups share one producer warpgroup (the producer needs about half a consumer's time per tile):
// consumer group g takes the CTA's tiles i = g, g + NCONS, ...; stage of tile i = i % NST for everybody.  With two
// consumer groups each scheduler holds two collision warps.  In double precision a collision warp needs ~216 registers,
// more than 64 K / 384 threads: REGS_P / REGS_C move registers from the producer to the consumers with setmaxnreg
// (0 = leave the allocation alone).
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <typename T, int MRT, int NST, int CTAS, int NCONS, int REGS_P, int REGS_C>
__global__ void __launch_bounds__((NCONS + 1) * COLLIDE_TILE, CTAS) k_collide_odd_ws(const Lattice<T> L, const int tile0, const int ntiles, const int bulk_skip) {   // tiles [tile0, ntiles)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef OddStage<T> Stage;
    Stage* stg = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * NST);
    uint64_t* empty = full + NST;
    const int tid = threadIdx.x;
    const int stride = gridDim.x;
    const long long NC = L.NC;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) { pipe::mbar_init(&full[s], 2 * COLLIDE_TILE); pipe::mbar_init(&empty[s], COLLIDE_TILE / 32); }
        pipe::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NCONS * COLLIDE_TILE) {
        // =========================== producer ===========================
        // The map entries, link ranks and site ids of the NEXT tile are ordinary loads into registers, issued one tile
        // ahead (the producer never calls a subroutine, so nothing waits on their scoreboards early): an LDG costs the
        // load/store pipe 1.8 cycles against 8 for an LDGSTS.
        if (REGS_P > 0) reg_dealloc<REGS_P>();
        const int p = tid - NCONS * COLLIDE_TILE, lane = p & 31, w = p >> 5;
        const int* __restrict__ cmap = L.cmap;
        const unsigned lanes_below = (1u << lane) - 1u;
        auto site_of = [&](const int tl) -> int { return (tl < ntiles && tl * COLLIDE_TILE + p < L.n_fluid) ? __ldg(L.fl_u + (tl * COLLIDE_TILE + p)) : -1; };
        auto load_raw = [&](const int tl, const int u, int (&c)[18], int& wb) {
            wb = (tl < ntiles && lane < 18) ? __ldg(L.wbase + ((tl * (COLLIDE_TILE / 32) + w) * 18 + lane)) : 0;
#pragma unroll
            for (int q = 1; q < 19; q++) c[q - 1] = u >= 0 ? __ldg(cmap + (u + L.off(q))) : 0;
        };
        int tile = tile0 + blockIdx.x;
        int uCur = site_of(tile), uNxt = site_of(tile + stride);
        int cN[18], wbN;
        load_raw(tile, uCur, cN, wbN);
        int st = 0;
        uint32_t ph_empty = 0;
        for (int i = 0; tile < ntiles; tile += stride, i++) {
            int c[18];
#pragma unroll
            for (int q = 0; q < 18; q++) c[q] = cN[q];
            const int wb = wbN;
            const int uNxt2 = site_of(tile + 2 * stride);
            load_raw(tile + stride, uNxt, cN, wbN);   // consumed in the next iteration
            if (i >= NST) pipe::mbar_wait(&empty[st], ph_empty);
            Stage& S = stg[st];
            const int t = tile * COLLIDE_TILE + p;
            const bool live = uCur >= 0;
            if (live) {   // the rest populations: direction 0 at the node itself
                pipe::cp_async<sizeof(Pair<T>)>(&S.val[0][p], L.pairs(0) + t);
                pipe::cp_async<sizeof(T)>(&S.cn[p], L.c_norm + uCur);
            }
            // slot entries of the 18 neighbour cells: the map entry itself for a non-solid neighbour, the mailbox entry
            // mb0 + rank for a wall link (core.cuh).  Neighbour j+1 is where slot opc(j+1) is pulled from: its two
            // gathers go out as soon as the entry is known, nothing is kept.
#pragma unroll
            for (int q = 1; q < 19; q++) {
                const int cc = c[q - 1];
                const unsigned walls = __ballot_sync(0xffffffffu, cc < 0);
                const int base = __shfl_sync(0xffffffffu, wb, q - 1);
                const int nbq = cc >= 0 ? cc : L.mb0 + base + __popc(walls & lanes_below);
                S.nb[q - 1][p] = nbq;
                if (live) pipe::cp_async<sizeof(Pair<T>)>(&S.val[opc(q)][p], L.pairs(opc(q)) + nbq);   // x - e_s = x + e_q for s = opc(q): row s there
            }
            S.u[p] = uCur;
            mbar_arrive(&full[st]);                  // release: nb / u are visible to whoever sees the phase complete
            cp_async_mbar_arrive_noinc(&full[st]);   // fires when every cp.async above has landed
            uCur = uNxt; uNxt = uNxt2;
            if (++st == NST) { st = 0; if (i >= NST) ph_empty ^= 1; }
        }
        pipe::cp_async_commit();
        pipe::cp_async_wait<0>();
        return;
    }

    // =========================== consumers ===========================
    if (REGS_C > 0) reg_alloc<REGS_C>();
    const int cg = tid / COLLIDE_TILE, ct = tid - cg * COLLIDE_TILE;   // consumer group, entry within the tile
    for (int i = cg; ; i += NCONS) {
        const int tile = tile0 + blockIdx.x + i * stride;
        if (tile >= ntiles) break;
        const int st = i % NST;
        Stage& S = stg[st];
        pipe::mbar_wait(&full[st], (uint32_t)((i / NST) & 1));
        const int t = tile * COLLIDE_TILE + ct;
        const int u = S.u[ct];
        const bool live = u >= 0;
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        T phi_keep = T(0);
        if (live) {
            const T cnorm = S.cn[ct];
            T cnx, cny, cnz, tmp;
            node_force(L, u, cnorm, bulk_skip != 0, cnx, cny, cnz, tmp);
            T g1[19], g2[19];
#pragma unroll
            for (int q = 0; q < 19; q++) { const Pair<T> v = S.val[q][ct]; g1[q] = v.a; g2[q] = v.b; }
            const T phi_loc = collide_node<T, MRT>(L, g1, g2, cnx, cny, cnz, tmp);
            L.phi[u] = phi_loc;
            L.pairs(0)[t] = Pair<T>{g1[0], g2[0]};
#pragma unroll
            for (int q = 1; q < 19; q++) {
                const int dst = S.nb[q - 1][ct];
                L.pairs(opc(q))[dst] = Pair<T>{g1[q], g2[q]};
            }
            phi_keep = phi_loc;
        }
        // every value and every entry of the stage has been consumed by an issued store (see the even kernel)
        __syncwarp();
        if ((ct & 31) == 0) mbar_arrive(&empty[st]);
        if (live && L.grp_p) raise_activity(L, t, phi_keep, live_mask);
    }
}

}  // namespace mflbm
