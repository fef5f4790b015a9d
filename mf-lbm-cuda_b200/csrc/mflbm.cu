// mflbm.cu — Solver<T> (device state, per-step sequencing, CUDA-graph replay) and the C ABI of include/mflbm.h.
// One translation unit, both precisions.  Build: see mf-lbm-cuda_b200/csrc/Makefile (sm_100a, -lineinfo).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mflbm.h"
#include "core.cuh"
#include "kernels_step.cuh"
#include "kernels_collide.cuh"
#include "kernels_chain.cuh"
#include "kernels_aux.cuh"
#include "scan.cuh"
#include "kernels_setup.cuh"

namespace mflbm {

static thread_local std::string g_last_error;

struct Error {
    std::string msg;
};
#define MF_FAIL(...)                                        \
    do {                                                    \
        char buf_[512];                                     \
        snprintf(buf_, sizeof buf_, __VA_ARGS__);           \
        throw Error{std::string(buf_)};                     \
    } while (0)
#define MF_CUDA(x)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) MF_FAIL("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #x); \
    } while (0)

template <typename T> struct ParamsOf;
template <> struct ParamsOf<float> { using type = mflbm_f32_params; };
template <> struct ParamsOf<double> { using type = mflbm_f64_params; };

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

template <typename T>
struct Solver {
    using Params = typename ParamsOf<T>::type;
    Params P{};
    mflbm_slab slab{};
    bool is_slab = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Lattice<T> L{};
    long long N1 = 0, N2 = 0, N4 = 0, NP = 0, PN = 0, NC = 0;   // cells of 1/2/4-ghost arrays, of an (NX1*NY1) plane, of the U grid, of a PDF slot
    // device state (layouts: core.cuh)
    T *d_pdf = nullptr, *d_phi = nullptr, *d_cnx = nullptr, *d_cny = nullptr, *d_cnz = nullptr, *d_cnorm = nullptr;
    T *d_Win = nullptr, *d_fconv = nullptr, *d_gconv = nullptr, *d_phiconv = nullptr;
    // device geometry
    signed char* d_types = nullptr;
    int *d_cmap = nullptr, *d_flu = nullptr, *d_zstart = nullptr, *d_wbase = nullptr;
    long long n_links = 0;   // wall links (core.cuh)
    int *d_list_phi = nullptr, *d_mask_phi = nullptr, *d_list_cn = nullptr, *d_mask_cn = nullptr, *d_list_alter = nullptr, *d_list_n = nullptr;
    T* d_sn[3] = {nullptr, nullptr, nullptr};   // solid-surface normals in d_list_alter order
    unsigned char *d_live_n = nullptr, *d_live_cn = nullptr;   // per list entry: outputs may be non-zero (k_normals, k_extrap_cn)
    unsigned char* d_near = nullptr;                           // per U site: a neighbour got a non-zero normal in this chain (kernels_step.cuh)
    // gradient chain evaluated brick by brick where an interface can be (kernels_chain.cuh; MFLBM_CHAIN=list selects the
    // four list kernels of kernels_step.cuh instead, MFLBM_ACT_SCAN=1 takes the brick flags from a scan of phi instead of the
    // collide kernels - both for cross-checks, results are identical)
    bool brick_chain = true, act_scan = false;
    bool chain_csr = true;               // k_chain_normals_csr (per-brick CSR of the box's solid-boundary sites); MFLBM_CHAIN=brick: k_chain_normals (sites collected per step)
    int *d_bx_start = nullptr, *d_bx_ent = nullptr;   // solid-boundary sites of the 10 x 6 x 6 box of every brick (k_brick_box_sites)
    BrickGrid bricks{};
    unsigned char* d_act = nullptr;      // [3 sets][P | M][bricks]: sets 0 / 1 are raised by the collide kernel of even / odd steps, set 2 by k_act_scan
    unsigned char* d_quiet = nullptr;    // [bricks] verdict of the previous chain
    int *d_active = nullptr, *d_n_active = nullptr;   // bricks to process in this chain; {their number, their cn-extrapolation entries} (one 64-bit counter)
    unsigned* d_chain_live = nullptr;                 // [bricks][4 warps] bit per site: outputs of k_chain_normals_csr may be non-zero in memory
    int* d_ent_off = nullptr;                         // [slot] first entry of the listed brick in the flat numbering of k_chain_extrap_cn_flat
    int n_sb = 0;                                     // entries of d_sb_list
    int *d_shell = nullptr, n_shell = 0;              // non-solid sites outside the real box
    int *d_bc_list = nullptr, *d_bc_mask = nullptr, n_bc = 0;   // solid-boundary sites whose phi a boundary kernel copies (k_chain_pre)
    int* d_alt_start = nullptr;                       // [bricks + 1] first fluid-boundary entry of every brick (d_list_alter, d_sn)
    int *d_sb_start = nullptr, *d_sb_list = nullptr, *d_sb_mask = nullptr;   // solid-boundary sites of [0..n+1]^3 brick by brick (k_chain_extrap_cn)
    unsigned char* d_grp = nullptr;                   // [P | M][groups of 32 fluid entries]: raised by the collide kernels, spread and cleared by k_chain_pre
    int *d_grp_start = nullptr, *d_grp_bricks = nullptr, n_groups = 0;     // bricks touched by every group (CSR)
    T* d_phi_base = nullptr;                          // allocations of phi / types with 16 guard elements on both sides
    signed char* d_types_base = nullptr;
    CUtensorMap tm_phi{};                             // 3-D tensor map of phi over the U grid, box 16 x 8 x 8 (k_chain_normals)
    int n_list_phi = 0, n_list_cn = 0, n_list_alter = 0, n_list_alter_all = 0, n_list_n = 0;
    long long counts[4] = {0, 0, 0, 0};
    long long n_fluid = 0;
    long long n_boundary = 0;   // fluid nodes of the columns that face a neighbour slab: the first entries of the fluid order
    long long launches = 0;
    bool have_geometry = false;
    // staging buffer for layout conversion at the boundary
    void* d_stage = nullptr;
    size_t stage_bytes = 0;
    // state upload / download: a ring of RING_NB staging buffers; the PCIe copies run on copy_stream, the layout kernels
    // on `stream`, chained with events, so that array n+1 crosses the bus while array n is being permuted (the first
    // version did copy -> kernel -> host synchronisation 38 times through one buffer)
    static constexpr int RING_NB = 3;
    void* d_ring[RING_NB] = {nullptr, nullptr, nullptr};
    size_t ring_bytes = 0;
    int ring_pos = 0;
    bool ring_used[RING_NB] = {false, false, false};
    cudaEvent_t ev_ring_copy[RING_NB] = {nullptr, nullptr, nullptr}, ev_ring_kernel[RING_NB] = {nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    // monitor
    double* d_mon = nullptr;
    double* h_mon = nullptr;
    T* d_phi_old = nullptr;   // steady-state monitor 2 (src/Monitor.cpp:279), per fluid entry, allocated on first use
    // halo buffers: [kind 0..2][side 0..1]
    T* d_send[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    T* d_recv[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    // halo messages through peer memory (kernels_aux.cuh): my incoming flags [kind*2+side], block tickets [8 + ...] and
    // message counters [16 + ...]; the neighbours' receive buffers / flags as peer pointers
    cudaStream_t bc_stream = nullptr;           // second lane of a step: distribution part of the boundary kernels next to the gradient chain
    cudaEvent_t ev_bc_fork = nullptr, ev_bc_join = nullptr;
    cudaStream_t aux_stream = nullptr;          // second lane for the right-hand neighbour's messages (fork/join with events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* d_face[4] = {nullptr, nullptr, nullptr, nullptr};   // slot entries of columns 0, 1, nx, nx+1 over the y-z plane (k_halo_pdf)
    unsigned char* d_p2p = nullptr;
    size_t p2p_bytes = 0;
    unsigned* d_flags = nullptr;
    T* peer_recv[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    unsigned* peer_flag[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    // CUDA graph of one (odd, even) or (even, odd) step pair
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    long long pair_launches = 0;
    bool no_overlap = false;   // MFLBM_NO_OVERLAP=1: halo messages after the whole collide launch (the pre-overlap schedule, for comparison)
    int variant = 0;   // kernel schedule variant (MFLBM_VARIANT, tuning only; results are identical)
    int max_ctas = 0;  // MFLBM_MAX_CTAS: cap on the persistent collide grids (tests only)
    int num_sms = 148;
    // true when c_norm == 0 implies cn_* == 0 at every fluid node: established by k_normals (every gradient_chain), not
    // guaranteed for arrays handed in through upload_state.  Lets the collide kernels skip cn/curvature in the bulk.
    bool cn_consistent = false;

    // ------------------------------------------------------------------------------------------------
    void zalloc(void** ptr, size_t bytes) {
        MF_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
        MF_CUDA(cudaMemsetAsync(*ptr, 0, bytes, stream));
    }
    template <typename X> void dfree(X*& q) { if (q) { cudaFree(q); q = nullptr; } }

    void* stage(size_t bytes) {
        if (bytes > stage_bytes) {
            if (d_stage) { MF_CUDA(cudaStreamSynchronize(stream)); cudaFree(d_stage); d_stage = nullptr; stage_bytes = 0; }
            MF_CUDA(cudaMalloc(&d_stage, bytes));
            stage_bytes = bytes;
        }
        return d_stage;
    }
    void drop_stage() { if (d_stage) { cudaStreamSynchronize(stream); cudaFree(d_stage); d_stage = nullptr; stage_bytes = 0; } }

    void ring_reserve(size_t bytes) {
        if (!copy_stream) {
            MF_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
            for (int b = 0; b < RING_NB; b++) {
                MF_CUDA(cudaEventCreateWithFlags(&ev_ring_copy[b], cudaEventDisableTiming));
                MF_CUDA(cudaEventCreateWithFlags(&ev_ring_kernel[b], cudaEventDisableTiming));
            }
        }
        if (bytes <= ring_bytes) return;
        ring_drain(true);
        for (int b = 0; b < RING_NB; b++) MF_CUDA(cudaMalloc(&d_ring[b], bytes));
        ring_bytes = bytes;
    }
    // host array -> staging buffer (copy_stream) -> launch(staging pointer) on `stream`; returns without waiting
    template <typename F>
    void ring_upload(const void* host, size_t bytes, F&& launch) {
        const int b = ring_pos;
        ring_pos = (ring_pos + 1) % RING_NB;
        if (ring_used[b]) MF_CUDA(cudaStreamWaitEvent(copy_stream, ev_ring_kernel[b], 0));   // its last reader has finished
        MF_CUDA(cudaMemcpyAsync(d_ring[b], host, bytes, cudaMemcpyHostToDevice, copy_stream));
        MF_CUDA(cudaEventRecord(ev_ring_copy[b], copy_stream));
        MF_CUDA(cudaStreamWaitEvent(stream, ev_ring_copy[b], 0));
        launch(d_ring[b]);
        MF_CUDA(cudaEventRecord(ev_ring_kernel[b], stream));
        ring_used[b] = true;
    }
    // launch(staging pointer) on `stream` -> staging buffer -> host array (copy_stream)
    template <typename F>
    void ring_download(void* host, size_t bytes, F&& launch) {
        const int b = ring_pos;
        ring_pos = (ring_pos + 1) % RING_NB;
        if (ring_used[b]) MF_CUDA(cudaStreamWaitEvent(stream, ev_ring_copy[b], 0));   // its last copy has left
        launch(d_ring[b]);
        MF_CUDA(cudaEventRecord(ev_ring_kernel[b], stream));
        MF_CUDA(cudaStreamWaitEvent(copy_stream, ev_ring_kernel[b], 0));
        MF_CUDA(cudaMemcpyAsync(host, d_ring[b], bytes, cudaMemcpyDeviceToHost, copy_stream));
        MF_CUDA(cudaEventRecord(ev_ring_copy[b], copy_stream));
        ring_used[b] = true;
    }
    void ring_drain(bool release) {
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (stream) cudaStreamSynchronize(stream);
        for (int b = 0; b < RING_NB; b++) ring_used[b] = false;
        ring_pos = 0;
        if (release) { for (int b = 0; b < RING_NB; b++) if (d_ring[b]) { cudaFree(d_ring[b]); d_ring[b] = nullptr; } ring_bytes = 0; }
    }

    void create(const Params* p, const mflbm_slab* sl, int dev, void* strm) {
        device = dev;
        MF_CUDA(cudaSetDevice(device));   // before any stream / event is created: they belong to the current device
        if (const char* v = getenv("MFLBM_VARIANT")) variant = atoi(v);
        if (const char* v = getenv("MFLBM_MAX_CTAS")) max_ctas = atoi(v);
        if (const char* v = getenv("MFLBM_CHAIN")) { brick_chain = strcmp(v, "list") != 0; chain_csr = brick_chain && strcmp(v, "brick") != 0; }
        if (const char* v = getenv("MFLBM_ACT_SCAN")) act_scan = atoi(v) != 0;
        if (const char* v = getenv("MFLBM_NO_OVERLAP")) no_overlap = atoi(v) != 0;
        MF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        if (strm) { stream = (cudaStream_t)strm; own_stream = false; }
        else { MF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); own_stream = true; }
        MF_CUDA(cudaStreamCreateWithFlags(&bc_stream, cudaStreamNonBlocking));
        MF_CUDA(cudaEventCreateWithFlags(&ev_bc_fork, cudaEventDisableTiming));
        MF_CUDA(cudaEventCreateWithFlags(&ev_bc_join, cudaEventDisableTiming));
        if (p->nx < 3 || p->ny < 3 || p->nz < 3) MF_FAIL("lattice too small");
        if (p->iper != 0) MF_FAIL("x-periodic boundaries are not supported (reference: src/IO_multiphase.cpp:210)");
        if (p->mrt < 1 || p->mrt > 4) MF_FAIL("mrt must be 1..4");
        is_slab = sl != nullptr;
        if (is_slab) slab = *sl;
        else { slab.x0 = 1; slab.nx_local = p->nx; slab.has_left = 0; slab.has_right = 0; }
        if (slab.nx_local < 4 && is_slab) MF_FAIL("slab narrower than the 4-column phi halo");
        L.nx = (int)slab.nx_local; L.ny = (int)p->ny; L.nz = (int)p->nz;
        L.x0 = (int)slab.x0; L.nx_global = (int)p->nx;
        L.NX1 = L.nx + 2; L.NY1 = L.ny + 2; L.NZ1 = L.nz + 2;
        L.PX = 16 * ceil_div(L.nx + 8, 16); L.PY = L.ny + 8; L.PZ = L.nz + 8;
        N1 = (long long)L.NX1 * L.NY1 * L.NZ1;
        N2 = (long long)(L.nx + 4) * (L.ny + 4) * (L.nz + 4);
        N4 = (long long)(L.nx + 8) * (L.ny + 8) * (L.nz + 8);
        NP = (long long)L.NX1 * L.NY1;
        PN = (long long)L.PX * L.PY * L.PZ;
        if (PN >= (1LL << 31)) MF_FAIL("lattice (or slab) exceeds 2^31 cells per field; decompose into more slabs");
        L.sy = L.PX; L.sz = L.PX * L.PY; L.NC = 0; L.n_fluid = 0;   // the PDF slots are sized by the geometry (finish_geometry)
        zalloc((void**)&d_phi_base, sizeof(T) * (PN + 32)); d_phi = d_phi_base + 16;
        zalloc((void**)&d_cnx, sizeof(T) * PN); zalloc((void**)&d_cny, sizeof(T) * PN); zalloc((void**)&d_cnz, sizeof(T) * PN);
        zalloc((void**)&d_cnorm, sizeof(T) * PN);
        zalloc((void**)&d_Win, sizeof(T) * NP);
        zalloc((void**)&d_fconv, sizeof(T) * NP * 19); zalloc((void**)&d_gconv, sizeof(T) * NP * 19); zalloc((void**)&d_phiconv, sizeof(T) * NP);
        zalloc((void**)&d_types_base, PN + 32); d_types = d_types_base + 16;
        zalloc((void**)&d_cmap, sizeof(int) * PN);
        bricks.nbx = L.PX / BR_X; bricks.nby = ceil_div(L.PY, BR_Y); bricks.nbz = ceil_div(L.PZ, BR_Z);
        L.nbx = bricks.nbx; L.nby = bricks.nby; L.nbz = bricks.nbz;
        L.dv_sz = make_fastdiv((unsigned)(L.PX * L.PY)); L.dv_px = make_fastdiv((unsigned)L.PX);
        L.dv_nbxy = make_fastdiv((unsigned)(bricks.nbx * bricks.nby)); L.dv_nbx = make_fastdiv((unsigned)bricks.nbx);
        L.act_p = nullptr; L.act_m = nullptr; L.grp_p = nullptr; L.grp_m = nullptr;
        if (brick_chain) {
            make_tile_map(&tm_phi, d_phi, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sizeof(T));
            const size_t nb = (size_t)bricks.count();
            zalloc((void**)&d_act, 6 * nb); zalloc((void**)&d_quiet, nb);
            zalloc((void**)&d_active, sizeof(int) * nb); zalloc((void**)&d_n_active, 2 * sizeof(int)); zalloc((void**)&d_ent_off, sizeof(int) * nb); zalloc((void**)&d_chain_live, sizeof(unsigned) * 4 * nb);
        }
        if (!brick_chain) zalloc((void**)&d_near, PN);
        zalloc((void**)&d_zstart, sizeof(int) * 2 * (L.nz + 2));
        zalloc((void**)&d_mon, sizeof(double) * MFLBM_MON_N * L.nz);
        MF_CUDA(cudaMallocHost((void**)&h_mon, sizeof(double) * MFLBM_MON_N * L.nz));
        if (is_slab) {
            int prio_lo = 0, prio_hi = 0;
            MF_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            MF_CUDA(cudaStreamCreateWithPriority(&aux_stream, cudaStreamNonBlocking, prio_hi));   // halo kernels go first when CTA slots free up
            MF_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            MF_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
            // everything a neighbour may write - arrival flags and the six receive buffers - is ONE allocation, so that a
            // single CUDA IPC handle plus offsets describes it (small cudaMalloc blocks are sub-allocated and cannot be
            // exported one by one)
            const long long npdf = 10LL * L.NY1 * L.NZ1, nphi = 4LL * L.PY * L.PZ;
            auto up256 = [](size_t b) { return (b + 255) / 256 * 256; };
            size_t off = 256, offs[3][2];
            for (int kind = 0; kind < 3; kind++)
                for (int side = 0; side < 2; side++) { offs[kind][side] = off; off += up256(sizeof(T) * (size_t)(kind == 2 ? nphi : npdf)); }
            p2p_bytes = std::max<size_t>(off, 4u << 20);
            static_assert(16 + 6 <= 64, "flag block");
            zalloc((void**)&d_p2p, p2p_bytes);
            d_flags = reinterpret_cast<unsigned*>(d_p2p);
            for (int kind = 0; kind < 3; kind++)
                for (int side = 0; side < 2; side++) {
                    d_recv[kind][side] = reinterpret_cast<T*>(d_p2p + offs[kind][side]);
                    zalloc((void**)&d_send[kind][side], sizeof(T) * (kind == 2 ? nphi : npdf));
                }
        }
        L.pdf = d_pdf; L.phi = d_phi; L.cn_x = d_cnx; L.cn_y = d_cny; L.cn_z = d_cnz; L.c_norm = d_cnorm;
        L.W_in = d_Win; L.f_convec = d_fconv; L.g_convec = d_gconv; L.phi_convec = d_phiconv;
        L.types = d_types; L.cmap = d_cmap; L.fl_u = nullptr;
        set_params(p);
        MF_CUDA(cudaStreamSynchronize(stream));
    }

    // tensor map of a U-grid array for the tiles of k_chain_normals (cuTensorMapEncodeTiled through the runtime's driver
    // entry point: no link against libcuda)
    void make_tile_map(CUtensorMap* map, void* base, CUtensorMapDataType type, size_t elem) {
        typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        MF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) MF_FAIL("cuTensorMapEncodeTiled is not available in this driver");
        const cuuint64_t dims[3] = {(cuuint64_t)L.PX, (cuuint64_t)L.PY, (cuuint64_t)L.PZ};
        const cuuint64_t strides[2] = {(cuuint64_t)L.PX * elem, (cuuint64_t)L.PX * L.PY * elem};
        const cuuint32_t box[3] = {TL_X, TL_Y, TL_Z}, estr[3] = {1, 1, 1};
        const CUresult r = ((Encode)fn)(map, type, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) MF_FAIL("cuTensorMapEncodeTiled failed (%d)", (int)r);
    }

    void destroy() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        drop_graphs();
        dfree(d_pdf); dfree(d_phi_base); d_phi = nullptr; dfree(d_cnx); dfree(d_cny); dfree(d_cnz); dfree(d_cnorm);
        dfree(d_Win); dfree(d_fconv); dfree(d_gconv); dfree(d_phiconv);
        dfree(d_types_base); d_types = nullptr; dfree(d_cmap); dfree(d_flu); dfree(d_zstart); dfree(d_wbase);
        dfree(d_list_phi); dfree(d_mask_phi); dfree(d_list_cn); dfree(d_mask_cn); dfree(d_list_alter); dfree(d_list_n);
        for (auto& q : d_sn) dfree(q);
        dfree(d_live_n); dfree(d_live_cn); dfree(d_near);
        dfree(d_act); dfree(d_quiet); dfree(d_active); dfree(d_n_active); dfree(d_ent_off); dfree(d_chain_live); dfree(d_shell); dfree(d_bc_list); dfree(d_bc_mask); dfree(d_alt_start); dfree(d_sb_start); dfree(d_sb_list); dfree(d_sb_mask); dfree(d_grp); dfree(d_grp_start); dfree(d_grp_bricks); dfree(d_bx_start); dfree(d_bx_ent);
        dfree(d_mon); dfree(d_phi_old); dfree(d_p2p); d_flags = nullptr;
        for (auto& q : d_face) dfree(q);
        if (d_stage) { cudaFree(d_stage); d_stage = nullptr; }
        ring_drain(true);
        for (int b = 0; b < RING_NB; b++) {
            if (ev_ring_copy[b]) { cudaEventDestroy(ev_ring_copy[b]); ev_ring_copy[b] = nullptr; }
            if (ev_ring_kernel[b]) { cudaEventDestroy(ev_ring_kernel[b]); ev_ring_kernel[b] = nullptr; }
        }
        if (copy_stream) { cudaStreamDestroy(copy_stream); copy_stream = nullptr; }
        for (int kind = 0; kind < 3; kind++) for (int side = 0; side < 2; side++) { dfree(d_send[kind][side]); d_recv[kind][side] = nullptr; }
        if (h_mon) { cudaFreeHost(h_mon); h_mon = nullptr; }
        if (aux_stream) { cudaStreamDestroy(aux_stream); aux_stream = nullptr; }
        if (bc_stream) { cudaStreamSynchronize(bc_stream); cudaStreamDestroy(bc_stream); bc_stream = nullptr; }
        if (ev_bc_fork) { cudaEventDestroy(ev_bc_fork); ev_bc_fork = nullptr; }
        if (ev_bc_join) { cudaEventDestroy(ev_bc_join); ev_bc_join = nullptr; }
        if (ev_fork) { cudaEventDestroy(ev_fork); ev_fork = nullptr; }
        if (ev_join) { cudaEventDestroy(ev_join); ev_join = nullptr; }
        if (own_stream && stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }

    void drop_graphs() {
        for (auto& g : graph_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    }

    // copyConstantData, src/main_iteration_GPU.cu:14-47
    void set_params(const Params* p) {
        if (have_geometry && (p->nx != P.nx || p->ny != P.ny || p->nz != P.nz)) MF_FAIL("lattice dimensions cannot change after create");
        P = *p;
        L.lbm_gamma = p->lbm_gamma; L.force_z = p->force_z; L.la_nui1 = p->la_nui1; L.la_nui2 = p->la_nui2; L.lbm_beta = p->lbm_beta;
        L.RK_weight2 = T(1) / std::sqrt(T(2)) / T(36);   // includes/Fluid_multiphase.h:32, evaluated on the host in T
        L.phi_inlet = p->phi_inlet; L.relaxation = p->relaxation; L.sa_inject = p->sa_inject; L.uin_avg = p->uin_avg;
        L.cos_theta = p->cos_theta; L.rho_in = p->rho_in; L.rho_out = p->rho_out;
        L.Z_porous_plate = p->Z_porous_plate; L.porous_plate_cmd = p->porous_plate_cmd;
        drop_graphs();   // kernel arguments are baked into captured graphs
    }

    // ------------------------------------------------------------------------------------------------
    // launch helpers
    dim3 block2() const { const int bx = std::min(128, 32 * ceil_div(L.nx + 2, 32)); return dim3(bx, std::max(1, 128 / bx), 1); }
    void count(int n = 1) { launches += n; }
    void check_launch() { MF_CUDA(cudaGetLastError()); }
    dim3 grid_box(int G, int bx) const { return dim3(ceil_div(L.nx + 2 * G, bx), L.ny + 2 * G, L.nz + 2 * G); }

    // reference-layout array (ghost width G) on the host <-> U-grid array on the device
    // (asynchronous: the caller drains the ring before the host arrays may be reused / read)
    template <typename S>
    void to_u(const S* host, S* d_u, int G, long long n) {
        ring_reserve(sizeof(S) * n);
        ring_upload(host, sizeof(S) * n, [&](void* st) { k_repitch<T, S, true><<<grid_box(G, 128), 128, 0, stream>>>(L, G, (S*)st, d_u); check_launch(); count(); });
    }
    template <typename S>
    void from_u(S* host, S* d_u, int G, long long n) {
        ring_reserve(sizeof(S) * n);
        ring_download(host, sizeof(S) * n, [&](void* st) { k_repitch<T, S, false><<<grid_box(G, 128), 128, 0, stream>>>(L, G, (S*)st, d_u); check_launch(); count(); });
    }

    // ------------------------------------------------------------------------------------------------
    // geometry.  d_wtype: walls_type in the reference's s4 layout; d_sn4: the three s4 normal arrays (device).
    // Builds node types on the U grid, the site permutation of the PDF slots, the wall-link ranks and the compact lists of
    // every list-driven kernel - all of it on the device (kernels_setup.cuh: flag, exclusive scan, fill).
    void finish_geometry(const int* d_wtype, T* const d_sn4[3]) {
        k_types_to_u<T><<<dim3(ceil_div(L.PX, 128), L.PY, L.PZ), 128, 0, stream>>>(L, d_wtype, d_types); check_launch(); count();
        dfree(d_phi_old);   // indexed by fluid entry: invalid for a new geometry
        const SetupInfo SI{slab.has_left, slab.has_right, open_z() ? 1 : 0, P.kper, P.jper};
        const dim3 gu(ceil_div(L.PX, 128), L.PY, L.PZ);
        int *d_pred = nullptr, *d_scan[3] = {nullptr, nullptr, nullptr};
        unsigned long long* d_counts = nullptr;
        int* d_links = nullptr;
        auto cleanup = [&]() { cudaFree(d_pred); for (auto q : d_scan) cudaFree(q); cudaFree(d_counts); cudaFree(d_links); };
        try {
            MF_CUDA(cudaMalloc((void**)&d_pred, sizeof(int) * (size_t)(PN + 1)));
            for (auto& q : d_scan) MF_CUDA(cudaMalloc((void**)&q, sizeof(int) * (size_t)(PN + 1)));
            MF_CUDA(cudaMemsetAsync(d_pred + PN, 0, sizeof(int), stream));
            // number of sites that satisfy `kind`, their ranks in d_scan[slot]
            auto rank = [&](int kind, int slot) -> long long {
                k_setup_flags<T><<<gu, 128, 0, stream>>>(L, SI, kind, d_pred); check_launch(); count();
                long long total = 0;
                MF_CUDA(exclusive_scan(d_pred, d_scan[slot], PN + 1, stream, &total));
                return total;
            };
            // ---- site permutation: [fluid nodes of neighbour-facing columns | other fluid nodes | other sites of the 1-ghost box]
            n_boundary = rank(P_FLUID_B, 0);
            const long long n_interior = rank(P_FLUID_I, 1);
            const long long n_passive = rank(P_PASSIVE, 2);
            n_fluid = n_boundary + n_interior;
            if (n_fluid + n_passive != N1) MF_FAIL("internal error: site permutation covers %lld of %lld sites", n_fluid + n_passive, N1);
            dfree(d_flu);
            MF_CUDA(cudaMalloc((void**)&d_flu, sizeof(int) * (size_t)std::max<long long>(n_fluid, 1)));
            k_setup_site_map<T><<<gu, 128, 0, stream>>>(L, SI, d_scan[0], d_scan[1], d_scan[2], (int)n_boundary, (int)n_fluid, d_cmap, d_flu); check_launch(); count();
            k_setup_zstart<T><<<ceil_div(L.nz + 2, 128), 128, 0, stream>>>(L, d_scan[0], d_scan[1], (int)n_boundary, d_zstart); check_launch(); count();
            L.n_fluid = (int)n_fluid; L.fl_u = d_flu;
            // ---- wall links: rank of every link inside its direction, per 32-entry group (core.cuh)
            const int ngrp = ceil_div((int)n_fluid, 32), nrows = ceil_div((int)n_fluid, COLLIDE_TILE) * (COLLIDE_TILE / 32) + 1;
            long long link_tot[19] = {0};
            dfree(d_wbase);
            MF_CUDA(cudaMalloc((void**)&d_wbase, sizeof(int) * (size_t)nrows * 18));
            if (ngrp) {
                MF_CUDA(cudaMalloc((void**)&d_links, sizeof(int) * ((size_t)ngrp * 18 + 1)));
                MF_CUDA(cudaMemsetAsync(d_links + (size_t)ngrp * 18, 0, sizeof(int), stream));
                k_setup_link_count<T><<<ceil_div(ngrp, 4), 128, 0, stream>>>(L, ngrp, d_links); check_launch(); count();
                long long total = 0;
                MF_CUDA(exclusive_scan(d_links, d_links, (long long)ngrp * 18 + 1, stream, &total));
                k_setup_wbase<<<ceil_div(nrows * 18, 128), 128, 0, stream>>>(d_links, ngrp, nrows, d_wbase); check_launch(); count();
                std::vector<int> ends(19);
                for (int q = 0; q <= 18; q++) MF_CUDA(cudaMemcpyAsync(&ends[(size_t)q], d_links + (size_t)q * ngrp, sizeof(int), cudaMemcpyDeviceToHost, stream));
                MF_CUDA(cudaStreamSynchronize(stream));
                for (int q = 1; q <= 18; q++) link_tot[q] = ends[(size_t)q] - ends[(size_t)q - 1];
            } else MF_CUDA(cudaMemsetAsync(d_wbase, 0, sizeof(int) * (size_t)nrows * 18, stream));
            L.wbase = d_wbase;
            n_links = 0;
            long long max_links = 0;
            for (int q = 1; q < 19; q++) { n_links += link_tot[q]; max_links = std::max(max_links, link_tot[q]); }
            // slot = [fluid nodes | other sites of the 1-ghost box | pad to whole 128-entry tiles (+1: the TMA row copies of the
            // last fluid tile stay inside the slot) | mailboxes of the one link direction the slot hosts]
            const long long mb0 = 128 * ((N1 + 127) / 128) + 128;
            NC = mb0 + 128 * ((max_links + 127) / 128);
            if (NC >= (1LL << 31) - 2) MF_FAIL("PDF slot exceeds 2^31 entries; decompose into more slabs");
            L.mb0 = (int)mb0; L.NC = NC;
            dfree(d_pdf);
            zalloc((void**)&d_pdf, sizeof(T) * NC * 38);
            L.pdf = d_pdf;
            // ---- compact site lists
            auto make_list = [&](int kind, int*& d_list, int** d_mask) -> int {
                const long long n = rank(kind, 0);
                dfree(d_list);
                MF_CUDA(cudaMalloc((void**)&d_list, sizeof(int) * (size_t)std::max<long long>(n, 1)));
                if (d_mask) { dfree(*d_mask); MF_CUDA(cudaMalloc((void**)d_mask, sizeof(int) * (size_t)std::max<long long>(n, 1))); }
                if (n) { k_setup_fill<T><<<gu, 128, 0, stream>>>(L, SI, kind, d_scan[0], d_list, d_mask ? *d_mask : nullptr); check_launch(); count(); }
                return (int)n;
            };
            n_list_phi = make_list(P_PHI, d_list_phi, &d_mask_phi);
            n_list_cn = n_list_n = n_shell = n_bc = 0;
            if (brick_chain) { n_shell = make_list(P_SHELL, d_shell, nullptr); n_bc = make_list(P_COPIED, d_bc_list, &d_bc_mask); }
            else { n_list_cn = make_list(P_CN, d_list_cn, &d_mask_cn); n_list_n = make_list(P_NORMALS, d_list_n, nullptr); }
            // ---- counters of the reference's geometry report
            MF_CUDA(cudaMalloc((void**)&d_counts, 4 * sizeof(unsigned long long)));
            MF_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(unsigned long long), stream));
            k_setup_counts<T><<<gu, 128, 0, stream>>>(L, SI, d_counts); check_launch(); count();
            unsigned long long hc[4];
            MF_CUDA(cudaMemcpyAsync(hc, d_counts, sizeof hc, cudaMemcpyDeviceToHost, stream));
            MF_CUDA(cudaStreamSynchronize(stream));
            for (int n = 0; n < 4; n++) counts[n] = (long long)hc[n];
        } catch (...) { cleanup(); throw; }
        cleanup();
        dfree(d_live_n); dfree(d_live_cn);
        MF_CUDA(cudaMalloc((void**)&d_live_n, std::max(n_list_n, 1))); MF_CUDA(cudaMalloc((void**)&d_live_cn, std::max(n_list_cn, 1)));
        // fluid-boundary sites of the whole U grid (wetting; the kernels check their own range [-1..n+2]^3, :814) and, for the
        // brick chain, solid-boundary sites of [0..n+1]^3 (:885), compacted brick by brick on the device: count per brick,
        // exclusive scan, fill (sites of a brick in z,y,x order)
        {
            const int nb = bricks.count();
            int* d_cnt = nullptr;
            MF_CUDA(cudaMalloc((void**)&d_cnt, sizeof(int) * (size_t)(nb + 1)));
            auto compact = [&](auto count_kernel, auto fill_kernel, int*& d_start, int*& d_list, int** d_mask) -> int {
                MF_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)(nb + 1), stream));
                count_kernel<<<nb, CHAIN_THREADS, 0, stream>>>(L, d_cnt, nullptr, nullptr); check_launch(); count();
                long long total = 0;
                if (!d_start) MF_CUDA(cudaMalloc((void**)&d_start, sizeof(int) * (size_t)(nb + 1)));
                MF_CUDA(exclusive_scan(d_cnt, d_start, nb + 1, stream, &total));
                dfree(d_list);
                MF_CUDA(cudaMalloc((void**)&d_list, sizeof(int) * (size_t)std::max<long long>(total, 1)));
                if (d_mask) { dfree(*d_mask); MF_CUDA(cudaMalloc((void**)d_mask, sizeof(int) * (size_t)std::max<long long>(total, 1))); }
                fill_kernel<<<nb, CHAIN_THREADS, 0, stream>>>(L, d_start, d_list, d_mask ? *d_mask : nullptr); check_launch(); count();
                return (int)total;
            };
            try {
                n_list_alter_all = compact(k_brick_sites<T, 0, 0>, k_brick_sites<T, 0, 1>, d_alt_start, d_list_alter, nullptr);
                n_list_alter = n_list_alter_all;
                if (brick_chain) n_sb = compact(k_brick_sites<T, 1, 0>, k_brick_sites<T, 1, 1>, d_sb_start, d_sb_list, &d_sb_mask);
                if (brick_chain && chain_csr) compact(k_brick_box_sites<T, 0>, k_brick_box_sites<T, 1>, d_bx_start, d_bx_ent, nullptr);
            } catch (...) { cudaFree(d_cnt); throw; }
            cudaFree(d_cnt);
        }
        mark_all_live();
        if (brick_chain) {   // bricks touched by every group of 32 fluid entries (kernels_chain.cuh, raise_activity / k_chain_pre)
            n_groups = ceil_div((int)n_fluid, 32);
            dfree(d_grp); dfree(d_grp_start); dfree(d_grp_bricks);
            zalloc((void**)&d_grp, 2 * (size_t)std::max(n_groups, 1));
            int* d_cnt = nullptr;
            MF_CUDA(cudaMalloc((void**)&d_cnt, sizeof(int) * (size_t)(n_groups + 1)));
            MF_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)(n_groups + 1), stream));
            MF_CUDA(cudaMalloc((void**)&d_grp_start, sizeof(int) * (size_t)(n_groups + 1)));
            long long total = 0;
            cudaError_t e = cudaSuccess;
            if (n_groups) { k_group_bricks<T, 0><<<ceil_div(n_groups, 4), 128, 0, stream>>>(L, n_groups, d_cnt, nullptr); count(); }
            e = exclusive_scan(d_cnt, d_grp_start, n_groups + 1, stream, &total);
            cudaFree(d_cnt);
            MF_CUDA(e);
            MF_CUDA(cudaMalloc((void**)&d_grp_bricks, sizeof(int) * (size_t)std::max<long long>(total, 1)));
            if (n_groups) { k_group_bricks<T, 1><<<ceil_div(n_groups, 4), 128, 0, stream>>>(L, n_groups, d_grp_start, d_grp_bricks); check_launch(); count(); }
        }
        for (int a = 0; a < 3; a++) {
            dfree(d_sn[a]);
            MF_CUDA(cudaMalloc((void**)&d_sn[a], sizeof(T) * std::max(n_list_alter_all, 1)));
            if (n_list_alter_all) { k_list_s4<T, true><<<ceil_div(n_list_alter_all, 128), 128, 0, stream>>>(L, d_list_alter, n_list_alter_all, d_sn4[a], d_sn[a]); check_launch(); count(); }
        }
        if (is_slab) {
            const int cols[4] = {0, 1, L.nx, L.nx + 1};
            for (int k = 0; k < 4; k++) {
                dfree(d_face[k]);
                MF_CUDA(cudaMalloc((void**)&d_face[k], sizeof(int) * (size_t)L.NY1 * L.NZ1));
                k_setup_face<T><<<dim3(ceil_div(L.NY1, 128), L.NZ1), 128, 0, stream>>>(L, cols[k], d_face[k]); check_launch(); count();
            }
        }
        MF_CUDA(cudaStreamSynchronize(stream));
        have_geometry = true;
        drop_graphs();
    }

    void upload_geometry(const int32_t* walls, const int32_t* wtype, const T* snx, const T* sny, const T* snz) {
        if (!walls || !wtype || !snx || !sny || !snz) MF_FAIL("upload_geometry: null array");
        // walls (s2) is redundant with walls_type (walls == (walls_type > 0) on its extents, Geometry_preprocessing.cpp:135-184)
        int* d_wt = nullptr; T* d_sn4[3] = {nullptr, nullptr, nullptr};
        MF_CUDA(cudaMalloc((void**)&d_wt, sizeof(int) * N4));
        MF_CUDA(cudaMemcpyAsync(d_wt, wtype, sizeof(int) * N4, cudaMemcpyHostToDevice, stream));
        const T* src[3] = {snx, sny, snz};
        for (int a = 0; a < 3; a++) {
            MF_CUDA(cudaMalloc((void**)&d_sn4[a], sizeof(T) * N4));
            MF_CUDA(cudaMemcpyAsync(d_sn4[a], src[a], sizeof(T) * N4, cudaMemcpyHostToDevice, stream));
        }
        try { finish_geometry(d_wt, d_sn4); }
        catch (...) { cudaFree(d_wt); for (auto q : d_sn4) cudaFree(q); throw; }
        cudaFree(d_wt); for (auto q : d_sn4) cudaFree(q);
    }

    void preprocess_geometry(const int8_t* interior_global) {
        if (!interior_global) MF_FAIL("preprocess_geometry: null array");
        GeoDims D{};
        D.nx = L.nx; D.ny = L.ny; D.nz = L.nz; D.TX = L.nx + 20; D.TY = L.ny + 20; D.TZ = L.nz + 20;
        D.x0 = L.x0; D.nxg = (int)P.nx; D.nyg = (int)P.ny; D.nzg = (int)P.nz; D.iper = P.iper; D.jper = P.jper; D.kper = P.kper;
        const size_t TN = (size_t)D.TX * D.TY * D.TZ, NG = (size_t)P.nx * P.ny * P.nz;
        int8_t *d_in = nullptr, *d_wt = nullptr, *d_ty = nullptr;
        T *d_ws1 = nullptr, *d_ws2 = nullptr;
        int *d_walls = nullptr, *d_wtype = nullptr;
        T* d_sn4[3] = {nullptr, nullptr, nullptr};
        auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_wt); cudaFree(d_ty); cudaFree(d_ws1); cudaFree(d_ws2); cudaFree(d_walls); cudaFree(d_wtype); for (auto q : d_sn4) cudaFree(q); };
        try {
            MF_CUDA(cudaMalloc((void**)&d_in, NG)); MF_CUDA(cudaMalloc((void**)&d_wt, TN)); MF_CUDA(cudaMalloc((void**)&d_ty, TN));
            MF_CUDA(cudaMalloc((void**)&d_ws1, TN * sizeof(T))); MF_CUDA(cudaMalloc((void**)&d_ws2, TN * sizeof(T)));
            MF_CUDA(cudaMalloc((void**)&d_walls, sizeof(int) * N2)); MF_CUDA(cudaMalloc((void**)&d_wtype, sizeof(int) * N4));
            for (int a = 0; a < 3; a++) { MF_CUDA(cudaMalloc((void**)&d_sn4[a], sizeof(T) * N4)); MF_CUDA(cudaMemsetAsync(d_sn4[a], 0, sizeof(T) * N4, stream)); }
            MF_CUDA(cudaMemcpyAsync(d_in, interior_global, NG, cudaMemcpyHostToDevice, stream));
            MF_CUDA(cudaMemsetAsync(d_ty, 0, TN, stream));
            const int bx = 128;
            k_geo_fill<T><<<dim3(ceil_div(D.TX, bx), D.TY, D.TZ), bx, 0, stream>>>(D, d_in, d_wt, d_ws1, d_ws2); check_launch(); count();
            const dim3 ginner(ceil_div(D.TX - 2, bx), D.TY - 2, D.TZ - 2);
            k_geo_classify<<<ginner, bx, 0, stream>>>(D, d_wt, d_ty); check_launch(); count();
            for (int it = 0; it < 4; it++) {   // :187-206
                k_geo_smooth<T><<<ginner, bx, 0, stream>>>(D, d_ws1, d_ws2); check_launch();
                k_geo_copy_inner<T><<<ginner, bx, 0, stream>>>(D, d_ws2, d_ws1); check_launch();
                count(2);
            }
            Iso8Tables tab;
            static const signed char TX_[34][3] = {{1,0,0},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{2,0,0},
                {2,1,0},{2,-1,0},{2,0,1},{2,0,-1},{1,2,0},{1,-2,0},{1,0,2},{1,0,-2},
                {2,1,1},{2,1,-1},{2,-1,1},{2,-1,-1},{1,2,1},{1,2,-1},{1,-2,1},{1,-2,-1},{1,1,2},{1,1,-2},{1,-1,2},{1,-1,-2},
                {2,2,0},{2,-2,0},{2,0,2},{2,0,-2}};
            static const signed char TY_[34][3] = {{0,1,0},{1,1,0},{-1,1,0},{0,1,1},{0,1,-1},{1,1,1},{1,1,-1},{-1,1,-1},{-1,1,1},{0,2,0},
                {2,1,0},{-2,1,0},{0,2,1},{0,2,-1},{1,2,0},{-1,2,0},{0,1,2},{0,1,-2},
                {2,1,1},{2,1,-1},{-2,1,1},{-2,1,-1},{1,2,1},{1,2,-1},{-1,2,1},{-1,2,-1},{1,1,2},{1,1,-2},{-1,1,2},{-1,1,-2},
                {2,2,0},{-2,2,0},{0,2,2},{0,2,-2}};
            static const signed char TZ_[34][3] = {{0,0,1},{0,1,1},{0,-1,1},{1,0,1},{-1,0,1},{1,1,1},{1,-1,1},{-1,1,1},{-1,-1,1},{0,0,2},
                {0,1,2},{0,-1,2},{2,0,1},{-2,0,1},{0,2,1},{0,-2,1},{1,0,2},{-1,0,2},
                {2,1,1},{2,-1,1},{-2,1,1},{-2,-1,1},{1,2,1},{1,-2,1},{-1,2,1},{-1,-2,1},{1,1,2},{1,-1,2},{-1,1,2},{-1,-1,2},
                {0,2,2},{0,-2,2},{2,0,2},{-2,0,2}};
            memcpy(tab.o[0], TX_, sizeof TX_); memcpy(tab.o[1], TY_, sizeof TY_); memcpy(tab.o[2], TZ_, sizeof TZ_);
            const T eps = (T)1.1920928955078125e-07f;   // includes/Module.h:10-16: float epsilon in both precisions
            k_geo_export<T><<<dim3(ceil_div(L.nx + 8, bx), L.ny + 8, L.nz + 8), bx, 0, stream>>>(D, tab, d_wt, d_ty, d_ws2, d_walls, d_wtype, d_sn4[0], d_sn4[1], d_sn4[2], eps);
            check_launch(); count();
            cudaFree(d_in); d_in = nullptr; cudaFree(d_wt); d_wt = nullptr; cudaFree(d_ty); d_ty = nullptr;
            finish_geometry(d_wtype, d_sn4);
        } catch (...) { cleanup(); throw; }
        cleanup();
    }

    void download_geometry(int32_t* walls, int32_t* wtype, T* snx, T* sny, T* snz, int64_t* cnt) {
        if (!have_geometry) MF_FAIL("download_geometry before geometry");
        if (walls || wtype) {
            int* st = (int*)stage(sizeof(int) * (N2 + N4));
            k_types_from_u<T><<<grid_box(4, 128), 128, 0, stream>>>(L, st, st + N2); check_launch(); count();
            if (walls) MF_CUDA(cudaMemcpyAsync(walls, st, sizeof(int) * N2, cudaMemcpyDeviceToHost, stream));
            if (wtype) MF_CUDA(cudaMemcpyAsync(wtype, st + N2, sizeof(int) * N4, cudaMemcpyDeviceToHost, stream));
            MF_CUDA(cudaStreamSynchronize(stream));
        }
        T* out[3] = {snx, sny, snz};
        for (int a = 0; a < 3; a++) {
            if (!out[a]) continue;
            T* st = (T*)stage(sizeof(T) * N4);
            MF_CUDA(cudaMemsetAsync(st, 0, sizeof(T) * N4, stream));   // zero except at fluid-boundary nodes (Geometry_preprocessing.cpp:229-386)
            if (n_list_alter_all) { k_list_s4<T, false><<<ceil_div(n_list_alter_all, 128), 128, 0, stream>>>(L, d_list_alter, n_list_alter_all, st, d_sn[a]); check_launch(); count(); }
            MF_CUDA(cudaMemcpyAsync(out[a], st, sizeof(T) * N4, cudaMemcpyDeviceToHost, stream));
            MF_CUDA(cudaStreamSynchronize(stream));
        }
        if (cnt) for (int n = 0; n < 4; n++) cnt[n] = counts[n];
        drop_stage();
    }

    // ------------------------------------------------------------------------------------------------
    // state
    void upload_state(const T* pdf, const T* phi, const T* cnx, const T* cny, const T* cnz, const T* cnorm, const T* curv,
                      const T* Win, const T* fconv, const T* gconv, const T* phiconv) {
        if (!have_geometry) MF_FAIL("upload_state before geometry (the PDF site order depends on it)");
        ring_reserve(sizeof(T) * (size_t)N4);
        if (pdf) {
            const dim3 g = grid_box(1, 128);
            for (int s = 0; s < 38; s++)
                ring_upload(pdf + (size_t)s * N1, sizeof(T) * N1, [&](void* st) {
                    k_pdf_slot<T, true><<<g, 128, 0, stream>>>(L, (T*)st, s); check_launch(); count();
                    if (s % 19 != 0 && n_fluid) { k_pdf_mail<T, true><<<ceil_div((int)n_fluid, 128), 128, 0, stream>>>(L, (T*)st, s); check_launch(); count(); }
                });
        }
        if (phi) to_u<T>(phi, d_phi, 4, N4);
        if (cnx) to_u<T>(cnx, d_cnx, 2, N2);
        if (cny) to_u<T>(cny, d_cny, 2, N2);
        if (cnz) to_u<T>(cnz, d_cnz, 2, N2);
        if (cnorm) to_u<T>(cnorm, d_cnorm, 2, N2);
        if (cnx || cny || cnz || cnorm) { cn_consistent = false; drop_graphs(); }   // bulk_skip is baked into captured launches
        if (phi || cnx || cny || cnz || cnorm) mark_all_live();
        if (cnx || cny || cnz || cnorm) {   // the caller's arrays are not trusted to hold zeros in solids
            k_zero_solid_normals<T><<<grid_box(2, 128), 128, 0, stream>>>(L); check_launch(); count();
        }
        (void)curv;   // curv is a pure function of cn_* (CSF_Forces, :908-1003): recomputed where it is consumed, never stored
        auto up = [&](T* d, const T* h, long long n) { if (h) MF_CUDA(cudaMemcpyAsync(d, h, sizeof(T) * n, cudaMemcpyHostToDevice, stream)); };
        up(d_Win, Win, NP); up(d_fconv, fconv, NP * 19); up(d_gconv, gconv, NP * 19); up(d_phiconv, phiconv, NP);
        ring_drain(true);
    }

    void download_state(T* pdf, T* phi, T* cnx, T* cny, T* cnz, T* cnorm, T* curv, T* fconv, T* gconv, T* phiconv) {
        if (!have_geometry) MF_FAIL("download_state before geometry");
        check_halo_error();   // a state behind a lost halo message must not reach a checkpoint
        ring_reserve(sizeof(T) * (size_t)N4);
        if (pdf) {
            const dim3 g = grid_box(1, 128);
            for (int s = 0; s < 38; s++)
                ring_download(pdf + (size_t)s * N1, sizeof(T) * N1, [&](void* st) {
                    k_pdf_slot<T, false><<<g, 128, 0, stream>>>(L, (T*)st, s); check_launch(); count();
                    if (s % 19 != 0 && n_fluid) { k_pdf_mail<T, false><<<ceil_div((int)n_fluid, 128), 128, 0, stream>>>(L, (T*)st, s); check_launch(); count(); }
                });
        }
        if (phi) {
            // the brick chain leaves phi at the solid-boundary sites of long-quiet bricks at its last evaluation (kernels_chain.cuh):
            // the array that leaves the library holds what extrapolate_phi_toSolid (:732-755) gives for the current phi
            if (brick_chain && n_list_phi) { k_extrap_phi<T><<<ceil_div(n_list_phi, 128), 128, 0, stream>>>(L, d_list_phi, d_mask_phi, n_list_phi); check_launch(); count(); }
            from_u<T>(phi, d_phi, 4, N4);
        }
        if (cnx) from_u<T>(cnx, d_cnx, 2, N2);
        if (cny) from_u<T>(cny, d_cny, 2, N2);
        if (cnz) from_u<T>(cnz, d_cnz, 2, N2);
        if (cnorm) from_u<T>(cnorm, d_cnorm, 2, N2);
        if (curv) {   // the stepping path keeps curv at fluid nodes only; give the caller the reference's dense array (:908-1003 over [1..n]^3)
            ring_download(curv, sizeof(T) * N1, [&](void* st) {
                MF_CUDA(cudaMemsetAsync(st, 0, sizeof(T) * N1, stream));
                k_curvature_dense<T><<<dim3(ceil_div(L.nx, 128), L.ny, L.nz), 128, 0, stream>>>(L, (T*)st); check_launch(); count();
            });
        }
        auto dn = [&](T* h, const T* d, long long n) { if (h) MF_CUDA(cudaMemcpyAsync(h, d, sizeof(T) * n, cudaMemcpyDeviceToHost, stream)); };
        dn(fconv, d_fconv, NP * 19); dn(gconv, d_gconv, NP * 19); dn(phiconv, d_phiconv, NP);
        ring_drain(true);
    }

    // the cn_* / c_norm arrays hold values the chain did not write (fresh geometry, upload, init): every entry must be stored once
    void mark_all_live() {
        if (d_live_n) MF_CUDA(cudaMemsetAsync(d_live_n, 1, std::max(n_list_n, 1), stream));
        if (d_live_cn) MF_CUDA(cudaMemsetAsync(d_live_cn, 1, std::max(n_list_cn, 1), stream));
        if (d_quiet) MF_CUDA(cudaMemsetAsync(d_quiet, 0, (size_t)bricks.count(), stream));   // every brick is processed by the next chain
        if (d_chain_live) MF_CUDA(cudaMemsetAsync(d_chain_live, 0xff, sizeof(unsigned) * 4 * (size_t)bricks.count(), stream));   // ... and stores every site
    }

    bool open_z() const { return P.kper == 0 && P.wall_z_min == 0 && P.wall_z_max == 0; }

    // phi_s4 != nullptr: the caller's phase field (4-ghost layout, e.g. the rand()-drawn distribution of option 6 with the
    // inlet ghost planes already filled, src/Init_multiphase.cpp:345-372) instead of one of the device patterns
    void init_state(int option, T interface_z0, const T* Win, const T* phi_s4 = nullptr) {
        if (!have_geometry) MF_FAIL("init_state before geometry");
        if (!phi_s4 && (option < 1 || option > 5)) MF_FAIL("initial_fluid_distribution_option %d not supported on the device (1..5)", option);
        const int bx = 128;
        MF_CUDA(cudaMemsetAsync(d_phi, 0, sizeof(T) * PN, stream));
        if (phi_s4) { to_u<T>(phi_s4, d_phi, 4, N4); ring_drain(true); }
        else { k_init_phi<T><<<grid_box(4, bx), bx, 0, stream>>>(L, option, interface_z0, (int)P.ny, (int)P.nz, open_z() ? 1 : 0); count(); }
        check_launch();
        k_init_pdf<T><<<grid_box(1, bx), bx, 0, stream>>>(L, P.outlet_BC == 1 ? 1 : 0); check_launch();
        count();
        if (Win) MF_CUDA(cudaMemcpyAsync(d_Win, Win, sizeof(T) * NP, cudaMemcpyHostToDevice, stream));
        MF_CUDA(cudaMemsetAsync(d_cnx, 0, sizeof(T) * PN, stream)); MF_CUDA(cudaMemsetAsync(d_cny, 0, sizeof(T) * PN, stream));
        MF_CUDA(cudaMemsetAsync(d_cnz, 0, sizeof(T) * PN, stream)); MF_CUDA(cudaMemsetAsync(d_cnorm, 0, sizeof(T) * PN, stream));
        mark_all_live();
        gradient_chain();
        MF_CUDA(cudaStreamSynchronize(stream));
    }

    // ------------------------------------------------------------------------------------------------
    // x ranges of plane kernels: extended over the ghost column on sides that face a neighbour slab
    int plo() const { return slab.has_left ? 0 : 1; }
    int phi_() const { return slab.has_right ? L.nx + 1 : L.nx; }

    // the colour-gradient kernels, call order of src/main_iteration_GPU.cu:2027-2049; the fifth (CSF_Forces, :2052) is
    // fused into the collide kernel of the next step.
    // parity: ntime & 1 of the step whose collide kernel raised the brick flags; -1: no collide precedes this chain (initial
    // state, restart, color_gradient through the ABI) - the flags then come from a scan of phi.
    void gradient_chain(int parity = -1) {
        if (!have_geometry) MF_FAIL("gradient chain before geometry");
        const int bl = 128;
        if (brick_chain) { gradient_chain_bricks(parity); return; }
        // the list chain (kernels_step.cuh): four kernels over compact site lists, every site every step
        if (n_list_phi) { k_extrap_phi<T><<<ceil_div(n_list_phi, bl), bl, 0, stream>>>(L, d_list_phi, d_mask_phi, n_list_phi); check_launch(); count(); }
        if (n_list_n) { k_normals<T><<<ceil_div(n_list_n, bl), bl, 0, stream>>>(L, d_list_n, d_live_n, d_near, n_list_n); check_launch(); count(); }
        if (n_list_alter) { k_alter<T><<<ceil_div(n_list_alter, bl), bl, 0, stream>>>(L, d_list_alter, d_sn[0], d_sn[1], d_sn[2], n_list_alter); check_launch(); count(); }
        if (n_list_cn) { k_extrap_cn<T><<<ceil_div(n_list_cn, bl), bl, 0, stream>>>(L, d_list_cn, d_mask_cn, d_live_cn, d_near, n_list_cn); check_launch(); count(); }
        cn_consistent = true;
    }

    unsigned char* act_set(int set, int which) const { return d_act + (size_t)bricks.count() * (size_t)(2 * set + which); }

    // kernels_chain.cuh: raise (collide kernels + k_act_shell, or k_act_scan) -> verdict -> process the listed bricks
    void gradient_chain_bricks(int parity) {
        const int bl = 128, nb = bricks.count();
        const bool scan = parity < 0 || act_scan;
        const int set = scan ? 2 : (parity & 1);
        Lattice<T> La = L;
        La.act_p = act_set(set, 0); La.act_m = act_set(set, 1);
        La.grp_p = d_grp; La.grp_m = d_grp + std::max(n_groups, 1);
        if (scan) {
            MF_CUDA(cudaMemsetAsync(La.act_p, 0, 2 * (size_t)nb, stream));
            k_act_scan<T><<<dim3(ceil_div(L.PX, bl), L.PY, L.PZ), bl, 0, stream>>>(L, La.act_p, La.act_m, d_n_active); check_launch(); count();
            if (parity < 0) MF_CUDA(cudaMemsetAsync(d_act, 0, 4 * (size_t)nb, stream));   // whatever step comes next starts from clean sets
            // sub-list refresh (see k_chain_pre); group flags a collide kernel may have raised are dropped with the scan's verdict
            if (n_bc) { k_chain_pre<T><<<ceil_div(n_bc, bl), bl, 0, stream>>>(La, d_shell, 0, d_bc_list, d_bc_mask, n_bc, d_grp_start, d_grp_bricks, 0, nullptr); check_launch(); count(); }
            if (n_groups) MF_CUDA(cudaMemsetAsync(d_grp, 0, 2 * (size_t)n_groups, stream));
        } else {
            k_chain_pre<T><<<std::max(1, ceil_div(n_shell + n_bc + n_groups, bl)), bl, 0, stream>>>(La, d_shell, n_shell, d_bc_list, d_bc_mask, n_bc, d_grp_start, d_grp_bricks, n_groups, d_n_active);
            check_launch(); count();
        }
        // the verdict kernel clears the set the next step's collide raises into
        unsigned char* clr = parity < 0 ? nullptr : act_set((parity & 1) ^ 1, 0);
        // csr chain: the verdict kernel also hands every listed brick its offset in a flat numbering of the cn-extrapolation entries
        k_act_verdict<<<ceil_div(nb, bl), bl, 0, stream>>>(bricks, La.act_p, La.act_m, d_quiet, d_active, d_n_active, clr, clr ? clr + nb : nullptr, d_sb_start, chain_csr ? d_ent_off : nullptr);
        check_launch(); count();
        BrickNormals<T> SN{d_alt_start, d_sn[0], d_sn[1], d_sn[2]};
        if (chain_csr) {
            // resident CTAs per SM: the single-precision kernel's 40 registers allow 12, the double-precision one needs 64 (8; with
            // 10 and 48 registers it spills a dozen words and was measured slower, profiles/r03a_chain_variants.json)
            constexpr int C = sizeof(T) == 8 ? 8 : 12;
            k_chain_normals_csr<T, C><<<std::max(1, std::min(nb, num_sms * C)), CHAIN_THREADS, 0, stream>>>(L, tm_phi, d_active, d_n_active, SN, d_bx_start, d_bx_ent, d_chain_live);
            check_launch(); count();
            k_chain_extrap_cn_flat<T><<<std::max(1, std::min(ceil_div(n_sb, CHAIN_THREADS), num_sms * 12)), CHAIN_THREADS, 0, stream>>>(L, d_active, d_n_active, d_ent_off, d_sb_start, d_sb_list, d_sb_mask);
            check_launch(); count();
        } else {
            k_chain_normals<T><<<std::max(1, std::min(nb, num_sms * 8)), CHAIN_THREADS, 0, stream>>>(L, tm_phi, d_active, d_n_active, SN); check_launch(); count();
            k_chain_extrap_cn<T><<<std::max(1, std::min(nb, num_sms * 12)), CHAIN_THREADS, 0, stream>>>(L, d_active, d_n_active, d_sb_start, d_sb_list, d_sb_mask); check_launch(); count();
        }
        cn_consistent = true;
    }

    // the lattice view a collide kernel gets: with the brick chain it raises the activity flags of its groups of 32 entries
    Lattice<T> collide_lattice(int) const {
        Lattice<T> La = L;
        if (brick_chain && !act_scan && d_grp) { La.grp_p = d_grp; La.grp_m = d_grp + std::max(n_groups, 1); }
        return La;
    }

    // persistent grid of the collide kernels: CTAS resident CTAs per SM.  MFLBM_MAX_CTAS caps it (tests: a small lattice
    // then walks many tiles per CTA, so the stage rings wrap as they do at full size)
    int collide_grid(int ntiles, int ctas) const {
        const int sms = std::max(1, num_sms - tile_reserve_sms);
        return std::max(1, std::min(ntiles, max_ctas > 0 ? std::min(max_ctas, sms * ctas) : sms * ctas));
    }
    // tile range of the collide launch being issued (set by phase_collide): [tile_first, tile_end) of the fluid order, and the
    // SMs its persistent grid leaves free for the halo kernels that run next to it (see step_p2p)
    int tile_first = 0, tile_end = 0, tile_reserve_sms = 0;

    // pipelined kernels (kernels_collide.cuh): persistent CTAs, TMA / cp.async staged PDF rows
    template <int MRT, int NST, int CTAS>
    void launch_even() {
        auto kern = k_collide_even_tma<T, MRT, NST, CTAS>;
        constexpr size_t smem = collide_even_smem<T, NST>();
        static thread_local int configured = -1;
        if (configured != device) { MF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = device; }
        halo_lane_fits(kern, COLLIDE_EVEN_THREADS * CTAS);
        kern<<<collide_grid(tile_end - tile_first, CTAS), COLLIDE_EVEN_THREADS, smem, stream>>>(collide_lattice(0), tile_first, tile_end, cn_consistent ? 1 : 0);
    }
    template <int MRT, int NST, int CTAS, int NCONS = 1, int REGS_P = 0, int REGS_C = 0>
    void launch_odd_ws() {
        auto kern = k_collide_odd_ws<T, MRT, NST, CTAS, NCONS, REGS_P, REGS_C>;
        constexpr size_t smem = collide_odd_ws_smem<T, NST>();
        static thread_local int configured = -1;
        if (configured != device) { MF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = device; }
        halo_lane_fits(kern, (NCONS + 1) * COLLIDE_TILE * CTAS);
        kern<<<collide_grid(tile_end - tile_first, CTAS), (NCONS + 1) * COLLIDE_TILE, smem, stream>>>(collide_lattice(1), tile_first, tile_end, cn_consistent ? 1 : 0);
    }
    // An interior collide launch that overlaps halo kernels (tile_reserve_sms < 0 = "decide here"): when the resident CTAs of
    // the kernel take (almost) the whole register file of an SM - the single-precision odd kernel: 2 x 256 threads x 128
    // registers - no halo CTA could start next to them before the launch retires, so a few SMs are kept out of its grid.
    template <typename K>
    void halo_lane_fits(K kern, int threads_per_sm) {
        if (tile_reserve_sms >= 0) return;
        cudaFuncAttributes a{};
        MF_CUDA(cudaFuncGetAttributes(&a, kern));
        tile_reserve_sms = ((long long)a.numRegs * threads_per_sm > 65536 - 128 * 32) ? 4 : 0;
    }

    // Stage counts are sized for the 227 KB of shared memory of an SM.  MFLBM_VARIANT = 100*e + o
    // selects other (even, odd) configurations for tuning runs, for the shipped MRT model only.
    template <int MRT>
    void launch_collide_default(bool odd) {
        if (tile_end <= tile_first) return;
        // MFLBM_VARIANT = 100 * even + odd selects the other configurations that were measured (profiles/README.md); they
        // are only instantiated for the shipped MRT model.  0 = the defaults.  Results are identical whatever the variant.
        bool done = false;
        if constexpr (MRT == 2) {
            const int e = (variant % 10000) / 100, o = variant % 100;
            done = true;
            if constexpr (sizeof(T) == 8) {
                if (odd && o == 1) launch_odd_ws<MRT, 2, 1>();
                else if (odd && o == 2) launch_odd_ws<MRT, 4, 1>();
                else if (odd && o == 3) launch_odd_ws<MRT, 4, 1, 2, 72, 216>();   // two consumer warpgroups, registers moved with setmaxnreg
                else if (!odd && e == 1) launch_even<MRT, 3, 1>();
                else done = false;
            } else {
                if (odd && o == 1) launch_odd_ws<MRT, 4, 1>();
                else if (odd && o == 2) launch_odd_ws<MRT, 3, 2>();
                else if (odd && o == 3) launch_odd_ws<MRT, 4, 1, 2>();
                else if (!odd && e == 1) launch_even<MRT, 8, 1>();
                else if (!odd && e == 2) launch_even<MRT, 2, 4>();
                else done = false;
            }
        }
        if (!done) {
            if constexpr (sizeof(T) == 8) { if (odd) launch_odd_ws<MRT, 3, 1>(); else launch_even<MRT, 2, 2>(); }
            else { if (odd) launch_odd_ws<MRT, 2, 2>(); else launch_even<MRT, 4, 2>(); }
        }
        check_launch(); count();
    }

    // part 0: every tile; 1: the tiles that hold the nodes of the neighbour-facing columns (the first n_boundary entries);
    // 2: the other tiles, launched while the halo messages of part 1 are on their way
    void phase_collide(int ntime, int part = 0) {
        const int ntiles = ceil_div((int)n_fluid, COLLIDE_TILE), nb_tiles = std::min(ntiles, ceil_div((int)n_boundary, COLLIDE_TILE));
        tile_first = part == 2 ? nb_tiles : 0;
        tile_end = part == 1 ? nb_tiles : ntiles;
        tile_reserve_sms = part == 2 ? -1 : 0;
        const bool odd = (ntime % 2) != 0;
        switch (P.mrt) {
            case 1: launch_collide_default<1>(odd); break;
            case 3: launch_collide_default<3>(odd); break;
            case 4: launch_collide_default<4>(odd); break;
            default: launch_collide_default<2>(odd); break;
        }
    }

    // boundary kernels in the reference's order (:1903-1958 even, :1969-2023 odd).
    // part 0: everything; 1: the phase-field part (ghost layers of phi, phi_convec); 2: the distribution part.  The parts are
    // independent (kernels_step.cuh, k_open_z): step() runs part 2 on a second lane next to the gradient chain.
    void phase_boundaries(int ntime, int part = 0) {
        const bool odd = (ntime % 2) != 0;
        const bool do_phi = part != 2, do_pdf = part != 1;
        const dim3 b = block2();
        const int lo = plo(), hi = phi_();
        const int ni = hi - lo + 1;
        const dim3 gz(ceil_div(ni, b.x), ceil_div(L.ny, b.y), 1), gy(ceil_div(ni, b.x), ceil_div(L.nz, b.y), 1), ge(ceil_div(ni, b.x), 1, 1);
        if (P.kper) {
            if (do_pdf) { if (odd) k_periodic_pdf<T, 2, true><<<gz, b, 0, stream>>>(L, lo, hi); else k_periodic_pdf<T, 2, false><<<gz, b, 0, stream>>>(L, lo, hi); count(); }
            if (do_phi) { k_periodic_phi<T, 2><<<gz, b, 0, stream>>>(L, lo, hi); count(); }
            check_launch();
        }
        if (P.jper) {
            if (do_pdf) { if (odd) k_periodic_pdf<T, 1, true><<<gy, b, 0, stream>>>(L, lo, hi); else k_periodic_pdf<T, 1, false><<<gy, b, 0, stream>>>(L, lo, hi); count(); }
            if (do_phi) { k_periodic_phi<T, 1><<<gy, b, 0, stream>>>(L, lo, hi); count(); }
            check_launch();
        }
        if (P.jper && P.kper) {
            const dim3 be(b.x, 1, 1);
            if (do_pdf) { if (odd) k_periodic_pdf_edges<T, true><<<ge, be, 0, stream>>>(L, lo, hi); else k_periodic_pdf_edges<T, false><<<ge, be, 0, stream>>>(L, lo, hi); count(); }
            if (do_phi) { k_periodic_phi<T, 3><<<ge, be, 0, stream>>>(L, lo, hi); count(); }
            check_launch();
        }
        const dim3 gp(ceil_div(L.nx, b.x), ceil_div(L.ny, b.y), 1);
        if (open_z()) {
            // inlet and outlet planes in one launch (k_open_z)
            const dim3 g2(gp.x, gp.y, 2);
            const int in = (P.inlet_BC == 1 || P.inlet_BC == 2) ? P.inlet_BC : 0, out = (P.outlet_BC == 1 || P.outlet_BC == 2) ? P.outlet_BC : 0;
            if (in || out) {
#define MF_OPEN_Z(I, O)                                                                                                                  \
                if (in == I && out == O) {                                                                                               \
                    if (part == 1) k_open_z<T, false, I, O, 1><<<g2, b, 0, stream>>>(L, 1, L.nx);                                        \
                    else if (part == 2) { if (odd) k_open_z<T, true, I, O, 2><<<g2, b, 0, stream>>>(L, 1, L.nx); else k_open_z<T, false, I, O, 2><<<g2, b, 0, stream>>>(L, 1, L.nx); } \
                    else { if (odd) k_open_z<T, true, I, O, 0><<<g2, b, 0, stream>>>(L, 1, L.nx); else k_open_z<T, false, I, O, 0><<<g2, b, 0, stream>>>(L, 1, L.nx); } \
                }
                MF_OPEN_Z(0, 1) MF_OPEN_Z(0, 2) MF_OPEN_Z(1, 0) MF_OPEN_Z(1, 1) MF_OPEN_Z(1, 2) MF_OPEN_Z(2, 0) MF_OPEN_Z(2, 1) MF_OPEN_Z(2, 2)
#undef MF_OPEN_Z
                count();
            }
            check_launch();
        }
        if (P.porous_plate_cmd != 0 && do_pdf) {
            if (odd) k_porous_plate<T, true><<<gz, b, 0, stream>>>(L, 1, L.nx, lo, hi); else k_porous_plate<T, false><<<gz, b, 0, stream>>>(L, 1, L.nx, lo, hi);
            check_launch(); count();
        }
    }

    // boundaries + gradient chain of one step on two lanes: phase-field part of the boundary kernels, then their distribution
    // part (second lane) next to the chain (first lane); joined before anything else is enqueued
    void boundaries_and_chain(int ntime, bool exchange_phi) {
        phase_boundaries(ntime, 1);
        MF_CUDA(cudaEventRecord(ev_bc_fork, stream));
        MF_CUDA(cudaStreamWaitEvent(bc_stream, ev_bc_fork, 0));
        std::swap(stream, bc_stream);
        try { phase_boundaries(ntime, 2); } catch (...) { std::swap(stream, bc_stream); throw; }
        std::swap(stream, bc_stream);
        MF_CUDA(cudaEventRecord(ev_bc_join, bc_stream));
        if (exchange_phi) exchange_p2p(2);
        gradient_chain(ntime & 1);
        MF_CUDA(cudaStreamWaitEvent(stream, ev_bc_join, 0));
    }

    void step_phase(int ntime, int phase) {
        if (!have_geometry) MF_FAIL("step before geometry");
        if (phase == 0) phase_collide(ntime, 0);
        else if (phase == 1) phase_boundaries(ntime);
        else if (phase == 2) gradient_chain(ntime & 1);
        else MF_FAIL("bad phase");
    }

    void step(int ntime) {
        if (!have_geometry) MF_FAIL("step before geometry");
        if (is_slab && (slab.has_left || slab.has_right)) {
            if (!p2p_ready()) MF_FAIL("a slab with neighbours steps through step_phase + halo exchange, or through step/run once halo_p2p_connect has been called for every neighbour");
            step_p2p(ntime);
            return;
        }
        phase_collide(ntime, 0);
        boundaries_and_chain(ntime, false);
    }

    // nsteps consecutive steps; pairs of steps are replayed from a captured CUDA graph (launch-bound small lattices)
    void run(int ntime_first, int nsteps) {
        if (nsteps <= 0) return;
        if (!have_geometry) MF_FAIL("step before geometry");
        if (is_slab && (slab.has_left || slab.has_right)) check_halo_error();   // fail fast once a message has been lost
        int nt = ntime_first, left = nsteps;
        // the collide launch carries bulk_skip = cn_consistent as an argument: step once outside the graph after arrays came
        // from outside, so that the captured pair has the fast setting in both of its collide launches
        if (!cn_consistent && left > 0) { step(nt); nt++; left--; }
        if (left >= 4) {
            const int par = nt & 1;
            // both step-pair graphs are built at the first call that needs one (capture does not execute): a later batch that
            // starts on the other parity - e.g. timed steps after an odd number of warm-up steps - finds its graph ready
            for (int k = 0; k < 2; k++) {
                const int pp = par ^ k;
                if (graph_exec[pp]) continue;
                const long long l0 = launches;
                cudaGraph_t g = nullptr;
                MF_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                try { step(nt + k); step(nt + k + 1); }
                catch (...) { cudaStreamEndCapture(stream, &g); if (g) cudaGraphDestroy(g); throw; }
                MF_CUDA(cudaStreamEndCapture(stream, &g));
                MF_CUDA(cudaGraphInstantiate(&graph_exec[pp], g, 0));
                MF_CUDA(cudaGraphDestroy(g));
                pair_launches = launches - l0;
                launches = l0;   // capture does not execute
            }
            while (left >= 2) { MF_CUDA(cudaGraphLaunch(graph_exec[par], stream)); launches += pair_launches; nt += 2; left -= 2; }
        }
        for (; left > 0; left--, nt++) step(nt);
    }

    // ------------------------------------------------------------------------------------------------
    // a halo message that never arrived (kernels_aux.cuh, HALO_TIMEOUT_CYCLES): surface it at the next host synchronisation
    void check_halo_error() {
        if (!d_flags) return;
        unsigned e = 0;
        MF_CUDA(cudaMemcpyAsync(&e, d_flags + 31, sizeof e, cudaMemcpyDeviceToHost, stream));
        MF_CUDA(cudaStreamSynchronize(stream));
        if (e) MF_FAIL("halo message %u (1 + 2*kind + side) from a neighbour slab did not arrive within the time-out", e);
    }

    void monitor(mflbm_monitor_out* out) {
        if (!have_geometry) MF_FAIL("monitor before geometry");
        if (!out) MF_FAIL("monitor: null output");
        check_halo_error();
        k_monitor<T><<<L.nz, 256, 0, stream>>>(L, d_zstart, d_mon); check_launch(); count();
        MF_CUDA(cudaMemcpyAsync(h_mon, d_mon, sizeof(double) * MFLBM_MON_N * L.nz, cudaMemcpyDeviceToHost, stream));
        MF_CUDA(cudaStreamSynchronize(stream));
        const int nz = L.nz;
        auto at = [&](int k, int n) { return h_mon[(size_t)(k - 1) * MFLBM_MON_N + n]; };
        double v1 = 0, v2 = 0, m1 = 0, m2 = 0, v1f = 0, v2f = 0, m1f = 0, m2f = 0, f1 = 0, f2 = 0, f1w = 0, f2w = 0, umax = 0, k1 = 0, k2 = 0, nanf = 0;
        double pw = 0, nw = 0, pnw = 0, nnw = 0;
        for (int k = 1; k <= nz; k++) {
            const bool in = k >= P.n_exclude_inlet + 1 && k <= nz - P.n_exclude_outlet;
            if (in) { m1 += at(k, 3); m2 += at(k, 4); v1 += at(k, 5); v2 += at(k, 6); f1 += at(k, 0); f2 += at(k, 1); }
            m1f += at(k, 3); m2f += at(k, 4); v1f += at(k, 5); v2f += at(k, 6); f1w += at(k, 0); f2w += at(k, 1);
            umax = std::max(umax, at(k, 7)); k1 += at(k, 8); k2 += at(k, 9); nanf = std::max(nanf, at(k, 10));
            pw += at(k, 11); nw += at(k, 12); pnw += at(k, 13); nnw += at(k, 14);
            if (out->fl1) out->fl1[k - 1] = at(k, 0);
            if (out->fl2) out->fl2[k - 1] = at(k, 1);
            if (out->pre) out->pre[k - 1] = at(k, 2);
            if (out->mass1) out->mass1[k - 1] = at(k, 3);
            if (out->mass2) out->mass2[k - 1] = at(k, 4);
            if (out->vol1) out->vol1[k - 1] = at(k, 5);
            if (out->vol2) out->vol2[k - 1] = at(k, 6);
        }
        out->vol1_sum = v1; out->vol2_sum = v2; out->mass1_sum = m1; out->mass2_sum = m2;
        out->vol1_full = v1f; out->vol2_full = v2f; out->mass1_full = m1f; out->mass2_full = m2f;
        out->saturation = v1 / (v1 + v2);
        out->saturation_full_domain = v1f / (v1f + v2f);
        const double nin = (double)(nz - P.n_exclude_outlet - P.n_exclude_inlet);
        out->fl1_avg = f1 / nin; out->fl2_avg = f2 / nin; out->fl1_avg_whole = f1w / nz; out->fl2_avg_whole = f2w / nz;
        out->ca = ((out->fl1_avg + out->fl2_avg) / (double)P.A_xy) * (double)P.la_nu1 / (double)P.lbm_gamma;   // Monitor.cpp:170-171
        out->umax = std::sqrt(umax);
        out->kinetic_energy[0] = 0.5 * k1; out->kinetic_energy[1] = 0.5 * k2;
        out->nan_detected = (nanf != 0.0) || std::isnan(out->saturation_full_domain) || std::isnan(out->ca);
        out->pre_w_sum = pw; out->pre_nw_sum = pnw; out->n_w = (int64_t)nw; out->n_nw = (int64_t)nnw;
        out->outlet_phase1_count = nz >= 2 ? (int64_t)at(nz - 1, 15) : 0;   // obs_z = nzGlobal - 1, Monitor.cpp:452
    }

    void phi_change(int seed, double* d_phi_max) {
        if (!have_geometry) MF_FAIL("phi_change before geometry");
        const bool fresh = d_phi_old == nullptr;
        if (fresh) zalloc((void**)&d_phi_old, sizeof(T) * std::max<long long>(n_fluid, 1));
        if (seed) {
            if (n_fluid) { k_phi_change<T, 1><<<std::min(ceil_div((int)n_fluid, 256), 148 * 8), 256, 0, stream>>>(L, d_phi_old, nullptr); check_launch(); count(); }
            if (d_phi_max) *d_phi_max = 0.0;
            return;
        }
        const int nb = std::max(1, std::min(ceil_div((int)n_fluid, 256), MFLBM_MON_N * L.nz));   // partials reuse the monitor buffers
        k_phi_change<T, 0><<<nb, 256, 0, stream>>>(L, d_phi_old, d_mon); check_launch(); count();
        MF_CUDA(cudaMemcpyAsync(h_mon, d_mon, sizeof(double) * nb, cudaMemcpyDeviceToHost, stream));
        MF_CUDA(cudaStreamSynchronize(stream));
        double m = 0.0;
        for (int b = 0; b < nb; b++) m = (h_mon[b] != h_mon[b] || m != m) ? (m != m ? m : h_mon[b]) : std::max(m, h_mon[b]);
        if (d_phi_max) *d_phi_max = m;
    }

    void download_macro(T* rho, T* u, T* v, T* w) {
        if (!have_geometry) MF_FAIL("download_macro before geometry");
        check_halo_error();
        T* out[4] = {rho, u, v, w};
        int n = 0;
        for (auto q : out) n += q != nullptr;
        if (!n) return;
        T* st = (T*)stage(sizeof(T) * N1 * n);
        MF_CUDA(cudaMemsetAsync(st, 0, sizeof(T) * N1 * n, stream));
        T* dev[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int a = 0, m = 0; a < 4; a++) if (out[a]) dev[a] = st + (size_t)N1 * m++;
        if (n_fluid) { k_macro<T><<<ceil_div((int)n_fluid, 128), 128, 0, stream>>>(L, dev[0], dev[1], dev[2], dev[3]); check_launch(); count(); }
        for (int a = 0; a < 4; a++) if (out[a]) MF_CUDA(cudaMemcpyAsync(out[a], dev[a], sizeof(T) * N1, cudaMemcpyDeviceToHost, stream));
        MF_CUDA(cudaStreamSynchronize(stream));
        drop_stage();
    }

    // ------------------------------------------------------------------------------------------------
    // halo exchange (x slabs)
    long long halo_count(int kind) const { return kind == 2 ? 4LL * L.PY * L.PZ : 10LL * L.NY1 * L.NZ1; }

    // pack the outgoing columns of message `kind`.  push = false: into my send buffers (the caller moves them, e.g. NCCL);
    // push = true: straight into the neighbours' receive buffers over NVLink, arrival published in their flags
    void halo_pack(int kind, bool push = false, int sides = 3, cudaStream_t on = nullptr) {
        cudaStream_t stream = on ? on : this->stream;
        const bool do_left = slab.has_left && (sides & 1), do_right = slab.has_right && (sides & 2);
        if (!is_slab) MF_FAIL("halo_pack on a non-slab solver");
        if (!have_geometry) MF_FAIL("halo_pack before geometry");
        if (kind < 0 || kind > 2) MF_FAIL("bad halo kind");
        const int bt = 128;
        const dim3 g1(ceil_div(L.NY1 * L.NZ1, bt), 5), g4(ceil_div(L.PY, bt), L.PZ);
        auto dst = [&](int side) -> T* {
            if (!push) return d_send[kind][side];
            if (!peer_recv[kind][side] || !peer_flag[kind][side]) MF_FAIL("halo_push: neighbour %d of message kind %d is not connected", side, kind);
            return peer_recv[kind][side];
        };
        auto sync = [&](int side) -> HaloSync {
            if (!push) return HaloSync{nullptr, nullptr, nullptr, nullptr};
            return HaloSync{peer_flag[kind][side], d_flags + 8 + kind * 2 + side, d_flags + 16 + kind * 2 + side, nullptr};
        };
        if (kind == 0) {   // after an even step: real boundary columns -> neighbour ghost columns
            if (do_left) { k_halo_pdf<T, false, true><<<g1, bt, 0, stream>>>(L, dst(0), d_face[1], sync(0)); count(); }      // ex=-1 slots of column 1
            if (do_right) { k_halo_pdf<T, true, true><<<g1, bt, 0, stream>>>(L, dst(1), d_face[2], sync(1)); count(); }      // ex=+1 slots of column nx
        } else if (kind == 1) {   // after an odd step: what was pushed into my ghost columns -> neighbour real columns
            if (do_left) { k_halo_pdf<T, true, true><<<g1, bt, 0, stream>>>(L, dst(0), d_face[0], sync(0)); count(); }       // ex=+1 slots of ghost column 0
            if (do_right) { k_halo_pdf<T, false, true><<<g1, bt, 0, stream>>>(L, dst(1), d_face[3], sync(1)); count(); }     // ex=-1 slots of ghost column nx+1
        } else {
            if (do_left) { k_halo_phi<T, true><<<g4, bt, 0, stream>>>(L, dst(0), 1, sync(0)); count(); }
            if (do_right) { k_halo_phi<T, true><<<g4, bt, 0, stream>>>(L, dst(1), L.nx - 3, sync(1)); count(); }
        }
        check_launch();
    }

    // unpack message `kind` from my receive buffers; wait = true: spin on my flags until the neighbours' pushes have landed
    // wait: 0 none (the caller moved the message, e.g. NCCL), 1 every CTA of the unpack kernel spins on my flag until the
    // neighbour's push has landed, 2 a one-CTA wait kernel first (overlapped schedule, see k_halo_wait)
    void halo_unpack(int kind, int wait = 0, int sides = 3, cudaStream_t on = nullptr) {
        cudaStream_t stream = on ? on : this->stream;
        const bool do_left = slab.has_left && (sides & 1), do_right = slab.has_right && (sides & 2);
        if (!is_slab) MF_FAIL("halo_unpack on a non-slab solver");
        if (!have_geometry) MF_FAIL("halo_unpack before geometry");
        if (kind < 0 || kind > 2) MF_FAIL("bad halo kind");
        const int bt = 128;
        const dim3 g1(ceil_div(L.NY1 * L.NZ1, bt), 5), g4(ceil_div(L.PY, bt), L.PZ);
        auto sync = [&](int side) -> HaloSync {
            if (!wait) return HaloSync{nullptr, nullptr, nullptr, nullptr};
            const HaloSync hs{d_flags + kind * 2 + side, nullptr, d_flags + 16 + kind * 2 + side, d_flags + 31};
            if (wait == 1) return hs;
            k_halo_wait<<<1, 32, 0, stream>>>(hs); count();
            return HaloSync{nullptr, nullptr, nullptr, d_flags + 31};
        };
        if (kind == 0) {   // neighbour's real boundary column -> my ghost column
            if (do_left) { k_halo_pdf<T, true, false><<<g1, bt, 0, stream>>>(L, d_recv[0][0], d_face[0], sync(0)); count(); }     // left's column nx (ex=+1) -> ghost 0
            if (do_right) { k_halo_pdf<T, false, false><<<g1, bt, 0, stream>>>(L, d_recv[0][1], d_face[3], sync(1)); count(); }   // right's column 1 (ex=-1) -> ghost nx+1
        } else if (kind == 1) {   // neighbour's ghost column -> my real boundary column
            if (do_left) { k_halo_pdf<T, false, false><<<g1, bt, 0, stream>>>(L, d_recv[1][0], d_face[1], sync(0)); count(); }    // left's ghost nx+1 (ex=-1) -> column 1
            if (do_right) { k_halo_pdf<T, true, false><<<g1, bt, 0, stream>>>(L, d_recv[1][1], d_face[2], sync(1)); count(); }    // right's ghost 0 (ex=+1) -> column nx
        } else {
            if (do_left) { k_halo_phi<T, false><<<g4, bt, 0, stream>>>(L, d_recv[2][0], -3, sync(0)); count(); }
            if (do_right) { k_halo_phi<T, false><<<g4, bt, 0, stream>>>(L, d_recv[2][1], L.nx + 1, sync(1)); count(); }
        }
        check_launch();
    }

    void halo_connect(int kind, int side, T* recv_of_peer, unsigned* flag_of_peer) {
        if (!is_slab) MF_FAIL("halo_connect on a non-slab solver");
        if (kind < 0 || kind > 2 || side < 0 || side > 1) MF_FAIL("halo_connect: bad kind/side");
        peer_recv[kind][side] = recv_of_peer;
        peer_flag[kind][side] = flag_of_peer;
        drop_graphs();
    }
    bool p2p_ready() const {
        for (int kind = 0; kind < 3; kind++) {
            if (slab.has_left && !(peer_recv[kind][0] && peer_flag[kind][0])) return false;
            if (slab.has_right && !(peer_recv[kind][1] && peer_flag[kind][1])) return false;
        }
        return true;
    }
    // one step of a slab with its two halo messages through peer memory (the sequence of mflbm/slab.py SlabStepper.step)
    // one message of `kind` to and from every neighbour.  With two neighbours the right-hand side runs on a second stream
    // (event fork / join, capturable): the four small latency-bound kernels of an interior slab overlap pairwise.
    void exchange_p2p(int kind) {
        if (slab.has_left && slab.has_right) {
            MF_CUDA(cudaEventRecord(ev_fork, stream));
            MF_CUDA(cudaStreamWaitEvent(aux_stream, ev_fork, 0));
            halo_pack(kind, true, 2, aux_stream); halo_unpack(kind, 1, 2, aux_stream);
            halo_pack(kind, true, 1); halo_unpack(kind, 1, 1);
            MF_CUDA(cudaEventRecord(ev_join, aux_stream));
            MF_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
        } else {
            halo_pack(kind, true); halo_unpack(kind, 1);
        }
    }
    // One step of a slab with its halo messages through peer memory.  The nodes of the neighbour-facing columns are the
    // first entries of the fluid order (finish_geometry): their collide tiles run first, then the PDF message - push into the
    // neighbours' buffers, wait for theirs, unpack - travels on the second lane while the collide launch of all other tiles
    // runs on the first.  Safe: in the AA pattern every PDF cell is read and written by ONE node per step; the cells a message
    // packs are written by boundary-column nodes only, the cells it unpacks into belong to ghost-column nodes (odd step) or
    // are not touched at all during an even step.  The boundary kernels need both and join the lanes.
    void step_p2p(int ntime) {
        if (!have_geometry) MF_FAIL("step before geometry");
        const int kind = (ntime % 2) ? 1 : 0;
        if (n_boundary > 0 && n_boundary < n_fluid && !no_overlap) {
            phase_collide(ntime, 1);
            MF_CUDA(cudaEventRecord(ev_fork, stream));
            MF_CUDA(cudaStreamWaitEvent(aux_stream, ev_fork, 0));
            halo_pack(kind, true, 3, aux_stream);
            halo_unpack(kind, 2, 3, aux_stream);
            MF_CUDA(cudaEventRecord(ev_join, aux_stream));
            phase_collide(ntime, 2);
            MF_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
        } else {
            phase_collide(ntime, 0);
            exchange_p2p(kind);
        }
        boundaries_and_chain(ntime, true);
    }

    void* device_ptr(const char* name) {
#define MF_PTR(n, p) if (!strcmp(name, n)) return (void*)(p)
        MF_PTR("pdf", d_pdf); MF_PTR("phi", d_phi); MF_PTR("cn_x", d_cnx); MF_PTR("cn_y", d_cny); MF_PTR("cn_z", d_cnz); MF_PTR("c_norm", d_cnorm);
        MF_PTR("W_in", d_Win); MF_PTR("f_convec", d_fconv); MF_PTR("g_convec", d_gconv); MF_PTR("phi_convec", d_phiconv);
        MF_PTR("types", d_types); MF_PTR("site_map", d_cmap); MF_PTR("fluid_sites", d_flu); MF_PTR("chain_bricks_processed", d_n_active);
#undef MF_PTR
        return nullptr;
    }
};

}  // namespace mflbm

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace mflbm;

#define MF_GUARD(body)                                                    \
    try { body; return 0; }                                               \
    catch (const Error& e) { g_last_error = e.msg; return 1; }            \
    catch (const std::exception& e) { g_last_error = e.what(); return 2; } \
    catch (...) { g_last_error = "unknown error"; return 3; }

// every entry point runs with the solver's device current and restores the caller's afterwards (a caller that drives
// several devices from one thread, e.g. mflbm_run with x-slabs, must not launch onto foreign streams)
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(int dev) : want(dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want && cudaSetDevice(want) != cudaSuccess) throw Error{"cudaSetDevice failed for the solver's device"};
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
};
#define MF_SOLVER(P, REAL) reinterpret_cast<Solver<REAL>*>(s)
#define MF_NEED(s) if (!(s)) throw Error{"null solver handle"}
#define MF_ENTER(P, REAL) MF_NEED(s); DeviceGuard device_guard_(MF_SOLVER(P, REAL)->device)

#define MFLBM_DEFINE_API(P, REAL)                                                                                                     \
    extern "C" int mflbm_##P##_create(const mflbm_##P##_params* params, const mflbm_slab* slab, int device, void* stream,                 \
                                      mflbm_##P##_solver** out) {                                                                         \
        MF_GUARD({                                                                                                                        \
            if (!params || !out) throw Error{"create: null argument"};                                                                    \
            auto* sv = new Solver<REAL>();                                                                                                \
            DeviceGuard device_guard_(device);                                                                                            \
            try { sv->create(params, slab, device, stream); } catch (...) { sv->destroy(); delete sv; throw; }                            \
            *out = reinterpret_cast<mflbm_##P##_solver*>(sv);                                                                             \
        })                                                                                                                                \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_destroy(mflbm_##P##_solver* s) {                                                                          \
        MF_GUARD({ if (s) { DeviceGuard device_guard_(MF_SOLVER(P, REAL)->device); MF_SOLVER(P, REAL)->destroy(); delete MF_SOLVER(P, REAL); } }) \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_set_params(mflbm_##P##_solver* s, const mflbm_##P##_params* params) {                                      \
        MF_GUARD({ MF_ENTER(P, REAL); if (!params) throw Error{"null params"}; MF_SOLVER(P, REAL)->set_params(params); })                        \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_upload_geometry(mflbm_##P##_solver* s, const int32_t* walls, const int32_t* walls_type, const REAL* s_nx,   \
                                               const REAL* s_ny, const REAL* s_nz) {                                                      \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->upload_geometry(walls, walls_type, s_nx, s_ny, s_nz); })                               \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_preprocess_geometry(mflbm_##P##_solver* s, const int8_t* w) {                                              \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->preprocess_geometry(w); })                                                             \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_download_geometry(mflbm_##P##_solver* s, int32_t* walls, int32_t* walls_type, REAL* s_nx, REAL* s_ny,       \
                                                 REAL* s_nz, int64_t* counts) {                                                           \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->download_geometry(walls, walls_type, s_nx, s_ny, s_nz, counts); })                     \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_upload_state(mflbm_##P##_solver* s, const REAL* pdf, const REAL* phi, const REAL* cn_x, const REAL* cn_y,   \
                                            const REAL* cn_z, const REAL* c_norm, const REAL* curv, const REAL* W_in, const REAL* f_convec, \
                                            const REAL* g_convec, const REAL* phi_convec) {                                               \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->upload_state(pdf, phi, cn_x, cn_y, cn_z, c_norm, curv, W_in, f_convec, g_convec, phi_convec); }) \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_init_state(mflbm_##P##_solver* s, int option, REAL interface_z0, const REAL* W_in) {                       \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->init_state(option, interface_z0, W_in); })                                             \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_init_state_from_phi(mflbm_##P##_solver* s, const REAL* phi, const REAL* W_in) {                            \
        MF_GUARD({ MF_ENTER(P, REAL); if (!phi) throw Error{"init_state_from_phi: null phi"}; MF_SOLVER(P, REAL)->init_state(0, REAL(0), W_in, phi); }) \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_download_state(mflbm_##P##_solver* s, REAL* pdf, REAL* phi, REAL* cn_x, REAL* cn_y, REAL* cn_z,             \
                                              REAL* c_norm, REAL* curv, REAL* f_convec, REAL* g_convec, REAL* phi_convec) {               \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->download_state(pdf, phi, cn_x, cn_y, cn_z, c_norm, curv, f_convec, g_convec, phi_convec); }) \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_step(mflbm_##P##_solver* s, int ntime) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->step(ntime); }) }       \
    extern "C" int mflbm_##P##_run(mflbm_##P##_solver* s, int ntime_first, int nsteps) {                                                  \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->run(ntime_first, nsteps); })                                                           \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_color_gradient(mflbm_##P##_solver* s) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->gradient_chain(); }) } \
    extern "C" int mflbm_##P##_monitor(mflbm_##P##_solver* s, mflbm_monitor_out* out) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->monitor(out); }) } \
    extern "C" int mflbm_##P##_phi_change(mflbm_##P##_solver* s, int seed, double* d_phi_max) {                                           \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->phi_change(seed, d_phi_max); })                                                        \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_download_macro(mflbm_##P##_solver* s, REAL* rho, REAL* u, REAL* v, REAL* w) {                              \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->download_macro(rho, u, v, w); })                                                       \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_sync(mflbm_##P##_solver* s) {                                                                              \
        MF_GUARD({ MF_ENTER(P, REAL); MF_CUDA(cudaStreamSynchronize(MF_SOLVER(P, REAL)->stream)); MF_SOLVER(P, REAL)->check_halo_error(); })      \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_halo_buffers(mflbm_##P##_solver* s, int kind, int side, REAL** send, REAL** recv, int64_t* count) {        \
        MF_GUARD({                                                                                                                        \
            MF_ENTER(P, REAL);                                                                                                                   \
            if (kind < 0 || kind > 2 || side < 0 || side > 1) throw Error{"halo_buffers: bad kind/side"};                                 \
            if (!MF_SOLVER(P, REAL)->is_slab) throw Error{"halo_buffers on a non-slab solver"};                                           \
            if (send) *send = MF_SOLVER(P, REAL)->d_send[kind][side];                                                                     \
            if (recv) *recv = MF_SOLVER(P, REAL)->d_recv[kind][side];                                                                     \
            if (count) *count = MF_SOLVER(P, REAL)->halo_count(kind);                                                                     \
        })                                                                                                                                \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_halo_pack(mflbm_##P##_solver* s, int kind) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->halo_pack(kind); }) } \
    extern "C" int mflbm_##P##_halo_unpack(mflbm_##P##_solver* s, int kind) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->halo_unpack(kind); }) } \
    extern "C" int mflbm_##P##_halo_p2p_local(mflbm_##P##_solver* s, int kind, int side, REAL** recv, uint32_t** flag) {                  \
        MF_GUARD({                                                                                                                        \
            MF_ENTER(P, REAL);                                                                                                                   \
            if (kind < 0 || kind > 2 || side < 0 || side > 1) throw Error{"halo_p2p_local: bad kind/side"};                               \
            if (!MF_SOLVER(P, REAL)->is_slab) throw Error{"halo_p2p_local on a non-slab solver"};                                         \
            if (recv) *recv = MF_SOLVER(P, REAL)->d_recv[kind][side];                                                                     \
            if (flag) *flag = MF_SOLVER(P, REAL)->d_flags + kind * 2 + side;                                                              \
        })                                                                                                                                \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_halo_p2p_region(mflbm_##P##_solver* s, void** base, int64_t* bytes) {                                      \
        MF_GUARD({                                                                                                                        \
            MF_ENTER(P, REAL);                                                                                                                   \
            if (!MF_SOLVER(P, REAL)->is_slab) throw Error{"halo_p2p_region on a non-slab solver"};                                        \
            if (base) *base = MF_SOLVER(P, REAL)->d_p2p;                                                                                  \
            if (bytes) *bytes = (int64_t)MF_SOLVER(P, REAL)->p2p_bytes;                                                                   \
        })                                                                                                                                \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_halo_p2p_connect(mflbm_##P##_solver* s, int kind, int side, REAL* peer_recv, uint32_t* peer_flag) {        \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->halo_connect(kind, side, peer_recv, peer_flag); })                                     \
    }                                                                                                                                     \
    extern "C" int mflbm_##P##_halo_push(mflbm_##P##_solver* s, int kind) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->halo_pack(kind, true); }) } \
    extern "C" int mflbm_##P##_halo_unpack_wait(mflbm_##P##_solver* s, int kind) { MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->halo_unpack(kind, 1); }) } \
    extern "C" int mflbm_##P##_step_phase(mflbm_##P##_solver* s, int ntime, int phase) {                                                  \
        MF_GUARD({ MF_ENTER(P, REAL); MF_SOLVER(P, REAL)->step_phase(ntime, phase); })                                                           \
    }                                                                                                                                     \
    extern "C" int64_t mflbm_##P##_num_fluid_nodes(mflbm_##P##_solver* s) { return s ? MF_SOLVER(P, REAL)->n_fluid : -1; }                \
    extern "C" int64_t mflbm_##P##_kernel_launches(mflbm_##P##_solver* s) { return s ? MF_SOLVER(P, REAL)->launches : -1; }               \
    extern "C" void* mflbm_##P##_stream(mflbm_##P##_solver* s) { return s ? (void*)MF_SOLVER(P, REAL)->stream : nullptr; }                \
    extern "C" void* mflbm_##P##_device_ptr(mflbm_##P##_solver* s, const char* name) {                                                    \
        return (s && name) ? MF_SOLVER(P, REAL)->device_ptr(name) : nullptr;                                                              \
    }

MFLBM_DEFINE_API(f32, float)
MFLBM_DEFINE_API(f64, double)

// CUDA IPC helpers so that a caller in another process can reach the buffers returned by halo_p2p_local
extern "C" int mflbm_ipc_export(void* device_ptr, void* handle64) {
    MF_GUARD({
        if (!device_ptr || !handle64) throw Error{"ipc_export: null argument"};
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        cudaIpcMemHandle_t h;
        MF_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
        memcpy(handle64, &h, 64);
    })
}
extern "C" int mflbm_ipc_import(const void* handle64, void** device_ptr) {
    MF_GUARD({
        if (!device_ptr || !handle64) throw Error{"ipc_import: null argument"};
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        MF_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    })
}
extern "C" int mflbm_ipc_release(void* device_ptr) { MF_GUARD({ if (device_ptr) MF_CUDA(cudaIpcCloseMemHandle(device_ptr)); }) }

// x-slabs of one lattice on several devices driven from ONE process (mflbm_run --gpus N): peer access both ways, so that the
// pointers returned by halo_p2p_local can be handed to halo_p2p_connect of the neighbour as they are
static void peer_enable(int device_a, int device_b) {
    if (device_a == device_b) return;
    int prev = 0, ok_ab = 0, ok_ba = 0;
    MF_CUDA(cudaGetDevice(&prev));
    MF_CUDA(cudaDeviceCanAccessPeer(&ok_ab, device_a, device_b));
    MF_CUDA(cudaDeviceCanAccessPeer(&ok_ba, device_b, device_a));
    if (!ok_ab || !ok_ba) MF_FAIL("devices %d and %d cannot access each other's memory", device_a, device_b);
    for (int k = 0; k < 2; k++) {
        const int from = k ? device_b : device_a, to = k ? device_a : device_b;
        MF_CUDA(cudaSetDevice(from));
        const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { cudaSetDevice(prev); MF_CUDA(e); }
    }
    MF_CUDA(cudaSetDevice(prev));
}
extern "C" int mflbm_peer_enable(int device_a, int device_b) { MF_GUARD(peer_enable(device_a, device_b)) }

extern "C" const char* mflbm_last_error(void) { return g_last_error.c_str(); }
extern "C" int mflbm_version(void) { return MFLBM_VERSION; }
