/* =====================================================================================================
 * mflbm.h — C ABI of the B200-native MF-LBM time-step library (libmflbm.so).
 *
 * Drop-in boundary for the GPU layer of MF-LBM-CUDA.  The reference has no FFI: its GPU layer is five free
 * functions working on global variables (paths relative to /root/reference):
 *
 *   includes/Init_multiphase_GPU.h:5-10   void initialization_GPU(); MemAllocate_geometry_GPU(int); MemAllocate_multi_GPU(int);
 *   includes/main_iteration_GPU.h:4-7     void main_iteration_kernel_GPU(); void copyConstantData();
 *
 * and an implicit post-condition: on "timer" steps the host arrays phi, curv, c_norm, cn_*, pdf hold the device
 * state (src/main_iteration_GPU.cu:2059-2076).  This header turns that de-facto API into explicit entry points:
 * every global the GPU layer reads becomes a field of mflbm_params or an argument, every array keeps the
 * reference's layout at the boundary (includes/Idx_gpu.cuh:52-70: 1-based, x fastest, ghost widths 0/1/2/4).
 *
 * Conventions
 *   - one symbol set per precision: mflbm_f32_* (real = float) and mflbm_f64_* (real = double), mirroring the
 *     reference's compile-time T_P (includes/solver_precision.h:8-22);
 *   - every call returns 0 on success, non-zero on failure; mflbm_last_error() gives the message of the last
 *     failure on the calling thread.  No exceptions or C++ types cross the boundary.  (The reference prints and
 *     exit()s instead: includes/utils_GPU.cuh:8-14.)
 *   - host arrays stay caller-owned; the library owns device memory;
 *   - calls are asynchronous on the solver's stream unless they return data to the host.
 *
 * Array shapes (element counts), with NXg = nx + 2g etc.:
 *   s0: nx*ny*nz            s1: NX1*NY1*NZ1        s2: NX2*NY2*NZ2        s4: NX4*NY4*NZ4
 *   pdf: 38 * s1, index x + NX1*(y + NY1*(z + NZ1*(e + 19*g)))            (Idx_gpu.cuh:66)
 *   f_convec/g_convec: 19*NX1*NY1  (Idx_gpu.cuh:70),  phi_convec, W_in: NX1*NY1
 * For an x-slab (mflbm_slab), nx in these formulas is the slab's local width.
 * ===================================================================================================== */
#ifndef MFLBM_H
#define MFLBM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFLBM_VERSION 100

/* x-slab of a global lattice owned by one GPU (new; the reference is single-GPU, README.md:119).
 * x0 = global index (1-based) of the slab's first real column, nx_local = its width.  has_left/has_right tell
 * whether ghost columns on that side belong to a neighbour slab (halo exchange) or to the domain wall. */
typedef struct mflbm_slab {
    int64_t x0;
    int64_t nx_local;
    int32_t has_left;
    int32_t has_right;
} mflbm_slab;

/* Monitor reductions (src/Monitor.cpp:17-171 computed on the device).  All sums are accumulated in double
 * regardless of the solver precision; profile arrays have nz entries each and are caller-allocated. */
typedef struct mflbm_monitor_out {
    double saturation;             /* vol1/(vol1+vol2) over k in [n_exclude_inlet+1, nz-n_exclude_outlet]   Monitor.cpp:111-118 */
    double saturation_full_domain; /* same over all k                                                       Monitor.cpp:136-143 */
    double vol1_sum, vol2_sum, mass1_sum, mass2_sum;   /* excluded-layer sums                               Monitor.cpp:111-116 */
    double vol1_full, vol2_full, mass1_full, mass2_full;
    double fl1_avg, fl2_avg, fl1_avg_whole, fl2_avg_whole; /* Monitor.cpp:146-167 */
    double ca;                     /* Monitor.cpp:170-171 */
    double umax;                   /* sqrt(max |u|^2)                                                       Monitor.cpp:103 */
    double kinetic_energy[2];      /* Monitor.cpp:100-101 */
    int32_t nan_detected;          /* any non-finite value met while reducing                               Monitor.cpp:244 */
    int32_t reserved;
    double* fl1;  double* fl2;  double* pre;  double* mass1;  double* mass2;  double* vol1;  double* vol2;   /* per z slice, may be NULL */
    /* steady-state monitors and breakthrough test (src/Monitor.cpp:357-411, 447-466) */
    double pre_w_sum, pre_nw_sum;          /* sum of rho over fluid nodes with phi < -0.99 / phi > 0.99             Monitor.cpp:379-388 */
    int64_t n_w, n_nw;                     /* their counts */
    int64_t outlet_phase1_count;           /* fluid nodes of slice nz-1 with phi > 0                                Monitor.cpp:452-459 */
} mflbm_monitor_out;

#define MFLBM_DECLARE_API(P, REAL)                                                                                        \
    /* every scalar the reference uploads with copyConstantData (src/main_iteration_GPU.cu:14-47) or passes as a    */ \
    /* kernel argument (rho_in, rho_out), plus the host-side switches main_iteration_kernel_GPU branches on         */ \
    /* (:1903-1954).  Derived values (la_nui*, cos_theta, phi_inlet, ...) are supplied by the caller exactly as the  */ \
    /* reference host computes them (src/Init_multiphase.cpp:173-214, src/IO_multiphase.cpp:204-206).                */ \
    typedef struct mflbm_##P##_params {                                                                                  \
        int64_t nx, ny, nz;            /* nxGlobal, nyGlobal, nzGlobal */                                                \
        int32_t iper, jper, kper;      /* iper must be 0 (src/IO_multiphase.cpp:210) */                                  \
        int32_t wall_z_min, wall_z_max;/* domain_wall_status_z_* : open z ends enable the inlet/outlet kernels */        \
        int32_t inlet_BC, outlet_BC;   /* 1 velocity / convective, 2 Zou-He pressure */                                  \
        int32_t porous_plate_cmd, Z_porous_plate;                                                                        \
        int32_t n_exclude_inlet, n_exclude_outlet;                                                                       \
        int32_t mrt;                   /* includes/preprocessor.h:4, 1..4; shipped 2 */                                  \
        REAL lbm_gamma, lbm_beta, la_nu1, la_nui1, la_nui2, cos_theta, force_z;                                          \
        REAL rho_in, rho_out, phi_inlet, sa_inject, uin_avg, relaxation;                                                 \
        REAL A_xy;                     /* la_x*la_y, only used by the monitor's capillary number */                      \
    } mflbm_##P##_params;                                                                                                \
    typedef struct mflbm_##P##_solver mflbm_##P##_solver;                                                                \
                                                                                                                         \
    /* replaces initialization_GPU (src/Init_multiphase_GPU.cu:13-40): selects `device`, allocates device state.    */ \
    /* slab == NULL: the whole lattice on one GPU.  stream == NULL: the library creates its own stream; otherwise a */ \
    /* cudaStream_t owned by the caller (e.g. a torch stream) on which all work is enqueued.                        */ \
    int mflbm_##P##_create(const mflbm_##P##_params* params, const mflbm_slab* slab, int device, void* stream,           \
                           mflbm_##P##_solver** out);                                                                    \
    int mflbm_##P##_destroy(mflbm_##P##_solver* s);                                                                      \
    /* copyConstantData (src/main_iteration_GPU.cu:14): re-read scalar parameters (force_z, rho_in, ...) */             \
    int mflbm_##P##_set_params(mflbm_##P##_solver* s, const mflbm_##P##_params* params);                                 \
                                                                                                                         \
    /* MemAllocate_geometry_GPU(1) (src/Init_multiphase_GPU.cu:43-60): walls s2 int32, walls_type s4 int32,        */ \
    /* s_nx/s_ny/s_nz s4 real, as produced by geometry_preprocessing_new (src/Geometry_preprocessing.cpp:29).      */ \
    int mflbm_##P##_upload_geometry(mflbm_##P##_solver* s, const int32_t* walls, const int32_t* walls_type,              \
                                    const REAL* s_nx, const REAL* s_ny, const REAL* s_nz);                               \
    /* GPU twin of geometry_preprocessing_new, bit-exact.  walls_interior: s0 int8, 1 = solid, after */ \
    /* set_walls (src/Misc.cpp:17-101).  For a slab pass the GLOBAL interior array; the slab cuts its own window.  */ \
    int mflbm_##P##_preprocess_geometry(mflbm_##P##_solver* s, const int8_t* walls_interior_global);                     \
    /* any pointer may be NULL; counts[4] = num_solid_boundary_global, num_fluid_boundary_global,                  */ \
    /* num_solid_boundary, num_fluid_boundary (src/Geometry_preprocessing.cpp:208-222,389-401)                      */ \
    int mflbm_##P##_download_geometry(mflbm_##P##_solver* s, int32_t* walls, int32_t* walls_type, REAL* s_nx,            \
                                      REAL* s_ny, REAL* s_nz, int64_t* counts);                                          \
                                                                                                                         \
    /* MemAllocate_multi_GPU(1) (src/Init_multiphase_GPU.cu:73-93).  f_convec/g_convec/phi_convec may be NULL     */ \
    /* unless outlet_BC == 1; W_in may be NULL unless inlet_BC == 1.                                                */ \
    int mflbm_##P##_upload_state(mflbm_##P##_solver* s, const REAL* pdf, const REAL* phi, const REAL* cn_x,              \
                                 const REAL* cn_y, const REAL* cn_z, const REAL* c_norm, const REAL* curv,               \
                                 const REAL* W_in, const REAL* f_convec, const REAL* g_convec, const REAL* phi_convec);  \
    /* device twin of initialization_new_multi + color_gradient for options 1..5 (src/Init_multiphase.cpp:299-496, */ \
    /* src/Phase_gradient.cpp:15).  W_in (inlet_BC == 1) still comes from the host, it needs libm.                  */ \
    int mflbm_##P##_init_state(mflbm_##P##_solver* s, int initial_fluid_distribution_option, REAL interface_z0,          \
                               const REAL* W_in);                                                                        \
    /* the same from a caller-supplied phase field (4-ghost layout): equilibrium PDFs at rest from phi                */ \
    /* (initialization_new_multi_pdf, src/Init_multiphase.cpp:393-496) + colour gradient.  For option 6, whose phi is  */ \
    /* drawn with rand() on the host (:345-351).                                                                       */ \
    int mflbm_##P##_init_state_from_phi(mflbm_##P##_solver* s, const REAL* phi, const REAL* W_in);                       \
    /* the implicit post-condition of main_iteration_kernel_GPU on timer steps (:2059-2076); NULL skips an array.  */ \
    int mflbm_##P##_download_state(mflbm_##P##_solver* s, REAL* pdf, REAL* phi, REAL* cn_x, REAL* cn_y, REAL* cn_z,      \
                                   REAL* c_norm, REAL* curv, REAL* f_convec, REAL* g_convec, REAL* phi_convec);          \
                                                                                                                         \
    /* main_iteration_kernel_GPU (src/main_iteration_GPU.cu:1890-2055) for time step `ntime` (parity = ntime % 2), */ \
    /* asynchronous, no host transfer.  run = nsteps consecutive steps ntime_first, ntime_first+1, ...              */ \
    int mflbm_##P##_step(mflbm_##P##_solver* s, int ntime);                                                              \
    int mflbm_##P##_run(mflbm_##P##_solver* s, int ntime_first, int nsteps);                                             \
    /* the five colour-gradient kernels (:2027-2055) on the current phi, results materialised in cn_*, c_norm, curv */ \
    int mflbm_##P##_color_gradient(mflbm_##P##_solver* s);                                                               \
    /* monitor() reductions on the device; valid after an even ntime (PDFs in natural slots).  Synchronises.       */ \
    int mflbm_##P##_monitor(mflbm_##P##_solver* s, mflbm_monitor_out* out);                                              \
    /* monitor_multiphase_steady_phasefield (src/Monitor.cpp:279-312): max |phi - phi_old| over fluid nodes, then   */ \
    /* phi_old <- phi.  phi_old starts as the phi of the first call's predecessor: zero (the reference's calloc) or */ \
    /* the initial phi when seed_from_current != 0 was passed once before (Init_multiphase.cpp:381-391).            */ \
    int mflbm_##P##_phi_change(mflbm_##P##_solver* s, int seed_from_current, double* d_phi_max);                         \
    /* compute_macro_vars (src/Misc.cpp:222-274) on the device: rho, u, v, w in the 1-ghost layout (zero at solid   */ \
    /* nodes and ghosts), valid after an even ntime.  NULL skips an array.                                          */ \
    int mflbm_##P##_download_macro(mflbm_##P##_solver* s, REAL* rho, REAL* u, REAL* v, REAL* w);                         \
    int mflbm_##P##_sync(mflbm_##P##_solver* s);                                                                         \
                                                                                                                         \
    /* ---- x-slab halo exchange (new).  The caller moves the packed buffers between neighbours (NCCL send/recv,   */ \
    /* cudaMemcpyPeer or peer pointers); all pointers returned are DEVICE pointers owned by the solver.             */ \
    /* kind: 0 = PDF halo after an even step, 1 = PDF halo after an odd step, 2 = phi halo (4 columns).            */ \
    /* side: 0 = left (x-), 1 = right (x+).  count = number of reals in the buffer.                                 */ \
    int mflbm_##P##_halo_buffers(mflbm_##P##_solver* s, int kind, int side, REAL** send, REAL** recv, int64_t* count);   \
    int mflbm_##P##_halo_pack(mflbm_##P##_solver* s, int kind);                                                          \
    int mflbm_##P##_halo_unpack(mflbm_##P##_solver* s, int kind);                                                        \
    /* ---- the same messages through peer memory, no NCCL on the data path.  halo_p2p_local returns my receive      */ \
    /* buffer and my arrival flag for (kind, side); the neighbour on that side passes them (as peer pointers: same   */ \
    /* process, or mflbm_ipc_export/import across processes) to ITS halo_p2p_connect for the opposite side.          */ \
    /* halo_push packs straight into the neighbours' buffers and publishes a sequence number in their flags;         */ \
    /* halo_unpack_wait spins on my flags until the pushes have landed, then unpacks.  Both are asynchronous on the  */ \
    /* solver's stream; every rank must call them in the same order.                                                 */ \
    int mflbm_##P##_halo_p2p_local(mflbm_##P##_solver* s, int kind, int side, REAL** recv, uint32_t** flag);             \
    /* the single allocation all of these pointers live in (export it once, address the rest by offset) */               \
    int mflbm_##P##_halo_p2p_region(mflbm_##P##_solver* s, void** base, int64_t* bytes);                                 \
    int mflbm_##P##_halo_p2p_connect(mflbm_##P##_solver* s, int kind, int side, REAL* peer_recv, uint32_t* peer_flag);   \
    int mflbm_##P##_halo_push(mflbm_##P##_solver* s, int kind);                                                          \
    int mflbm_##P##_halo_unpack_wait(mflbm_##P##_solver* s, int kind);                                                   \
    /* split step for overlap: phase 0 = collide (+pack), phase 1 = boundary kernels (after the PDF halo landed),  */ \
    /* phase 2 = gradient chain (after the phi halo landed).  mflbm_step == phases 0,1,2 with no exchange.          */ \
    int mflbm_##P##_step_phase(mflbm_##P##_solver* s, int ntime, int phase);                                             \
                                                                                                                         \
    /* bookkeeping for bench/roofline: number of fluid nodes (walls == 0 in [1..n]^3), kernels launched so far     */ \
    int64_t mflbm_##P##_num_fluid_nodes(mflbm_##P##_solver* s);                                                          \
    int64_t mflbm_##P##_kernel_launches(mflbm_##P##_solver* s);                                                          \
    void* mflbm_##P##_stream(mflbm_##P##_solver* s);                                                                     \
    /* raw device pointer of a named internal array ("pdf","phi","cn_x",...) for zero-copy interop; NULL if unknown */ \
    void* mflbm_##P##_device_ptr(mflbm_##P##_solver* s, const char* name);

MFLBM_DECLARE_API(f32, float)
MFLBM_DECLARE_API(f64, double)

/* CUDA IPC for the pointers of halo_p2p_local: export fills a 64-byte handle, import opens it in another process */
int mflbm_ipc_export(void* device_ptr, void* handle64);
int mflbm_ipc_import(const void* handle64, void** device_ptr);
int mflbm_ipc_release(void* device_ptr);
/* x-slabs of one lattice on several devices driven from ONE process (mflbm_run --gpus N, host/domain.hpp): enables peer
 * access between two devices in both directions, after which the pointers of halo_p2p_local can be passed to the
 * neighbour's halo_p2p_connect as they are.  New (the reference is single-GPU, README.md:119). */
int mflbm_peer_enable(int device_a, int device_b);

const char* mflbm_last_error(void);
int mflbm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MFLBM_H */
