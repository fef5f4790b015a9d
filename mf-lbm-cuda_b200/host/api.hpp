// api.hpp — precision dispatch onto the C ABI (include/mflbm.h): Api<float> binds mflbm_f32_*, Api<double> mflbm_f64_*.
// A non-zero status becomes a Fatal carrying mflbm_last_error(); the reference prints and exits at the same places
// (includes/utils_GPU.cuh:8-14 in /root/reference).
#pragma once
#include "../../include/mflbm.h"
#include "control.hpp"

namespace mfhost {

inline void api_check(int status, const char* what) {
    if (status != 0) throw Fatal(std::string(what) + ": " + mflbm_last_error());
}

template <typename T> struct Api;

#define MFHOST_API(P, REAL)                                                                                                         \
    template <> struct Api<REAL> {                                                                                                  \
        using Params = mflbm_##P##_params;                                                                                          \
        using Handle = mflbm_##P##_solver;                                                                                          \
        static constexpr const char* name = #P;                                                                                     \
        static void create(const Params& p, int device, Handle** h) { api_check(mflbm_##P##_create(&p, nullptr, device, nullptr, h), "create"); } \
        static void create_slab(const Params& p, const mflbm_slab& s, int device, Handle** h) {                                     \
            api_check(mflbm_##P##_create(&p, &s, device, nullptr, h), "create (slab)");                                             \
        }                                                                                                                           \
        /* my message of `kind` through `side` lands in the neighbour's buffer of side `nside` */                                  \
        static void connect(Handle* me, int kind, int side, Handle* neighbour, int nside) {                                         \
            REAL* recv = nullptr; uint32_t* flag = nullptr;                                                                         \
            api_check(mflbm_##P##_halo_p2p_local(neighbour, kind, nside, &recv, &flag), "halo_p2p_local");                          \
            api_check(mflbm_##P##_halo_p2p_connect(me, kind, side, recv, flag), "halo_p2p_connect");                                \
        }                                                                                                                           \
        static void halo_push(Handle* h, int kind) { api_check(mflbm_##P##_halo_push(h, kind), "halo_push"); }                      \
        static void halo_unpack_wait(Handle* h, int kind) { api_check(mflbm_##P##_halo_unpack_wait(h, kind), "halo_unpack_wait"); } \
        static void destroy(Handle* h) { mflbm_##P##_destroy(h); }                                                                  \
        static void set_params(Handle* h, const Params& p) { api_check(mflbm_##P##_set_params(h, &p), "set_params"); }              \
        static void preprocess_geometry(Handle* h, const int8_t* w) { api_check(mflbm_##P##_preprocess_geometry(h, w), "preprocess_geometry"); } \
        static void geometry_counts(Handle* h, int64_t* c) { api_check(mflbm_##P##_download_geometry(h, nullptr, nullptr, nullptr, nullptr, nullptr, c), "download_geometry"); } \
        static void init_state(Handle* h, int opt, REAL z0, const REAL* W) { api_check(mflbm_##P##_init_state(h, opt, z0, W), "init_state"); } \
        static void init_state_from_phi(Handle* h, const REAL* phi, const REAL* W) { api_check(mflbm_##P##_init_state_from_phi(h, phi, W), "init_state_from_phi"); } \
        static void upload_pdf(Handle* h, const REAL* pdf) {                                                                        \
            api_check(mflbm_##P##_upload_state(h, pdf, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "upload_state"); \
        }                                                                                                                           \
        static void upload_restart(Handle* h, const REAL* pdf, const REAL* phi, const REAL* W, const REAL* fc, const REAL* gc, const REAL* pc) { \
            api_check(mflbm_##P##_upload_state(h, pdf, phi, nullptr, nullptr, nullptr, nullptr, nullptr, W, fc, gc, pc), "upload_state"); \
        }                                                                                                                           \
        static void color_gradient(Handle* h) { api_check(mflbm_##P##_color_gradient(h), "color_gradient"); }                       \
        static void run(Handle* h, int first, int n) { api_check(mflbm_##P##_run(h, first, n), "run"); }                            \
        static void monitor(Handle* h, mflbm_monitor_out* m) { api_check(mflbm_##P##_monitor(h, m), "monitor"); }                   \
        static void phi_change(Handle* h, int seed, double* d) { api_check(mflbm_##P##_phi_change(h, seed, d), "phi_change"); }     \
        static void download_macro(Handle* h, REAL* r, REAL* u, REAL* v, REAL* w) { api_check(mflbm_##P##_download_macro(h, r, u, v, w), "download_macro"); } \
        static void download(Handle* h, REAL* pdf, REAL* phi, REAL* cx, REAL* cy, REAL* cz, REAL* cn, REAL* fc, REAL* gc, REAL* pc) { \
            api_check(mflbm_##P##_download_state(h, pdf, phi, cx, cy, cz, cn, nullptr, fc, gc, pc), "download_state");              \
        }                                                                                                                           \
        static void sync(Handle* h) { api_check(mflbm_##P##_sync(h), "sync"); }                                                     \
        static long long fluid_nodes(Handle* h) { return mflbm_##P##_num_fluid_nodes(h); }                                          \
    };

MFHOST_API(f32, float)
MFHOST_API(f64, double)
#undef MFHOST_API

}  // namespace mfhost
