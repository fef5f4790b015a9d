// integration/mflbm_shim.cpp — the reference-side binding of libmflbm.so.
//
// A maintainer of MF-LBM-CUDA drops this file into src/ IN PLACE OF src/main_iteration_GPU.cu and src/Init_multiphase_GPU.cu:
// it defines the five entry points src/main.cpp calls (includes/Init_multiphase_GPU.h:5-10, includes/main_iteration_GPU.h:4-7)
// on top of the C ABI of include/mflbm.h, built with g++ like the other host sources and linked with
// -L<repo>/mf-lbm-cuda_b200/lib -lmflbm.  Nothing else of the reference changes.
// oracle/build_ref.sh builds the reference's unmodified src/main.cpp + CPU sources around it (oracle/_ref/MF_LBM_CUDA_shim_*)
// and tests/test_host_driver.py compares that program's output files with the stock program's.
#include "externLib.h"
#include "solver_precision.h"
#include "preprocessor.h"
#include "Module_extern.h"
#include "Fluid_singlephase_extern.h"
#include "Fluid_multiphase_extern.h"
#include "utils.h"
static const int mrt_model = mrt;   // includes/preprocessor.h:4 defines `mrt` as a macro; mflbm.h has a struct field of that name
#undef mrt
#include "mflbm.h"

#if (PRECISION == SINGLE_PRECISION)
#define MF(name) mflbm_f32_##name
typedef mflbm_f32_params mf_params; typedef mflbm_f32_solver mf_solver;
#else
#define MF(name) mflbm_f64_##name
typedef mflbm_f64_params mf_params; typedef mflbm_f64_solver mf_solver;
#endif
#define MF_CHECK(call) do { if ((call) != 0) { ERROR(mflbm_last_error()); } } while (0)

static mf_solver* g_solver = nullptr;

static mf_params collect_params() {            // the globals copyConstantData uploads (src/main_iteration_GPU.cu:14-47)
    mf_params p{};
    p.nx = nxGlobal; p.ny = nyGlobal; p.nz = nzGlobal;
    p.iper = iper; p.jper = jper; p.kper = kper;
    p.wall_z_min = domain_wall_status_z_min; p.wall_z_max = domain_wall_status_z_max;
    p.inlet_BC = inlet_BC; p.outlet_BC = outlet_BC;
    p.porous_plate_cmd = porous_plate_cmd; p.Z_porous_plate = Z_porous_plate;
    p.n_exclude_inlet = n_exclude_inlet; p.n_exclude_outlet = n_exclude_outlet;
    p.mrt = mrt_model;
    p.lbm_gamma = lbm_gamma; p.lbm_beta = lbm_beta; p.la_nu1 = la_nu1; p.la_nui1 = la_nui1; p.la_nui2 = la_nui2;
    p.cos_theta = cos_theta; p.force_z = force_z; p.rho_in = rho_in; p.rho_out = rho_out; p.phi_inlet = phi_inlet;
    p.sa_inject = sa_inject; p.uin_avg = uin_avg; p.relaxation = relaxation; p.A_xy = A_xy;
    return p;
}

void initialization_GPU() {                     // src/main.cpp:97
    mf_params p = collect_params();
    MF_CHECK(MF(create)(&p, nullptr, 0, nullptr, &g_solver));
    MF_CHECK(MF(upload_geometry)(g_solver, walls, walls_type, s_nx, s_ny, s_nz));
    MF_CHECK(MF(upload_state)(g_solver, pdf, phi, cn_x, cn_y, cn_z, c_norm, curv, W_in,
                              f_convec_bc, g_convec_bc, phi_convec_bc));
}
void copyConstantData() { mf_params p = collect_params(); MF_CHECK(MF(set_params)(g_solver, &p)); }   // :99

void main_iteration_kernel_GPU() {              // src/main.cpp:147
    MF_CHECK(MF(step)(g_solver, ntime));
    if (ntime % ntime_monitor == 0 || ntime % ntime_animation == 0 || ntime % ntime_visual == 0 || ntime % ntime_clock_sum == 0)
        MF_CHECK(MF(download_state)(g_solver, pdf, phi, cn_x, cn_y, cn_z, c_norm, curv,                 // :2059-2076
                                    f_convec_bc, g_convec_bc, phi_convec_bc));
}
void MemAllocate_geometry_GPU(int flag) { if (flag != 1 && g_solver) { MF_CHECK(MF(destroy)(g_solver)); g_solver = nullptr; } }
void MemAllocate_multi_GPU(int) {}
