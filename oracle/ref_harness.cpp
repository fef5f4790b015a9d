// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Driver around the UNMODIFIED reference sources (compiled from where they lie under /root/reference by
// oracle/build_ref.sh; nothing from that tree is copied into this repository).  It replaces only
// src/main.cpp: it performs the same start-up sequence (src/main.cpp:71-99), then calls the reference's
// own per-step entry point main_iteration_kernel_GPU() (src/main_iteration_GPU.cu:1890) and writes the
// reference's global arrays to raw binary files so that tests can compare our CUDA path and the C oracle
// against the real thing.
//
// Two builds:
//   ref_cpu_f32 / ref_cpu_f64   g++ only (-DREF_NO_GPU): geometry, initial state, CPU color_gradient.
//   ref_gpu_f32 / ref_gpu_f64   nvcc, links the reference's two .cu files: adds time stepping.
//
// Usage (cwd must contain input/simulation_control.txt, input/job_status.txt, ...):
//   ref_xxx <outdir> [--dump n1,n2,...] [--monitor n1,n2,...] [--time WARMUP STEPS] [--e2e K] [--block X,Y,Z]
//
// --e2e K (after everything else): the end-to-end cost of K steps for a caller that owns HOST arrays, measured with the
// reference's own transfer code: MemAllocate_multi_GPU(1) (cudaMalloc + H2D of the whole state,
// src/Init_multiphase_GPU.cu:72-97, the state half of initialization_GPU()), K calls of main_iteration_kernel_GPU() with
// the timers set so that its D2H block (src/main_iteration_GPU.cu:2058-2076) fires on the last step only.  Geometry
// stays resident.  Written to e2e.txt.
// --block X,Y,Z overrides block_Threads_X/Y/Z of the control file (tuning runs of the baseline).
//
// The include block below mirrors src/main.cpp:11-27 because the reference keeps all state in globals
// that are *defined* by these headers.
#include "externLib.h"
#include "solver_precision.h"
#include "preprocessor.h"
#include "utils.h"
#include "Module_extern.h"
#include "Module.h"
#include "Fluid_singlephase_extern.h"
#include "Fluid_singlephase.h"
#include "Fluid_multiphase_extern.h"
#include "Fluid_multiphase.h"
#include "Init_multiphase.h"
#include "Misc.h"
#include "Phase_gradient.h"
#include "Monitor.h"
#include "IO_multiphase.h"
#ifndef REF_NO_GPU
#include "Init_multiphase_GPU.h"
#include "main_iteration_GPU.h"
#endif
#include "Idx_cpu.h"

#include <vector>
#include <set>

static void put(const std::string& dir, const char* name, const void* p, long long bytes) {
    if (!p) return;
    std::string f = dir + "/" + name;
    FILE* fp = fopen(f.c_str(), "wb");
    if (!fp) { fprintf(stderr, "cannot open %s\n", f.c_str()); exit(2); }
    if (bytes > 0 && fwrite(p, 1, (size_t)bytes, fp) != (size_t)bytes) { fprintf(stderr, "short write %s\n", f.c_str()); exit(2); }
    fclose(fp);
}

static void dump_geometry(const std::string& dir) {
    fs::create_directories(dir);
    put(dir, "walls.i32", walls, mem_size_s2_int);
    put(dir, "walls_type.i32", walls_type, mem_size_s4_int);
    put(dir, "walls_global.i32", walls_global, mem_size_s0_int);
    put(dir, "s_nx.real", s_nx, mem_size_s4_TP);
    put(dir, "s_ny.real", s_ny, mem_size_s4_TP);
    put(dir, "s_nz.real", s_nz, mem_size_s4_TP);
    put(dir, "W_in.real", W_in, NXG1 * NYG1 * (long long)sizeof(T_P));
    put(dir, "pore_profile_z.i32", pore_profile_z, nzGlobal * (long long)sizeof(int));
}

static void dump_state(const std::string& dir) {
    fs::create_directories(dir);
    put(dir, "pdf.real", pdf, mem_size_f1_TP);
    put(dir, "phi.real", phi, mem_size_s4_TP);
    put(dir, "cn_x.real", cn_x, mem_size_s2_TP);
    put(dir, "cn_y.real", cn_y, mem_size_s2_TP);
    put(dir, "cn_z.real", cn_z, mem_size_s2_TP);
    put(dir, "c_norm.real", c_norm, mem_size_s2_TP);
    put(dir, "curv.real", curv, mem_size_s1_TP);
    if (outlet_BC == 1) {
        put(dir, "f_convec_bc.real", f_convec_bc, NXG1 * NYG1 * 19 * (long long)sizeof(T_P));
        put(dir, "g_convec_bc.real", g_convec_bc, NXG1 * NYG1 * 19 * (long long)sizeof(T_P));
        put(dir, "phi_convec_bc.real", phi_convec_bc, NXG1 * NYG1 * (long long)sizeof(T_P));
    }
}

static void dump_meta(const std::string& dir) {
    std::string f = dir + "/meta.txt";
    FILE* fp = fopen(f.c_str(), "w");
    fprintf(fp, "real_bytes %d\n", (int)sizeof(T_P));
    fprintf(fp, "mrt %d\n", (int)mrt);
    fprintf(fp, "nx %lld\nny %lld\nnz %lld\n", nxGlobal, nyGlobal, nzGlobal);
    fprintf(fp, "nx_sample %lld\nny_sample %lld\nnz_sample %lld\n", nx_sample, ny_sample, nz_sample);
    fprintf(fp, "iper %d\njper %d\nkper %d\n", iper, jper, kper);
    fprintf(fp, "wall_x_min %d\nwall_x_max %d\nwall_y_min %d\nwall_y_max %d\nwall_z_min %d\nwall_z_max %d\n",
            domain_wall_status_x_min, domain_wall_status_x_max, domain_wall_status_y_min, domain_wall_status_y_max,
            domain_wall_status_z_min, domain_wall_status_z_max);
    fprintf(fp, "inlet_BC %d\noutlet_BC %d\n", inlet_BC, outlet_BC);
    fprintf(fp, "porous_plate_cmd %d\nZ_porous_plate %d\n", porous_plate_cmd, Z_porous_plate);
    fprintf(fp, "n_exclude_inlet %d\nn_exclude_outlet %d\n", n_exclude_inlet, n_exclude_outlet);
    fprintf(fp, "initial_fluid_distribution_option %d\n", initial_fluid_distribution_option);
    fprintf(fp, "ntime0 %d\n", ntime0);
    fprintf(fp, "num_solid_boundary_global %lld\nnum_fluid_boundary_global %lld\nnum_solid_boundary %lld\nnum_fluid_boundary %lld\n",
            num_solid_boundary_global, num_fluid_boundary_global, num_solid_boundary, num_fluid_boundary);
    fprintf(fp, "pore_sum %lld\npore_sum_effective %lld\n", pore_sum, pore_sum_effective);
#define PUTR(name) fprintf(fp, #name " %.17g\n", (double)(name))
    PUTR(la_nu1); PUTR(la_nu2); PUTR(la_nui1); PUTR(la_nui2); PUTR(lbm_gamma); PUTR(lbm_beta); PUTR(theta); PUTR(cos_theta);
    PUTR(force_z); PUTR(force_z0); PUTR(rho_in); PUTR(rho_out); PUTR(phi_inlet); PUTR(sa_inject); PUTR(uin_avg); PUTR(uin_avg_0);
    PUTR(relaxation); PUTR(interface_z0); PUTR(ca_0); PUTR(flowrate); PUTR(A_xy); PUTR(A_xy_effective); PUTR(la_x); PUTR(la_y); PUTR(la_z);
    PUTR(porosity_full); PUTR(porosity_effective); PUTR(saturation_full_domain); PUTR(vol1_sum); PUTR(vol2_sum); PUTR(eps);
    PUTR(RK_weight2); PUTR(mrt_e2_coef2);
#undef PUTR
    fclose(fp);
}

static std::set<int> parse_list(const char* s) {
    std::set<int> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) out.insert(std::stoi(tok));
    return out;
}

#ifndef REF_NO_GPU
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } } while (0)
// Same transfers as the reference's conditional D2H block (src/main_iteration_GPU.cu:2059-2076), plus the
// convective-outlet buffers, which the reference never copies back (they only live on the device).
static void device_to_host() {
    CK(cudaMemcpy(phi, phi_d, mem_size_s4_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(curv, curv_d, mem_size_s1_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c_norm, c_norm_d, mem_size_s2_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cn_x, cn_x_d, mem_size_s2_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cn_y, cn_y_d, mem_size_s2_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cn_z, cn_z_d, mem_size_s2_TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pdf, pdf_d, mem_size_f1_TP, cudaMemcpyDeviceToHost));
    if (outlet_BC == 1) {
        CK(cudaMemcpy(f_convec_bc, f_convec_bc_d, NXG1 * NYG1 * 19 * sizeof(T_P), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_convec_bc, g_convec_bc_d, NXG1 * NYG1 * 19 * sizeof(T_P), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(phi_convec_bc, phi_convec_bc_d, NXG1 * NYG1 * sizeof(T_P), cudaMemcpyDeviceToHost));
    }
}
#endif

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <outdir> [--dump a,b,..] [--monitor a,b,..] [--time W K] [--e2e K] [--block X,Y,Z]\n", argv[0]); return 1; }
    std::string outdir = argv[1];
    std::set<int> dumps, monitors;
    int time_warm = -1, time_steps = 0, e2e_steps = 0;
    std::vector<int> block_override;
    for (int a = 2; a < argc; a++) {
        std::string s = argv[a];
        if (s == "--dump" && a + 1 < argc) dumps = parse_list(argv[++a]);
        else if (s == "--monitor" && a + 1 < argc) monitors = parse_list(argv[++a]);
        else if (s == "--time" && a + 2 < argc) { time_warm = atoi(argv[++a]); time_steps = atoi(argv[++a]); }
        else if (s == "--e2e" && a + 1 < argc) { e2e_steps = atoi(argv[++a]); }
        else if (s == "--block" && a + 1 < argc) { std::stringstream ss(argv[++a]); std::string tok; while (std::getline(ss, tok, ',')) block_override.push_back(std::stoi(tok)); }
        else { fprintf(stderr, "bad arg %s\n", s.c_str()); return 1; }
    }
    fs::create_directories(outdir);

    // ---- start-up, as src/main.cpp:48-99 ----
    simulation_end_indicator = 0;
    relaxation = prc(1.);
    auto t0 = chrono::steady_clock::now();
    initialization_basic_multi();
    auto t1 = chrono::steady_clock::now();
    if (job_status == "new_simulation") initialization_new_multi();
    else if (job_status == "continue_simulation") initialization_old_multi();
    else ERROR("job_status");
    if (change_inlet_fluid_phase_cmd != 0) change_inlet_fluid_phase();
    ntime = ntime0;
    color_gradient();
    cal_saturation();
    auto t2 = chrono::steady_clock::now();

    dump_geometry(outdir + "/geometry");
    dump_state(outdir + "/step0");
    dump_meta(outdir);
    {
        std::string f = outdir + "/host_timing.txt";
        FILE* fp = fopen(f.c_str(), "w");
        fprintf(fp, "initialization_basic_multi_s %.6f\ninit_fields_and_color_gradient_s %.6f\n",
                chrono::duration<double>(t1 - t0).count(), chrono::duration<double>(t2 - t1).count());
        fclose(fp);
    }

#ifndef REF_NO_GPU
    int last = 0;
    for (int d : dumps) last = std::max(last, d);
    for (int d : monitors) last = std::max(last, d);
    if (time_warm >= 0) last = std::max(last, time_warm + time_steps);
    if (last > 0 || e2e_steps > 0) {
        if (block_override.size() == 3) { block_Threads_X = block_override[0]; block_Threads_Y = block_override[1]; block_Threads_Z = block_override[2]; }
        initialization_GPU();
        copyConstantData();
        FILE* fmon = nullptr;
        if (!monitors.empty()) { std::string f = outdir + "/monitor.txt"; fmon = fopen(f.c_str(), "w"); }
        // step numbering as src/main.cpp:145: ntime runs ntime0, ntime0+1, ...; "after n steps" == ntime0+n-1 done.
        int done = 0;
        chrono::steady_clock::time_point tstart;
        double timed_s = -1.;
        for (ntime = ntime0; done < last; ntime++) {
            if (time_warm >= 0 && done == time_warm) { CK(cudaDeviceSynchronize()); tstart = chrono::steady_clock::now(); }
            main_iteration_kernel_GPU();
            done++;
            if (time_warm >= 0 && done == time_warm + time_steps) {
                CK(cudaDeviceSynchronize());
                timed_s = chrono::duration<double>(chrono::steady_clock::now() - tstart).count();
            }
            bool want_dump = dumps.count(done) > 0, want_mon = monitors.count(done) > 0;
            if (want_dump || want_mon) device_to_host();
            if (want_dump) dump_state(outdir + "/step" + std::to_string(done));
            if (want_mon) {
                monitor();   // src/Monitor.cpp:17 (needs natural-slot PDFs: only meaningful after an even ntime)
                fprintf(fmon, "%d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", done, ntime, (double)saturation,
                        (double)saturation_full_domain, (double)vol1_sum, (double)vol2_sum, (double)mass1_sum, (double)mass2_sum,
                        (double)ca, (double)umax_global);
                fflush(fmon);
            }
        }
        if (fmon) fclose(fmon);
        if (timed_s >= 0.) {
            std::string f = outdir + "/timing.txt";
            FILE* fp = fopen(f.c_str(), "w");
            double mlups = (double)nxGlobal * nyGlobal * nzGlobal * time_steps / (1e6 * timed_s);
            fprintf(fp, "steps %d\nseconds %.9f\nms_per_step %.6f\nmlups %.3f\n", time_steps, timed_s, 1e3 * timed_s / time_steps, mlups);
            fclose(fp);
            printf("REF_TIMING steps=%d seconds=%.6f mlups=%.3f\n", time_steps, timed_s, mlups);
        }
        if (e2e_steps > 0) {
            // the host arrays take the device state (as after a timer step), the device copy of the state is dropped
            device_to_host();
            MemAllocate_multi_GPU(0);
            CK(cudaDeviceSynchronize());
            const int first = ntime, lastn = ntime + e2e_steps - 1;
            if (e2e_steps >= first) { fprintf(stderr, "--e2e: K must be smaller than the current step index\n"); return 1; }
            ntime_monitor = ntime_animation = ntime_visual = 2000000000;
            ntime_clock_sum = lastn;   // ntime % ntime_clock_sum == 0 only at ntime == lastn inside the window
            auto e0 = chrono::steady_clock::now();
            MemAllocate_multi_GPU(1);
            CK(cudaDeviceSynchronize());
            auto e1 = chrono::steady_clock::now();
            for (ntime = first; ntime <= lastn; ntime++) main_iteration_kernel_GPU();
            CK(cudaDeviceSynchronize());
            auto e2 = chrono::steady_clock::now();
            const double s_up = chrono::duration<double>(e1 - e0).count(), s_all = chrono::duration<double>(e2 - e0).count();
            auto e3 = chrono::steady_clock::now();
            monitor();   // the reference's monitor is a host loop over the downloaded arrays (src/Monitor.cpp:17); timed apart
            const double s_mon = chrono::duration<double>(chrono::steady_clock::now() - e3).count();
            const long long plane = NXG1 * NYG1 * (long long)sizeof(T_P);
            const long long conv = outlet_BC == 1 ? 39 * plane : 0;
            const long long h2d = 2 * mem_size_s1_TP + plane + mem_size_f1_TP + 4 * mem_size_s2_TP + conv + mem_size_s4_TP;
            const long long d2h = mem_size_s4_TP + mem_size_s1_TP + 4 * mem_size_s2_TP + mem_size_f1_TP;
            std::string f = outdir + "/e2e.txt";
            FILE* fp = fopen(f.c_str(), "w");
            const double mlups = (double)nxGlobal * nyGlobal * nzGlobal * e2e_steps / (1e6 * s_all);
            fprintf(fp, "steps %d\nseconds %.9f\nupload_seconds %.9f\nhost_monitor_seconds %.9f\nmlups %.3f\nh2d_bytes %lld\nd2h_bytes %lld\n",
                    e2e_steps, s_all, s_up, s_mon, mlups, h2d, d2h);
            fclose(fp);
            printf("REF_E2E steps=%d seconds=%.6f upload=%.6f mlups=%.3f\n", e2e_steps, s_all, s_up, mlups);
        }
    }
#else
    (void)time_warm; (void)time_steps;
    if (!dumps.empty() || !monitors.empty()) { fprintf(stderr, "CPU-only build cannot step\n"); return 1; }
#endif
    return 0;
}
