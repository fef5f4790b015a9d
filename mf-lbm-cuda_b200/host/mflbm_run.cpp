// mflbm_run — host driver of the B200-native MF-LBM time-step library.
//
// Drop-in for the reference program (/root/reference/src/main.cpp:31-367): run it in a directory that holds
// input/simulation_control.txt, input/job_status.txt and (for external geometries) input/Geometry_File_Path.txt and it
// produces results/out1.output, out2.checkpoint and out3.field_data in the reference's formats.  Everything numerical
// happens behind the C ABI of include/mflbm.h; there is no CPU solver path in this program.
//
// Differences from the reference, all on the host side:
//   * precision is a run-time switch (--prec f32|f64) instead of a rebuild (includes/solver_precision.h:8);
//   * the time loop hands the GPU whole batches of steps up to the next output event (monitor / VTK / timer /
//     checkpoint step) instead of synchronising eight times per step (src/main_iteration_GPU.cu:1903-2055), and the state
//     only crosses PCIe when a checkpoint or a VTK file is written: the monitor reductions run on the device;
//   * --gpus N (or --devices a,b,..) cuts the lattice into N x-slabs, one per GPU of the box (host/domain.hpp); every file
//     is written in the reference's single-domain format whatever the decomposition;
//   * geometry_preprocess_cmd 1 (unimplemented in the reference too, src/Geometry_preprocessing.cpp:426) is rejected.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <memory>

#include "api.hpp"
#include "domain.hpp"
#include "case.hpp"
#include "control.hpp"
#include "output.hpp"

using namespace mfhost;
using Clock = std::chrono::steady_clock;
static double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

struct Options {
    std::string dir = ".";
    std::string prec = "f64";
    int device = 0;
    std::vector<int> devices;   // --gpus N / --devices a,b,...: x-slabs over several GPUs of the box (host/domain.hpp)
    int mrt = 2;             // includes/preprocessor.h:4
    bool quirk = true;       // swapped-stride geometry read, src/Misc.cpp:171
    bool check_input = false;
};

template <typename T>
class Run {
  public:
    explicit Run(const Options& o) : opt(o) {}

    int main() {
        std::cout << "Solver precision: " << (sizeof(T) == 4 ? "Single" : "Double") << " precision" << std::endl;
        std::cout << "***************************** Initialization **********************************" << std::endl;
        cs.dir = opt.dir;
        cs.geometry_index_quirk = opt.quirk;
        cs.ctl = Control<T>::read(opt.dir);
        if (!cs.ctl.validate()) {
            for (const auto& p : cs.ctl.problems) std::cout << "Error: " << p << "! Exiting program!" << std::endl;
            throw Fatal("Exit Program!");
        }
        const auto& c = cs.ctl;
        if (c.geometry_preprocess_cmd == 1) throw Fatal("geometry_preprocess_cmd 1: loading precomputed boundary info is not implemented (nor in the reference)");
        if (c.geometry_preprocess_cmd != 0) throw Fatal("Wrong value of geometry_preprocess_cmd! Stop program!");
        if (c.extreme_large_sim_cmd != 0) throw Fatal("Extreme large simulation domain (extreme_large_sim_cmd = 1) is not supported yet!");
        if (!opt.check_input) make_result_dirs(opt.dir);
        cs.set_walls();
        cs.derive();
        if (opt.check_input) { print_check(); return 0; }

        typename Api<T>::Params P;
        cs.fill_params(P, opt.mrt);
        h.create(P, opt.devices.empty() ? std::vector<int>{opt.device} : opt.devices);
        if (h.slabs() > 1) std::cout << "Lattice cut into " << h.slabs() << " x-slabs, one per GPU" << std::endl;
        std::cout << "------ Start processing boundary nodes info ------" << std::endl;
        auto t0 = Clock::now();
        h.preprocess_geometry(cs.walls.data());
        int64_t counts[4];
        h.geometry_counts(counts);
        std::cout << "Total number of solid boundary nodes = " << counts[0] << std::endl << "Total number of fluid boundary nodes = " << counts[1] << std::endl;
        std::cout << "------ End processing boundary nodes info -------- (" << seconds_since(t0) << " s on the device)" << std::endl;
        write_info(counts);

        if (c.job_status == "new_simulation") {
            if (c.initial_fluid_distribution_option < 1 || c.initial_fluid_distribution_option > 6)
                throw Fatal("Input parameter initial_fluid_distribution_option error! Stop program!!!");
            cs.ntime0 = 1;
            const T* W = cs.W_in.empty() ? nullptr : cs.W_in.data();
            if (c.initial_fluid_distribution_option == 6) { const std::vector<T> phi = random_phi(); h.init_state_from_phi(phi.data(), W); }
            else h.init_state(c.initial_fluid_distribution_option, c.interface_z0, W);
            if (c.steady_state_option == 2) h.phi_change(1, nullptr);   // phi_old = phi, src/Init_multiphase.cpp:381-391
        } else {
            load_checkpoint(P);
        }
        if (c.change_inlet_fluid_phase_cmd != 0) {
            change_inlet_fluid_phase();
            std::cout << "Inlet fluid phase was changed (1 to fluid1; 2 to fluid2): " << c.change_inlet_fluid_phase_cmd << std::endl;
        }
        std::cout << "************************** Initialization ends ********************************" << std::endl;

        mflbm_monitor_out m;
        monitor_raw(m);
        T saturation_old = c.job_status == "new_simulation" ? T(m.saturation_full_domain) : T(-1.);
        (void)saturation_old;
        if (c.job_status == "new_simulation") {
            if (c.benchmark_cmd == 0) { write_vtk(cs.ntime0, 2); write_walls_vtk(); }
        } else if (c.open_z() && c.inlet_BC == 2 && c.rho_in_new) {   // new pressure BC value overrides the loaded one, src/main.cpp:123-135
            cs.pressure_inlet();
            cs.fill_params(P, opt.mrt);
            h.set_params(P);
        }
        std::cout << "Initial saturation: " << T(m.saturation_full_domain) << std::endl;
        loop();
        finish();
        std::cout << std::endl << "Code Finished " << std::endl;
        return 0;
    }

  private:
    Options opt;
    Case<T> cs;
    Domain<T> h;
    int ntime = 0, end_indicator = 0;
    T monitor_previous = 0, monitor_current = 0;
    std::vector<double> prof[7];   // fl1, fl2, pre, mass1, mass2, vol1, vol2 per z slice

    long long N(int g) const { return (cs.nx() + 2 * g) * (cs.ny() + 2 * g) * (cs.nz() + 2 * g); }
    long long NP() const { return (cs.nx() + 2) * (cs.ny() + 2); }
    bool convective() const { return cs.ctl.outlet_BC == 1; }

    void print_check() const {
        auto hex = [](T v) { char b[64]; snprintf(b, sizeof b, "%a", (double)v); return std::string(b); };
        unsigned long long hash = 1469598103934665603ULL;   // FNV-1a over the interior wall array
        for (auto w : cs.walls) { hash ^= (unsigned char)w; hash *= 1099511628211ULL; }
        std::cout << "CHECK prec " << Api<T>::name << std::endl;
        std::cout << "CHECK dims " << cs.nx() << " " << cs.ny() << " " << cs.nz() << std::endl;
        std::cout << "CHECK walls_fnv1a " << hash << std::endl;
        std::cout << "CHECK pore_sum " << cs.pore_sum << " " << cs.pore_sum_effective << std::endl;
        std::cout << "CHECK timers " << cs.ntime_max << " " << cs.ntime_monitor << " " << cs.ntime_monitor_profile << " " << cs.ntime_animation << " " << cs.ntime_visual << std::endl;
        const char* names[] = {"cos_theta", "la_nui1", "la_nui2", "phi_inlet", "force_z", "rho_in", "rho_out", "uin_avg", "A_xy", "flowrate", "porosity_full", "porosity_effective"};
        const T vals[] = {cs.cos_theta, cs.la_nui1, cs.la_nui2, cs.phi_inlet, cs.force_z, cs.rho_in, cs.rho_out, cs.uin_avg, cs.A_xy, cs.flowrate, cs.porosity_full, cs.porosity_effective};
        for (int n = 0; n < 12; n++) std::cout << "CHECK " << names[n] << " " << hex(vals[n]) << std::endl;
        if (!cs.W_in.empty()) {
            unsigned long long hw = 1469598103934665603ULL;
            const unsigned char* b = reinterpret_cast<const unsigned char*>(cs.W_in.data());
            for (size_t n = 0; n < cs.W_in.size() * sizeof(T); n++) { hw ^= b[n]; hw *= 1099511628211ULL; }
            std::cout << "CHECK W_in_fnv1a " << hw << std::endl;
        }
    }

    // initial distribution 6 (src/Init_multiphase.cpp:303-372): fluid 1 with probability sa_target per site of [0..n+1]^3,
    // drawn with rand() seeded by the wall clock like the reference (so not reproducible there either), then the inlet
    // ghost planes
    std::vector<T> random_phi() const {
        const long long NX4 = cs.nx() + 8, NY4 = cs.ny() + 8;
        std::vector<T> phi((size_t)N(4), T(0));
        auto at = [&](long long i, long long j, long long k) -> T& { return phi[(size_t)((i + 3) + NX4 * ((j + 3) + NY4 * (k + 3)))]; };
        srand((unsigned)time(NULL));
        for (long long k = 0; k <= cs.nz() + 1; k++)
            for (long long j = 0; j <= cs.ny() + 1; j++)
                for (long long i = 0; i <= cs.nx() + 1; i++) {
                    const T r = (T)rand() / RAND_MAX;
                    at(i, j, k) = r > cs.ctl.sa_target ? T(-1.) : T(1.);
                }
        if (cs.ctl.open_z())
            for (long long k = -3; k <= 0; k++)
                for (long long j = -3; j <= cs.ny() + 4; j++)
                    for (long long i = -3; i <= cs.nx() + 4; i++) at(i, j, k) = cs.phi_inlet;
        return phi;
    }

    // change_inlet_fluid_phase (src/Misc.cpp:277-384): for k <= interface_z0 and i in [0 .. nx-1] (the reference's loop bound)
    // the populations of one component are added to the other's and zeroed.  Host-side, once, on the downloaded PDFs.
    void change_inlet_fluid_phase() {
        const long long NX1 = cs.nx() + 2, NY1 = cs.ny() + 2, n1 = N(1);
        std::vector<T> pdf((size_t)(38 * n1));
        h.download(pdf.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        const int to = cs.ctl.change_inlet_fluid_phase_cmd == 1 ? 0 : 1, from = 1 - to;
        if (cs.ctl.change_inlet_fluid_phase_cmd == 1 || cs.ctl.change_inlet_fluid_phase_cmd == 2)
            for (long long k = 0; k <= cs.nz() + 1; k++) {
                if (!((T)k <= cs.ctl.interface_z0)) continue;
                for (long long j = 0; j <= cs.ny() + 1; j++)
                    for (long long i = 0; i <= cs.nx() - 1; i++)
                        for (int q = 0; q < 19; q++) {
                            const size_t cell = (size_t)(i + NX1 * (j + NY1 * k));
                            pdf[(size_t)(q + 19 * to) * n1 + cell] += pdf[(size_t)(q + 19 * from) * n1 + cell];
                            pdf[(size_t)(q + 19 * from) * n1 + cell] = T(0.);
                        }
            }
        h.upload_pdf(pdf.data());
    }

    // results/out1.output/info.txt, src/Init_multiphase.cpp:128-165
    void write_info(const int64_t* counts) {
        std::ofstream f((opt.dir + "/results/out1.output/info.txt").c_str(), std::ios_base::out);
        if (!f.good()) throw Fatal("Couldn't open the file results/out1.output/info.txt !");
        f << " Grid information:" << std::endl;
        f << "nxGlobal = " << cs.nx() << " , nyGlobal = " << cs.ny() << " , nzGlobal = " << cs.nz() << std::endl;
        f << "Inlet open cross sectional area = " << cs.A_xy << std::endl;
        f << " Pore information:" << std::endl;
        f << "Total number of pore nodes = " << cs.pore_sum << std::endl;
        f << "Total number of solid boundary nodes = " << counts[0] << std::endl;
        f << "Total number of fluid boundary nodes = " << counts[1] << std::endl;
        f << "Total number of effective pore nodes (excluding inlet/outlet) = " << cs.pore_sum_effective << std::endl;
        f << "Full domain porosity = " << cs.porosity_full << std::endl;
        f << "Effective domain porosity  (excluding inlet/outlet) = " << cs.porosity_effective << std::endl;
        std::cout << "Total number of pore nodes = " << cs.pore_sum << std::endl << "Full domain porosity = " << cs.porosity_full << std::endl;
    }

    // initialization_old_multi (src/Init_multiphase.cpp:501-538) + the CPU colour gradient the reference runs before upload
    void load_checkpoint(typename Api<T>::Params& P) {
        std::cout << "loading checkpoint data ... " << std::flush;
        Checkpoint<T> ck;
        ck.pdf.resize((size_t)(38 * N(1))); ck.phi.resize((size_t)N(4));
        if (convective()) { ck.f_convec.resize((size_t)(19 * NP())); ck.g_convec.resize((size_t)(19 * NP())); ck.phi_convec.resize((size_t)NP()); }
        read_checkpoint(checkpoint_path(opt.dir, false), ck, convective());
        cs.ntime0 = ck.ntime_next - 1;   // the last step index is executed again, like the reference does
        cs.force_z = ck.force_z; cs.rho_in = ck.rho_in;
        cs.fill_params(P, opt.mrt);
        h.set_params(P);
        h.upload_restart(ck.pdf.data(), ck.phi.data(), cs.W_in.empty() ? nullptr : cs.W_in.data(),
                               convective() ? ck.f_convec.data() : nullptr, convective() ? ck.g_convec.data() : nullptr,
                               convective() ? ck.phi_convec.data() : nullptr);
        h.color_gradient();
        std::cout << "Complete" << std::endl;
    }

    // save_checkpoint, src/IO_multiphase.cpp:252-305.  The device state is downloaded here (the reference writes whatever
    // host copy the last timer step left behind).
    void save_checkpoint(bool secondary) {
        Checkpoint<T> ck;
        ck.ntime_next = ntime + 1; ck.force_z = cs.force_z; ck.rho_in = cs.rho_in;
        ck.pdf.resize((size_t)(38 * N(1))); ck.phi.resize((size_t)N(4));
        if (convective()) { ck.f_convec.resize((size_t)(19 * NP())); ck.g_convec.resize((size_t)(19 * NP())); ck.phi_convec.resize((size_t)NP()); }
        h.download(ck.pdf.data(), ck.phi.data(), nullptr, nullptr, nullptr, nullptr, convective() ? ck.f_convec.data() : nullptr,
                         convective() ? ck.g_convec.data() : nullptr, convective() ? ck.phi_convec.data() : nullptr);
        write_checkpoint(checkpoint_path(opt.dir, secondary), ck, convective());
        if (!secondary) { std::cout << "Saving checkpoint data completed!" << std::endl; write_status("continue_simulation"); }
        else std::cout << "Saving secondary checkpoint data completed!" << std::endl;
    }

    // the reference writes ./job_status.txt, not input/job_status.txt
    void write_status(const char* s) {
        std::ofstream f((opt.dir + "/job_status.txt").c_str(), std::ios_base::out);
        if (!f.good()) throw Fatal("Could not open ./job_status.txt");
        f << s << std::endl;
    }

    // VTK_legacy_writer_3D, src/IO_multiphase.cpp:500-716: type 1 full flow field, type 2 phase field (float), type 3 CSF vectors
    void write_vtk(int nt, int type) {
        const long long nx = cs.nx(), ny = cs.ny(), nz = cs.nz();
        std::vector<T> phi((size_t)N(4));
        if (type == 2) {
            h.download(nullptr, phi.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            std::vector<float> f4;
            crop(phi, 4, nx, ny, nz, f4);
            for (size_t n = 0; n < f4.size(); n++) if (cs.walls[n]) f4[n] = 0.f;   // phi zeroed in solids, :558-566
            VtkFile v(vtk_name(opt.dir, "phase_distribution/small_", nt), nx, ny, nz);
            v.scalars("phi", "float", f4);
            return;
        }
        const char* fmt = cs.ctl.output_fieldData_precision_cmd == 0 ? "float" : "double";   // declared type; the data is always T
        std::vector<T> a((size_t)N(type == 1 ? 1 : 2)), b(a.size()), c(a.size()), d(a.size());
        if (type == 1) { h.download_macro(d.data(), a.data(), b.data(), c.data()); h.download(nullptr, phi.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); }
        else h.download(nullptr, phi.data(), a.data(), b.data(), c.data(), d.data(), nullptr, nullptr, nullptr);
        const int g = type == 1 ? 1 : 2;
        std::vector<T> out;
        VtkFile v(vtk_name(opt.dir, type == 1 ? "full_flow_field/full_flow_field_" : "full_flow_field/force_vector_", nt), nx, ny, nz);
        crop(phi, 4, nx, ny, nz, out);
        if (type == 1) for (size_t n = 0; n < out.size(); n++) if (cs.walls[n]) out[n] = T(0);   // compute_macro_vars zeroes the host phi in solids (src/Misc.cpp:222-274)
        v.scalars("phi", fmt, out);
        crop(d, g, nx, ny, nz, out); v.scalars("density", fmt, out);
        crop(a, g, nx, ny, nz, out); v.scalars("velocity_X", fmt, out);
        crop(b, g, nx, ny, nz, out); v.scalars("velocity_Y", fmt, out);
        crop(c, g, nx, ny, nz, out); v.scalars("velocity_Z", fmt, out);
    }

    void write_walls_vtk() {   // VTK_walls_bin, src/IO_multiphase.cpp:800-841
        std::vector<int32_t> w(cs.walls.begin(), cs.walls.end());
        VtkFile v(opt.dir + "/results/out3.field_data/walls_bin.vtk", cs.nx(), cs.ny(), cs.nz());
        v.scalars("walls", "int", w);
    }

    void monitor_raw(mflbm_monitor_out& m) {
        std::memset(&m, 0, sizeof m);
        for (auto& p : prof) p.assign((size_t)cs.nz(), 0.0);
        m.fl1 = prof[0].data(); m.fl2 = prof[1].data(); m.pre = prof[2].data(); m.mass1 = prof[3].data(); m.mass2 = prof[4].data();
        m.vol1 = prof[5].data(); m.vol2 = prof[6].data();
        h.monitor(&m);
    }

    bool fresh_file() const { return cs.ntime0 == 1 && ntime == cs.ntime_monitor; }

    // monitor(), src/Monitor.cpp:17-273: the reductions come from the device, files and tests are the reference's
    void monitor_displacement() {
        const auto& c = cs.ctl;
        mflbm_monitor_out m;
        monitor_raw(m);
        const bool fresh = fresh_file();
        const long long nz = cs.nz();
        RowFile(opt.dir, "saturation.dat", fresh) << ntime << " " << T(m.saturation) << " " << T(m.vol1_sum) << " " << T(m.vol2_sum) << " " << T(m.mass1_sum) << " " << T(m.mass2_sum) << "\n";
        const T fl_avg_whole = T(m.fl1_avg_whole) + T(m.fl2_avg_whole), fl_avg = T(m.fl1_avg) + T(m.fl2_avg);
        RowFile(opt.dir, "Ca_number.dat", fresh) << ntime << " " << T(m.ca) << " " << T(m.umax) << " " << T(m.kinetic_energy[0]) << " " << T(m.kinetic_energy[1]) << "\n";
        RowFile(opt.dir, "flowrate_time.dat", fresh) << ntime << " " << fl_avg_whole << " " << T(m.fl1_avg_whole) << " " << T(m.fl2_avg_whole) << " " << fl_avg << " " << T(m.fl1_avg) << " " << T(m.fl2_avg) << "\n";
        RowFile(opt.dir, "saturation_full_domain.dat", fresh) << ntime << " " << T(m.saturation_full_domain) << " " << T(m.vol1_full) << " " << T(m.vol2_full) << " " << T(m.mass1_full) << " " << T(m.mass2_full) << "\n";
        if (c.open_z()) {
            const size_t ki = (size_t)c.n_exclude_inlet, ko = (size_t)(nz - c.n_exclude_outlet - 1);   // slices 1 + n_exclude_inlet and nz - n_exclude_outlet
            const T p1 = T(prof[2][ki]) / cs.pore_profile_z[ki], p2 = T(prof[2][ko]) / cs.pore_profile_z[ko];
            RowFile(opt.dir, "pre.dat", fresh) << ntime << " " << p1 << " " << p2 << " " << T(p1 - p2) << "\n";
        }
        if (cs.ntime_monitor_profile > 0 && ntime % cs.ntime_monitor_profile == 0) {
            std::ostringstream name;
            name << opt.dir << "/results/out1.output/profile/monitor " << std::setfill('0') << std::setw(8) << ntime;   // the blank is the reference's
            std::ofstream f(name.str().c_str(), std::ios_base::out);
            // the reference's loop runs k = 0 .. nz-1 over slice k (its row 0 reads one element before the arrays)
            f << T(0) << " " << T(0) << " " << T(0) << " " << T(0) << " " << T(0) << std::endl;
            for (long long k = 1; k < nz; k++) {
                const size_t s = (size_t)k - 1;
                f << T(k) << " " << T(prof[0][s]) << " " << T(prof[1][s]) << " " << T(prof[3][s]) / (T(prof[3][s]) + T(prof[4][s])) << " "
                  << T(prof[2][s]) / (T(cs.pore_profile_z[s]) + cs.eps) << std::endl;
            }
        }
        if (std::isnan(m.saturation_full_domain) || std::isnan(m.ca) || m.nan_detected) {
            std::cout << "Simulation failed due to NAN of saturation or capillary number" << std::endl;
            end_indicator = 3;
        } else if (T(m.umax) > T(0.5)) {
            std::cout << "Simulation failed due to maximum velocity larger than 0.5 !" << std::endl;
            end_indicator = 3;
        }
        if (c.steady_state_option == 3) {
            monitor_previous = monitor_current;
            monitor_current = T(m.saturation_full_domain);
            const T err = std::fabs(monitor_current - monitor_previous) / (std::fabs(monitor_current) + cs.eps);
            RowFile(opt.dir, "steady_monitor_saturation_error.dat", fresh) << ntime << " " << err << "\n";
            if (err < c.convergence_criteria && ntime > cs.ntime_monitor) {
                std::cout << "Simulation converged based on saturation! Relative error = " << err << " ; maximum velocity = " << T(m.umax) << std::endl;
                end_indicator = 1;
            }
        }
        if (c.breakthrough_check == 1 && m.outlet_phase1_count >= 1) {   // monitor_breakthrough, src/Monitor.cpp:447-466
            end_indicator = 1;
            std::cout << "Breakthrough point reached! Exiting program!" << std::endl;
        }
    }

    // monitor_multiphase_steady_capillarypressure, src/Monitor.cpp:357-441
    void monitor_capillary_pressure() {
        mflbm_monitor_out m;
        monitor_raw(m);
        const T pre_w = T(m.pre_w_sum) / (m.n_w + cs.eps), pre_nw = T(m.pre_nw_sum) / (m.n_nw + cs.eps);
        const T pc = (pre_nw - pre_w) / T(3.);
        monitor_previous = monitor_current;
        monitor_current = pc;
        const T err = std::abs(monitor_current - monitor_previous) / (std::abs(monitor_current) + cs.eps);
        RowFile(opt.dir, "steady_monitor_capillary_pressure_error.dat", fresh_file()) << ntime << " " << err << " " << pc << " " << pre_w << " " << pre_nw << " " << err << " " << T(m.umax) << "\n";
        if (std::isnan(m.umax) || std::isnan(pc) || m.nan_detected) { std::cout << "Simulation failed due to NAN!" << std::endl; end_indicator = 3; }
        else if (T(m.umax) < T(0.5)) {
            if (err < cs.ctl.convergence_criteria && ntime > cs.ntime_monitor) {
                std::cout << "Simulation converged based on change of capillary pressure! Relative error =   " << err << " ; maximum velocity = " << T(m.umax) << std::endl;
                end_indicator = 1;
            }
        } else { std::cout << "maximum velocity larger than 0.5, simulation failed!!!" << std::endl; end_indicator = 3; }
    }

    // monitor_multiphase_steady_phasefield, src/Monitor.cpp:279-352 (ntime_relaxation is never set in the reference: 0)
    void monitor_phase_field() {
        mflbm_monitor_out m;
        monitor_raw(m);
        double d = 0.0;
        h.phi_change(0, &d);
        if (ntime <= 0) return;
        if (std::isnan(m.umax) || std::isnan(d) || m.nan_detected) { std::cout << "Simulation failed due to NAN!" << std::endl; end_indicator = 3; return; }
        RowFile(opt.dir, "steady_monitor_max_phi_change.dat", fresh_file()) << ntime << " " << T(d) << " " << T(m.umax) << "\n";
        if (T(m.umax) < 0.5) {
            if (T(d) < cs.ctl.convergence_criteria && ntime > cs.ntime_monitor) {
                std::cout << "Simulation converged based on maximum local change of phi! Relative error =    " << T(d) << "; maximum velocity =  " << T(m.umax) << std::endl;
                end_indicator = 1;
            }
        } else { std::cout << "Maximum velocity larger than 0.5, simulation failed!!!" << std::endl; end_indicator = 3; }
    }

    static bool due(int nt, int timer) { return timer > 0 && nt % timer == 0; }
    // first step >= nt at which any output timer fires, capped at `last`
    int next_event(int nt, int last) const {
        const int timers[] = {cs.ctl.ntime_clock_sum, cs.ntime_monitor, cs.ctl.ntime_display_steps, cs.ntime_animation, cs.ntime_visual};
        int e = last;
        for (int t : timers) if (t > 0) { const long long n = ((long long)(nt + t - 1) / t) * t; if (n < e) e = (int)n; }
        return e;
    }

    // the main loop of src/main.cpp:145-263, batched between events
    Clock::time_point t_loop;
    void loop() {
        const auto& c = cs.ctl;
        std::cout << "************************** Entering main loop *********************************" << std::endl;
        int counter_ck = 1, counter_ck2 = 1;
        auto t_clock = Clock::now(), t_display = Clock::now();
        t_loop = Clock::now();
        const int last = cs.ntime0 + cs.ntime_max;
        long long since_display = 0;
        ntime = cs.ntime0;
        while (ntime <= last) {
            const int e = next_event(ntime, last);
            h.run(ntime, e - ntime + 1);
            since_display += e - ntime + 1;
            ntime = e;
            const bool any = due(e, c.ntime_clock_sum) || due(e, cs.ntime_monitor) || due(e, c.ntime_display_steps) || due(e, cs.ntime_animation) || due(e, cs.ntime_visual);
            if (any) h.sync();
            if (due(e, c.ntime_clock_sum)) {
                std::ofstream f((opt.dir + "/results/out1.output/time.dat").c_str(), (cs.ntime0 == 1 && e == c.ntime_clock_sum) ? std::ios_base::out : std::ios_base::app);
                if (!f.good()) throw Fatal("Could not open results/out1.output/time.dat");
                f << e << " " << seconds_since(t_clock) << " sec" << std::endl;
                t_clock = Clock::now();
            }
            if (due(e, cs.ntime_monitor)) {
                if (c.steady_state_option == 0 || c.steady_state_option == 3) monitor_displacement();
                else if (c.steady_state_option == 1) monitor_capillary_pressure();
                else if (c.steady_state_option == 2) monitor_phase_field();
            }
            if (due(e, c.ntime_display_steps)) {
                const double speed = double(cs.nx() * cs.ny() * cs.nz()) * double(since_display) / (1e6 * seconds_since(t_display));
                std::cout << "ntime =\t" << e << std::endl << "simulation speed: " << speed << " MLUPS" << std::endl;
                t_display = Clock::now(); since_display = 0;
            }
            if (due(e, cs.ntime_animation)) write_vtk(e, 2);
            if (due(e, cs.ntime_visual)) write_vtk(e, 1);
            if (due(e, c.ntime_clock_sum)) {   // wall-clock driven checkpoints, src/main.cpp:226-259
                const double hours = seconds_since(t_loop) * 2.77777777777e-4;
                std::cout << "simulation has run\t" << hours << "\thours" << std::endl;
                if (hours >= c.simulation_duration_timer) { std::cout << "Time to save checkpoint data and exit the program!" << std::endl; end_indicator = 2; }
                if (hours >= counter_ck * c.checkpoint_save_timer) { std::cout << "Time to save checkpoint data!" << std::endl; save_checkpoint(false); counter_ck++; }
                if (hours >= counter_ck2 * c.checkpoint_2rd_save_timer) { std::cout << "Time to save secondary checkpoint data!" << std::endl; save_checkpoint(true); counter_ck2++; }
            }
            if (end_indicator > 0) break;
            ntime = e + 1;
        }
        h.sync();
        std::cout << "************************** Exiting main iteration *********************************" << std::endl;
        const double total = seconds_since(t_loop);
        const double speed = double(cs.nx() * cs.ny() * cs.nz()) * double((ntime - cs.ntime0) - 1) / (1e6 * total);
        std::cout << "Code speed:\t" << speed << " MLUPS" << std::endl;
    }

    // end-of-run state machine, src/main.cpp:273-357
    void finish() {
        if (end_indicator == 0) {
            ntime = ntime - 1;
            save_checkpoint(false);
            write_vtk(ntime, 2); write_vtk(ntime, 1);
            std::cout << "Simulation ended after\t" << ntime << " iterations which reached the maximum time step!" << std::endl;
            write_status("simulation_reached_max_step");
        } else if (end_indicator == 1) {
            save_checkpoint(false);
            write_vtk(ntime, 2); write_vtk(ntime, 1);
            std::cout << "Simulation ended successfully after\t" << ntime << " iterations!" << std::endl;
            write_status("simulation_done");
        } else if (end_indicator == 2) {
            save_checkpoint(false);
            std::cout << "Simulation data saved but not finished after\t" << ntime << " iterations!" << std::endl;
            write_status("continue_simulation");
        } else {
            write_vtk(ntime, 1);
            std::cout << "Simulation failed after\t" << ntime << " iterations!" << std::endl;
            write_status("simulation_failed");
        }
    }
};

int main(int argc, char** argv) {
    Options o;
    for (int a = 1; a < argc; a++) {
        const std::string s = argv[a];
        auto next = [&]() -> std::string { if (a + 1 >= argc) { std::cerr << "missing value after " << s << std::endl; exit(2); } return argv[++a]; };
        if (s == "--dir") o.dir = next();
        else if (s == "--prec") o.prec = next();
        else if (s == "--device") o.device = std::stoi(next());
        else if (s == "--gpus") { const int n = std::stoi(next()); o.devices.clear(); for (int d = 0; d < n; d++) o.devices.push_back(d); }
        else if (s == "--devices") { std::stringstream ss(next()); std::string tok; o.devices.clear(); while (std::getline(ss, tok, ',')) if (!tok.empty()) o.devices.push_back(std::stoi(tok)); }
        else if (s == "--mrt") o.mrt = std::stoi(next());
        else if (s == "--no-geometry-quirk") o.quirk = false;
        else if (s == "--check-input") o.check_input = true;
        else {
            std::cerr << "usage: mflbm_run [--dir CASE_DIR] [--prec f32|f64] [--device N | --gpus N | --devices a,b,..] [--mrt 1..4] [--no-geometry-quirk] [--check-input]" << std::endl;
            return s == "--help" || s == "-h" ? 0 : 2;
        }
    }
    std::cout << "==============================================================================" << std::endl;
    std::cout << "MF-LBM time-step path, B200-native build (libmflbm " << mflbm_version() << ")" << std::endl;
    std::cout << "==============================================================================" << std::endl;
    try {
        if (o.prec == "f32") return Run<float>(o).main();
        if (o.prec == "f64") return Run<double>(o).main();
        std::cerr << "--prec must be f32 or f64" << std::endl;
        return 2;
    } catch (const Fatal& e) {
        std::cout << "Error: " << e.what() << std::endl;   // the reference's ERROR(): message, exit(1) (src/utils.cpp:8-13)
        return 1;
    }
}
