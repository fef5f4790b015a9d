// case.hpp — everything the host derives from the input files before the GPU layer takes over: scalar parameters in the
// solver precision, the interior wall array, the pore profile and the velocity-inlet profile.
//
// Arithmetic is done in T exactly where the reference does it in T_P, in the same order, so that the scalars handed to
// the GPU layer are bit-identical to the reference's (src/Init_multiphase.cpp:128-284, src/IO_multiphase.cpp:200-206,
// src/Misc.cpp:17-217,387-419; paths relative to /root/reference).  Built with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "../../include/mflbm.h"
#include "control.hpp"

namespace mfhost {

template <typename T> struct Api;   // C-ABI dispatch by precision, see api.hpp

template <typename T> inline T m_cos(T x) { return std::cos(x); }
template <typename T> inline T m_tanh(T x) { return std::tanh(x); }
template <typename T> inline T m_exp(T x) { return std::exp(x); }
template <typename T> inline T m_pow(T x, T y) { return std::pow(x, y); }

template <typename T>
struct Case {
    Control<T> ctl;
    std::string dir;
    // derived scalars
    T Pi = T(3.14159265358979323846L);                       // includes/Module.h:9
    T eps = std::numeric_limits<float>::epsilon();           // includes/Module.h:15: float epsilon in both precisions
    T cos_theta = 0, la_x = 0, la_y = 0, la_z = 0, A_xy = 0, A_xy_effective = 0, volume_sample = 0;
    T la_nui1 = 0, la_nui2 = 0, phi_inlet = 0, force_z = 0, rho_in = 0, rho_out = 1, uin_avg = 0, uin_avg_0 = 0, flowrate = 0, relaxation = 1;
    T porosity_full = 0, porosity_effective = 0;
    int ntime0 = 1, ntime_max = 0, ntime_monitor = 0, ntime_monitor_profile = 0, ntime_animation = 0, ntime_visual = 0;
    long long pore_sum = 0, pore_sum_effective = 0;
    long long nx_sample = 0, ny_sample = 0, nz_sample = 0;
    std::vector<int8_t> walls;          // interior [nz][ny][nx], 1 = solid, after set_walls
    std::vector<int> pore_profile_z;    // fluid nodes per z slice
    std::vector<T> W_in;                // (nx+2)*(ny+2), inlet_BC == 1 only
    bool geometry_index_quirk = true;   // replicate the swapped-stride read of src/Misc.cpp:171

    long long nx() const { return ctl.nxGlobal; }
    long long ny() const { return ctl.nyGlobal; }
    long long nz() const { return ctl.nzGlobal; }
    size_t i0(long long i, long long j, long long k) const { return (size_t)((i - 1) + nx() * ((j - 1) + ny() * (k - 1))); }

    // ---- read_walls, src/Misc.cpp:141-196 -------------------------------------------------------------------------
    void read_walls(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw Fatal("Could not open the geometry file ! (" + path + ")");
        bool ok = true;
        if (ctl.geometry_dims_type_size == 4) {
            int32_t d[3] = {0, 0, 0};
            ok = fread(d, sizeof(int32_t), 3, f) == 3;
            nx_sample = d[0]; ny_sample = d[1]; nz_sample = d[2];
        } else {
            int64_t d[3] = {0, 0, 0};
            ok = fread(d, sizeof(int64_t), 3, f) == 3;
            nx_sample = d[0]; ny_sample = d[1]; nz_sample = d[2];
        }
        if (!ok || nx_sample <= 0 || ny_sample <= 0 || nz_sample <= 0) { fclose(f); throw Fatal("Could not load from the geometry file!"); }
        std::cout << "Porous media sample size: nx = " << nx_sample << ", ny = " << ny_sample << ", nz = " << nz_sample << std::endl;
        if (nx() < nx_sample || ny() < ny_sample || nz() < nz_sample) { fclose(f); throw Fatal("Error! Domain size is smaller than porous media sample size! Exiting program!"); }
        const size_t ns = (size_t)nx_sample * ny_sample * nz_sample;
        std::vector<int8_t> raw(ns);
        ok = fread(raw.data(), 1, ns, f) == ns;
        fclose(f);
        if (!ok) throw Fatal("Could not load from the geometry file!");
        // The reference indexes the sample with the GLOBAL extents and with x/y strides swapped (src/Misc.cpp:171):
        // right only when the domain equals the sample and nx == ny.  Kept (default) so that both programs see the same
        // geometry; an index past the sample - undefined behaviour in the reference - reads as fluid here.
        for (long long k = 1; k <= nz(); k++)
            for (long long j = 1; j <= ny(); j++)
                for (long long i = 1; i <= nx(); i++) {
                    size_t src;
                    if (geometry_index_quirk) src = (size_t)((i - 1) + ny() * ((j - 1) + nx() * (k - 1)));
                    else if (i <= nx_sample && j <= ny_sample && k <= nz_sample) src = (size_t)((i - 1) + nx_sample * ((j - 1) + ny_sample * (k - 1)));
                    else src = ns;
                    walls[i0(i, j, k)] = src < ns ? raw[src] : 0;
                }
        if (ctl.wall_x_max == 1 && ctl.wall_y_max == 1)   // pad with solid walls
            for (long long k = 1; k <= nz(); k++)
                for (long long j = 1; j <= ny(); j++)
                    for (long long i = 1; i <= nx(); i++)
                        if (j >= ny_sample || i >= nx_sample) walls[i0(i, j, k)] = 1;
    }

    // the sample obstacle of modify_geometry (src/Misc.cpp:103-138): a sphere in a tube with 10 open layers at each end
    void sample_obstacle() {
        const double xc = 0.5 * double(nx() + 1), yc = 0.5 * double(ny() + 1), zc = 0.5 * double(nz() + 1);
        const double r1 = 0.25 * ny(), r2 = ny() * 0.5;
        const long long buffer = 10;
        for (long long k = 1; k <= nz(); k++)
            for (long long j = 1; j <= ny(); j++)
                for (long long i = 1; i <= nx(); i++) {
                    const double d2 = (i - xc) * (i - xc) + (j - yc) * (j - yc) + (k - zc) * (k - zc);
                    if (d2 < r1 * r1) walls[i0(i, j, k)] = 1;
                    if (d2 > r2 * r2 && k > buffer && k < nz() - buffer + 1) walls[i0(i, j, k)] = 1;
                }
        std::cout << "Internal geometry modified!" << std::endl;
    }

    // ---- set_walls + pore_profile, src/Misc.cpp:17-101,198-217 ----------------------------------------------------
    void set_walls() {
        walls.assign((size_t)nx() * ny() * nz(), 0);
        if (ctl.external_geometry_read_cmd == 1) {
            std::cout << "This simulation uses external geometry data!" << std::endl;
            read_walls(geometry_path(dir, "geo_file_path_"));
        } else {
            std::cout << "This simulation does not use external geometry data!" << std::endl;
            nx_sample = nx(); ny_sample = ny(); nz_sample = nz();
        }
        if (ctl.modify_geometry_cmd == 1) sample_obstacle();
        for (long long k = 1; k <= nz(); k++)
            for (long long j = 1; j <= ny(); j++) {
                if (ctl.wall_x_min == 1) walls[i0(1, j, k)] = 1;
                if (ctl.wall_x_max == 1) walls[i0(nx(), j, k)] = 1;
            }
        for (long long k = 1; k <= nz(); k++)
            for (long long i = 1; i <= nx(); i++) {
                if (ctl.wall_y_min == 1) walls[i0(i, 1, k)] = 1;
                if (ctl.wall_y_max == 1) walls[i0(i, ny(), k)] = 1;
            }
        for (long long j = 1; j <= ny(); j++)
            for (long long i = 1; i <= nx(); i++) {
                if (ctl.wall_z_min == 1) walls[i0(i, j, 1)] = 1;
                if (ctl.wall_z_max == 1) walls[i0(i, j, nz())] = 1;
            }
        for (auto w : walls) if (w != 0 && w != 1) throw Fatal("geometry file holds values other than 0 (fluid) / 1 (solid)");
        long long open = 0;
        for (long long j = 1; j <= ny(); j++) for (long long i = 1; i <= nx(); i++) open += walls[i0(i, j, 1)] <= 0;
        A_xy_effective = T(open);
        std::cout << "Inlet effective open area =    " << A_xy_effective << std::endl;
        pore_profile_z.assign((size_t)nz(), 0);
        pore_sum = pore_sum_effective = 0;
        for (long long k = 1; k <= nz(); k++) {
            int n = 0;
            const int8_t* p = &walls[i0(1, 1, k)];
            for (long long m = 0; m < nx() * ny(); m++) n += p[m] <= 0;
            pore_profile_z[(size_t)k - 1] = n;
            pore_sum += n;
            if (k >= 1 + ctl.n_exclude_inlet && k <= nz() - ctl.n_exclude_outlet) pore_sum_effective += n;
        }
    }

    // ---- analytical duct profile, src/Init_multiphase.cpp:258-296 + src/Misc.cpp:387-419 --------------------------
    void velocity_inlet() {
        uin_avg_0 = ctl.ca_0 * ctl.lbm_gamma / ctl.la_nu1;
        uin_avg = uin_avg_0;
        flowrate = uin_avg_0 * A_xy;
        std::cout << "Inlet average velocity = " << uin_avg_0 << std::endl;
        std::cout << "Inlet flowrate = " << flowrate << std::endl;
        const long long NX1 = nx() + 2;
        W_in.assign((size_t)(NX1 * (ny() + 2)), T(0));
        auto W = [&](long long i, long long j) -> T& { return W_in[(size_t)(i + NX1 * j)]; };
        for (long long j = 2; j < ny(); j++) for (long long i = 2; i < nx(); i++) W(i, j) = uin_avg;
        const int terms = 1000;
        const T a = T(0.5) * la_x, b = T(0.5) * la_y;
        T t1 = T(0);
        for (long long n = 1; n <= terms; n += 2) t1 += (m_tanh<T>(T(0.5) * T(n) * Pi * b / a)) / m_pow<T>(T(n), T(5));
        T t2 = T(1.) - T(192.) / m_pow<T>(Pi, T(5)) * (a / b) * t1;
        t2 = T(-3.) * uin_avg_0 / (t2 * m_pow<T>(a, T(2)));
        for (long long j = 2; j < ny(); j++)
            for (long long i = 2; i < nx(); i++) {
                const T xx = i - T(1.5) - a, yy = j - T(1.5) - b;
                T t3 = T(0);
                for (long long n = 1; n <= terms; n += 2)
                    t3 += m_pow<T>(T(-1.), T(0.5) * T(n - 1)) * m_cos<T>(T(0.5) * T(n) * Pi * xx / a) / m_pow<T>(T(n), T(3))
                          * (T(1.) - (m_exp<T>(T(0.5) * T(n) * Pi * (yy - b) / a) + m_exp<T>(T(0.5) * T(n) * Pi * (-yy - b) / a)) /
                                         (T(1.) + m_exp<T>(T(0.5) * T(n) * Pi * (-b - b) / a)));
                W(i, j) = t3 * (T(-16.) * t2 * m_pow<T>(a, T(2)) * m_pow<T>(Pi, T(-3)));
            }
        if (ctl.target_inject_pore_volume > 0) {
            ntime_max = int(T(ctl.target_inject_pore_volume * pore_sum) / flowrate);
            if (ntime_max % 2 == 1) ntime_max++;
            std::cout << "Maximum time step is modified based on target inject volume! ntime_max = " << ntime_max << std::endl;
        }
    }

    // ---- scalar derivations of initialization_basic_multi, src/Init_multiphase.cpp:128-254 ------------------------
    void derive() {
        T theta = T(180.) - ctl.theta;         // measured through the defending phase, src/IO_multiphase.cpp:204-206
        theta = theta * Pi / T(180.);
        cos_theta = m_cos<T>(theta);
        la_z = T(nz() - 1);
        la_y = T(ny() - 1) - T(0.5) - T(0.5);
        la_x = T(nx() - 1) - T(0.5) - T(0.5);
        A_xy = la_x * la_y;
        volume_sample = A_xy * la_z;
        porosity_full = T(pore_sum) / T(((nx() - 2) * (ny() - 2) * (nz())));
        porosity_effective = T(pore_sum_effective) / ((nx() - 2) * (ny() - 2)) / T(nz() - ctl.n_exclude_outlet - ctl.n_exclude_inlet);
        la_nui1 = T(1.) / ctl.la_nu1;
        la_nui2 = T(1.) / ctl.la_nu2;
        phi_inlet = T(2.) * ctl.sa_inject - T(1.);
        force_z = ctl.force_z0;
        rho_out = T(1.);
        ntime_max = ctl.ntime_max; ntime_monitor = ctl.ntime_monitor; ntime_animation = ctl.ntime_animation; ntime_visual = ctl.ntime_visual;
        ntime_monitor_profile = ctl.ntime_monitor_profile_ratio * ntime_monitor;
        if (ctl.open_z()) {
            if (ctl.inlet_BC == 1) { force_z = T(0.); velocity_inlet(); }
            else if (ctl.inlet_BC == 2) { force_z = T(0.); pressure_inlet(); }
        }
        // output timers by injected volume, src/Init_multiphase.cpp:216-248 (odd timers go to the next even step: after an
        // odd step the AA pattern holds the PDFs in swapped slots)
        const bool vin = ctl.inlet_BC == 1 && ctl.open_z();
        auto by_volume = [&](T d_vol) { int n = int(T(d_vol * pore_sum) / flowrate); return n % 2 == 1 ? n + 1 : n; };
        if (ctl.d_vol_animation > T(0.) && vin) ntime_animation = by_volume(ctl.d_vol_animation);
        if (ctl.d_vol_detail > T(0.) && vin) ntime_visual = by_volume(ctl.d_vol_detail);
        if (ctl.d_vol_monitor > T(0.) && vin) { ntime_monitor = by_volume(ctl.d_vol_monitor); ntime_monitor_profile = ctl.ntime_monitor_profile_ratio * ntime_monitor; }
    }

    // pressure BC densities from the pressure gradient, src/Init_multiphase.cpp:199-211 and src/main.cpp:124-134
    void pressure_inlet() {
        const T p_gradient = -ctl.force_z0 / T(3.);
        if (ctl.rho_out_BC) rho_out = T(1.) - p_gradient * nz();
        else rho_in = rho_out - p_gradient * nz();
    }

    // the scalars of copyConstantData + kernel arguments, as the C ABI wants them
    template <typename P> void fill_params(P& p, int mrt) const {
        std::memset(&p, 0, sizeof p);
        p.nx = nx(); p.ny = ny(); p.nz = nz();
        p.iper = ctl.iper; p.jper = ctl.jper; p.kper = ctl.kper;
        p.wall_z_min = ctl.wall_z_min; p.wall_z_max = ctl.wall_z_max;
        p.inlet_BC = ctl.inlet_BC; p.outlet_BC = ctl.outlet_BC;
        p.porous_plate_cmd = ctl.porous_plate_cmd; p.Z_porous_plate = ctl.Z_porous_plate;
        p.n_exclude_inlet = ctl.n_exclude_inlet; p.n_exclude_outlet = ctl.n_exclude_outlet;
        p.mrt = mrt;
        p.lbm_gamma = ctl.lbm_gamma; p.lbm_beta = ctl.lbm_beta; p.la_nu1 = ctl.la_nu1; p.la_nui1 = la_nui1; p.la_nui2 = la_nui2;
        p.cos_theta = cos_theta; p.force_z = force_z; p.rho_in = rho_in; p.rho_out = rho_out; p.phi_inlet = phi_inlet;
        p.sa_inject = ctl.sa_inject; p.uin_avg = uin_avg; p.relaxation = relaxation; p.A_xy = A_xy;
    }
};

}  // namespace mfhost
