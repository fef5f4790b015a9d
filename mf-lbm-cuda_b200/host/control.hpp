// control.hpp — input side of the host driver: simulation_control.txt, job_status.txt, Geometry_File_Path.txt.
//
// File compatibility with the reference (paths relative to /root/reference):
//   * key/value text, '#' in column 0 starts a comment, the first line whose first token equals the key wins, a missing
//     key is fatal (src/utils.cpp:18-155; the reference re-opens and rescans the file for every key, we scan it once);
//   * the keys and their order of validation follow read_parameter_multi (src/IO_multiphase.cpp:16-247);
//   * job_status.txt is compared as a whole file, so a trailing newline is NOT tolerated (src/IO_multiphase.cpp:25-36).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mfhost {

struct Fatal : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline std::string slurp(const std::string& path, const char* what) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.good()) throw Fatal(std::string(what) + " (" + path + ")");
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// One scan of a key/value file; lookups have the reference's first-match semantics.
class KeyFile {
  public:
    explicit KeyFile(const std::string& path, const char* missing_msg) : path_(path) {
        std::istringstream in(slurp(path, missing_msg));
        std::string line;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream ls(line);
            std::string key, value;
            ls >> key;
            if (key.empty()) continue;
            ls >> value;
            rows_.emplace_back(key, value);
        }
    }
    const std::string& raw(const std::string& key) const {
        for (const auto& r : rows_) if (r.first == key) return r.second;
        throw Fatal("ERROR! Couldn't find the variable named: (" + key + ") in the file " + path_);
    }
    bool has(const std::string& key) const {
        for (const auto& r : rows_) if (r.first == key) return true;
        return false;
    }
    int get_int(const std::string& key) const { return conv(key, [](const std::string& s) { return std::stoi(s); }); }
    long long get_ll(const std::string& key) const { return conv(key, [](const std::string& s) { return std::stoll(s); }); }
    // READ_T_P: stof in the single-precision build, stod in the double-precision build (includes/utils.h:33-37)
    template <typename T> T get_real(const std::string& key) const {
        if (sizeof(T) == 4) return (T)conv(key, [](const std::string& s) { return std::stof(s); });
        return (T)conv(key, [](const std::string& s) { return std::stod(s); });
    }

  private:
    template <typename F> auto conv(const std::string& key, F f) const -> decltype(f(std::string())) {
        const std::string& v = raw(key);
        try { return f(v); }
        catch (const std::exception&) { throw Fatal("bad value '" + v + "' for key " + key + " in " + path_); }
    }
    std::string path_;
    std::vector<std::pair<std::string, std::string>> rows_;
};

// Every key of input/simulation_control.txt the reference reads (src/IO_multiphase.cpp:46-190).  The CUDA block-shape
// keys are parsed for compatibility and ignored: launch shapes are fixed by the kernels.
template <typename T>
struct Control {
    int initial_fluid_distribution_option, breakthrough_check, steady_state_option, benchmark_cmd, output_fieldData_precision_cmd;
    int extreme_large_sim_cmd, modify_geometry_cmd, external_geometry_read_cmd, geometry_dims_type_size, geometry_preprocess_cmd;
    int porous_plate_cmd, change_inlet_fluid_phase_cmd;
    long long nxGlobal, nyGlobal, nzGlobal;
    int n_exclude_inlet, n_exclude_outlet;
    int wall_x_min, wall_x_max, wall_y_min, wall_y_max, wall_z_min, wall_z_max;
    int iper, jper, kper;
    T convergence_criteria, la_nu1, la_nu2, lbm_gamma, theta, lbm_beta;
    int inlet_BC, outlet_BC;
    T target_inject_pore_volume, ca_0, sa_inject, interface_z0, force_z0, sa_target;
    int Z_porous_plate;
    int ntime_max, ntime_max_benchmark, ntime_visual, ntime_animation, ntime_monitor, ntime_monitor_profile_ratio, ntime_clock_sum, ntime_display_steps;
    T checkpoint_save_timer, checkpoint_2rd_save_timer, simulation_duration_timer;
    T d_vol_animation, d_vol_detail, d_vol_monitor;
    int rho_in_new, rho_out_BC;
    int block_Threads_X, block_Threads_Y, block_Threads_Z;
    std::string job_status;
    std::vector<std::string> problems;   // filled by validate()

    static Control read(const std::string& dir) {
        Control c{};
        c.job_status = slurp(dir + "/input/job_status.txt", "Missing job status file! Exiting program!");
        if (c.job_status != "new_simulation" && c.job_status != "continue_simulation") throw Fatal("Wrong simlation status! Exiting program!");
        KeyFile f(dir + "/input/simulation_control.txt", "Missing simulation control file! Exiting program!");
#define MF_INT(field, key) c.field = f.get_int(key)
#define MF_REAL(field, key) c.field = f.template get_real<T>(key)
        MF_INT(initial_fluid_distribution_option, "initial_fluid_distribution_option");
        MF_INT(breakthrough_check, "breakthrough_check");
        MF_INT(steady_state_option, "steady_state_option");
        MF_REAL(convergence_criteria, "convergence_criteria");
        MF_INT(benchmark_cmd, "benchmark_cmd");
        MF_INT(output_fieldData_precision_cmd, "output_fieldData_precision_cmd");
        MF_INT(extreme_large_sim_cmd, "extreme_large_sim_cmd");
        MF_INT(modify_geometry_cmd, "modify_geometry_cmd");
        MF_INT(external_geometry_read_cmd, "external_geometry_read_cmd");
        MF_INT(geometry_dims_type_size, "geometry_dims_type_size");
        if (c.geometry_dims_type_size != 4 && c.geometry_dims_type_size != 8) throw Fatal("Incorrect geometry_dims_type_size!");
        MF_INT(geometry_preprocess_cmd, "geometry_preprocess_cmd");
        MF_INT(porous_plate_cmd, "porous_plate_cmd");
        MF_INT(change_inlet_fluid_phase_cmd, "change_inlet_fluid_phase_cmd");
        c.nxGlobal = f.get_ll("nxGlobal"); c.nyGlobal = f.get_ll("nyGlobal"); c.nzGlobal = f.get_ll("nzGlobal");
        MF_INT(n_exclude_inlet, "n_exclude_inlet"); MF_INT(n_exclude_outlet, "n_exclude_outlet");
        MF_INT(wall_x_min, "domain_wall_status_x_min"); MF_INT(wall_x_max, "domain_wall_status_x_max");
        MF_INT(wall_y_min, "domain_wall_status_y_min"); MF_INT(wall_y_max, "domain_wall_status_y_max");
        MF_INT(wall_z_min, "domain_wall_status_z_min"); MF_INT(wall_z_max, "domain_wall_status_z_max");
        MF_INT(iper, "iper"); MF_INT(jper, "jper"); MF_INT(kper, "kper");
        MF_REAL(la_nu1, "fluid1_viscosity"); MF_REAL(la_nu2, "fluid2_viscosity"); MF_REAL(lbm_gamma, "surface_tension");
        MF_REAL(theta, "theta"); MF_REAL(lbm_beta, "RK_beta");
        MF_INT(inlet_BC, "inlet_BC"); MF_INT(outlet_BC, "outlet_BC");
        MF_REAL(target_inject_pore_volume, "target_inject_pore_volume"); MF_REAL(ca_0, "capillary_number");
        MF_REAL(sa_inject, "saturation_injection"); MF_REAL(interface_z0, "initial_interface_position");
        MF_REAL(force_z0, "body_force_0"); MF_REAL(sa_target, "target_fluid1_saturation");
        MF_INT(Z_porous_plate, "Z_porous_plate");
        MF_INT(ntime_max, "max_time_step"); MF_INT(ntime_max_benchmark, "max_time_step_benchmark");
        MF_INT(ntime_visual, "ntime_visual"); MF_INT(ntime_animation, "ntime_animation");
        MF_INT(ntime_monitor, "monitor_timer"); MF_INT(ntime_monitor_profile_ratio, "monitor_profile_timer_ratio");
        MF_INT(ntime_clock_sum, "computation_time_timer"); MF_INT(ntime_display_steps, "display_steps_timer");
        MF_REAL(checkpoint_save_timer, "checkpoint_save_timer"); MF_REAL(checkpoint_2rd_save_timer, "checkpoint_2rd_save_timer");
        MF_REAL(simulation_duration_timer, "simulation_duration_timer");
        MF_REAL(d_vol_animation, "d_vol_animation"); MF_REAL(d_vol_detail, "d_vol_detail"); MF_REAL(d_vol_monitor, "d_vol_monitor");
        MF_INT(rho_in_new, "rho_in_new"); MF_INT(rho_out_BC, "rho_out_BC");
        MF_INT(block_Threads_X, "block_Threads_X"); MF_INT(block_Threads_Y, "block_Threads_Y"); MF_INT(block_Threads_Z, "block_Threads_Z");
#undef MF_INT
#undef MF_REAL
        return c;
    }

    // the parameter checks of src/IO_multiphase.cpp:196-243 (messages kept recognisable); true when the case may run
    bool validate() {
        problems.clear();
        if (theta > T(90.)) problems.push_back("contact angle is larger than 90 degrees");
        if (iper == 1 || wall_x_max == 0 || wall_x_min == 0) problems.push_back("X direction periodic BC enabled or non-slip BC not applied at x = xmin or x = xmax");
        if (jper == 0 && (wall_y_max == 0 || wall_y_min == 0)) problems.push_back("non-slip BC not applied at y = ymin or y = ymax while y direction periodic BC not enabled");
        if (jper == 1 && (wall_y_max == 1 || wall_y_min == 1)) problems.push_back("non-slip BC applied at y = ymin or y = ymax while y direction periodic BC enabled");
        if (kper == 1 && (wall_z_max == 1 || wall_z_min == 1)) problems.push_back("non-slip BC applied at z = zmin or z = zmax while z direction periodic BC enabled");
        if (outlet_BC == 1 && inlet_BC == 2) problems.push_back("Inlet pressure + outlet convective BC is not supported");
        return problems.empty();
    }

    bool open_z() const { return kper == 0 && wall_z_min == 0 && wall_z_max == 0; }
};

// Geometry_File_Path.txt: value of `key`, trailing non-digits stripped, ".dat" appended (src/Misc.cpp:26-41).
// An all-non-digit name underflows in the reference; here it is an error.
inline std::string geometry_path(const std::string& dir, const char* key) {
    KeyFile f(dir + "/input/Geometry_File_Path.txt", "Error opening Geometry_File_Path.txt!");
    std::string p = f.raw(key);
    while (!p.empty() && !(p.back() >= '0' && p.back() <= '9')) p.pop_back();
    if (p.empty()) throw Fatal(std::string("geometry file name without a digit: ") + f.raw(key));
    p += ".dat";
    if (!p.empty() && p[0] != '/') p = dir + "/" + p;
    return p;
}

}  // namespace mfhost
