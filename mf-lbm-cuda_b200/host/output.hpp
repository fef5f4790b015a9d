// output.hpp — results/out1.output, results/out2.checkpoint, results/out3.field_data in the reference's formats.
//
// Formats (paths relative to /root/reference):
//   checkpoint  src/IO_multiphase.cpp:252-305   int ntime+1 | T force_z | T rho_in | pdf[38*s1] | phi[s4] |
//                                               (outlet_BC == 1) f_convec[19*NX1*NY1] g_convec[...] phi_convec[NX1*NY1]
//   legacy VTK  src/IO_multiphase.cpp:500-716   STRUCTURED_POINTS, BINARY big-endian; type 1 phi/density/velocity_X/Y/Z in
//                                               T (declared "float" when output_fieldData_precision_cmd == 0 whatever T
//                                               is), type 2 phi as float with solids zeroed
//   walls VTK   src/IO_multiphase.cpp:800-841   walls as big-endian int32
//   monitors    src/Monitor.cpp:119-241         one text row per monitor step, default ostream formatting of T
// The reference converts and writes element by element; here a field is byte-swapped into one buffer and written with
// a single call.
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "control.hpp"

namespace mfhost {

inline void make_dir(const std::string& p) {
    struct stat st;
    if (stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) return;
    if (mkdir(p.c_str(), 0777) != 0) throw Fatal("Could not Create " + p + " folder !");
}

// the directory tree of src/Init_multiphase.cpp:22-92
inline void make_result_dirs(const std::string& dir) {
    for (const char* d : {"results", "results/out2.checkpoint", "results/out2.checkpoint/2rd_backup", "results/out1.output",
                          "results/out1.output/profile", "results/out3.field_data", "results/out3.field_data/phase_distribution",
                          "results/out3.field_data/full_flow_field"})
        make_dir(dir + "/" + d);
}

template <typename V> inline V byteswap(V v) {
    unsigned char* b = reinterpret_cast<unsigned char*>(&v);
    std::reverse(b, b + sizeof(V));
    return v;
}

// rows appended to results/out1.output/<name>; `fresh` truncates (first monitor step of a new simulation)
class RowFile {
  public:
    RowFile(const std::string& dir, const std::string& name, bool fresh)
        : out_((dir + "/results/out1.output/" + name).c_str(), fresh ? std::ios_base::out : std::ios_base::app) {
        if (!out_.good()) throw Fatal("Could not open results/out1.output/" + name);
    }
    template <typename A> RowFile& operator<<(const A& a) { out_ << a; return *this; }
    void end() { out_ << std::endl; }

  private:
    std::ofstream out_;
};

template <typename T>
struct Checkpoint {
    int ntime_next = 0;
    T force_z = 0, rho_in = 0;
    std::vector<T> pdf, phi, f_convec, g_convec, phi_convec;
};

inline std::string checkpoint_path(const std::string& dir, bool secondary) {
    return dir + (secondary ? "/results/out2.checkpoint/2rd_backup/id0000" : "/results/out2.checkpoint/id0000");
}

template <typename T>
void write_checkpoint(const std::string& path, const Checkpoint<T>& c, bool convective) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Fatal("Could not create save_checkpoint file!");
    bool ok = fwrite(&c.ntime_next, sizeof(int), 1, f) == 1 && fwrite(&c.force_z, sizeof(T), 1, f) == 1 && fwrite(&c.rho_in, sizeof(T), 1, f) == 1;
    auto put = [&](const std::vector<T>& v) { ok = ok && fwrite(v.data(), sizeof(T), v.size(), f) == v.size(); };
    put(c.pdf); put(c.phi);
    if (convective) { put(c.f_convec); put(c.g_convec); put(c.phi_convec); }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Fatal("short write on " + path);
}

// sizes must be set by the caller (vectors pre-sized); src/Init_multiphase.cpp:501-538
template <typename T>
void read_checkpoint(const std::string& path, Checkpoint<T>& c, bool convective) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Fatal("Checkpoint data not found! Exiting program!");
    bool ok = fread(&c.ntime_next, sizeof(int), 1, f) == 1 && fread(&c.force_z, sizeof(T), 1, f) == 1 && fread(&c.rho_in, sizeof(T), 1, f) == 1;
    auto get = [&](std::vector<T>& v) { ok = ok && fread(v.data(), sizeof(T), v.size(), f) == v.size(); };
    get(c.pdf); get(c.phi);
    if (convective) { get(c.f_convec); get(c.g_convec); get(c.phi_convec); }
    fclose(f);
    if (!ok) throw Fatal("Could not load from data check_point_file file!");
}

class VtkFile {
  public:
    VtkFile(const std::string& path, long long nx, long long ny, long long nz) : out_(path.c_str(), std::ios_base::out | std::ios::binary), n_((size_t)(nx * ny * nz)) {
        if (!out_.good()) throw Fatal("Could not open vtk output file " + path);
        out_ << "# vtk DataFile Version 3.0" << std::endl << "vtk output" << std::endl << "BINARY" << std::endl << "DATASET STRUCTURED_POINTS" << std::endl;
        out_ << "DIMENSIONS " << nx << " " << ny << " " << nz << std::endl;
        out_ << "ORIGIN " << 1 << " " << 1 << " " << 1 << std::endl << "SPACING " << 1 << " " << 1 << " " << 1 << std::endl;
        out_ << "POINT_DATA " << nx * ny * nz << std::endl;
    }
    // `interior`: n values in x-fastest order; written big-endian with one call
    template <typename V> void scalars(const char* name, const char* type, const std::vector<V>& interior) {
        out_ << "SCALARS " << name << " " << type << std::endl << "LOOKUP_TABLE default" << std::endl;
        std::vector<V> be(interior.size());
        for (size_t n = 0; n < interior.size(); n++) be[n] = byteswap(interior[n]);
        out_.write(reinterpret_cast<const char*>(be.data()), (std::streamsize)(sizeof(V) * be.size()));
    }
    size_t points() const { return n_; }

  private:
    std::ofstream out_;
    size_t n_;
};

// interior [1..n]^3 of an array stored with G ghost layers, x fastest
template <typename S, typename D>
void crop(const std::vector<S>& src, int G, long long nx, long long ny, long long nz, std::vector<D>& dst) {
    dst.resize((size_t)(nx * ny * nz));
    const long long NX = nx + 2 * G, NY = ny + 2 * G;
    for (long long k = 0; k < nz; k++)
        for (long long j = 0; j < ny; j++) {
            const S* s = &src[(size_t)(G + NX * ((j + G) + NY * (k + G)))];
            D* d = &dst[(size_t)(nx * (j + ny * k))];
            for (long long i = 0; i < nx; i++) d[i] = D(s[i]);
        }
}

inline std::string vtk_name(const std::string& dir, const char* stem, int nt) {
    std::ostringstream s;
    s << dir << "/results/out3.field_data/" << stem << std::setfill('0') << std::setw(10) << nt << ".vtk";
    return s.str();
}

}  // namespace mfhost
