// domain.hpp — one lattice on one GPU, or cut into x-slabs over several GPUs of the box, behind one interface.
//
// The reference is single-GPU (README.md:119 of /root/reference lists multi-GPU as future work); the x-slab decomposition
// is new.  With one device every call goes straight to the C ABI.  With N devices this class owns N slab
// solvers (mflbm_*_create with an mflbm_slab) in THIS process: neighbours are connected through plain peer pointers
// (mflbm_peer_enable + halo_p2p_local / halo_p2p_connect - no IPC handles, no NCCL), every slab's run() is enqueued
// asynchronously from the one host thread and the halo messages synchronise the devices among themselves (arrival flags
// in device memory, include/mflbm.h).  Host arrays keep the reference's GLOBAL layouts: uploads cut each slab's window
// (own columns + ghost columns), downloads gather the columns a slab owns.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "api.hpp"
#include "slabs.hpp"

namespace mfhost {

template <typename T>
class Domain {
  public:
    using Params = typename Api<T>::Params;
    using Handle = typename Api<T>::Handle;

    ~Domain() { for (auto h : hs) if (h) Api<T>::destroy(h); }

    int slabs() const { return (int)hs.size(); }

    void create(const Params& p, const std::vector<int>& devices) {
        P = p;
        const int n = (int)devices.size();
        nxg = (long long)p.nx; ny = (long long)p.ny; nz = (long long)p.nz;
        if (n <= 1) {
            hs.assign(1, nullptr);
            Api<T>::create(p, devices.empty() ? 0 : devices[0], &hs[0]);
            cut.assign(1, mflbm_slab{1, (int64_t)nxg, 0, 0});
            return;
        }
        if (nxg / n < 4) throw Fatal("a slab must be at least 4 columns wide (phi halo): fewer GPUs or a wider lattice");
        for (int a = 0; a + 1 < n; a++)
            if (devices[a] != devices[a + 1]) api_check(mflbm_peer_enable(devices[a], devices[a + 1]), "peer access between neighbouring devices");
        hs.assign((size_t)n, nullptr);
        cut = cut_slabs(nxg, n);
        for (int r = 0; r < n; r++) Api<T>::create_slab(p, cut[(size_t)r], devices[r], &hs[(size_t)r]);
    }
    // neighbours' receive buffers and arrival flags: a message sent through my right face lands in my right neighbour's
    // left-side buffer.  After the geometry (the buffers exist from create on, the call order only mirrors mflbm/slab.py).
    void connect() {
        for (int r = 0; r < slabs(); r++)
            for (int kind = 0; kind < 3; kind++) {
                if (cut[r].has_left) Api<T>::connect(hs[r], kind, 0, hs[r - 1], 1);
                if (cut[r].has_right) Api<T>::connect(hs[r], kind, 1, hs[r + 1], 0);
            }
    }

    void set_params(const Params& p) { P = p; for (auto h : hs) Api<T>::set_params(h, p); }
    void preprocess_geometry(const int8_t* walls_global) { for (auto h : hs) Api<T>::preprocess_geometry(h, walls_global); if (slabs() > 1) connect(); }
    void geometry_counts(int64_t* c) {   // every slab counts the sites of the columns it owns
        for (int n = 0; n < 4; n++) c[n] = 0;
        for (auto h : hs) { int64_t m[4]; Api<T>::geometry_counts(h, m); for (int n = 0; n < 4; n++) c[n] += m[n]; }
    }
    long long fluid_nodes() { long long s = 0; for (auto h : hs) s += Api<T>::fluid_nodes(h); return s; }

    void init_state(int opt, T z0, const T* W) {
        for (int r = 0; r < slabs(); r++) { window(W, 1, 1, 1, r, w_); Api<T>::init_state(hs[r], opt, z0, W ? w_.data() : nullptr); }
        last_ntime = 0;
    }
    void init_state_from_phi(const T* phi, const T* W) {
        std::vector<T> ph;
        for (int r = 0; r < slabs(); r++) { window(W, 1, 1, 1, r, w_); window(phi, 4, 1, nz + 8, r, ph); Api<T>::init_state_from_phi(hs[r], ph.data(), W ? w_.data() : nullptr); }
        last_ntime = 0;
    }
    void upload_pdf(const T* pdf) {
        std::vector<T> a;
        for (int r = 0; r < slabs(); r++) { window(pdf, 1, 38, nz + 2, r, a); Api<T>::upload_pdf(hs[r], a.data()); }
    }
    void upload_restart(const T* pdf, const T* phi, const T* W, const T* fc, const T* gc, const T* pc) {
        if (slabs() == 1) { Api<T>::upload_restart(hs[0], pdf, phi, W, fc, gc, pc); return; }
        std::vector<T> a, b, c, d, e;
        for (int r = 0; r < slabs(); r++) {
            window(pdf, 1, 38, nz + 2, r, a); window(phi, 4, 1, nz + 8, r, b); window(W, 1, 1, 1, r, w_);
            window(fc, 1, 19, 1, r, c); window(gc, 1, 19, 1, r, d); window(pc, 1, 1, 1, r, e);
            Api<T>::upload_restart(hs[r], a.data(), b.data(), W ? w_.data() : nullptr, fc ? c.data() : nullptr, gc ? d.data() : nullptr, pc ? e.data() : nullptr);
        }
    }
    void color_gradient() { for (auto h : hs) Api<T>::color_gradient(h); }
    // every slab's batch is enqueued before any of them is waited for: the devices run concurrently
    void run(int first, int n) { for (auto h : hs) Api<T>::run(h, first, n); if (n > 0) last_ntime = first + n - 1; }
    void sync() { for (auto h : hs) Api<T>::sync(h); }

    void phi_change(int seed, double* d) {
        double m = 0.0;
        for (auto h : hs) { double v = 0.0; Api<T>::phi_change(h, seed, &v); m = (v != v || m != m) ? (m != m ? m : v) : std::max(m, v); }
        if (d) *d = m;
    }

    // per-slab sums combined as src/Monitor.cpp:111-171 does for the whole lattice (mflbm/slab.py reduce_monitor is the
    // torch.distributed twin of this)
    void monitor(mflbm_monitor_out* out) {
        if (slabs() == 1) { Api<T>::monitor(hs[0], out); return; }
        double* prof[7] = {out->fl1, out->fl2, out->pre, out->mass1, out->mass2, out->vol1, out->vol2};
        std::vector<double> tmp[7];
        mflbm_monitor_out acc;
        std::memset(&acc, 0, sizeof acc);
        for (int r = 0; r < slabs(); r++) {
            mflbm_monitor_out m;
            std::memset(&m, 0, sizeof m);
            double** dst[7] = {&m.fl1, &m.fl2, &m.pre, &m.mass1, &m.mass2, &m.vol1, &m.vol2};
            for (int k = 0; k < 7; k++) if (prof[k]) { tmp[k].assign((size_t)nz, 0.0); *dst[k] = tmp[k].data(); }
            Api<T>::monitor(hs[r], &m);
            for (int k = 0; k < 7; k++) if (prof[k]) for (long long z = 0; z < nz; z++) prof[k][z] = (r == 0 ? 0.0 : prof[k][z]) + tmp[k][(size_t)z];
            acc.vol1_sum += m.vol1_sum; acc.vol2_sum += m.vol2_sum; acc.mass1_sum += m.mass1_sum; acc.mass2_sum += m.mass2_sum;
            acc.vol1_full += m.vol1_full; acc.vol2_full += m.vol2_full; acc.mass1_full += m.mass1_full; acc.mass2_full += m.mass2_full;
            acc.fl1_avg += m.fl1_avg; acc.fl2_avg += m.fl2_avg; acc.fl1_avg_whole += m.fl1_avg_whole; acc.fl2_avg_whole += m.fl2_avg_whole;
            acc.kinetic_energy[0] += m.kinetic_energy[0]; acc.kinetic_energy[1] += m.kinetic_energy[1];
            acc.umax = std::max(acc.umax, m.umax); acc.nan_detected |= m.nan_detected;
            acc.pre_w_sum += m.pre_w_sum; acc.pre_nw_sum += m.pre_nw_sum; acc.n_w += m.n_w; acc.n_nw += m.n_nw; acc.outlet_phase1_count += m.outlet_phase1_count;
        }
        acc.saturation = acc.vol1_sum / (acc.vol1_sum + acc.vol2_sum);
        acc.saturation_full_domain = acc.vol1_full / (acc.vol1_full + acc.vol2_full);
        acc.ca = ((acc.fl1_avg + acc.fl2_avg) / (double)P.A_xy) * (double)P.la_nu1 / (double)P.lbm_gamma;
        acc.nan_detected = acc.nan_detected || std::isnan(acc.saturation_full_domain) || std::isnan(acc.ca);
        acc.fl1 = out->fl1; acc.fl2 = out->fl2; acc.pre = out->pre; acc.mass1 = out->mass1; acc.mass2 = out->mass2; acc.vol1 = out->vol1; acc.vol2 = out->vol2;
        *out = acc;
    }

    void download(T* pdf, T* phi, T* cx, T* cy, T* cz, T* cn, T* fc, T* gc, T* pc) {
        if (slabs() == 1) { Api<T>::download(hs[0], pdf, phi, cx, cy, cz, cn, fc, gc, pc); return; }
        settle();
        std::vector<T> a, b, c2[4], e, f, g;
        for (int r = 0; r < slabs(); r++) {
            const long long w = cut[r].nx_local;
            if (pdf) a.resize((size_t)(38 * (nz + 2) * (ny + 2) * (w + 2)));
            if (phi) b.resize((size_t)((nz + 8) * (ny + 8) * (w + 8)));
            T* two[4] = {cx, cy, cz, cn};
            for (int k = 0; k < 4; k++) if (two[k]) c2[k].resize((size_t)((nz + 4) * (ny + 4) * (w + 4)));
            if (fc) e.resize((size_t)(19 * (ny + 2) * (w + 2)));
            if (gc) f.resize((size_t)(19 * (ny + 2) * (w + 2)));
            if (pc) g.resize((size_t)((ny + 2) * (w + 2)));
            Api<T>::download(hs[r], pdf ? a.data() : nullptr, phi ? b.data() : nullptr, cx ? c2[0].data() : nullptr, cy ? c2[1].data() : nullptr,
                             cz ? c2[2].data() : nullptr, cn ? c2[3].data() : nullptr, fc ? e.data() : nullptr, gc ? f.data() : nullptr, pc ? g.data() : nullptr);
            gather(pdf, a, 1, 38, nz + 2, r); gather(phi, b, 4, 1, nz + 8, r);
            for (int k = 0; k < 4; k++) gather(two[k], c2[k], 2, 1, nz + 4, r);
            gather(fc, e, 1, 19, 1, r); gather(gc, f, 1, 19, 1, r); gather(pc, g, 1, 1, 1, r);
        }
    }
    void download_macro(T* rho, T* u, T* v, T* w) {
        if (slabs() == 1) { Api<T>::download_macro(hs[0], rho, u, v, w); return; }
        std::vector<T> a[4];
        T* out[4] = {rho, u, v, w};
        for (int r = 0; r < slabs(); r++) {
            for (int k = 0; k < 4; k++) if (out[k]) a[k].resize((size_t)((nz + 2) * (ny + 2) * (cut[r].nx_local + 2)));
            Api<T>::download_macro(hs[r], out[0] ? a[0].data() : nullptr, out[1] ? a[1].data() : nullptr, out[2] ? a[2].data() : nullptr, out[3] ? a[3].data() : nullptr);
            for (int k = 0; k < 4; k++) gather(out[k], a[k], 1, 1, nz + 2, r);
        }
    }

  private:
    std::vector<Handle*> hs;
    std::vector<mflbm_slab> cut;
    Params P{};
    long long nxg = 0, ny = 0, nz = 0;
    int last_ntime = 0;
    std::vector<T> w_;

    // Make every slab's OWN columns complete before they are gathered.  After an even step the "before odd" boundary kernels
    // of my neighbour have written the populations its next pull needs into ITS ghost copy of my boundary column; sending the
    // ghost columns back to their owner (the odd-step message) makes the owner's copy equal to the single-domain array
    // (mflbm/slab.py SlabStepper.settle).  After an odd step the regular exchange has already done it.
    void settle() {
        if (last_ntime <= 0 || last_ntime % 2 != 0) return;
        for (auto h : hs) Api<T>::halo_push(h, 1);
        for (auto h : hs) Api<T>::halo_unpack_wait(h, 1);
    }

    // arrays are [outer][z][y][x] with x fastest (slabs.hpp)
    void window(const T* global, int g, long long outer, long long zext, int r, std::vector<T>& local) const {
        if (!global) { local.clear(); return; }
        slab_window(global, outer * zext * (ny + 2 * g), g, nxg, cut[(size_t)r], local);
    }
    void gather(T* global, const std::vector<T>& local, int g, long long outer, long long zext, int r) const {
        if (!global) return;
        slab_gather(global, local.data(), outer * zext * (ny + 2 * g), g, nxg, cut[(size_t)r]);
    }
};

}  // namespace mfhost
