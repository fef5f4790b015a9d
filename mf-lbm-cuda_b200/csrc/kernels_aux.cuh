// kernels_aux.cuh — set-up and diagnostics kernels: bit-exact geometry preprocessing, initial state, monitor
// reductions, x-slab halo pack/unpack.  Reference citations relative to /root/reference.
#pragma once
#include "core.cuh"

namespace mflbm {

// IEEE operations that the compiler must not contract into FMAs: the host reference (g++ -O3, x86-64 baseline,
// Makefile:47-49) evaluates these stages with separately rounded multiplies and adds, and the results are
// specified to be bit-exact.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// =====================================================================================================
// geometry preprocessing: src/Geometry_preprocessing.cpp:29-403 on the device
// Work arrays carry G = 10 ghost layers like the reference's temporaries (:38-47).
// =====================================================================================================
struct GeoDims {
    int nx, ny, nz;            // local real extents
    int TX, TY, TZ;            // nx + 20, ...
    int x0, nxg, nyg, nzg;     // global offset of local column 1 and global extents
    int iper, jper, kper;
    __device__ __forceinline__ long long t10(int x, int y, int z) const { return (x + 9) + (long long)TX * ((y + 9) + (long long)TY * (z + 9)); }
};

__device__ __forceinline__ int wrap_or_clamp(int c, int n, int per) {
    if (per) { while (c < 1) c += n; while (c > n) c -= n; return c; }
    return c < 1 ? 1 : (c > n ? n : c);
}

// ghost fill (:59-132): the sequential z, y, x fills of the reference compose to a per-axis clamp (non-periodic)
// or wrap (periodic) of the source coordinate.  Writes the int flags and both float copies (:144-151).
template <typename T>
__global__ void k_geo_fill(GeoDims D, const int8_t* __restrict__ interior_global, int8_t* __restrict__ wt, T* __restrict__ ws1, T* __restrict__ ws2) {
    const int x = -9 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -9 + (int)blockIdx.y, z = -9 + (int)blockIdx.z;
    if (x > D.nx + 10) return;
    const int gx = wrap_or_clamp(x + D.x0 - 1, D.nxg, D.iper);
    const int gy = wrap_or_clamp(y, D.nyg, D.jper), gz = wrap_or_clamp(z, D.nzg, D.kper);
    const int8_t v = interior_global[(gx - 1) + (long long)D.nxg * ((gy - 1) + (long long)D.nyg * (gz - 1))];
    const long long n = D.t10(x, y, z);
    wt[n] = v; ws1[n] = (T)v; ws2[n] = (T)v;
}

// node classification (:154-175): solid with a fluid D3Q18 neighbour -> 2, fluid with a solid neighbour -> -1.
// The reference updates in place; the tests it applies (<= 0, >= 1) do not distinguish updated from original
// values, so an out-of-place pass is identical.
__global__ void k_geo_classify(GeoDims D, const int8_t* __restrict__ wt, int8_t* __restrict__ type) {
    const int x = -8 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -8 + (int)blockIdx.y, z = -8 + (int)blockIdx.z;
    if (x > D.nx + 9) return;
    const long long n = D.t10(x, y, z);
    const int8_t me = wt[n];
    int8_t out = me;
    bool hit = false;
#pragma unroll
    for (int q = 1; q < 19; q++) {
        const int8_t nb = wt[D.t10(x + ex(q), y + ey(q), z + ez(q))];
        hit = hit || (me == 1 ? nb <= 0 : nb >= 1);
    }
    if (hit) out = me == 1 ? 2 : -1;
    type[n] = out;
}

// one 27-point smoothing pass (:187-198), summation order and weights of the reference, no FMA contraction
template <typename T>
__global__ void k_geo_smooth(GeoDims D, const T* __restrict__ src, T* __restrict__ dst) {
    const int x = -8 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -8 + (int)blockIdx.y, z = -8 + (int)blockIdx.z;
    if (x > D.nx + 9) return;
    constexpr int iex[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    constexpr int iey[27] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
    constexpr int iez[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1};
    constexpr T we[4] = {T(8) / T(27), T(2) / T(27), T(1) / T(54), T(1) / T(216)};
    T acc = T(0);
#pragma unroll
    for (int n = 0; n < 27; n++) {
        const int m = iex[n] * iex[n] + iey[n] * iey[n] + iez[n] * iez[n];
        acc = add_rn(acc, mul_rn(src[D.t10(x + iex[n], y + iey[n], z + iez[n])], we[m]));
    }
    dst[D.t10(x, y, z)] = acc;
}
// copy-back of the interior region (:199-205): the outermost layer keeps the raw flags
template <typename T>
__global__ void k_geo_copy_inner(GeoDims D, const T* __restrict__ src, T* __restrict__ dst) {
    const int x = -8 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -8 + (int)blockIdx.y, z = -8 + (int)blockIdx.z;
    if (x > D.nx + 9) return;
    const long long n = D.t10(x, y, z);
    dst[n] = src[n];
}

// ISO8 two-ring gradient tables (src/Geometry_preprocessing.cpp:233-375): every term is w(x+o) - w(x-o); the
// "+o" offsets are listed in source order, seven groups per axis with sizes ISO8_N.
struct Iso8Tables { signed char o[3][34][3]; };

// solid-surface normals at fluid-boundary nodes + export of walls / walls_type (:135-141, :178-184, :229-386)
template <typename T>
__global__ void k_geo_export(GeoDims D, Iso8Tables tab, const int8_t* __restrict__ wt, const int8_t* __restrict__ type,
                             const T* __restrict__ ws2, int* __restrict__ walls, int* __restrict__ walls_type,
                             T* __restrict__ s_nx, T* __restrict__ s_ny, T* __restrict__ s_nz, T eps) {
    const int x = -3 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -3 + (int)blockIdx.y, z = -3 + (int)blockIdx.z;
    if (x > D.nx + 4) return;
    const long long n = D.t10(x, y, z);
    const int NX4 = D.nx + 8, NY4 = D.ny + 8, NX2 = D.nx + 4, NY2 = D.ny + 4;
    const long long c4 = (x + 3) + (long long)NX4 * ((y + 3) + (long long)NY4 * (z + 3));
    const int ty = type[n];
    walls_type[c4] = ty;
    if (x >= -1 && x <= D.nx + 2 && y >= -1 && y <= D.ny + 2 && z >= -1 && z <= D.nz + 2)
        walls[(x + 1) + (long long)NX2 * ((y + 1) + (long long)NY2 * (z + 1))] = wt[n];
    if (ty != -1) return;
    constexpr int ISO8_N[7] = {1, 4, 4, 1, 8, 12, 4};
    constexpr T ISO8[7] = {T(4) / T(45), T(1) / T(21), T(2) / T(105), T(5) / T(504), T(1) / T(315), T(1) / T(630), T(1) / T(5040)};
    T nw[3];
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        T total = T(0);
        int pos = 0;
#pragma unroll
        for (int g = 0; g < 7; g++) {
            T s = T(0);
            for (int m = 0; m < ISO8_N[g]; m++, pos++) {
                const int dx = tab.o[a][pos][0], dy = tab.o[a][pos][1], dz = tab.o[a][pos][2];
                const T plus = ws2[D.t10(x + dx, y + dy, z + dz)], minus = ws2[D.t10(x - dx, y - dy, z - dz)];
                s = (m == 0) ? sub_rn(plus, minus) : sub_rn(add_rn(s, plus), minus);
            }
            total = (g == 0) ? mul_rn(ISO8[g], s) : add_rn(total, mul_rn(ISO8[g], s));
        }
        nw[a] = total;
    }
    const T n2 = add_rn(add_rn(mul_rn(nw[0], nw[0]), mul_rn(nw[1], nw[1])), mul_rn(nw[2], nw[2]));
    const T tmp = div_rn(T(1), add_rn(sqrt_rn(n2), eps));   // :377
    s_nx[c4] = mul_rn(nw[0], tmp); s_ny[c4] = mul_rn(nw[1], tmp); s_nz[c4] = mul_rn(nw[2], tmp);
}

// =====================================================================================================
// layout conversion between the reference's boundary layouts (includes/Idx_gpu.cuh:52-70, ghost width G, x fastest,
// dense) and the internal U grid / permuted PDF slots (core.cuh).  Threads run over the G-ghost box.
// =====================================================================================================
template <typename T, typename S, bool TO_U>
__global__ void k_repitch(const Lattice<T> L, const int G, S* __restrict__ ref, S* __restrict__ ugrid) {
    const int x = 1 - G + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = 1 - G + (int)blockIdx.y, z = 1 - G + (int)blockIdx.z;
    if (x > L.nx + G) return;
    const long long r = (x + G - 1) + (long long)(L.nx + 2 * G) * ((y + G - 1) + (long long)(L.ny + 2 * G) * (z + G - 1));
    if (TO_U) ugrid[L.u(x, y, z)] = ref[r]; else ref[r] = ugrid[L.u(x, y, z)];
}

// walls_type (s4, int32) -> node types (U, int8); padding cells of the U rows are marked solid
template <typename T>
__global__ void k_types_to_u(const Lattice<T> L, const int* __restrict__ wtype_s4, signed char* __restrict__ types) {
    const int px = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int py = (int)blockIdx.y, pz = (int)blockIdx.z;
    if (px >= L.PX) return;
    const int NX4 = L.nx + 8, NY4 = L.ny + 8;
    signed char t = 1;
    if (px < NX4) t = (signed char)wtype_s4[px + (long long)NX4 * (py + (long long)NY4 * pz)];
    types[px + L.PX * (py + L.PY * pz)] = t;
}

// node types (U) -> the reference's walls (s2) and walls_type (s4) arrays (download_geometry)
template <typename T>
__global__ void k_types_from_u(const Lattice<T> L, int* __restrict__ walls_s2, int* __restrict__ wtype_s4) {
    const int x = -3 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = -3 + (int)blockIdx.y, z = -3 + (int)blockIdx.z;
    if (x > L.nx + 4) return;
    const int t = L.types[L.u(x, y, z)];
    wtype_s4[(x + 3) + (long long)(L.nx + 8) * ((y + 3) + (long long)(L.ny + 8) * (z + 3))] = t;
    if (x >= -1 && x <= L.nx + 2 && y >= -1 && y <= L.ny + 2 && z >= -1 && z <= L.nz + 2)
        walls_s2[(x + 1) + (long long)(L.nx + 4) * ((y + 1) + (long long)(L.ny + 4) * (z + 1))] = t > 0 ? 1 : 0;
}

// s4 array <-> values in list order (solid-surface normals live only on the fluid-boundary list)
template <typename T, bool GATHER>
__global__ void k_list_s4(const Lattice<T> L, const int* __restrict__ list, const int count, T* __restrict__ s4, T* __restrict__ compact) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int u = list[t];
    const int px = u % L.PX, r = u / L.PX, py = r % L.PY, pz = r / L.PY;
    const long long c4 = px + (long long)(L.nx + 8) * (py + (long long)(L.ny + 8) * pz);
    if (GATHER) compact[t] = s4[c4]; else s4[c4] = compact[t];
}

// one PDF slot: dense 1-ghost array (reference order) <-> internal storage, in two kernels.
// k_pdf_slot moves every cell between the dense array and SLOT storage (the entry of its site).  The cells of wall links live
// in mailboxes instead (core.cuh): k_pdf_mail, one thread per fluid entry, moves those - the mailbox entry of a link is the
// rank the odd collide kernel computes (wbase + ballot / popc), the dense cell is the solid neighbour's.  On the way in the
// second kernel runs after the first (a mailbox cell also got a harmless copy in the solid site's unused slot storage), on the
// way out it overwrites what the first kernel read from there.  (A first version resolved mailboxes per SITE through
// Lattice::f: every solid site next to a fluid node walked its 32-entry group, 3 ms per slot - the state upload was bound by
// this kernel, not by PCIe.)
template <typename T, bool TO_SLOT>
__global__ void __launch_bounds__(128) k_pdf_slot(const Lattice<T> L, T* __restrict__ dense_s1, const int slot) {
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    if (x > L.nx + 1) return;
    const long long c1 = x + (long long)L.NX1 * (y + (long long)L.NY1 * z);
    T& cell = L.f_raw(slot % 19, slot / 19, L.u(x, y, z));
    if (TO_SLOT) cell = dense_s1[c1]; else dense_s1[c1] = cell;
}
template <typename T, bool TO_SLOT>
__global__ void __launch_bounds__(128) k_pdf_mail(const Lattice<T> L, T* __restrict__ dense_s1, const int slot) {
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x), lane = threadIdx.x & 31;
    const int q = slot % 19, g = slot / 19, o = opc(q);   // the cell (x + e_o, slot q) is the mailbox of link (x, o)
    const bool live = t < L.n_fluid;
    const int u = live ? L.fl_u[t] : 0;
    const int cc = live ? L.cmap[u + L.off(o)] : 0;
    const unsigned walls = __ballot_sync(0xffffffffu, cc < 0);
    if (cc >= 0) return;
    const int entry = L.mb0 + L.wbase[(t >> 5) * 18 + (o - 1)] + __popc(walls & ((1u << lane) - 1u));
    const unsigned Z = fastdiv((unsigned)u, L.dv_sz), r = (unsigned)u - Z * (unsigned)L.sz, Y = fastdiv(r, L.dv_px), X = r - Y * (unsigned)L.PX;
    const int x = (int)X - 3 + ex(o), y = (int)Y - 3 + ey(o), z = (int)Z - 3 + ez(o);   // the solid neighbour, reference coordinates
    const long long c1 = x + (long long)L.NX1 * (y + (long long)L.NY1 * z);
    T& cell = L.at(q, g, entry);
    if (TO_SLOT) cell = dense_s1[c1]; else dense_s1[c1] = cell;
}

// =====================================================================================================
// initial state: src/Init_multiphase.cpp:299-496 (options 1-5), u = v = w = 0, rho = 1
// =====================================================================================================
template <typename T>
__global__ void k_init_phi(const Lattice<T> L, int option, T interface_z0, int nyg, int nzg, int open_z) {
    // phi: [-3..n+4]^3 threads; pattern written on [0..n+1]^3 (:308-356), then phi_inlet for z <= 0 (:358-372)
    const int i = -3 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int j = -3 + (int)blockIdx.y, k = -3 + (int)blockIdx.z;
    if (i > L.nx + 4) return;
    const int gi = i + L.x0 - 1;   // global x
    T ph = T(0);
    bool set = false;
    // slab note: columns 0 and nx+1 of an interior slab are real columns of the neighbour, the pattern formula is
    // the same there; beyond the global range [0..nxg+1] the reference array keeps its calloc'ed 0
    if (gi >= 0 && gi <= L.nx_global + 1 && j >= 0 && j <= nyg + 1 && k >= 0 && k <= nzg + 1) {
        const T x = (T)gi, y = (T)j, z = (T)k;
        set = true;
        if (option == 1) { ph = T(-1); if (z <= interface_z0) ph = T(1); }
        else if (option == 2) { ph = T(1); if (z <= interface_z0) ph = T(-1); }
        else {
            const T cx = option == 5 ? lit<T>(0.5) : lit<T>(0.);
            const T dx = x - (L.nx_global + 1) * cx, dz = z - (nzg + 1) * lit<T>(0.5), dy = y - (nyg + 1) * lit<T>(0.5);
            // pow(a,2) on the host == a*a correctly rounded; sums left to right as :328
            const T d = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dz, dz)), mul_rn(dy, dy));
            const bool inside = d <= mul_rn(interface_z0, interface_z0);
            ph = option == 4 ? (inside ? T(-1) : T(1)) : (inside ? T(1) : T(-1));
        }
    }
    if (open_z && k <= 0) { ph = L.phi_inlet; set = true; }
    if (set) L.phi[L.u(i, j, k)] = ph;
}

template <typename T>
__global__ void k_init_pdf(const Lattice<T> L, int outlet_convective) {
    // equilibrium at rest on [0..n+1]^3 (:393-442): pdf_q = rho_g * w_q  (+ rho_g*w_q*(-1.5*0) == same value)
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int j = (int)blockIdx.y, k = (int)blockIdx.z;
    if (i > L.nx + 1) return;
    const int u = L.u(i, j, k);
    const T ph = L.phi[u];
    const T rho1 = mul_rn(mul_rn(T(1), add_rn(T(1), ph)), lit<T>(0.5));
    const T rho2 = mul_rn(mul_rn(T(1), sub_rn(T(1), ph)), lit<T>(0.5));
#pragma unroll 1
    for (int q = 0; q < 19; q++) {
        L.f(q, 0, u) = mul_rn(rho1, w_equ<T>(q));
        L.f(q, 1, u) = mul_rn(rho2, w_equ<T>(q));
    }
    if (outlet_convective && k == L.nz) {   // :445-494
        const int cb = L.iplane(i, j), plane = L.NX1 * L.NY1;
#pragma unroll
        for (int q = 0; q < 19; q++) {
            L.f_convec[cb + plane * q] = mul_rn(rho1, w_equ<T>(q));
            L.g_convec[cb + plane * q] = mul_rn(rho2, w_equ<T>(q));
        }
        L.phi_convec[cb] = ph;
    }
}

// =====================================================================================================
// monitor: compute_macro_vars (src/Misc.cpp:222-274) + per-slice sums (src/Monitor.cpp:34-80) in one pass, on the
// device (the reference copies the whole state to the host and loops there).  One block per z slice; the fluid
// nodes of a slice are two contiguous ranges of the permuted order - [zstart[k-1], zstart[k]) among the nodes of a slab's
// neighbour-facing columns (empty for a full lattice), [zstart[nz+2+k-1], zstart[nz+2+k]) among the others - so every PDF
// read is a contiguous row.  Warp-shuffle + block reduction, sums in double (the reference sums sequentially in T; see
// DESIGN.md).
// out layout per slice k-1: [0..6] fl1, fl2, pre, mass1, mass2, vol1, vol2, [7] max |u|^2, [8] usq1, [9] usq2, [10] nan flag,
// [11] sum rho where phi < -0.99, [12] their count, [13] sum rho where phi > 0.99, [14] their count, [15] count phi > 0
// (src/Monitor.cpp:376-390, 452-459)
// =====================================================================================================
#define MFLBM_MON_N 16
template <typename T>
__global__ void __launch_bounds__(256) k_monitor(const Lattice<T> L, const int* __restrict__ zstart, double* __restrict__ out) {
    const int k = 1 + blockIdx.x;
    double acc[MFLBM_MON_N];
#pragma unroll
    for (int n = 0; n < MFLBM_MON_N; n++) acc[n] = 0.0;
    for (int seg = 0; seg < 2; seg++) {
    const int t0 = zstart[seg * (L.nz + 2) + k - 1], t1 = zstart[seg * (L.nz + 2) + k];
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const int u = L.fl_u[t];
        T ft[19];
#pragma unroll
        for (int q = 0; q < 19; q++) { const Pair<T> v = L.pairs(q)[t]; ft[q] = v.a + v.b; }
        T rho = ft[0];
#pragma unroll
        for (int q = 1; q < 19; q++) rho = rho + ft[q];
        const T tmp = lit<T>(0.5) * L.lbm_gamma * curvature_at(L, u) * L.c_norm[u];
        const T fx = tmp * L.cn_x[u], fy = tmp * L.cn_y[u], fz = tmp * L.cn_z[u] + L.force_z;
        const T uu = ft[1] - ft[2] + ft[7] - ft[8] + ft[9] - ft[10] + ft[11] - ft[12] + ft[13] - ft[14] - lit<T>(0.5) * fx;
        const T v = ft[3] - ft[4] + ft[7] + ft[8] - ft[9] - ft[10] + ft[15] - ft[16] + ft[17] - ft[18] - lit<T>(0.5) * fy;
        const T w = ft[5] - ft[6] + ft[11] + ft[12] - ft[13] - ft[14] + ft[15] + ft[16] - ft[17] - ft[18] - lit<T>(0.5) * fz;
        const T ph = L.phi[u];
        const T usq = uu * uu + v * v + w * w;
        const double hp = (double)(lit<T>(0.5) * (lit<T>(1.) + ph)), hm = (double)(lit<T>(0.5) * (lit<T>(1.) - ph));
        acc[0] += (double)w * hp; acc[1] += (double)w * hm; acc[2] += (double)rho;
        acc[3] += (double)rho * hp; acc[4] += (double)rho * hm; acc[5] += hp; acc[6] += hm;
        acc[7] = fmax(acc[7], (double)usq);
        if (ph > lit<T>(0.999)) acc[8] += (double)usq; else if (ph < lit<T>(-0.999)) acc[9] += (double)usq;
        if (!(isfinite((double)usq) && isfinite((double)rho) && isfinite((double)ph))) acc[10] = 1.0;
        if (ph < lit<T>(-0.99)) { acc[11] += (double)rho; acc[12] += 1.0; }
        if (ph > lit<T>(0.99)) { acc[13] += (double)rho; acc[14] += 1.0; }
        if (ph > lit<T>(0.)) acc[15] += 1.0;
    }
    }
    __shared__ double sm[MFLBM_MON_N][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < MFLBM_MON_N; n++) {
        double v = acc[n];
        const bool is_max = (n == 7 || n == 10);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmax(v, other) : v + other;
        }
        if (lane == 0) sm[n][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < MFLBM_MON_N) {
        const int n = threadIdx.x;
        const bool is_max = (n == 7 || n == 10);
        double v = sm[n][0];
        for (int wv = 1; wv < (int)(blockDim.x >> 5); wv++) v = is_max ? fmax(v, sm[n][wv]) : v + sm[n][wv];
        out[(long long)(k - 1) * MFLBM_MON_N + n] = v;
    }
}

// monitor_multiphase_steady_phasefield (src/Monitor.cpp:288-312): max |phi - phi_old| over the real fluid nodes, then
// phi_old <- phi there.  phi_old is indexed by fluid entry (the reference's dense copy is only ever read at fluid nodes).
// MODE 0: reduce + update, MODE 1: seed phi_old from the current phi (Init_multiphase.cpp:381-391).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_phi_change(const Lattice<T> L, T* __restrict__ phi_old, double* __restrict__ out) {
    double m = 0.0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L.n_fluid; t += gridDim.x * blockDim.x) {
        const T ph = L.phi[L.fl_u[t]];
        if (MODE == 0) {
            const T d = fabs(ph - phi_old[t]);
            // the reference's `d_phi_max < tmp2` never selects a NaN; report it instead of hiding it
            m = (d != d) ? (double)d : fmax(m, (double)d);
        }
        phi_old[t] = ph;
    }
    if (MODE != 0) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double other = __shfl_xor_sync(0xffffffffu, m, o); m = (other != other || m != m) ? (m != m ? m : other) : fmax(m, other); }
    __shared__ double sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = sm[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = (sm[w] != sm[w] || v != v) ? (v != v ? v : sm[w]) : fmax(v, sm[w]);
        out[blockIdx.x] = v;
    }
}

// compute_macro_vars (src/Misc.cpp:222-274) for field output: rho, u, v, w at the real fluid nodes, written into dense
// 1-ghost arrays (pre-zeroed: the reference multiplies by (1 - wall_indicator) and never touches ghosts).
template <typename T>
__global__ void __launch_bounds__(128) k_macro(const Lattice<T> L, T* __restrict__ rho_s1, T* __restrict__ u_s1, T* __restrict__ v_s1, T* __restrict__ w_s1) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.n_fluid) return;
    const int u = L.fl_u[t];
    T ft[19];
#pragma unroll
    for (int q = 0; q < 19; q++) { const Pair<T> v = L.pairs(q)[t]; ft[q] = v.a + v.b; }
    T rho = ft[0];
#pragma unroll
    for (int q = 1; q < 19; q++) rho = rho + ft[q];
    const T tmp = lit<T>(0.5) * L.lbm_gamma * curvature_at(L, u) * L.c_norm[u];
    const T fx = tmp * L.cn_x[u], fy = tmp * L.cn_y[u], fz = tmp * L.cn_z[u] + L.force_z;
    const int px = u % L.PX, r = u / L.PX, py = r % L.PY, pz = r / L.PY;   // U -> (x+3, y+3, z+3)
    const long long c1 = (px - 3) + (long long)L.NX1 * ((py - 3) + (long long)L.NY1 * (pz - 3));
    if (rho_s1) rho_s1[c1] = rho;
    if (u_s1) u_s1[c1] = ft[1] - ft[2] + ft[7] - ft[8] + ft[9] - ft[10] + ft[11] - ft[12] + ft[13] - ft[14] - lit<T>(0.5) * fx;
    if (v_s1) v_s1[c1] = ft[3] - ft[4] + ft[7] + ft[8] - ft[9] - ft[10] + ft[15] - ft[16] + ft[17] - ft[18] - lit<T>(0.5) * fy;
    if (w_s1) w_s1[c1] = ft[5] - ft[6] + ft[11] + ft[12] - ft[13] - ft[14] + ft[15] + ft[16] - ft[17] - ft[18] - lit<T>(0.5) * fz;
}

// =====================================================================================================
// x-slab halo exchange (new).  Buffers are dense [item][z][y] planes.
// PDF halos carry the ten populations with ex = +1 (slots 1,7,9,11,13) or ex = -1 (2,8,10,12,14) of both
// components over the full (ny+2)(nz+2) plane; the phi halo carries 4 columns over (ny+8)(nz+8).
// =====================================================================================================
__host__ __device__ constexpr int slot_exp(int n) { constexpr int t[5] = {1, 7, 9, 11, 13}; return t[n]; }   // ex = +1
__host__ __device__ constexpr int slot_exm(int n) { constexpr int t[5] = {2, 8, 10, 12, 14}; return t[n]; }  // ex = -1

// ---- halo messages over peer memory (NVLink) --------------------------------------------------------------------
// A slab's pack kernel can write straight into the neighbour's receive buffer (a peer pointer, cudaIpc*) instead of a
// local send buffer that NCCL then moves: the message is one kernel, and its arrival is a sequence number in the
// receiver's memory.  Sender: every block fences its stores system-wide and takes a ticket; the last one publishes
// `seq` in the receiver's flag.  Receiver: the unpack kernel spins on its own flag, then reads the buffer past L1.
// Overwriting a buffer is safe without double buffering: the next message of a kind is only packed after data that
// depends on the receiver having unpacked the previous one has travelled back.
// The sequence number of a (kind, side) message lives in device memory (`seq`, incremented by the push kernel): both
// kernels take static arguments and a whole step pair including its halo messages replays from a CUDA graph.  An
// exchange is symmetric - I push message n of a kind and then wait for the neighbour's message n of that kind - so the
// number I wait for is the one my own push just wrote.
struct HaloSync {
    unsigned* flag;      // push: the peer's flag; unpack: my flag; nullptr = no synchronisation (NCCL path)
    unsigned* counter;   // push only: block tickets (own memory, returns to 0)
    unsigned* seq;       // my message counter for this (kind, side)
    unsigned* error;     // unpack only: set to 1 + kind*2 + side when a message did not arrive within HALO_TIMEOUT_CYCLES
};
constexpr long long HALO_TIMEOUT_CYCLES = 120000000000LL;   // ~60 s at 2 GHz: a dead neighbour must not hang this GPU for ever
__device__ __forceinline__ void halo_publish(const HaloSync& hs) {
    if (!hs.flag) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(hs.counter, 1u) == nblocks - 1u) {
            *hs.counter = 0u;
            const unsigned n = *hs.seq + 1u;
            *hs.seq = n;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(hs.flag) = n;
        }
    }
}
// false: the message did not arrive (now or earlier - the error word is sticky, so every later wait gives up at once and
// the state stops changing through halos); the caller then leaves the lattice untouched and the host sees the error at
// its next synchronisation, download or run call (Solver::check_halo_error)
__device__ __forceinline__ bool halo_await(const HaloSync& hs) {
    if (!hs.flag) return true;
    __shared__ int ok_s;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int ok = 1;
        if (hs.error && *reinterpret_cast<volatile unsigned*>(hs.error) != 0u) ok = 0;
        else {
            const unsigned want = *reinterpret_cast<volatile unsigned*>(hs.seq);
            const long long t0 = clock64();
            // sequence numbers only grow; the signed difference survives wrap-around
            while ((int)(*reinterpret_cast<volatile unsigned*>(hs.flag) - want) < 0) {
                __nanosleep(64);
                if (clock64() - t0 > HALO_TIMEOUT_CYCLES) { if (hs.error) *hs.error = 1u + (unsigned)(hs.flag - (hs.error - 31)); ok = 0; break; }
            }
        }
        __threadfence_system();
        ok_s = ok;
    }
    __syncthreads();
    return ok_s != 0;
}

// The wait on its own, one CTA: in the overlapped schedule (Solver::step_p2p) the unpack kernel follows it on the same lane.
// (An unpack kernel whose every CTA waits fills the SMs' thread slots with spinning CTAs and keeps the interior collide
// launch from starting until the message is there.)
__global__ void k_halo_wait(const HaloSync hs) { halo_await(hs); }
// unpack without a wait of its own: leave the lattice alone once a message has been lost
__device__ __forceinline__ bool halo_ok(const HaloSync& hs) { return !(hs.error && *reinterpret_cast<volatile unsigned*>(hs.error) != 0u); }

// slot entries of one x column over the (ny+2) x (nz+2) plane, built once per geometry: a halo kernel then needs no site-map
// look-up (a scattered 4-byte gather per value before)
template <typename T>
__global__ void __launch_bounds__(128) k_setup_face(const Lattice<T> L, const int col, int* __restrict__ face) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (y < L.NY1) face[y + L.NY1 * z] = Lattice<T>::entry_of(L.cmap[L.u(col, y, z)]);
}

// copy one x column of the ten slots (PLUS ? ex=+1 : ex=-1) between the lattice and a buffer; face = the column's slot entries.
// Slot storage proper: wall links that end in a neighbour-facing ghost column are not mailboxes (Solver::finish_geometry), so
// everything a neighbour slab reads or writes lives in slot storage and is exchanged as is; a solid site of the real boundary
// column can additionally be the far end of a link of THIS slab's nodes, which keep that cell in a mailbox the exchange does
// not touch.  One thread per (population, y, z): the two components of a population sit side by side (core.cuh), one 8/16-byte
// access moves both.  blockIdx.y = n, the population's index among the five.
template <typename T, bool PLUS, bool PACK>
__global__ void __launch_bounds__(128) k_halo_pdf(const Lattice<T> L, T* __restrict__ buf, const int* __restrict__ face, const HaloSync hs) {
    if (!PACK && !(hs.flag ? halo_await(hs) : halo_ok(hs))) return;
    const int plane = L.NY1 * L.NZ1;
    const int p = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (p < plane) {
        const int q = PLUS ? slot_exp(n) : slot_exm(n);
        Pair<T>* cell = L.pairs(q) + face[p];
        T* b0 = buf + (long long)n * plane + p;
        T* b1 = buf + (long long)(5 + n) * plane + p;
        if (PACK) { const Pair<T> v = *cell; *b0 = v.a; *b1 = v.b; }
        else *cell = Pair<T>{__ldcg(b0), __ldcg(b1)};
    }
    if (PACK) halo_publish(hs);
}

// copy 4 phi columns starting at local column `col0` between the lattice and a buffer
template <typename T, bool PACK>
__global__ void k_halo_phi(const Lattice<T> L, T* __restrict__ buf, const int col0, const HaloSync hs) {
    if (!PACK && !(hs.flag ? halo_await(hs) : halo_ok(hs))) return;
    const int y = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;   // 0-based over the 4-ghost extents
    if (y < L.PY) {
        const int plane = L.PY * L.PZ;
#pragma unroll
        for (int n = 0; n < 4; n++) {
            T* cell = L.phi + ((col0 + n + 3) + L.PX * (y + L.PY * z));
            T* b = buf + (long long)n * plane + (y + L.PY * z);
            if (PACK) *b = *cell; else *cell = __ldcg(b);
        }
    }
    if (PACK) halo_publish(hs);
}

}  // namespace mflbm
