// kernels_step.cuh — the per-time-step kernels: AA-pattern collide+stream, colour-gradient chain, open/periodic
// boundary kernels, porous plate.  sm_100a, plain SIMT (nothing on this path is a dense contraction).
// Reference citations are relative to /root/reference/src/main_iteration_GPU.cu unless a file is named.
#pragma once
#include "core.cuh"

namespace mflbm {

// =====================================================================================================
// collide + stream, AA pattern.  ODD: pull f_q from x-e_q (slot q), collide, push f_q* to x+e_q (slot opc(q))
// (:56-388).  EVEN: read local slot opc(q) as f_q, collide, write local slot q (:395-726).
// One thread per node, x fastest => every one of the 38 loads / 38 stores of a warp is one coalesced row segment.
// Solid and ghost storage is live: fluid nodes write into / read from solid neighbours, which is how the reference
// realises (two-step-delayed) bounce-back; this kernel keeps that data flow bit for bit.
// =====================================================================================================
template <typename T, int MRT, bool ODD>
__global__ void __launch_bounds__(128) k_collide(const Lattice<T> L, const int ilo, const int ihi) {
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    if (i > ihi || j > L.ny) return;
    const int c1 = L.i1(i, j, k);
    if (L.solid1[c1]) return;

    T g1[19], g2[19];
    const long long N1 = L.N1;
    const T* __restrict__ p0 = L.pdf;
    if (ODD) {
#pragma unroll
        for (int q = 0; q < 19; q++) {
            const int src = c1 - (ex(q) + L.NX1 * (ey(q) + L.NY1 * ez(q)));
            g1[q] = p0[(long long)q * N1 + src];
            g2[q] = p0[(long long)(q + 19) * N1 + src];
        }
    } else {
#pragma unroll
        for (int q = 0; q < 19; q++) {
            g1[q] = p0[(long long)opc(q) * N1 + c1];
            g2[q] = p0[(long long)(opc(q) + 19) * N1 + c1];
        }
    }
    const int c2 = L.i2(i, j, k);
    const T cnx = L.cn_x[c2], cny = L.cn_y[c2], cnz = L.cn_z[c2];
    const T tmp = lit<T>(0.5) * L.lbm_gamma * L.curv[c1] * L.c_norm[c2];   // :147

    const T phi_loc = collide_node<T, MRT>(L, g1, g2, cnx, cny, cnz, tmp);
    L.phi[L.i4(i, j, k)] = phi_loc;

    T* __restrict__ po = L.pdf;
    if (ODD) {
#pragma unroll
        for (int q = 0; q < 19; q++) {
            const int dst = c1 + (ex(q) + L.NX1 * (ey(q) + L.NY1 * ez(q)));
            po[(long long)opc(q) * N1 + dst] = g1[q];
            po[(long long)(opc(q) + 19) * N1 + dst] = g2[q];
        }
    } else {
#pragma unroll
        for (int q = 0; q < 19; q++) {
            po[(long long)q * N1 + c1] = g1[q];
            po[(long long)(q + 19) * N1 + c1] = g2[q];
        }
    }
}

// =====================================================================================================
// colour-gradient chain (:732-1003).  Boundary-node stages run over compact index lists built once per geometry
// (the reference scans the whole volume for them).
// =====================================================================================================
// decode an s4 linear index into 1-based coordinates
template <typename T>
__device__ __forceinline__ void decode4(const Lattice<T>& L, int n, int& i, int& j, int& k) {
    const int x = n % L.NX4;
    const int r = n / L.NX4;
    i = x - 3; j = r % L.NY4 - 3; k = r / L.NY4 - 3;
}

// phi at solid-boundary nodes <- weighted mean over D3Q18 neighbours that are fluid (:732-755)
template <typename T>
__global__ void k_extrap_phi(const Lattice<T> L, const int* __restrict__ list, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int n = list[t];
    T phi_sum = T(0), weight_sum = T(0);
#pragma unroll
    for (int q = 1; q < 19; q++) {
        const int nb = n + ex(q) + L.NX4 * (ey(q) + L.NY4 * ez(q));
        if (L.walls_type[nb] <= 0) { phi_sum += L.phi[nb] * w_equ<T>(q); weight_sum += w_equ<T>(q); }
    }
    L.phi[n] = phi_sum / weight_sum;
}

// the three 18-point isotropic first-derivative patterns used by :765-791 and :916-996
template <typename T, int A>
__device__ __forceinline__ T iso4(const T* __restrict__ p, const int c, const int sy, const int sz) {
    constexpr int D[3][4][3] = {{{1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1}},
                                {{1, 1, 0}, {-1, 1, 0}, {0, 1, 1}, {0, 1, -1}},
                                {{1, 0, 1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, 1}}};
    constexpr int a0 = (A == 0) ? 1 : 0, a1 = (A == 1) ? 1 : 0, a2 = (A == 2) ? 1 : 0;
    const int oa = a0 + sy * a1 + sz * a2;
    const T axis = p[c + oa] - p[c - oa];
    T s = T(0);
#pragma unroll
    for (int n = 0; n < 4; n++) {
        const int o = D[A][n][0] + sy * D[A][n][1] + sz * D[A][n][2];
        const T plus = p[c + o], minus = p[c - o];
        s = (n == 0) ? (plus - minus) : (s + plus - minus);
    }
    constexpr T ISO4_0 = T(1) / T(6), ISO4_1 = T(1) / T(12);   // includes/Fluid_multiphase.h:34
    return ISO4_0 * axis + ISO4_1 * s;
}

// interface normals from the phase-field gradient (:757-807), over [-1..n+2]^3 exactly (the reference's guard
// over-runs by one; not replicated).  SKIP_SOLID: leave solid nodes untouched (they already hold 0
// and are never read before being rewritten) instead of re-zeroing them every step.
template <typename T, bool SKIP_SOLID>
__global__ void __launch_bounds__(128) k_normals(const Lattice<T> L, const int ilo, const int ihi) {
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = -1 + (int)(blockIdx.y * blockDim.y + threadIdx.y);
    const int k = -1 + (int)blockIdx.z;
    if (i > ihi || j > L.ny + 2) return;
    const int c2 = L.i2(i, j, k);
    const bool solid = L.walls[c2] == 1;
    if (SKIP_SOLID && solid) return;
    const int c4 = L.i4(i, j, k);
    T gx = iso4<T, 0>(L.phi, c4, L.NX4, L.NX4 * L.NY4);
    T gy = iso4<T, 1>(L.phi, c4, L.NX4, L.NX4 * L.NY4);
    T gz = iso4<T, 2>(L.phi, c4, L.NX4, L.NX4 * L.NY4);
    T nrm = sqrt(gx * gx + gy * gy + gz * gz);
    if (nrm < lit<T>(1e-6) || solid) { gx = T(0); gy = T(0); gz = T(0); nrm = T(0); }
    else { gx = gx / nrm; gy = gy / nrm; gz = gz / nrm; }
    L.cn_x[c2] = gx; L.cn_y[c2] = gy; L.cn_z[c2] = gz; L.c_norm[c2] = nrm;
}

// geometrical wetting model on fluid-boundary nodes: rotate cn so that n_w . cn = cos(theta), <= 4 secant
// iterations (:809-878)
template <typename T>
__global__ void k_alter(const Lattice<T> L, const int* __restrict__ list, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int n4 = list[t];
    int i, j, k;
    decode4(L, n4, i, j, k);
    const int c2 = L.i2(i, j, k);
    const T lambda = lit<T>(0.5), local_eps = lit<T>(1e-6), ct = L.cos_theta;
    if (!(L.c_norm[c2] > local_eps)) return;
    const T nwx = L.s_nx[n4], nwy = L.s_ny[n4], nwz = L.s_nz[n4];
    T vcx0 = L.cn_x[c2], vcy0 = L.cn_y[c2], vcz0 = L.cn_z[c2];
    T vcx1 = vcx0 - lambda * (vcx0 + nwx), vcy1 = vcy0 - lambda * (vcy0 + nwy), vcz1 = vcz0 - lambda * (vcz0 + nwz);
    T vcx2, vcy2, vcz2, err0, err1, err2, tmp;
    err0 = (nwx * vcx0 + nwy * vcy0 + nwz * vcz0) - ct;
    if ((fabs(vcx0 + nwx) + fabs(vcy0 + nwy) + fabs(vcz0 + nwz) > local_eps ||
         fabs(vcx0 - nwx) + fabs(vcy0 - nwy) + fabs(vcz0 - nwz) > local_eps) && err0 > local_eps) {
        err1 = (nwx * vcx1 + nwy * vcy1 + nwz * vcz1) - sqrt(vcx1 * vcx1 + vcy1 * vcy1 + vcz1 * vcz1) * ct;
        tmp = lit<T>(1.) / (err1 - err0);
        vcx2 = tmp * (vcx0 * err1 - vcx1 * err0); vcy2 = tmp * (vcy0 * err1 - vcy1 * err0); vcz2 = tmp * (vcz0 * err1 - vcz1 * err0);
        err2 = (nwx * vcx2 + nwy * vcy2 + nwz * vcz2) - sqrt(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2) * ct;
        if (err2 > local_eps) {
            for (int it = 2; it <= 4; it++) {
                vcx0 = vcx1; vcy0 = vcy1; vcz0 = vcz1;
                vcx1 = vcx2; vcy1 = vcy2; vcz1 = vcz2;
                err0 = (nwx * vcx0 + nwy * vcy0 + nwz * vcz0) - sqrt(vcx0 * vcx0 + vcy0 * vcy0 + vcz0 * vcz0) * ct;
                err1 = (nwx * vcx1 + nwy * vcy1 + nwz * vcz1) - sqrt(vcx1 * vcx1 + vcy1 * vcy1 + vcz1 * vcz1) * ct;
                tmp = lit<T>(1.) / (err1 - err0);
                if (isinf(tmp)) break;
                vcx2 = tmp * (vcx0 * err1 - vcx1 * err0); vcy2 = tmp * (vcy0 * err1 - vcy1 * err0); vcz2 = tmp * (vcz0 * err1 - vcz1 * err0);
                err2 = (nwx * vcx2 + nwy * vcy2 + nwz * vcz2) - sqrt(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2) * ct;
            }
        }
        tmp = lit<T>(1.) / ((lit<T>(1e-30)) + sqrt(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2));
        L.cn_x[c2] = vcx2 * tmp; L.cn_y[c2] = vcy2 * tmp; L.cn_z[c2] = vcz2 * tmp;
    }
}

// cn at solid-boundary nodes <- weighted mean of the fluid neighbours' cn (:880-906)
template <typename T>
__global__ void k_extrap_cn(const Lattice<T> L, const int* __restrict__ list, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int n4 = list[t];
    int i, j, k;
    decode4(L, n4, i, j, k);
    const int c2 = L.i2(i, j, k);
    T sx = T(0), sy = T(0), sz = T(0), wsum = T(0);
#pragma unroll
    for (int q = 1; q < 19; q++) {
        if (L.walls_type[n4 + ex(q) + L.NX4 * (ey(q) + L.NY4 * ez(q))] <= 0) {
            const int nb = c2 + ex(q) + L.NX2 * (ey(q) + L.NY2 * ez(q));
            sx += L.cn_x[nb] * w_equ<T>(q); sy += L.cn_y[nb] * w_equ<T>(q); sz += L.cn_z[nb] * w_equ<T>(q); wsum += w_equ<T>(q);
        }
    }
    L.cn_x[c2] = sx / wsum; L.cn_y[c2] = sy / wsum; L.cn_z[c2] = sz / wsum;
}

// interface curvature from nine derivatives of cn (:908-1003).  FLUID_ONLY: only where the value is consumed
// (collide and the monitor read curv at fluid nodes only); the dense variant reproduces the reference array.
// cn*cn replaces the reference's pow(cn, 2) (<= 2 ulp apart in double).
template <typename T, bool FLUID_ONLY>
__global__ void __launch_bounds__(128) k_curvature(const Lattice<T> L, const int ilo, const int ihi) {
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    if (i > ihi || j > L.ny) return;
    const int c1 = L.i1(i, j, k);
    if (FLUID_ONLY && L.solid1[c1]) return;
    const int c2 = L.i2(i, j, k);
    const int sy = L.NX2, sz = L.NX2 * L.NY2;
    const T kxx = iso4<T, 0>(L.cn_x, c2, sy, sz), kyy = iso4<T, 1>(L.cn_y, c2, sy, sz), kzz = iso4<T, 2>(L.cn_z, c2, sy, sz);
    const T kxy = iso4<T, 1>(L.cn_x, c2, sy, sz), kxz = iso4<T, 2>(L.cn_x, c2, sy, sz);
    const T kyx = iso4<T, 0>(L.cn_y, c2, sy, sz), kyz = iso4<T, 2>(L.cn_y, c2, sy, sz);
    const T kzx = iso4<T, 0>(L.cn_z, c2, sy, sz), kzy = iso4<T, 1>(L.cn_z, c2, sy, sz);
    const T cx = L.cn_x[c2], cy = L.cn_y[c2], cz = L.cn_z[c2];
    L.curv[c1] = (cx * cx - lit<T>(1.)) * kxx + (cy * cy - lit<T>(1.)) * kyy + (cz * cz - lit<T>(1.)) * kzz +
                 cx * cy * (kxy + kyx) + cx * cz * (kxz + kzx) + cy * cz * (kzy + kyz);
}

// =====================================================================================================
// inlet / outlet kernels (:1009-1524).  One thread per (i,j) of the z-plane.  AFTER = false: the "before odd"
// variant launched after an even step (fills the ghost plane the odd pull will read); AFTER = true: launched
// after an odd step (patches the swapped slots at k = 1 / k = nz).
// Unknown directions at the inlet: QIN = {5,11,12,15,16} (ez = +1); at the outlet their opposites.
// =====================================================================================================
__host__ __device__ constexpr int qin(int n) { constexpr int t[5] = {5, 11, 12, 15, 16}; return t[n]; }

template <typename T> __device__ __forceinline__ T blend(T newv, T oldv, int wi) { return newv * (1 - wi) + oldv * wi; }

template <typename T>
__device__ __forceinline__ void inlet_phi(const Lattice<T>& L, int i, int j, int wi) {  // :1021-1024
    const int c = L.i4(i, j, 0), sz = L.NX4 * L.NY4;
    const T v = L.phi_inlet * (1 - wi) + L.phi[c] * wi;
    L.phi[c] = v; L.phi[c - sz] = v; L.phi[c - 2 * sz] = v; L.phi[c - 3 * sz] = v;
}

#define MFLBM_PLANE_IJ()                                          \
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;    \
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;      \
    if (i > ihi || j > L.ny) return;

template <typename T, bool AFTER>
__global__ void k_inlet_velocity(const Lattice<T> L, const int ilo, const int ihi) {  // :1009-1077
    MFLBM_PLANE_IJ();
    const int wi = L.walls[L.i2(i, j, 1)];
    inlet_phi(L, i, j, wi);
    T tmp2 = L.W_in[L.i1(i, j, 0)] * L.relaxation;
    const T tmp1 = tmp2 * L.sa_inject;
    tmp2 = tmp2 - tmp1;
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const T t = g == 0 ? tmp1 : tmp2;
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int q = qin(n), o = opc(q);
            const T wgt = n == 0 ? T(1) / T(18) : T(1) / T(36);
            T* gh = L.slot(q, g) + L.i1(i - ex(q), j - ey(q), 0);   // ghost plane, slot q at x - e_q
            T* in = L.slot(o, g) + L.i1(i, j, 1);                   // first real plane, slot opc(q)
            if (!AFTER) *gh = blend(*in + lit<T>(6.0) * wgt * t, *gh, wi);
            else *in = blend(*gh + lit<T>(6.0) * wgt * t, *in, wi);
        }
    }
}

// f_q arriving at (i,j,k) for the nine in-plane directions; v is indexed by direction
template <typename T, bool AFTER>
__device__ __forceinline__ void zh_inplane(const Lattice<T>& L, int i, int j, int k, int g, T (&v)[11]) {
    constexpr int QS[9] = {0, 1, 2, 3, 4, 7, 8, 9, 10};
#pragma unroll
    for (int n = 0; n < 9; n++) {
        const int q = QS[n];
        v[q] = AFTER ? L.slot(opc(q), g)[L.i1(i, j, k)] : L.slot(q, g)[L.i1(i - ex(q), j - ey(q), k)];
    }
}
// operand orders of the reference's sums: the "after odd" kernels list local slots 0,2,1,4,3,8,7,10,9
template <typename T, bool AFTER> __device__ __forceinline__ T zh_sum9(const T (&v)[11]) {
    return AFTER ? (v[0] + v[1] + v[2] + v[3] + v[4] + v[9] + v[10] + v[7] + v[8])
                 : (v[0] + v[1] + v[2] + v[3] + v[4] + v[7] + v[8] + v[9] + v[10]);
}
template <typename T, bool AFTER> __device__ __forceinline__ T zh_tnx(const T (&v)[11]) {
    return AFTER ? lit<T>(0.5) * (v[1] + v[9] + v[7] - (v[2] + v[10] + v[8]))
                 : lit<T>(0.5) * (v[1] + v[7] + v[9] - (v[2] + v[8] + v[10]));
}

template <typename T, bool AFTER>
__global__ void k_inlet_pressure(const Lattice<T> L, const int ilo, const int ihi) {  // :1085-1241
    MFLBM_PLANE_IJ();
    const int wi = L.walls[L.i2(i, j, 1)];
    inlet_phi(L, i, j, wi);
    T rho2 = L.rho_in;
    const T rho1 = L.rho_in * L.sa_inject;
    rho2 = rho2 - rho1;
#pragma unroll
    for (int g = 0; g < 2; g++) {
        T v[11], out[5];
        zh_inplane<T, AFTER>(L, i, j, 1, g, v);
#pragma unroll
        for (int n = 0; n < 5; n++) {   // known outgoing f_opc(q): pulled from k = 2, or the swapped local slot q
            const int q = qin(n), o = opc(q);
            out[n] = AFTER ? L.slot(q, g)[L.i1(i, j, 1)] : L.slot(o, g)[L.i1(i - ex(o), j - ey(o), 2)];
        }
        const T tr = g == 0 ? rho1 : rho2;
        const T t = (tr - (zh_sum9<T, AFTER>(v) + lit<T>(2.) * (out[0] + out[1] + out[2] + out[3] + out[4]))) * L.relaxation;
        const T tnx = zh_tnx<T, AFTER>(v);
        const T tny = AFTER ? lit<T>(0.5) * (v[3] + v[8] + v[7] - (v[4] + v[9] + v[10]))
                            : lit<T>(0.5) * (v[3] + v[7] + v[8] - (v[4] + v[10] + v[9]));
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int q = qin(n), o = opc(q);
            T val;
            if (n == 0) val = out[n] + lit<T>(0.333333333333333333) * t;
            else {
                const T corr = (q == 11) ? -tnx : (q == 12) ? tnx : (q == 15) ? -tny : tny;
                val = out[n] + lit<T>(0.166666666666666667) * t + corr;
            }
            T* dst = AFTER ? L.slot(o, g) + L.i1(i, j, 1) : L.slot(q, g) + L.i1(i - ex(q), j - ey(q), 0);
            *dst = blend(val, *dst, wi);
        }
    }
}

template <typename T, bool AFTER>
__global__ void k_outlet_convective(const Lattice<T> L, const int ilo, const int ihi) {  // :1246-1357
    MFLBM_PLANE_IJ();
    const int nz = L.nz;
    const T u_convec = L.uin_avg;
    const T temp = lit<T>(1.) / (lit<T>(1.) + u_convec);
    const int wi = L.walls[L.i2(i, j, nz)];
    const int c = L.i4(i, j, nz), sz = L.NX4 * L.NY4, cb = L.i1(i, j, 0);
    const T ph = ((L.phi_convec[cb] + u_convec * L.phi[c]) * temp) * (1 - wi) + L.phi[c + sz] * wi;
    L.phi[c + sz] = ph; L.phi_convec[cb] = ph; L.phi[c + 2 * sz] = ph; L.phi[c + 3 * sz] = ph; L.phi[c + 4 * sz] = ph;
    const int plane = L.NX1 * L.NY1;
#pragma unroll
    for (int g = 0; g < 2; g++) {
        T* buf = g == 0 ? L.f_convec : L.g_convec;
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int o = opc(qin(n));   // unknown incoming direction at the outlet (ez = -1)
            T* dst; const T* inner;
            if (!AFTER) { dst = L.slot(o, g) + L.i1(i - ex(o), j - ey(o), nz + 1); inner = L.slot(o, g) + L.i1(i - ex(o), j - ey(o), nz); }
            else { dst = L.slot(opc(o), g) + L.i1(i, j, nz); inner = L.slot(opc(o), g) + L.i1(i, j, nz - 1); }
            const T val = ((buf[cb + plane * o] + u_convec * *inner) * temp) * (1 - wi) + *dst * wi;
            *dst = val;
            buf[cb + plane * o] = val;
        }
    }
}

template <typename T, bool AFTER>
__global__ void k_outlet_pressure(const Lattice<T> L, const int ilo, const int ihi) {  // :1363-1524
    MFLBM_PLANE_IJ();
    const int nz = L.nz;
    const int wi = L.walls[L.i2(i, j, nz)];
    const int c = L.i4(i, j, nz), sz = L.NX4 * L.NY4;
    const T phn = L.phi[c];
    L.phi[c + sz] = phn; L.phi[c + 2 * sz] = phn; L.phi[c + 3 * sz] = phn; L.phi[c + 4 * sz] = phn;
    T v0[11], v1[11], o0[5], o1[5];
    zh_inplane<T, AFTER>(L, i, j, nz, 0, v0);
    zh_inplane<T, AFTER>(L, i, j, nz, 1, v1);
#pragma unroll
    for (int n = 0; n < 5; n++) {   // known outgoing f_q (ez = +1): pulled from k = nz-1, or the swapped local slot opc(q)
        const int q = qin(n);
        o0[n] = AFTER ? L.slot(opc(q), 0)[L.i1(i, j, nz)] : L.slot(q, 0)[L.i1(i - ex(q), j - ey(q), nz - 1)];
        o1[n] = AFTER ? L.slot(opc(q), 1)[L.i1(i, j, nz)] : L.slot(q, 1)[L.i1(i - ex(q), j - ey(q), nz - 1)];
    }
    T tmp1;
    if (!AFTER)
        tmp1 = (v0[0] + v0[1] + v0[2] + v0[3] + v0[4] + v0[7] + v0[8] + v0[9] + v0[10] + lit<T>(2.) * (o0[0] + o0[1] + o0[2] + o0[3] + o0[4]) +
                v1[0] + v1[1] + v1[2] + v1[3] + v1[4] + v1[7] + v1[8] + v1[9] + v1[10] + lit<T>(2.) * (o1[0] + o1[1] + o1[2] + o1[3] + o1[4])) - L.rho_out;
    else
        tmp1 = (v0[0] + v0[1] + v0[2] + v0[3] + v0[4] + v0[9] + v0[10] + v0[7] + v0[8] + lit<T>(2.) * (o0[0] + o0[1] + o0[2] + o0[3] + o0[4]) +
                v1[0] + v1[1] + v1[2] + v1[3] + v1[4] + v1[9] + v1[10] + v1[7] + v1[8] + lit<T>(2.) * (o1[0] + o1[1] + o1[2] + o1[3] + o1[4])) - L.rho_out;
    const T tmp2 = tmp1 * lit<T>(0.5) * (lit<T>(1.) - phn);
    tmp1 = tmp1 - tmp2;
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const T(&vv)[11] = g == 0 ? v0 : v1;
        const T(&oo)[5] = g == 0 ? o0 : o1;
        const T t = g == 0 ? tmp1 : tmp2;
        const T tnx = zh_tnx<T, AFTER>(vv);
        const T tny = lit<T>(0.5) * (vv[3] + vv[7] + vv[8] - (vv[4] + vv[10] + vv[9]));
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int q = qin(n), o = opc(q);
            T val;
            if (n == 0) val = oo[n] - lit<T>(0.333333333333333333) * t;
            else {
                const T corr = (o == 13) ? -tnx : (o == 14) ? tnx : (o == 17) ? -tny : tny;
                val = oo[n] - lit<T>(0.166666666666666667) * t + corr;
            }
            T* dst = AFTER ? L.slot(q, g) + L.i1(i, j, nz) : L.slot(o, g) + L.i1(i - ex(o), j - ey(o), nz + 1);
            *dst = blend(val, *dst, wi);
        }
    }
}

// =====================================================================================================
// periodic kernels (:1529-1733)
// =====================================================================================================
// AXIS 1 = y (threads over (i,k)), 2 = z (threads over (i,j)).  Even: real boundary layer -> opposite ghost layer
// for the 5 directions leaving through that face; odd: the reverse copies.
template <typename T, int AXIS, bool ODD>
__global__ void k_periodic_pdf(const Lattice<T> L, const int ilo, const int ihi) {
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    const int m = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int n = AXIS == 1 ? L.ny : L.nz, mlim = AXIS == 1 ? L.nz : L.ny;
    if (i > ihi || m > mlim) return;
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int s = AXIS == 1 ? ey(q) : ez(q);
            if (s == 0) continue;
            const int a = s < 0 ? 1 : n, b = s < 0 ? n + 1 : 0;
            const int src = ODD ? b : a, dst = ODD ? a : b;
            T* p = L.slot(q, g);
            if (AXIS == 1) p[L.i1(i, dst, m)] = p[L.i1(i, src, m)];
            else p[L.i1(i, m, dst)] = p[L.i1(i, m, src)];
        }
    }
}

template <typename T, bool ODD>
__global__ void k_periodic_pdf_edges(const Lattice<T> L, const int ilo, const int ihi) {  // :1593-1608, :1674-1689
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ihi) return;
    constexpr int QE[4] = {18, 16, 17, 15};
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int q = QE[n];
            const int ja = ey(q) < 0 ? 1 : L.ny, jb = ey(q) < 0 ? L.ny + 1 : 0;
            const int ka = ez(q) < 0 ? 1 : L.nz, kb = ez(q) < 0 ? L.nz + 1 : 0;
            T* p = L.slot(q, g);
            if (!ODD) p[L.i1(i, jb, kb)] = p[L.i1(i, ja, ka)];
            else p[L.i1(i, ja, ka)] = p[L.i1(i, jb, kb)];
        }
    }
}

// WHICH 1 = y faces (threads over (i,k)), 2 = z faces (threads over (i,j)), 3 = the four y-z edges (threads over i)
template <typename T, int WHICH>
__global__ void k_periodic_phi(const Lattice<T> L, const int ilo, const int ihi) {  // :1691-1733, overlap 4
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    const int m = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ihi) return;
    const int ny = L.ny, nz = L.nz;
    constexpr int ov = 4;
    if (WHICH == 2) {
        if (m > ny) return;
        for (int k = 1; k <= ov; k++) {
            L.phi[L.i4(i, m, k + nz)] = L.phi[L.i4(i, m, k)];
            L.phi[L.i4(i, m, k - ov)] = L.phi[L.i4(i, m, nz + k - ov)];
        }
    } else if (WHICH == 1) {
        if (m > nz) return;
        for (int j = 1; j <= ov; j++) {
            L.phi[L.i4(i, j + ny, m)] = L.phi[L.i4(i, j, m)];
            L.phi[L.i4(i, j - ov, m)] = L.phi[L.i4(i, ny + j - ov, m)];
        }
    } else {
        if (m > 1) return;
        for (int k = 1; k <= ov; k++)
            for (int j = 1; j <= ov; j++) {
                L.phi[L.i4(i, j - ov, k - ov)] = L.phi[L.i4(i, ny + j - ov, nz + k - ov)];
                L.phi[L.i4(i, j + ny, k - ov)] = L.phi[L.i4(i, j, nz + k - ov)];
                L.phi[L.i4(i, j + ny, k + nz)] = L.phi[L.i4(i, j, k)];
                L.phi[L.i4(i, j - ov, k + nz)] = L.phi[L.i4(i, ny + j - ov, k)];
            }
    }
}

// =====================================================================================================
// porous plate at z = Z_porous_plate (:1744-1882): bounce-back across the plane for the blocked component,
// pass-through copies for the other
// =====================================================================================================
template <typename T, bool AFTER>
__global__ void k_porous_plate(const Lattice<T> L, const int ilo, const int ihi) {
    MFLBM_PLANE_IJ();
    const int zp = L.Z_porous_plate, cmd = L.porous_plate_cmd;
    if (!(zp >= 1 && zp <= L.nz) || (cmd != 1 && cmd != 2)) return;
    const int gb = cmd == 1 ? 0 : 1, gp = 1 - gb;
#pragma unroll
    for (int n = 0; n < 5; n++) {
        const int q = qin(n), o = opc(q);
        T* sq = L.slot(q, gb); T* so = L.slot(o, gb);
        if (!AFTER) {   // :1758-1768
            so[L.i1(i - ex(o), j - ey(o), zp)] = sq[L.i1(i, j, zp - 1)];
            sq[L.i1(i - ex(q), j - ey(q), zp)] = so[L.i1(i, j, zp + 1)];
        } else {        // :1828-1838
            sq[L.i1(i, j, zp - 1)] = so[L.i1(i - ex(o), j - ey(o), zp)];
            so[L.i1(i, j, zp + 1)] = sq[L.i1(i - ex(q), j - ey(q), zp)];
        }
        T* pq = L.slot(q, gp); T* po = L.slot(o, gp);
        if (!AFTER) {   // :1770-1780
            po[L.i1(i, j, zp)] = po[L.i1(i, j, zp + 1)];
            pq[L.i1(i, j, zp)] = pq[L.i1(i, j, zp - 1)];
        } else {        // :1840-1850
            pq[L.i1(i, j, zp - 1)] = pq[L.i1(i, j, zp)];
            po[L.i1(i, j, zp + 1)] = po[L.i1(i, j, zp)];
        }
    }
}

}  // namespace mflbm
