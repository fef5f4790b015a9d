// kernels_step.cuh — the per-time-step kernels: AA-pattern collide+stream, colour-gradient chain, open/periodic
// boundary kernels, porous plate.  sm_100a, plain SIMT (nothing on this path is a dense contraction).
// Reference citations are relative to /root/reference/src/main_iteration_GPU.cu unless a file is named.
// Index spaces (U grid, permuted PDF slots) are described in core.cuh.
#pragma once
#include "core.cuh"

namespace mflbm {

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

// interface curvature from nine derivatives of cn (:908-1003) at U index c2.
// cn*cn replaces the reference's pow(cn, 2) (<= 2 ulp apart in double).
template <typename T>
__device__ __forceinline__ T curvature_at(const Lattice<T>& L, const int c2) {
    const int sy = L.sy, sz = L.sz;
    const T kxx = iso4<T, 0>(L.cn_x, c2, sy, sz), kyy = iso4<T, 1>(L.cn_y, c2, sy, sz), kzz = iso4<T, 2>(L.cn_z, c2, sy, sz);
    const T kxy = iso4<T, 1>(L.cn_x, c2, sy, sz), kxz = iso4<T, 2>(L.cn_x, c2, sy, sz);
    const T kyx = iso4<T, 0>(L.cn_y, c2, sy, sz), kyz = iso4<T, 2>(L.cn_y, c2, sy, sz);
    const T kzx = iso4<T, 0>(L.cn_z, c2, sy, sz), kzy = iso4<T, 1>(L.cn_z, c2, sy, sz);
    const T cx = L.cn_x[c2], cy = L.cn_y[c2], cz = L.cn_z[c2];
    return (cx * cx - lit<T>(1.)) * kxx + (cy * cy - lit<T>(1.)) * kyy + (cz * cz - lit<T>(1.)) * kzz +
           cx * cy * (kxy + kyx) + cx * cz * (kxz + kzx) + cy * cz * (kzy + kyz);
}

// (the AA collide + stream kernels are in kernels_collide.cuh)

// =====================================================================================================
// colour-gradient chain (:732-1003).  Every stage runs over a compact site list built once per geometry (the
// reference scans the whole volume five times per step); lists are in z,y,x order so that consecutive threads
// touch consecutive x.
// =====================================================================================================

// phi at solid-boundary nodes <- weighted mean over the D3Q18 neighbours that are fluid (:732-755).
// mask bit (q-1) = neighbour q has walls_type <= 0 (precomputed: replaces 18 flag loads per node and step).
template <typename T>
__global__ void k_extrap_phi(const Lattice<T> L, const int* __restrict__ list, const int* __restrict__ mask, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int n = list[t];
    const int m = mask[t];
    T phi_sum = T(0), weight_sum = T(0);
#pragma unroll
    for (int q = 1; q < 19; q++) {
        if (m & (1 << (q - 1))) { phi_sum += L.phi[n + L.off(q)] * w_equ<T>(q); weight_sum += w_equ<T>(q); }
    }
    L.phi[n] = phi_sum / weight_sum;
}

// interface normals from the phase-field gradient (:757-807) at the non-solid sites of [-1..n+2]^3 (the reference's
// guard over-runs by one; not replicated).  Solid sites hold 0 from k_zero_solid_normals and are
// never written again, which is what the reference stores there every step.
//
// live[t] != 0 says that the four outputs of list entry t may be non-zero in memory.  Away from interfaces the result is
// zero step after step (the reference rewrites those zeros every step, 4 stores per site); here an entry that was zero
// and stays zero stores nothing.  Arrays that came from outside (upload_state) start with live = 1 everywhere.
// near[u'] = 1 is raised at the 18 neighbours of every site that gets a non-zero normal: k_extrap_cn reads that one byte
// instead of probing the c_norm of 18 scattered neighbours, and clears it.
template <typename T>
__device__ __forceinline__ void raise_near(const Lattice<T>& L, unsigned char* __restrict__ near, const int u) {
#pragma unroll
    for (int q = 1; q < 19; q++) near[u + L.off(q)] = 1;
}

template <typename T>
__global__ void __launch_bounds__(128, 16) k_normals(const Lattice<T> L, const int* __restrict__ list, unsigned char* __restrict__ live, unsigned char* __restrict__ near, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int u = list[t];
    T gx = iso4<T, 0>(L.phi, u, L.sy, L.sz);
    T gy = iso4<T, 1>(L.phi, u, L.sy, L.sz);
    T gz = iso4<T, 2>(L.phi, u, L.sy, L.sz);
    T nrm = sqrt(gx * gx + gy * gy + gz * gz);
    if (nrm < lit<T>(1e-6)) {
        if (!live[t]) return;
        live[t] = 0;
        gx = T(0); gy = T(0); gz = T(0); nrm = T(0);
    } else {
        gx = gx / nrm; gy = gy / nrm; gz = gz / nrm;
        live[t] = 1;
        raise_near(L, near, u);
    }
    L.cn_x[u] = gx; L.cn_y[u] = gy; L.cn_z[u] = gz; L.c_norm[u] = nrm;
}

// State of the solid sites of [-1..n+2]^3 after one pass of the reference chain: normalDirectionsOfInterfaces zeroes
// cn_* and c_norm at every solid site (:795-800), then extrapolateNormalToSolid overwrites cn_* at the solid-boundary
// sites of [0..n+1]^3 (:880-906).  Run once after arrays were uploaded, so that k_normals never has to touch solids:
// zero everything the chain leaves at zero, keep the extrapolated normals (the next collide's curvature reads them).
template <typename T>
__global__ void k_zero_solid_normals(const Lattice<T> L) {
    const int i = -1 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int j = -1 + (int)blockIdx.y, k = -1 + (int)blockIdx.z;
    if (i > L.nx + 2) return;
    const int u = L.u(i, j, k);
    const int ty = L.types[u];
    if (ty <= 0) return;
    L.c_norm[u] = T(0);
    const bool extrapolated = ty == 2 && i >= 0 && i <= L.nx + 1 && j >= 0 && j <= L.ny + 1 && k >= 0 && k <= L.nz + 1;
    if (!extrapolated) { L.cn_x[u] = T(0); L.cn_y[u] = T(0); L.cn_z[u] = T(0); }
}

// geometrical wetting model on fluid-boundary nodes: rotate cn so that n_w . cn = cos(theta), <= 4 secant
// iterations (:809-878).
// on values in registers:
template <typename T>
__device__ __forceinline__ void alter_values(const T ct, const T nwx, const T nwy, const T nwz, T& cx, T& cy, T& cz) {
    const T lambda = lit<T>(0.5), local_eps = lit<T>(1e-6);
    T vcx0 = cx, vcy0 = cy, vcz0 = cz;
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
        cx = vcx2 * tmp; cy = vcy2 * tmp; cz = vcz2 * tmp;
    }
}


// list version: fluid-boundary sites of the whole U grid in brick order (Solver::finish_geometry), snx/sny/snz their solid-surface
// normals in list order; the kernel's range is [-1 .. n+2]^3 (:814)
template <typename T>
__global__ void k_alter(const Lattice<T> L, const int* __restrict__ list, const T* __restrict__ snx, const T* __restrict__ sny,
                        const T* __restrict__ snz, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int c2 = list[t];
    const int Z = c2 / L.sz, r = c2 - Z * L.sz, Y = r / L.PX, X = r - Y * L.PX;
    if (X < 2 || X > L.nx + 5 || Y < 2 || Y > L.ny + 5 || Z < 2 || Z > L.nz + 5) return;
    if (!(L.c_norm[c2] > lit<T>(1e-6))) return;
    T cx = L.cn_x[c2], cy = L.cn_y[c2], cz = L.cn_z[c2];
    alter_values<T>(L.cos_theta, snx[t], sny[t], snz[t], cx, cy, cz);
    L.cn_x[c2] = cx; L.cn_y[c2] = cy; L.cn_z[c2] = cz;
}

// cn at solid-boundary nodes <- weighted mean of the fluid neighbours' cn (:880-906); mask as in k_extrap_phi.
// A neighbour with c_norm == 0 has cn == 0 (k_normals), so where every contributing neighbour is interface-free the mean
// is exactly +0: 18 c_norm loads decide that, and an entry that was zero and stays zero stores nothing (live, as above).
// near[c2] (raised by the normals kernel of this chain, cleared here) says whether any of them has c_norm != 0.
template <typename T>
__global__ void k_extrap_cn(const Lattice<T> L, const int* __restrict__ list, const int* __restrict__ mask, unsigned char* __restrict__ live,
                            unsigned char* __restrict__ near, const int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int c2 = list[t];
    const bool any = near[c2] != 0;
    if (any) near[c2] = 0;
    if (!any) {
        if (!live[t]) return;
        live[t] = 0;
        L.cn_x[c2] = T(0); L.cn_y[c2] = T(0); L.cn_z[c2] = T(0);
        return;
    }
    const int m = mask[t];
    T sx = T(0), sy = T(0), sz = T(0), wsum = T(0);
#pragma unroll
    for (int q = 1; q < 19; q++) {
        if (m & (1 << (q - 1))) {
            const int nb = c2 + L.off(q);
            sx += L.cn_x[nb] * w_equ<T>(q); sy += L.cn_y[nb] * w_equ<T>(q); sz += L.cn_z[nb] * w_equ<T>(q); wsum += w_equ<T>(q);
        }
    }
    live[t] = 1;
    L.cn_x[c2] = sx / wsum; L.cn_y[c2] = sy / wsum; L.cn_z[c2] = sz / wsum;
}

// boundary array: the reference's dense curv over [1..n]^3 in its own 1-ghost layout (download_state only)
template <typename T>
__global__ void __launch_bounds__(128) k_curvature_dense(const Lattice<T> L, T* __restrict__ curv_s1) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > L.nx) return;
    curv_s1[i + L.NX1 * (j + (long long)L.NY1 * k)] = curvature_at(L, L.u(i, j, k));
}

// =====================================================================================================
// inlet / outlet kernels (:1009-1524).  One thread per (i,j) of the z-plane.  AFTER = false: the "before odd"
// variant launched after an even step (fills the ghost plane the odd pull will read); AFTER = true: launched
// after an odd step (patches the swapped slots at k = 1 / k = nz).
// Unknown directions at the inlet: QIN = {5,11,12,15,16} (ez = +1); at the outlet their opposites.
// The reference blends every result with the old value by the wall flag of the boundary node,
// new*(1-w) + old*w: for w = 1 that is the old value, so threads of solid boundary nodes only do the phi part.
// =====================================================================================================
__host__ __device__ constexpr int qin(int n) { constexpr int t[5] = {5, 11, 12, 15, 16}; return t[n]; }

template <typename T>
__device__ __forceinline__ void inlet_phi(const Lattice<T>& L, int i, int j, int wi) {  // :1021-1024
    const int c = L.u(i, j, 0), sz = L.sz;
    const T v = L.phi_inlet * (1 - wi) + L.phi[c] * wi;
    L.phi[c] = v; L.phi[c - sz] = v; L.phi[c - 2 * sz] = v; L.phi[c - 3 * sz] = v;
}

#define MFLBM_PLANE_IJ()                                          \
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;    \
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;      \
    if (i > ihi || j > L.ny) return;

template <typename T, bool AFTER, int PART>
__device__ __forceinline__ void inlet_velocity_plane(const Lattice<T>& L, const int ilo, const int ihi) {  // :1009-1077
    MFLBM_PLANE_IJ();
    const int wi = L.solid(L.u(i, j, 1)) ? 1 : 0;
    if (PART != 2) inlet_phi(L, i, j, wi);
    if (wi || PART == 1) return;
    T tmp2 = L.W_in[L.iplane(i, j)] * L.relaxation;
    const T tmp1 = tmp2 * L.sa_inject;
    tmp2 = tmp2 - tmp1;
    // three rounds - all addresses, all loads, all stores - instead of ten dependent look-up -> load -> store chains
    // (the cells are distinct, but behind references the compiler has to assume they alias and serialises them)
    T* gh[10]; T* in[10]; T v[10];
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int q = qin(n), o = opc(q);
            gh[5 * g + n] = &L.f(q, g, L.u(i - ex(q), j - ey(q), 0));   // ghost plane, slot q at x - e_q
            in[5 * g + n] = &L.f(o, g, L.u(i, j, 1));                   // first real plane, slot opc(q)
        }
#pragma unroll
    for (int m = 0; m < 10; m++) v[m] = AFTER ? *gh[m] : *in[m];
#pragma unroll
    for (int m = 0; m < 10; m++) {
        const T t = m < 5 ? tmp1 : tmp2;
        const T wgt = (m % 5) == 0 ? T(1) / T(18) : T(1) / T(36);
        const T r = v[m] + lit<T>(6.0) * wgt * t;
        if (!AFTER) *gh[m] = r; else *in[m] = r;
    }
}

// f_q arriving at (i,j,k) for the nine in-plane directions; v is indexed by direction
template <typename T, bool AFTER>
__device__ __forceinline__ void zh_inplane(const Lattice<T>& L, int i, int j, int k, int g, T (&v)[11]) {
    constexpr int QS[9] = {0, 1, 2, 3, 4, 7, 8, 9, 10};
#pragma unroll
    for (int n = 0; n < 9; n++) {
        const int q = QS[n];
        v[q] = AFTER ? L.f(opc(q), g, L.u(i, j, k)) : L.f(q, g, L.u(i - ex(q), j - ey(q), k));
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

template <typename T, bool AFTER, int PART>
__device__ __forceinline__ void inlet_pressure_plane(const Lattice<T>& L, const int ilo, const int ihi) {  // :1085-1241
    MFLBM_PLANE_IJ();
    const int wi = L.solid(L.u(i, j, 1)) ? 1 : 0;
    if (PART != 2) inlet_phi(L, i, j, wi);
    if (wi || PART == 1) return;
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
            out[n] = AFTER ? L.f(q, g, L.u(i, j, 1)) : L.f(o, g, L.u(i - ex(o), j - ey(o), 2));
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
            if (AFTER) L.f(o, g, L.u(i, j, 1)) = val;
            else L.f(q, g, L.u(i - ex(q), j - ey(q), 0)) = val;
        }
    }
}

template <typename T, bool AFTER, int PART>
__device__ __forceinline__ void outlet_convective_plane(const Lattice<T>& L, const int ilo, const int ihi) {  // :1246-1357
    MFLBM_PLANE_IJ();
    const int nz = L.nz;
    const T u_convec = L.uin_avg;
    const T temp = lit<T>(1.) / (lit<T>(1.) + u_convec);
    const int wi = L.solid(L.u(i, j, nz)) ? 1 : 0;
    const int c = L.u(i, j, nz), sz = L.sz, cb = L.iplane(i, j);
    if (PART != 2) {
        const T ph = ((L.phi_convec[cb] + u_convec * L.phi[c]) * temp) * (1 - wi) + L.phi[c + sz] * wi;
        L.phi[c + sz] = ph; L.phi_convec[cb] = ph; L.phi[c + 2 * sz] = ph; L.phi[c + 3 * sz] = ph; L.phi[c + 4 * sz] = ph;
    }
    if (PART == 1) return;
    const int plane = L.NX1 * L.NY1;
    // addresses, loads, stores in three rounds (see k_inlet_velocity)
    T* dst[10]; const T* inner[10]; T* rec[10]; T a[10], b[10];
#pragma unroll
    for (int g = 0; g < 2; g++) {
        T* buf = g == 0 ? L.f_convec : L.g_convec;
#pragma unroll
        for (int n = 0; n < 5; n++) {
            const int o = opc(qin(n));   // unknown incoming direction at the outlet (ez = -1)
            const int m = 5 * g + n;
            if (!AFTER) { dst[m] = &L.f(o, g, L.u(i - ex(o), j - ey(o), nz + 1)); inner[m] = &L.f(o, g, L.u(i - ex(o), j - ey(o), nz)); }
            else { dst[m] = &L.f(opc(o), g, L.u(i, j, nz)); inner[m] = &L.f(opc(o), g, L.u(i, j, nz - 1)); }
            rec[m] = buf + (cb + plane * o);
        }
    }
#pragma unroll
    for (int m = 0; m < 10; m++) { a[m] = wi ? *dst[m] : *rec[m]; b[m] = wi ? T(0) : *inner[m]; }
#pragma unroll
    for (int m = 0; m < 10; m++) {
        if (wi) { *rec[m] = a[m]; continue; }   // solid boundary node: the blend keeps *dst and records it
        const T val = (a[m] + u_convec * b[m]) * temp;
        *dst[m] = val;
        *rec[m] = val;
    }
}

template <typename T, bool AFTER, int PART>
__device__ __forceinline__ void outlet_pressure_plane(const Lattice<T>& L, const int ilo, const int ihi) {  // :1363-1524
    MFLBM_PLANE_IJ();
    const int nz = L.nz;
    const int wi = L.solid(L.u(i, j, nz)) ? 1 : 0;
    const int c = L.u(i, j, nz), sz = L.sz;
    const T phn = L.phi[c];
    if (PART != 2) { L.phi[c + sz] = phn; L.phi[c + 2 * sz] = phn; L.phi[c + 3 * sz] = phn; L.phi[c + 4 * sz] = phn; }
    if (wi || PART == 1) return;
    T v0[11], v1[11], o0[5], o1[5];
    zh_inplane<T, AFTER>(L, i, j, nz, 0, v0);
    zh_inplane<T, AFTER>(L, i, j, nz, 1, v1);
#pragma unroll
    for (int n = 0; n < 5; n++) {   // known outgoing f_q (ez = +1): pulled from k = nz-1, or the swapped local slot opc(q)
        const int q = qin(n);
        o0[n] = AFTER ? L.f(opc(q), 0, L.u(i, j, nz)) : L.f(q, 0, L.u(i - ex(q), j - ey(q), nz - 1));
        o1[n] = AFTER ? L.f(opc(q), 1, L.u(i, j, nz)) : L.f(q, 1, L.u(i - ex(q), j - ey(q), nz - 1));
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
            if (AFTER) L.f(q, g, L.u(i, j, nz)) = val;
            else L.f(o, g, L.u(i - ex(o), j - ey(o), nz + 1)) = val;
        }
    }
}

// One launch for both ends of an open z axis: blockIdx.z = 0 runs the inlet plane, 1 the outlet plane (two independent
// latency-bound plane kernels; they touch k <= 2 and k >= nz - 1).  INLET / OUTLET: 0 none, 1 velocity / convective,
// 2 pressure (the reference's inlet_BC / outlet_BC).  PART: 0 everything; 1 the phase-field part only (ghost planes of phi,
// phi_convec); 2 the distribution part only.  The two parts are independent of each other - the distribution part reads phi at
// real fluid nodes only - so Solver::step runs part 2 on a second lane NEXT TO the gradient chain, which needs part 1 only.
template <typename T, bool AFTER, int INLET, int OUTLET, int PART>
__global__ void k_open_z(const Lattice<T> L, const int ilo, const int ihi) {
    if (blockIdx.z == 0) {
        if (INLET == 1) inlet_velocity_plane<T, AFTER, PART>(L, ilo, ihi);
        else if (INLET == 2) inlet_pressure_plane<T, AFTER, PART>(L, ilo, ihi);
    } else {
        if (OUTLET == 1) outlet_convective_plane<T, AFTER, PART>(L, ilo, ihi);
        else if (OUTLET == 2) outlet_pressure_plane<T, AFTER, PART>(L, ilo, ihi);
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
    // the four sites involved: real layers 1 and n, ghost layers 0 and n+1
    const int u1 = AXIS == 1 ? L.u(i, 1, m) : L.u(i, m, 1), un = AXIS == 1 ? L.u(i, n, m) : L.u(i, m, n);
    const int u0 = AXIS == 1 ? L.u(i, 0, m) : L.u(i, m, 0), up = AXIS == 1 ? L.u(i, n + 1, m) : L.u(i, m, n + 1);
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int s = AXIS == 1 ? ey(q) : ez(q);
            if (s == 0) continue;
            const int a = s < 0 ? u1 : un, b = s < 0 ? up : u0;   // real layer a <-> ghost layer b
            if (ODD) L.f(q, g, a) = L.f(q, g, b); else L.f(q, g, b) = L.f(q, g, a);
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
            if (!ODD) L.f(q, g, L.u(i, jb, kb)) = L.f(q, g, L.u(i, ja, ka));
            else L.f(q, g, L.u(i, ja, ka)) = L.f(q, g, L.u(i, jb, kb));
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
            L.phi[L.u(i, m, k + nz)] = L.phi[L.u(i, m, k)];
            L.phi[L.u(i, m, k - ov)] = L.phi[L.u(i, m, nz + k - ov)];
        }
    } else if (WHICH == 1) {
        if (m > nz) return;
        for (int j = 1; j <= ov; j++) {
            L.phi[L.u(i, j + ny, m)] = L.phi[L.u(i, j, m)];
            L.phi[L.u(i, j - ov, m)] = L.phi[L.u(i, ny + j - ov, m)];
        }
    } else {
        if (m > 1) return;
        for (int k = 1; k <= ov; k++)
            for (int j = 1; j <= ov; j++) {
                L.phi[L.u(i, j - ov, k - ov)] = L.phi[L.u(i, ny + j - ov, nz + k - ov)];
                L.phi[L.u(i, j + ny, k - ov)] = L.phi[L.u(i, j, nz + k - ov)];
                L.phi[L.u(i, j + ny, k + nz)] = L.phi[L.u(i, j, k)];
                L.phi[L.u(i, j - ov, k + nz)] = L.phi[L.u(i, ny + j - ov, k)];
            }
    }
}

// =====================================================================================================
// porous plate at z = Z_porous_plate (:1744-1882): bounce-back across the plane for the blocked component,
// pass-through copies for the other
// =====================================================================================================
// x range: the bounce-back part runs on the real columns [ilo..ihi]; the pass-through copies are local to a column
// and also run on the ghost columns [plo..phi] that face a neighbour slab, because they change real-plane values after
// the PDF halo of the step has been sent (mflbm/slab.py).
template <typename T, bool AFTER>
__global__ void k_porous_plate(const Lattice<T> L, const int ilo, const int ihi, const int plo, const int phi) {
    const int i = plo + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > phi || j > L.ny) return;
    const int zp = L.Z_porous_plate, cmd = L.porous_plate_cmd;
    if (!(zp >= 1 && zp <= L.nz) || (cmd != 1 && cmd != 2)) return;
    const int gb = cmd == 1 ? 0 : 1, gp = 1 - gb;
    const bool real_col = i >= ilo && i <= ihi;
#pragma unroll
    for (int n = 0; n < 5; n++) {
        const int q = qin(n), o = opc(q);
        if (real_col) {
            if (!AFTER) {   // :1758-1768
                L.f(o, gb, L.u(i - ex(o), j - ey(o), zp)) = L.f(q, gb, L.u(i, j, zp - 1));
                L.f(q, gb, L.u(i - ex(q), j - ey(q), zp)) = L.f(o, gb, L.u(i, j, zp + 1));
            } else {        // :1828-1838
                L.f(q, gb, L.u(i, j, zp - 1)) = L.f(o, gb, L.u(i - ex(o), j - ey(o), zp));
                L.f(o, gb, L.u(i, j, zp + 1)) = L.f(q, gb, L.u(i - ex(q), j - ey(q), zp));
            }
        }
        if (!AFTER) {   // :1770-1780
            L.f(o, gp, L.u(i, j, zp)) = L.f(o, gp, L.u(i, j, zp + 1));
            L.f(q, gp, L.u(i, j, zp)) = L.f(q, gp, L.u(i, j, zp - 1));
        } else if (real_col) {        // :1840-1850
            L.f(q, gp, L.u(i, j, zp - 1)) = L.f(q, gp, L.u(i, j, zp));
            L.f(o, gp, L.u(i, j, zp + 1)) = L.f(o, gp, L.u(i, j, zp));
        }
    }
}

}  // namespace mflbm
