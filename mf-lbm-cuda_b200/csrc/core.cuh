// core.cuh — lattice tables, device-side lattice view and the collision operator.
//
// Hand-written for sm_100a.  Everything here is templated on the real type T (float | double); the reference
// selects it at compile time (includes/solver_precision.h:8-22), we build both into one library.
// Reference citations are relative to /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mflbm {

// ---- D3Q19 lattice: includes/Module.h:98-101 ----
__host__ __device__ constexpr int ex(int q) { constexpr int t[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0}; return t[q]; }
__host__ __device__ constexpr int ey(int q) { constexpr int t[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return t[q]; }
__host__ __device__ constexpr int ez(int q) { constexpr int t[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1}; return t[q]; }
__host__ __device__ constexpr int opc(int q) { constexpr int t[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15}; return t[q]; }

// Literals: the reference writes prc(x) = x##f in single precision.  (T)x from a double literal is the same float
// for every literal used on this path; the static_asserts pin that (a failure here would be a parity bug).
template <typename T> __host__ __device__ constexpr T lit(double x) { return (T)x; }
static_assert((float)1.19 == 1.19f && (float)1.4 == 1.4f && (float)1.2 == 1.2f && (float)1.98 == 1.98f, "literal");
static_assert((float)0.666666666666666667 == 0.666666666666666667f && (float)0.333333333333333333 == 0.333333333333333333f, "literal");
static_assert((float)0.166666666666666667 == 0.166666666666666667f && (float)0.1 == 0.1f && (float)0.025 == 0.025f, "literal");
static_assert((float)5.5 == 5.5f && (float)1e-6 == 1e-6f && (float)1e-30 == 1e-30f && (float)0.999 == 0.999f, "literal");

// weights, includes/Module.h:114-122 (computed in T, like the reference's static initialisers)
template <typename T> __host__ __device__ constexpr T w_equ(int q) { return q == 0 ? T(1) / T(3) : (q < 7 ? T(1) / T(18) : T(1) / T(36)); }

// Device-side view of one lattice (or x-slab).
//
// Memory layout in HBM (internal; the C ABI converts from/to the reference's layouts, includes/Idx_gpu.cuh:52-70):
//
//  * "U" grid: ONE dense index space for every per-site scalar (phi, cn_x/y/z, c_norm, node types, the site->PDF
//    map): 1-based coordinates with 4 ghost layers, x fastest, row pitch PX padded to a multiple of 16 elements:
//        u(x,y,z) = (x+3) + PX*((y+3) + PY*(z+3)),   neighbour in direction q = u + ex + PX*(ey + PY*ez).
//    The reference keeps four different ghost widths (0/1/2/4) and hence four index spaces; one space means a kernel
//    decodes a site once and every stencil offset is the same constant for every field.
//  * PDFs: struct-of-arrays over the 19 directions; the two components of a direction sit side by side,
//        pdf[2*(q*NC + cmap[u]) + g]          (g = 0, 1: the reference's slots q and q + 19)
//    because every kernel that touches a population touches both components of it: one 8/16-byte request moves the
//    pair, which halves the number of gather (LDGSTS) and scatter (STG) requests of the odd step - the load/store pipe
//    is what bounds that kernel.  The SITE order inside a direction row is permuted: the n_fluid
//    real fluid nodes first (in z,y,x order), then every other site of the 1-ghost box.
//    The collide kernels run one thread per entry of the fluid range: warps are full, the even (local) step reads and
//    writes perfectly contiguous, aligned rows, and no DRAM sector is shared between fluid and solid sites.  Solid
//    and ghost sites keep their storage (the reference realises bounce-back through it).
//  * Wall-link mailboxes.  In the reference a fluid node x next to a solid site s = x + e_o pushes f_o* into
//    (s, slot opc(o)) on an odd step and pulls it back as f_opc(o) two steps later; no other thread ever touches that
//    cell (its only reader and writer is the node at s + e_slot = x).  At porosity 0.44 there are 1.2 such links per
//    fluid node, and as isolated 4/8-byte accesses into solid storage they cost a third of the odd step's DRAM traffic
//    (32 B sector fetched for the read, fetched again to merge the partial write).  They live in a compact region at
//    the END of the direction row instead: entry mb0 + rank_o(x) of row opc(o) (the slot the cell has in the reference; a
//    row only ever hosts links of one direction), rank_o = number of fluid entries before x that have a link in direction
//    o, = wbase[(t>>5)*18 + o-1] + popc(ballot & lanes below) inside the odd kernel.  Consecutive threads hit
//    consecutive elements, and a wall link is addressed exactly like a fluid neighbour (slot base + entry).  Every
//    other kernel reaches the same cells through Lattice::f(); the arrays handed back by download_state place them
//    at the reference's addresses.
//  * curv is not stored: the collide kernel (and the monitor) evaluate it from cn_* where it is consumed; the
//    reference's dense curv array is produced on demand by download_state.
// the two components of one population
template <typename T>
struct alignas(2 * sizeof(T)) Pair { T a, b; };

// n / d for n < 2^31 and a divisor fixed at set-up time: (n * mul) >> shift with mul = ceil(2^shift / d), shift = 31 + ceil(log2 d)
// (exact for every n < 2^31; mul < 2^32).  Used to turn a U index back into coordinates inside the hot kernels.
struct FastDiv { unsigned mul; unsigned shift; };
inline FastDiv make_fastdiv(unsigned d) {
    unsigned l = 0;
    while ((1ull << l) < d) l++;
    const unsigned shift = 31 + l;
    const unsigned long long mul = ((1ull << shift) + d - 1) / d;
    return FastDiv{(unsigned)mul, shift};
}
__host__ __device__ __forceinline__ unsigned fastdiv(unsigned n, FastDiv dv) { return (unsigned)(((unsigned long long)n * dv.mul) >> dv.shift); }

template <typename T>
struct Lattice {
    int nx, ny, nz;          // real nodes of this lattice (slab-local nx)
    int NX1, NY1, NZ1;       // 1-ghost extents (reference layout of W_in / convective buffers / the boundary arrays)
    int PX, PY, PZ;          // U grid extents (PX padded)
    int sy, sz;              // U strides: PX, PX*PY
    int x0;                  // global x of local column 1 (1 for a full lattice)
    int nx_global;
    int n_fluid;             // real fluid nodes = threads of the collide kernels
    long long NC;            // entries per PDF slot: fluid nodes, other sites of the 1-ghost box, wall-link mailboxes (multiple of 128)
    // state
    T* pdf; T* phi; T* cn_x; T* cn_y; T* cn_z; T* c_norm;
    T* W_in; T* f_convec; T* g_convec; T* phi_convec;
    // geometry
    const signed char* types;   // U: 0 fluid, -1 fluid boundary, 1 solid, 2 solid boundary (walls_type, Geometry_preprocessing.cpp:154-175)
    const int* cmap;            // U: entry of the site inside a PDF slot: e >= 0 for non-solid sites, -(e + 2) for solid-type
                                //    sites (they keep an entry, see f()), -1 outside the 1-ghost box
    const int* fl_u;            // [n_fluid] U index of the t-th fluid node
    // wall-link mailboxes (see above)
    int mb0;                    // first mailbox entry of every slot
    const int* wbase;           // [ceil(n_fluid/32)][18] rank of the first link of each 32-entry group
    // constants uploaded by copyConstantData in the reference (src/main_iteration_GPU.cu:14-47)
    T lbm_gamma, force_z, la_nui1, la_nui2, lbm_beta, RK_weight2, phi_inlet, relaxation, sa_inject, uin_avg, cos_theta;
    T rho_in, rho_out;
    int Z_porous_plate, porous_plate_cmd;
    // interface-activity bricks (kernels_chain.cuh): bricks per axis, divisors that decode a U index / a brick index, the
    // brick flags (k_chain_pre, k_act_scan) and the flags per group of 32 fluid entries the collide kernels raise
    // (nullptr: the list chain is in use, nothing is raised)
    int nbx, nby, nbz;
    FastDiv dv_sz, dv_px, dv_nbxy, dv_nbx;
    unsigned char* act_p; unsigned char* act_m;
    unsigned char* grp_p; unsigned char* grp_m;

    __device__ __forceinline__ int u(int x, int y, int z) const { return (x + 3) + PX * ((y + 3) + PY * (z + 3)); }
    __device__ __forceinline__ int off(int q) const { return ex(q) + sy * ey(q) + sz * ez(q); }
    __device__ __forceinline__ int iplane(int x, int y) const { return x + NX1 * y; }          // W_in, *_convec
    __device__ __forceinline__ bool solid(int uu) const { return types[uu] > 0; }
    __device__ __forceinline__ T& at(int q, int g, long long entry) const { return pdf[(((long long)q * NC + entry) << 1) + g]; }
    __device__ __forceinline__ Pair<T>* pairs(int q) const { return reinterpret_cast<Pair<T>*>(pdf) + (long long)q * NC; }
    __device__ __forceinline__ static int entry_of(int c) { return c >= 0 ? c : -c - 2; }
    // entry of the mailbox of link (fluid entry t, direction o) in slot opc(o); slow path (walks the 32-entry group),
    // for the plane kernels and layout conversion only
    __device__ int mail_index(int t, int o) const {
        int r = mb0 + wbase[(t >> 5) * 18 + (o - 1)];
        const int t0 = t & ~31, n = t & 31, oo = off(o);
        // all 31 candidates at once: two rounds of independent loads instead of a chain of up to 62 dependent ones (the
        // plane kernels spent ~20 us in the few threads that sit next to a wall)
#pragma unroll
        for (int k = 0; k < 31; k++) {
            const int uu = fl_u[t0 + (k < n ? k : 0)];
            r += (k < n && cmap[uu + oo] < -1) ? 1 : 0;
        }
        return r;
    }
    // PDF of slot (q,g) at the site with U index uu (must lie in the 1-ghost box), wherever it is stored
    __device__ __forceinline__ T& f(int q, int g, int uu) const {
        const int c = cmap[uu];
        if (c >= 0) return at(q, g, c);
        if (q != 0) {   // solid-type site: the cell may be the private mailbox of the fluid node at uu + e_q
            const int co = cmap[uu + off(q)];
            if (co >= 0 && co < n_fluid) return at(q, g, mail_index(co, opc(q)));
        }
        return at(q, g, -c - 2);
    }
    // the cell in slot storage proper, never a mailbox (x-slab halo columns, see k_halo_pdf)
    __device__ __forceinline__ T& f_raw(int q, int g, int uu) const { return at(q, g, entry_of(cmap[uu])); }
};

// MRT relaxation rates, src/main_iteration_GPU.cu:157-186.  MRT is a template parameter so that the compiler folds
// the same constant sub-expressions the reference's #if mrt==... build folds.
template <typename T, int MRT>
__device__ __forceinline__ void mrt_rates(T omega, T& s_e, T& s_e2, T& s_q, T& s_pi, T& s_t) {
    if (MRT == 1) { s_e = omega; s_e2 = omega; s_pi = omega; s_q = lit<T>(8.) * (lit<T>(2.) - omega) / (lit<T>(8.) - omega); s_t = s_q; }
    else if (MRT == 3) { s_e = omega; s_e2 = omega; s_pi = omega; s_q = omega; s_t = omega; }
    else if (MRT == 4) { s_e = omega; s_e2 = omega; s_pi = omega; s_q = (lit<T>(6.) - lit<T>(3.) * omega) / (lit<T>(3.) - omega); s_t = omega; }
    else { s_e = lit<T>(1.19); s_e2 = lit<T>(1.4); s_pi = lit<T>(1.4); s_q = lit<T>(1.2); s_t = lit<T>(1.98); }
}

// Colour-gradient collision of one node: MRT on the bulk distribution with the CSF force, then R-K recolouring.
// Follows src/main_iteration_GPU.cu:115-345 operation by operation (same association, so that nvcc's FMA
// contraction sees the same expression trees).  g1/g2: component PDFs in natural direction order, overwritten with
// the post-collision values.  Returns phi.
struct NoHook { __device__ __forceinline__ void operator()(float) const {} __device__ __forceinline__ void operator()(double) const {} };

// after_phi(phi) runs as soon as the order parameter is known, i.e. once all 38 inputs have been consumed by real
// arithmetic (the even kernel stores phi and hands its shared-memory stage back from there).
template <typename T, int MRT, typename Hook = NoHook>
__device__ __forceinline__ T collide_node(const Lattice<T>& L, T (&g1)[19], T (&g2)[19], T cnx, T cny, T cnz, T curv_cnorm_half_gamma, Hook after_phi = Hook()) {
    T f[19];
#pragma unroll
    for (int q = 0; q < 19; q++) f[q] = g1[q] + g2[q];
    T rho1 = g1[0], rho2 = g2[0];
#pragma unroll
    for (int q = 1; q < 19; q++) { rho1 = rho1 + g1[q]; rho2 = rho2 + g2[q]; }
    const T phi_loc = (rho1 - rho2) / (rho1 + rho2);
    after_phi(phi_loc);

    T tmp = curv_cnorm_half_gamma;  // 0.5 * gamma * curv * c_norm, formed by the caller in the reference's order
    const T fx = tmp * cnx, fy = tmp * cny, fz = tmp * cnz + L.force_z;

    const T omega = lit<T>(1.) / (lit<T>(6.) / ((lit<T>(1.0) + phi_loc) * L.la_nui1 + (lit<T>(1.0) - phi_loc) * L.la_nui2) + lit<T>(0.5));
    const T s_nu = omega;
    T s_e, s_e2, s_q, s_pi, s_t;
    mrt_rates<T, MRT>(omega, s_e, s_e2, s_q, s_pi, s_t);
    // includes/Module.h:104-110
    constexpr T mrt_coef1 = T(1) / T(19), mrt_coef2 = T(1) / T(2394), mrt_coef3 = T(1) / T(252), mrt_coef4 = T(1) / T(72);
    constexpr T mrt_e2_coef1 = T(0), mrt_e2_coef2 = T(-475) / T(63), mrt_omega_xx = T(0);

    const T den = rho1 + rho2;
    const T ux = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14] + lit<T>(0.5) * fx;
    const T uy = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18] + lit<T>(0.5) * fy;
    const T uz = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18] + lit<T>(0.5) * fz;
    const T u2 = ux * ux + uy * uy + uz * uz;

    T sum1 = f[1] + f[2] + f[3] + f[4] + f[5] + f[6];
    T sum2 = f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    T sum3 = f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    T sum4 = f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    T sum5 = f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    T sum6 = lit<T>(2.) * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6];
    T sum7 = f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - lit<T>(2.) * (f[15] + f[16] + f[17] + f[18]);
    T sum8 = f[3] + f[4] - f[5] - f[6];
    T sum9 = f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];

    T m_rho = den;
    T m_e = lit<T>(-30.) * f[0] - lit<T>(11.) * sum1 + lit<T>(8.) * sum2;
    T m_e2 = lit<T>(12.) * f[0] - lit<T>(4.) * sum1 + sum2;
    T m_jx = f[1] - f[2] + sum3;
    T m_qx = lit<T>(-4.) * (f[1] - f[2]) + sum3;
    T m_jy = f[3] - f[4] + sum4;
    T m_qy = lit<T>(-4.) * (f[3] - f[4]) + sum4;
    T m_jz = f[5] - f[6] + sum5;
    T m_qz = lit<T>(-4.) * (f[5] - f[6]) + sum5;
    T m_3pxx = sum6 + sum7;
    T m_3pixx = lit<T>(-2.) * sum6 + sum7;
    T m_pww = sum8 + sum9;
    T m_piww = lit<T>(-2.) * sum8 + sum9;
    T m_pxy = f[7] - f[8] - f[9] + f[10];
    T m_pyz = f[15] - f[16] - f[17] + f[18];
    T m_pzx = f[11] - f[12] - f[13] + f[14];
    T m_tx = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    T m_ty = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    T m_tz = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];

    // relaxation in moment space with forcing, :228-246
    m_e = m_e - s_e * (m_e - (lit<T>(-11.0) * den + lit<T>(19.0) * u2)) + (lit<T>(38.) - lit<T>(19.) * s_e) * (fx * ux + fy * uy + fz * uz);
    m_e2 = m_e2 - s_e2 * (m_e2 - (mrt_e2_coef1 * den + mrt_e2_coef2 * u2)) + (lit<T>(-11.) + lit<T>(5.5) * s_e2) * (fx * ux + fy * uy + fz * uz);
    m_jx = m_jx + fx;
    m_qx = m_qx - s_q * (m_qx - (lit<T>(-0.666666666666666667) * ux)) + (lit<T>(-0.666666666666666667) + lit<T>(0.333333333333333333) * s_q) * fx;
    m_jy = m_jy + fy;
    m_qy = m_qy - s_q * (m_qy - (lit<T>(-0.666666666666666667) * uy)) + (lit<T>(-0.666666666666666667) + lit<T>(0.333333333333333333) * s_q) * fy;
    m_jz = m_jz + fz;
    m_qz = m_qz - s_q * (m_qz - (lit<T>(-0.666666666666666667) * uz)) + (lit<T>(-0.666666666666666667) + lit<T>(0.333333333333333333) * s_q) * fz;
    m_3pxx = m_3pxx - s_nu * (m_3pxx - (lit<T>(3.) * ux * ux - u2)) + (lit<T>(2.) - s_nu) * (lit<T>(2.) * fx * ux - fy * uy - fz * uz);
    m_3pixx = m_3pixx - s_pi * (m_3pixx - mrt_omega_xx * (lit<T>(3.) * ux * ux - u2)) + (lit<T>(1.) - lit<T>(0.5) * s_pi) * (lit<T>(-2.) * fx * ux + fy * uy + fz * uz);
    m_pww = m_pww - s_nu * (m_pww - (uy * uy - uz * uz)) + (lit<T>(2.) - s_nu) * (fy * uy - fz * uz);
    m_piww = m_piww - s_pi * (m_piww - mrt_omega_xx * (uy * uy - uz * uz)) + (lit<T>(1.) - lit<T>(0.5) * s_pi) * (-fy * uy + fz * uz);
    m_pxy = m_pxy - s_nu * (m_pxy - (ux * uy)) + (lit<T>(1.) - lit<T>(0.5) * s_nu) * (fx * uy + fy * ux);
    m_pyz = m_pyz - s_nu * (m_pyz - (uy * uz)) + (lit<T>(1.) - lit<T>(0.5) * s_nu) * (fy * uz + fz * uy);
    m_pzx = m_pzx - s_nu * (m_pzx - (ux * uz)) + (lit<T>(1.) - lit<T>(0.5) * s_nu) * (fx * uz + fz * ux);
    m_tx = m_tx - s_t * (m_tx);
    m_ty = m_ty - s_t * (m_ty);
    m_tz = m_tz - s_t * (m_tz);

    // back to distribution space, :250-297
    m_rho = mrt_coef1 * m_rho;
    m_e = mrt_coef2 * m_e;
    m_e2 = mrt_coef3 * m_e2;
    m_jx = lit<T>(0.1) * m_jx;  m_qx = lit<T>(0.025) * m_qx;
    m_jy = lit<T>(0.1) * m_jy;  m_qy = lit<T>(0.025) * m_qy;
    m_jz = lit<T>(0.1) * m_jz;  m_qz = lit<T>(0.025) * m_qz;
    m_3pxx = lit<T>(2.) * mrt_coef4 * m_3pxx;
    m_3pixx = mrt_coef4 * m_3pixx;
    m_pww = lit<T>(6.) * mrt_coef4 * m_pww;
    m_piww = lit<T>(3.) * mrt_coef4 * m_piww;
    m_pxy = lit<T>(0.25) * m_pxy;  m_pyz = lit<T>(0.25) * m_pyz;  m_pzx = lit<T>(0.25) * m_pzx;
    m_tx = lit<T>(0.125) * m_tx;  m_ty = lit<T>(0.125) * m_ty;  m_tz = lit<T>(0.125) * m_tz;
    sum1 = m_rho - lit<T>(11.) * m_e - lit<T>(4.) * m_e2;
    sum2 = lit<T>(2.) * m_3pxx - lit<T>(4.) * m_3pixx;
    sum3 = m_pww - lit<T>(2.) * m_piww;
    sum4 = m_rho + lit<T>(8.) * m_e + m_e2;
    sum5 = m_jx + m_qx;
    sum6 = m_jy + m_qy;
    sum7 = m_jz + m_qz;
    sum8 = m_3pxx + m_3pixx;
    sum9 = m_pww + m_piww;

    f[0] = m_rho - lit<T>(30.) * m_e + lit<T>(12.) * m_e2;
    f[1] = sum1 + m_jx - lit<T>(4.) * m_qx + sum2;
    f[2] = sum1 - m_jx + lit<T>(4.) * m_qx + sum2;
    f[3] = sum1 + m_jy - lit<T>(4.) * m_qy - lit<T>(0.5) * sum2 + sum3;
    f[4] = sum1 - m_jy + lit<T>(4.) * m_qy - lit<T>(0.5) * sum2 + sum3;
    f[5] = sum1 + m_jz - lit<T>(4.) * m_qz - lit<T>(0.5) * sum2 - sum3;
    f[6] = sum1 - m_jz + lit<T>(4.) * m_qz - lit<T>(0.5) * sum2 - sum3;
    f[7] = sum4 + sum5 + sum6 + sum8 + sum9 + m_pxy + m_tx - m_ty;
    f[8] = sum4 - sum5 + sum6 + sum8 + sum9 - m_pxy - m_tx - m_ty;
    f[9] = sum4 + sum5 - sum6 + sum8 + sum9 - m_pxy + m_tx + m_ty;
    f[10] = sum4 - sum5 - sum6 + sum8 + sum9 + m_pxy - m_tx + m_ty;
    f[11] = sum4 + sum5 + sum7 + sum8 - sum9 + m_pzx - m_tx + m_tz;
    f[12] = sum4 - sum5 + sum7 + sum8 - sum9 - m_pzx + m_tx + m_tz;
    f[13] = sum4 + sum5 - sum7 + sum8 - sum9 - m_pzx - m_tx - m_tz;
    f[14] = sum4 - sum5 - sum7 + sum8 - sum9 + m_pzx + m_tx - m_tz;
    f[15] = sum4 + sum6 + sum7 - sum8 * lit<T>(2.) + m_pyz + m_ty - m_tz;
    f[16] = sum4 - sum6 + sum7 - sum8 * lit<T>(2.) - m_pyz - m_ty - m_tz;
    f[17] = sum4 + sum6 - sum7 - sum8 * lit<T>(2.) - m_pyz + m_ty + m_tz;
    f[18] = sum4 - sum6 - sum7 - sum8 * lit<T>(2.) + m_pyz - m_ty + m_tz;

    // R-K recolouring, :302-345
    const T tmp1 = rho1 / den;
    g1[0] = tmp1 * f[0];
    g2[0] = f[0] * (lit<T>(1.) - tmp1);
    tmp = rho1 * rho2 * L.lbm_beta / den;
    constexpr T w1 = T(1) / T(18);
    const T rk2 = L.RK_weight2;
    g1[1] = tmp1 * f[1] + w1 * tmp * (cnx);
    g1[2] = tmp1 * f[2] + w1 * tmp * (-cnx);
    g1[3] = tmp1 * f[3] + w1 * tmp * (cny);
    g1[4] = tmp1 * f[4] + w1 * tmp * (-cny);
    g1[5] = tmp1 * f[5] + w1 * tmp * (cnz);
    g1[6] = tmp1 * f[6] + w1 * tmp * (-cnz);
    g1[7] = tmp1 * f[7] + rk2 * tmp * (cnx + cny);
    g1[8] = tmp1 * f[8] + rk2 * tmp * (-cnx + cny);
    g1[9] = tmp1 * f[9] + rk2 * tmp * (cnx - cny);
    g1[10] = tmp1 * f[10] + rk2 * tmp * (-cnx - cny);
    g1[11] = tmp1 * f[11] + rk2 * tmp * (cnx + cnz);
    g1[12] = tmp1 * f[12] + rk2 * tmp * (-cnx + cnz);
    g1[13] = tmp1 * f[13] + rk2 * tmp * (cnx - cnz);
    g1[14] = tmp1 * f[14] + rk2 * tmp * (-cnx - cnz);
    g1[15] = tmp1 * f[15] + rk2 * tmp * (cny + cnz);
    g1[16] = tmp1 * f[16] + rk2 * tmp * (-cny + cnz);
    g1[17] = tmp1 * f[17] + rk2 * tmp * (cny - cnz);
    g1[18] = tmp1 * f[18] + rk2 * tmp * (-cny - cnz);
#pragma unroll
    for (int q = 1; q < 19; q++) g2[q] = f[q] - g1[q];
    return phi_loc;
}

}  // namespace mflbm
