/* =====================================================================================================
 * mflbm_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's time-step path (MF-LBM-CUDA, colour-gradient two-phase
 * D3Q19 LBM) and of the host-side set-up that produces its inputs.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; the product (mf-lbm-cuda_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every stage below against arrays dumped
 * by the real reference (oracle/_ref/ref_cpu_* built from /root/reference by oracle/build_ref.sh; GPU-kernel
 * outputs after 1/2/100 steps from oracle/_ref/ref_gpu_* run on a B200, committed under tests/golden/).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).  The arithmetic
 * keeps the reference's association order so that results agree to rounding; it is compiled without FMA
 * contraction (-ffp-contract=off), whereas nvcc contracts the reference kernels, hence "within tolerance"
 * rather than bit-exact for the floating-point stages.  Integer stages (walls, walls_type, counters) and the
 * host-only floating-point stages (solid normals, initial state) are bit-exact.
 *
 * Build:  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC [-DORC_F32] mflbm_oracle.c -o liboracle_f{32,64}.so -lm
 * ===================================================================================================== */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_F32
typedef float real;
#define R(x) x##f
#define M(fn) fn##f
#else
typedef double real;
#define R(x) x
#define M(fn) fn
#endif

typedef long long i64;

/* ---- lattice tables: includes/Module.h:98-101 ---- */
static const int EX[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int EY[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int EZ[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int OPC[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

/* primary inputs = the control-file keys the path depends on (src/IO_multiphase.cpp:46-190) */
typedef struct orc_params {
    i64 nx, ny, nz;
    int iper, jper, kper;
    int wall_x_min, wall_x_max, wall_y_min, wall_y_max, wall_z_min, wall_z_max;
    int inlet_BC, outlet_BC;
    int porous_plate_cmd, Z_porous_plate;
    int n_exclude_inlet, n_exclude_outlet;
    int mrt;        /* includes/preprocessor.h:4 (compile-time 1..4 in the reference; shipped 2) */
    int rho_out_BC; /* control key rho_out_BC */
    real la_nu1, la_nu2, lbm_gamma, theta_deg, lbm_beta, sa_inject, ca_0, force_z0;
} orc_params;

typedef struct orc_ctx {
    orc_params p;
    i64 NX1, NY1, NZ1, NX2, NY2, NZ2, NX4, NY4, NZ4;
    /* derived scalars (src/Init_multiphase.cpp:128-214, src/IO_multiphase.cpp:200-206) */
    real la_nui1, la_nui2, cos_theta, force_z, rho_in, rho_out, phi_inlet, uin_avg, uin_avg_0, flowrate, relaxation;
    real la_x, la_y, la_z, A_xy, A_xy_effective, eps;
    /* geometry */
    int *walls_global, *walls, *walls_type, *pore_profile_z;
    real *s_nx, *s_ny, *s_nz;
    i64 num_solid_boundary_global, num_fluid_boundary_global, num_solid_boundary, num_fluid_boundary;
    i64 pore_sum, pore_sum_effective;
    /* state */
    real *pdf, *phi, *cn_x, *cn_y, *cn_z, *c_norm, *curv, *W_in, *f_convec, *g_convec, *phi_convec;
    /* host-side macroscopic fields (monitor) */
    real *u, *v, *w, *rho;
} orc_ctx;

/* ---- index helpers: includes/Idx_cpu.h:52-69 (1-based, ghost-aware, x fastest) ---- */
#define S0(c, x, y, z) (((x) - 1) + (c)->p.nx * (((y) - 1) + (c)->p.ny * ((i64)(z) - 1)))
#define S1(c, x, y, z) ((x) + (c)->NX1 * ((y) + (c)->NY1 * (i64)(z)))
#define S2(c, x, y, z) (((x) + 1) + (c)->NX2 * (((y) + 1) + (c)->NY2 * ((i64)(z) + 1)))
#define S4(c, x, y, z) (((x) + 3) + (c)->NX4 * (((y) + 3) + (c)->NY4 * ((i64)(z) + 3)))
#define F1(c, x, y, z, e, g) ((x) + (c)->NX1 * ((y) + (c)->NY1 * ((i64)(z) + (c)->NZ1 * ((e) + 19 * (g)))))
#define CNV(c, x, y, e) ((x) + (c)->NX1 * ((y) + (c)->NY1 * (i64)(e)))

static const real W0 = R(1.) / R(3.), W1 = R(1.) / R(18.), W2 = R(1.) / R(36.); /* Module.h:120-122 */
static real wq(int q) { return q == 0 ? W0 : (q < 7 ? W1 : W2); }                  /* Module.h:114-119 */

/* =====================================================================================================
 * creation / parameters
 * ===================================================================================================== */
int orc_real_bytes(void) { return (int)sizeof(real); }

orc_ctx* orc_create(const orc_params* p) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->p = *p;
    c->NX1 = p->nx + 2; c->NY1 = p->ny + 2; c->NZ1 = p->nz + 2;
    c->NX2 = p->nx + 4; c->NY2 = p->ny + 4; c->NZ2 = p->nz + 4;
    c->NX4 = p->nx + 8; c->NY4 = p->ny + 8; c->NZ4 = p->nz + 8;
    i64 n0 = p->nx * p->ny * p->nz, n1 = c->NX1 * c->NY1 * c->NZ1, n2 = c->NX2 * c->NY2 * c->NZ2, n4 = c->NX4 * c->NY4 * c->NZ4;
    /* src/Init_multiphase.cpp:545-608 (calloc => zero initial content matters: ghost phi, solid cn, ...) */
    c->walls_global = (int*)calloc(n0, sizeof(int));
    c->walls = (int*)calloc(n2, sizeof(int));
    c->walls_type = (int*)calloc(n4, sizeof(int));
    c->pore_profile_z = (int*)calloc(p->nz, sizeof(int));
    c->s_nx = (real*)calloc(n4, sizeof(real)); c->s_ny = (real*)calloc(n4, sizeof(real)); c->s_nz = (real*)calloc(n4, sizeof(real));
    c->pdf = (real*)calloc(n1 * 38, sizeof(real));
    c->phi = (real*)calloc(n4, sizeof(real));
    c->cn_x = (real*)calloc(n2, sizeof(real)); c->cn_y = (real*)calloc(n2, sizeof(real)); c->cn_z = (real*)calloc(n2, sizeof(real));
    c->c_norm = (real*)calloc(n2, sizeof(real));
    c->curv = (real*)calloc(n1, sizeof(real));
    c->W_in = (real*)calloc(c->NX1 * c->NY1, sizeof(real));
    c->f_convec = (real*)calloc(c->NX1 * c->NY1 * 19, sizeof(real));
    c->g_convec = (real*)calloc(c->NX1 * c->NY1 * 19, sizeof(real));
    c->phi_convec = (real*)calloc(c->NX1 * c->NY1, sizeof(real));
    c->u = (real*)calloc(n1, sizeof(real)); c->v = (real*)calloc(n1, sizeof(real)); c->w = (real*)calloc(n1, sizeof(real));
    c->rho = (real*)calloc(n1, sizeof(real));
    for (i64 i = 0; i < n1; i++) c->rho[i] = R(1.);

    /* includes/Module.h:9-16: eps is float epsilon in both precisions */
    c->eps = (real)1.1920928955078125e-07f;
    c->relaxation = R(1.); /* src/main.cpp:57 */

    /* src/IO_multiphase.cpp:204-206 */
    real Pi = R(3.14159265358979323846);
    real theta = R(180.) - p->theta_deg;
    theta = theta * Pi / R(180.);
    c->cos_theta = M(cos)(theta);

    /* src/Init_multiphase.cpp:128-132 */
    c->la_z = (real)(p->nz - 1);
    c->la_y = (real)(p->ny - 1) - R(0.5) - R(0.5);
    c->la_x = (real)(p->nx - 1) - R(0.5) - R(0.5);
    c->A_xy = c->la_x * c->la_y;

    /* src/Init_multiphase.cpp:173-214 */
    c->la_nui1 = R(1.) / p->la_nu1;
    c->la_nui2 = R(1.) / p->la_nu2;
    c->phi_inlet = R(2.) * p->sa_inject - R(1.);
    c->force_z = p->force_z0;
    c->rho_out = R(1.);
    c->rho_in = R(0.); /* global, zero-initialised; only assigned on the pressure-inlet branch */
    if (p->kper == 0 && p->wall_z_min == 0 && p->wall_z_max == 0) {
        if (p->inlet_BC == 1) {
            c->force_z = R(0.);
            /* src/Init_multiphase.cpp:263-265 */
            c->uin_avg_0 = p->ca_0 * p->lbm_gamma / p->la_nu1;
            c->uin_avg = c->uin_avg_0;
            c->flowrate = c->uin_avg_0 * c->A_xy;
        } else if (p->inlet_BC == 2) {
            c->force_z = R(0.);
            real p_gradient = -p->force_z0 / R(3.);
            if (p->rho_out_BC) c->rho_out = R(1.) - p_gradient * p->nz;
            else c->rho_in = c->rho_out - p_gradient * p->nz;
        }
    }
    return c;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    free(c->walls_global); free(c->walls); free(c->walls_type); free(c->pore_profile_z);
    free(c->s_nx); free(c->s_ny); free(c->s_nz);
    free(c->pdf); free(c->phi); free(c->cn_x); free(c->cn_y); free(c->cn_z); free(c->c_norm); free(c->curv);
    free(c->W_in); free(c->f_convec); free(c->g_convec); free(c->phi_convec);
    free(c->u); free(c->v); free(c->w); free(c->rho);
    free(c);
}

/* array access for the Python side: name -> pointer (element counts follow the layouts above) */
void* orc_array(orc_ctx* c, const char* name) {
#define A(n) if (!strcmp(name, #n)) return (void*)c->n
    A(walls_global); A(walls); A(walls_type); A(pore_profile_z); A(s_nx); A(s_ny); A(s_nz);
    A(pdf); A(phi); A(cn_x); A(cn_y); A(cn_z); A(c_norm); A(curv); A(W_in); A(f_convec); A(g_convec); A(phi_convec);
    A(u); A(v); A(w); A(rho);
#undef A
    return NULL;
}

double orc_scalar(orc_ctx* c, const char* name) {
#define S(n) if (!strcmp(name, #n)) return (double)c->n
    S(la_nui1); S(la_nui2); S(cos_theta); S(force_z); S(rho_in); S(rho_out); S(phi_inlet); S(uin_avg); S(uin_avg_0);
    S(flowrate); S(relaxation); S(la_x); S(la_y); S(la_z); S(A_xy); S(A_xy_effective); S(eps);
    S(num_solid_boundary_global); S(num_fluid_boundary_global); S(num_solid_boundary); S(num_fluid_boundary);
    S(pore_sum); S(pore_sum_effective);
#undef S
    return NAN;
}

void orc_set_scalar(orc_ctx* c, const char* name, double v) {
#define S(n) if (!strcmp(name, #n)) { c->n = (real)v; return; }
    S(force_z); S(rho_in); S(rho_out); S(phi_inlet); S(uin_avg); S(relaxation); S(cos_theta); S(la_nui1); S(la_nui2);
#undef S
}

/* =====================================================================================================
 * geometry: src/Misc.cpp:17-217
 * ===================================================================================================== */
/* read_walls (src/Misc.cpp:139-192) + set_walls (src/Misc.cpp:59-101).  `sample` holds nxs*nys*nzs int8 voxels,
 * x fastest, or NULL for "no external geometry".  compat_stride=1 reproduces the reference's indexing
 * geo[(i-1) + nyGlobal*((j-1) + nxGlobal*(k-1))] (src/Misc.cpp:171: strides swapped and Global instead of
 * sample dims; right only when nx==ny==nx_sample==ny_sample); compat_stride=0 indexes the sample properly.
 * Reads beyond the sample (undefined behaviour in the reference) yield 0 here. */
void orc_set_walls(orc_ctx* c, const signed char* sample, i64 nxs, i64 nys, i64 nzs, int compat_stride, int modify_geometry) {
    const i64 nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    memset(c->walls_global, 0, sizeof(int) * nx * ny * nz);
    if (sample) {
        const i64 nsamp = nxs * nys * nzs;
        for (i64 k = 1; k <= nz; k++)
            for (i64 j = 1; j <= ny; j++)
                for (i64 i = 1; i <= nx; i++) {
                    i64 src;
                    if (compat_stride) src = (i - 1) + ny * ((j - 1) + nx * (k - 1));
                    else src = (i <= nxs && j <= nys && k <= nzs) ? (i - 1) + nxs * ((j - 1) + nys * (k - 1)) : -1;
                    c->walls_global[S0(c, i, j, k)] = (src >= 0 && src < nsamp) ? sample[src] : 0;
                }
        if (c->p.wall_x_max == 1 && c->p.wall_y_max == 1)
            for (i64 k = 1; k <= nz; k++)
                for (i64 j = 1; j <= ny; j++)
                    for (i64 i = 1; i <= nx; i++)
                        if (j >= nys || i >= nxs) c->walls_global[S0(c, i, j, k)] = 1;
    }
    if (modify_geometry) { /* src/Misc.cpp:105-136 (double arithmetic in both precisions) */
        double xc = 0.5 * (double)(nx + 1), yc = 0.5 * (double)(ny + 1), zc = 0.5 * (double)(nz + 1);
        double r1 = 0.25 * ny, r2 = ny * 0.5;
        i64 buffer = 10;
        for (i64 k = 1; k <= nz; k++)
            for (i64 j = 1; j <= ny; j++)
                for (i64 i = 1; i <= nx; i++) {
                    double d2 = pow((i - xc), 2) + pow((j - yc), 2) + pow((k - zc), 2);
                    if (d2 < pow(r1, 2)) c->walls_global[S0(c, i, j, k)] = 1;
                    if (d2 > pow(r2, 2) && k > buffer && k < nz - buffer + 1) c->walls_global[S0(c, i, j, k)] = 1;
                }
    }
    /* domain walls: src/Misc.cpp:67-78 */
    for (i64 k = 1; k <= nz; k++)
        for (i64 j = 1; j <= ny; j++)
            for (i64 i = 1; i <= nx; i++) {
                if (c->p.wall_z_min == 1) c->walls_global[S0(c, i, j, 1)] = 1;
                if (c->p.wall_z_max == 1) c->walls_global[S0(c, i, j, nz)] = 1;
                if (c->p.wall_x_min == 1) c->walls_global[S0(c, 1, j, k)] = 1;
                if (c->p.wall_x_max == 1) c->walls_global[S0(c, nx, j, k)] = 1;
                if (c->p.wall_y_min == 1) c->walls_global[S0(c, i, 1, k)] = 1;
                if (c->p.wall_y_max == 1) c->walls_global[S0(c, i, ny, k)] = 1;
            }
    /* src/Misc.cpp:81-87 */
    for (i64 k = 1; k <= nz; k++)
        for (i64 j = 1; j <= ny; j++)
            for (i64 i = 1; i <= nx; i++)
                c->walls[S2(c, i, j, k)] = c->walls_type[S4(c, i, j, k)] = c->walls_global[S0(c, i, j, k)];
    /* src/Misc.cpp:90-98 */
    i64 icount = 0;
    for (i64 j = 1; j <= ny; j++)
        for (i64 i = 1; i <= nx; i++)
            if (c->walls_global[S0(c, i, j, 1)] <= 0) icount++;
    c->A_xy_effective = (real)icount;
    /* pore_profile: src/Misc.cpp:195-217 */
    c->pore_sum = 0; c->pore_sum_effective = 0;
    for (i64 k = 1; k <= nz; k++) {
        int s = 0;
        for (i64 j = 1; j <= ny; j++)
            for (i64 i = 1; i <= nx; i++)
                if (c->walls[S2(c, i, j, k)] <= 0) s++;
        c->pore_profile_z[k - 1] = s;
        c->pore_sum += s;
    }
    for (i64 k = 1 + c->p.n_exclude_inlet; k <= nz - c->p.n_exclude_outlet; k++) c->pore_sum_effective += c->pore_profile_z[k - 1];
}

/* ---- ISO8 two-ring gradient tables, transcribed term by term from src/Geometry_preprocessing.cpp:233-375.
 * Every term there is  w(x+o) - w(x-o); a group is summed left to right; the seven group products are then
 * added left to right.  Offsets listed are the "+o" of each term, in source order. ---- */
typedef struct { signed char d[3]; } off3;
static const int ISO8_N[7] = {1, 4, 4, 1, 8, 12, 4};
static const off3 ISO8_X[34] = {
    {{1, 0, 0}},
    {{1, 1, 0}}, {{1, -1, 0}}, {{1, 0, 1}}, {{1, 0, -1}},
    {{1, 1, 1}}, {{1, 1, -1}}, {{1, -1, 1}}, {{1, -1, -1}},
    {{2, 0, 0}},
    {{2, 1, 0}}, {{2, -1, 0}}, {{2, 0, 1}}, {{2, 0, -1}}, {{1, 2, 0}}, {{1, -2, 0}}, {{1, 0, 2}}, {{1, 0, -2}},
    {{2, 1, 1}}, {{2, 1, -1}}, {{2, -1, 1}}, {{2, -1, -1}}, {{1, 2, 1}}, {{1, 2, -1}}, {{1, -2, 1}}, {{1, -2, -1}},
    {{1, 1, 2}}, {{1, 1, -2}}, {{1, -1, 2}}, {{1, -1, -2}},
    {{2, 2, 0}}, {{2, -2, 0}}, {{2, 0, 2}}, {{2, 0, -2}}};
static const off3 ISO8_Y[34] = {
    {{0, 1, 0}},
    {{1, 1, 0}}, {{-1, 1, 0}}, {{0, 1, 1}}, {{0, 1, -1}},
    {{1, 1, 1}}, {{1, 1, -1}}, {{-1, 1, -1}}, {{-1, 1, 1}},
    {{0, 2, 0}},
    {{2, 1, 0}}, {{-2, 1, 0}}, {{0, 2, 1}}, {{0, 2, -1}}, {{1, 2, 0}}, {{-1, 2, 0}}, {{0, 1, 2}}, {{0, 1, -2}},
    {{2, 1, 1}}, {{2, 1, -1}}, {{-2, 1, 1}}, {{-2, 1, -1}}, {{1, 2, 1}}, {{1, 2, -1}}, {{-1, 2, 1}}, {{-1, 2, -1}},
    {{1, 1, 2}}, {{1, 1, -2}}, {{-1, 1, 2}}, {{-1, 1, -2}},
    {{2, 2, 0}}, {{-2, 2, 0}}, {{0, 2, 2}}, {{0, 2, -2}}};
static const off3 ISO8_Z[34] = {
    {{0, 0, 1}},
    {{0, 1, 1}}, {{0, -1, 1}}, {{1, 0, 1}}, {{-1, 0, 1}},
    {{1, 1, 1}}, {{1, -1, 1}}, {{-1, 1, 1}}, {{-1, -1, 1}},
    {{0, 0, 2}},
    {{0, 1, 2}}, {{0, -1, 2}}, {{2, 0, 1}}, {{-2, 0, 1}}, {{0, 2, 1}}, {{0, -2, 1}}, {{1, 0, 2}}, {{-1, 0, 2}},
    {{2, 1, 1}}, {{2, -1, 1}}, {{-2, 1, 1}}, {{-2, -1, 1}}, {{1, 2, 1}}, {{1, -2, 1}}, {{-1, 2, 1}}, {{-1, -2, 1}},
    {{1, 1, 2}}, {{1, -1, 2}}, {{-1, 1, 2}}, {{-1, -1, 2}},
    {{0, 2, 2}}, {{0, -2, 2}}, {{2, 0, 2}}, {{-2, 0, 2}}};

/* geometry_preprocessing_new: src/Geometry_preprocessing.cpp:29-403 */
void orc_geometry_preprocess(orc_ctx* c) {
    const i64 nx = c->p.nx, ny = c->p.ny, nz = c->p.nz, G = 10;
    const i64 TX = nx + 2 * G, TY = ny + 2 * G, TZ = nz + 2 * G, TN = TX * TY * TZ;
#define T10(x, y, z) (((x) + 9) + TX * (((y) + 9) + TY * ((i64)(z) + 9)))
    /* 27-point neighbour order and weights: src/Geometry_preprocessing.cpp:31-34 */
    static const int iex[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    static const int iey[27] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
    static const int iez[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1};
    const real we[4] = {R(8.) / R(27.), R(2.) / R(27.), R(1.) / R(54.), R(1.) / R(216.)};
    /* includes/Fluid_multiphase.h:35 */
    const real ISO8[7] = {R(4.) / R(45.), R(1.) / R(21.), R(2.) / R(105.), R(5.) / R(504.), R(1.) / R(315.), R(1.) / R(630.), R(1.) / R(5040.)};

    real* ws1 = (real*)calloc(TN, sizeof(real));
    real* ws2 = (real*)calloc(TN, sizeof(real));
    int* wt = (int*)calloc(TN, sizeof(int));

    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) wt[T10(i, j, k)] = c->walls[S2(c, i, j, k)];
    /* ghost fill, z then y then x: :59-132 */
    for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        for (i64 k = 1 - G; k <= 0; k++) wt[T10(i, j, k)] = c->p.kper == 0 ? wt[T10(i, j, 1)] : wt[T10(i, j, nz + k)];
        for (i64 k = nz + 1; k <= nz + G; k++) wt[T10(i, j, k)] = c->p.kper == 0 ? wt[T10(i, j, nz)] : wt[T10(i, j, k - nz)];
    }
    for (i64 k = 1 - G; k <= nz + G; k++) for (i64 i = 1; i <= nx; i++) {
        for (i64 j = 1 - G; j <= 0; j++) wt[T10(i, j, k)] = c->p.jper == 0 ? wt[T10(i, 1, k)] : wt[T10(i, ny + j, k)];
        for (i64 j = ny + 1; j <= ny + G; j++) wt[T10(i, j, k)] = c->p.jper == 0 ? wt[T10(i, ny, k)] : wt[T10(i, j - ny, k)];
    }
    for (i64 k = 1 - G; k <= nz + G; k++) for (i64 j = 1 - G; j <= ny + G; j++) {
        for (i64 i = 1 - G; i <= 0; i++) wt[T10(i, j, k)] = c->p.iper == 0 ? wt[T10(1, j, k)] : wt[T10(nx + i, j, k)];
        for (i64 i = nx + 1; i <= nx + G; i++) wt[T10(i, j, k)] = c->p.iper == 0 ? wt[T10(nx, j, k)] : wt[T10(i - nx, j, k)];
    }
    /* :135-151 */
    for (i64 k = -1; k <= nz + 2; k++) for (i64 j = -1; j <= ny + 2; j++) for (i64 i = -1; i <= nx + 2; i++) c->walls[S2(c, i, j, k)] = wt[T10(i, j, k)];
    for (i64 n = 0; n < TN; n++) { ws1[n] = (real)wt[n]; ws2[n] = (real)wt[n]; }
    /* node classification, in place as in the reference (order-independent): :154-175 */
    for (i64 k = 2 - G; k <= nz + G - 1; k++) for (i64 j = 2 - G; j <= ny + G - 1; j++) for (i64 i = 2 - G; i <= nx + G - 1; i++) {
        if (wt[T10(i, j, k)] == 1) {
            for (int n = 1; n <= 18; n++) if (wt[T10(i + EX[n], j + EY[n], k + EZ[n])] <= 0) { wt[T10(i, j, k)] = 2; break; }
        }
        if (wt[T10(i, j, k)] == 0) {
            for (int n = 1; n <= 18; n++) if (wt[T10(i + EX[n], j + EY[n], k + EZ[n])] >= 1) { wt[T10(i, j, k)] = -1; break; }
        }
    }
    for (i64 k = -3; k <= nz + 4; k++) for (i64 j = -3; j <= ny + 4; j++) for (i64 i = -3; i <= nx + 4; i++) c->walls_type[S4(c, i, j, k)] = wt[T10(i, j, k)];
    /* four smoothing passes: :187-206 */
    for (int it = 1; it <= 4; it++) {
#pragma omp parallel for schedule(static)
        for (i64 k = 2 - G; k <= nz + G - 1; k++) for (i64 j = 2 - G; j <= ny + G - 1; j++) for (i64 i = 2 - G; i <= nx + G - 1; i++) {
            real acc = R(0.);
            for (int n = 0; n <= 26; n++) {
                int m = iex[n] * iex[n] + iey[n] * iey[n] + iez[n] * iez[n];
                acc += ws1[T10(i + iex[n], j + iey[n], k + iez[n])] * we[m];
            }
            ws2[T10(i, j, k)] = acc;
        }
#pragma omp parallel for schedule(static)
        for (i64 k = 2 - G; k <= nz + G - 1; k++) for (i64 j = 2 - G; j <= ny + G - 1; j++) for (i64 i = 2 - G; i <= nx + G - 1; i++) ws1[T10(i, j, k)] = ws2[T10(i, j, k)];
    }
    /* counters: :208-222 and :389-401 */
    c->num_solid_boundary_global = c->num_fluid_boundary_global = 0;
    for (i64 k = -3; k <= nz + 4; k++) for (i64 j = -3; j <= ny + 4; j++) for (i64 i = -3; i <= nx + 4; i++) {
        if (wt[T10(i, j, k)] == 2) c->num_solid_boundary_global++;
        if (wt[T10(i, j, k)] == -1) c->num_fluid_boundary_global++;
    }
    c->num_solid_boundary = c->num_fluid_boundary = 0;
    for (i64 k = -2; k <= nz + 3; k++) for (i64 j = -2; j <= ny + 3; j++) for (i64 i = -2; i <= nx + 3; i++) {
        if (wt[T10(i, j, k)] == 2) c->num_solid_boundary++;
        if (wt[T10(i, j, k)] == -1) c->num_fluid_boundary++;
    }
    /* solid-surface normals at fluid-boundary nodes: :229-386 */
#pragma omp parallel for schedule(static)
    for (i64 k = -3; k <= nz + 4; k++) for (i64 j = -3; j <= ny + 4; j++) for (i64 i = -3; i <= nx + 4; i++) {
        if (wt[T10(i, j, k)] != -1) continue;
        real nw[3];
        const off3* tabs[3] = {ISO8_X, ISO8_Y, ISO8_Z};
        for (int a = 0; a < 3; a++) {
            const off3* t = tabs[a];
            real total = R(0.);
            int pos = 0;
            for (int g = 0; g < 7; g++) {
                real s = R(0.);
                for (int n = 0; n < ISO8_N[g]; n++, pos++) {
                    int dx = t[pos].d[0], dy = t[pos].d[1], dz = t[pos].d[2];
                    real plus = ws2[T10(i + dx, j + dy, k + dz)], minus = ws2[T10(i - dx, j - dy, k - dz)];
                    s = (n == 0) ? (plus - minus) : (s + plus - minus);
                }
                total = (g == 0) ? (ISO8[g] * s) : (total + ISO8[g] * s);
            }
            nw[a] = total;
        }
        real tmp = R(1.) / (M(sqrt)(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]) + c->eps);
        c->s_nx[S4(c, i, j, k)] = nw[0] * tmp;
        c->s_ny[S4(c, i, j, k)] = nw[1] * tmp;
        c->s_nz[S4(c, i, j, k)] = nw[2] * tmp;
    }
    free(ws1); free(ws2); free(wt);
#undef T10
}

/* =====================================================================================================
 * initial state: src/Init_multiphase.cpp:258-496, src/Misc.cpp:387-419
 * ===================================================================================================== */
/* inlet_vel_profile_rectangular + the uniform default: src/Init_multiphase.cpp:271-284, src/Misc.cpp:387-419 */
void orc_init_inlet_velocity_profile(orc_ctx* c) {
    const i64 nx = c->p.nx, ny = c->p.ny;
    const real Pi = R(3.14159265358979323846);
    for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        c->W_in[S1(c, i, j, 0)] = R(0.);
        if (i > 1 && i < nx && j > 1 && j < ny) c->W_in[S1(c, i, j, 0)] = c->uin_avg;
    }
    const int num_terms = 1000;
    real a = R(0.5) * c->la_x, b = R(0.5) * c->la_y, tmp1 = R(0.), tmp2, tmp3;
    for (i64 n = 1; n <= num_terms; n += 2) tmp1 += (M(tanh)(R(0.5) * (real)n * Pi * b / a)) / M(pow)((real)n, 5);
    tmp2 = R(1.) - R(192.) / M(pow)(Pi, 5) * (a / b) * tmp1;
    tmp2 = R(-3.) * c->uin_avg_0 / (tmp2 * M(pow)(a, 2));
    for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        if (i > 1 && i < nx && j > 1 && j < ny) {
            real xx = i - R(1.5) - a, yy = j - R(1.5) - b;
            tmp3 = R(0.);
            for (i64 n = 1; n <= num_terms; n += 2) {
                tmp3 += M(pow)(R(-1.), R(0.5) * (real)(n - 1)) * M(cos)(R(0.5) * (real)n * Pi * xx / a) / M(pow)((real)n, 3)
                        * (R(1.) - (M(exp)(R(0.5) * (real)n * Pi * (yy - b) / a) + M(exp)(R(0.5) * (real)n * Pi * (-yy - b) / a)) /
                                       (R(1.) + M(exp)(R(0.5) * (real)n * Pi * (-b - b) / a)));
            }
            c->W_in[S1(c, i, j, 0)] = tmp3 * (R(-16.) * tmp2 * M(pow)(a, 2) * M(pow)(Pi, -3));
        }
    }
}

/* initialization_new_multi + initialization_new_multi_pdf: src/Init_multiphase.cpp:299-496.
 * option 6 (random) is not restated (it seeds rand() from the wall clock, src/Init_multiphase.cpp:306). */
int orc_init_new(orc_ctx* c, int option, double interface_z0_in) {
    const i64 nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    const real interface_z0 = (real)interface_z0_in;
    for (i64 k = 0; k <= nz + 1; k++) for (i64 j = 0; j <= ny + 1; j++) for (i64 i = 0; i <= nx + 1; i++) {
        real x = (real)i, y = (real)j, z = (real)k, ph;
        if (option == 1) { ph = R(-1.); if (z <= interface_z0) ph = R(1.); }
        else if (option == 2) { ph = R(1.); if (z <= interface_z0) ph = R(-1.); }
        else if (option == 3 || option == 4 || option == 5) {
            real cx = (option == 5) ? R(0.5) : R(0.);
            real d = M(pow)((x - (nx + 1) * cx), 2) + M(pow)((z - (nz + 1) * R(0.5)), 2) + M(pow)((y - (ny + 1) * R(0.5)), 2);
            int inside = d <= M(pow)(interface_z0, 2);
            if (option == 4) ph = inside ? R(-1.) : R(1.);
            else ph = inside ? R(1.) : R(-1.);
        } else return 1;
        c->phi[S4(c, i, j, k)] = ph;
    }
    if (c->p.kper == 0 && c->p.wall_z_min == 0 && c->p.wall_z_max == 0)
        for (i64 k = -3; k <= 0; k++) for (i64 j = -3; j <= ny + 4; j++) for (i64 i = -3; i <= nx + 4; i++) c->phi[S4(c, i, j, k)] = c->phi_inlet;
    /* equilibrium at u=v=w=0, rho=1 (host arrays are calloc'ed / filled with 1: src/Init_multiphase.cpp:569-572) */
    for (i64 k = 0; k <= nz + 1; k++) for (i64 j = 0; j <= ny + 1; j++) for (i64 i = 0; i <= nx + 1; i++) {
        real uu = c->u[S1(c, i, j, k)], vv = c->v[S1(c, i, j, k)], ww = c->w[S1(c, i, j, k)];
        real usqrt = uu * uu + vv * vv + ww * ww;
        real rho1 = c->rho[S1(c, i, j, k)] * (R(1.0) + c->phi[S4(c, i, j, k)]) * R(0.5);
        real rho2 = c->rho[S1(c, i, j, k)] * (R(1.0) - c->phi[S4(c, i, j, k)]) * R(0.5);
        for (int g = 0; g < 2; g++) {
            real rr = g == 0 ? rho1 : rho2;
            c->pdf[F1(c, i, j, k, 0, g)] = rr * W0 + rr * W0 * (R(-1.5) * usqrt);
            for (int q = 1; q < 19; q++) {
                real wgt = q < 7 ? W1 : W2;
                /* e.u written as the reference does: a single component (possibly negated) or a sum of two */
                real eu;
                if (q < 7) { real comp = EX[q] ? uu : (EY[q] ? vv : ww); int s = EX[q] + EY[q] + EZ[q]; eu = s > 0 ? comp : -comp; }
                else {
                    real a1, a2;
                    if (EZ[q] == 0) { a1 = EX[q] > 0 ? uu : -uu; a2 = EY[q] > 0 ? vv : -vv; }
                    else if (EY[q] == 0) { a1 = EX[q] > 0 ? uu : -uu; a2 = EZ[q] > 0 ? ww : -ww; }
                    else { a1 = EY[q] > 0 ? vv : -vv; a2 = EZ[q] > 0 ? ww : -ww; }
                    /* reference forms e.g. (-u + w), (u - w), (-u - w): first operand signed, second added/subtracted */
                    eu = a1 + a2;
                }
                real sq = (q < 7) ? ((EX[q] ? uu * uu : (EY[q] ? vv * vv : ww * ww))) : eu * eu;
                c->pdf[F1(c, i, j, k, q, g)] = rr * wgt + rr * wgt * (R(3.0) * eu + R(4.5) * sq - R(1.5) * usqrt);
            }
        }
    }
    if (c->p.outlet_BC == 1)
        for (i64 j = 0; j <= ny + 1; j++) for (i64 i = 0; i <= nx + 1; i++) {
            for (int q = 0; q < 19; q++) {
                c->f_convec[CNV(c, i, j, q)] = c->pdf[F1(c, i, j, nz, q, 0)];
                c->g_convec[CNV(c, i, j, q)] = c->pdf[F1(c, i, j, nz, q, 1)];
            }
            c->phi_convec[S1(c, i, j, 0)] = c->phi[S4(c, i, j, nz)];
        }
    if (c->p.kper == 0 && c->p.wall_z_min == 0 && c->p.wall_z_max == 0 && c->p.inlet_BC == 1) orc_init_inlet_velocity_profile(c);
    return 0;
}

/* =====================================================================================================
 * colour-gradient chain: src/main_iteration_GPU.cu:732-1003 (GPU) == src/Phase_gradient.cpp:15-290 (CPU twin)
 * ranges are the reference's; [ilo,ihi] restricts x for the slab tests (pass 1-g .. nx+g for the full range)
 * ===================================================================================================== */
void orc_extrapolate_phi_to_solid(orc_ctx* c, i64 ilo, i64 ihi) { /* :732-755, range [-2..n+3] */
    const i64 ny = c->p.ny, nz = c->p.nz;
#pragma omp parallel for schedule(static)
    for (i64 k = -2; k <= nz + 3; k++) for (i64 j = -2; j <= ny + 3; j++) for (i64 i = ilo; i <= ihi; i++) {
        if (c->walls_type[S4(c, i, j, k)] != 2) continue;
        real phi_sum = R(0.), weight_sum = R(0.);
        for (int q = 1; q < 19; q++) {
            i64 n = S4(c, i + EX[q], j + EY[q], k + EZ[q]);
            if (c->walls_type[n] <= 0) { phi_sum += c->phi[n] * wq(q); weight_sum += wq(q); }
        }
        c->phi[S4(c, i, j, k)] = phi_sum / weight_sum;
    }
}

/* the three ISO4 derivative patterns shared by :765-791 and :916-996 ("+o" offsets of the four diagonal pairs) */
static const signed char ISO4_DIAG[3][4][3] = {
    {{1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1}},
    {{1, 1, 0}, {-1, 1, 0}, {0, 1, 1}, {0, 1, -1}},
    {{1, 0, 1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, 1}}};
#define ISO4_0 (R(1.) / R(6.))  /* includes/Fluid_multiphase.h:34 */
#define ISO4_1 (R(1.) / R(12.))

static real iso4_phi(const orc_ctx* c, i64 i, i64 j, i64 k, int a) {
    const real* f = c->phi;
    int ax = a == 0, ay = a == 1, az = a == 2;
    real axis = f[S4(c, i + ax, j + ay, k + az)] - f[S4(c, i - ax, j - ay, k - az)];
    real s = R(0.);
    for (int n = 0; n < 4; n++) {
        int dx = ISO4_DIAG[a][n][0], dy = ISO4_DIAG[a][n][1], dz = ISO4_DIAG[a][n][2];
        real plus = f[S4(c, i + dx, j + dy, k + dz)], minus = f[S4(c, i - dx, j - dy, k - dz)];
        s = n == 0 ? (plus - minus) : (s + plus - minus);
    }
    return ISO4_0 * axis + ISO4_1 * s;
}
static real iso4_cn(const orc_ctx* c, const real* f, i64 i, i64 j, i64 k, int a) {
    int ax = a == 0, ay = a == 1, az = a == 2;
    real axis = f[S2(c, i + ax, j + ay, k + az)] - f[S2(c, i - ax, j - ay, k - az)];
    real s = R(0.);
    for (int n = 0; n < 4; n++) {
        int dx = ISO4_DIAG[a][n][0], dy = ISO4_DIAG[a][n][1], dz = ISO4_DIAG[a][n][2];
        real plus = f[S2(c, i + dx, j + dy, k + dz)], minus = f[S2(c, i - dx, j - dy, k - dz)];
        s = n == 0 ? (plus - minus) : (s + plus - minus);
    }
    return ISO4_0 * axis + ISO4_1 * s;
}

void orc_normal_directions(orc_ctx* c, i64 ilo, i64 ihi) { /* :757-807, range [-1..n+2] */
    const i64 ny = c->p.ny, nz = c->p.nz;
#pragma omp parallel for schedule(static)
    for (i64 k = -1; k <= nz + 2; k++) for (i64 j = -1; j <= ny + 2; j++) for (i64 i = ilo; i <= ihi; i++) {
        real gx = iso4_phi(c, i, j, k, 0), gy = iso4_phi(c, i, j, k, 1), gz = iso4_phi(c, i, j, k, 2);
        real nrm = M(sqrt)(gx * gx + gy * gy + gz * gz);
        i64 n = S2(c, i, j, k);
        if (nrm < R(1e-6) || c->walls[n] == 1) { c->cn_x[n] = c->cn_y[n] = c->cn_z[n] = R(0.); c->c_norm[n] = R(0.); }
        else { c->cn_x[n] = gx / nrm; c->cn_y[n] = gy / nrm; c->cn_z[n] = gz / nrm; c->c_norm[n] = nrm; }
    }
}

void orc_alter_color_gradient(orc_ctx* c, i64 ilo, i64 ihi) { /* :809-878, range [-1..n+2] */
    const i64 ny = c->p.ny, nz = c->p.nz;
    const real lambda = R(0.5), local_eps = R(1e-6), ct = c->cos_theta;
#pragma omp parallel for schedule(static)
    for (i64 k = -1; k <= nz + 2; k++) for (i64 j = -1; j <= ny + 2; j++) for (i64 i = ilo; i <= ihi; i++) {
        if (c->walls_type[S4(c, i, j, k)] != -1) continue;
        i64 n = S2(c, i, j, k);
        if (!(c->c_norm[n] > local_eps)) continue;
        real nwx = c->s_nx[S4(c, i, j, k)], nwy = c->s_ny[S4(c, i, j, k)], nwz = c->s_nz[S4(c, i, j, k)];
        real vcx0 = c->cn_x[n], vcy0 = c->cn_y[n], vcz0 = c->cn_z[n];
        real vcx1 = vcx0 - lambda * (vcx0 + nwx), vcy1 = vcy0 - lambda * (vcy0 + nwy), vcz1 = vcz0 - lambda * (vcz0 + nwz);
        real vcx2, vcy2, vcz2, err0, err1, err2, tmp;
        err0 = (nwx * vcx0 + nwy * vcy0 + nwz * vcz0) - ct;
        if ((M(fabs)(vcx0 + nwx) + M(fabs)(vcy0 + nwy) + M(fabs)(vcz0 + nwz) > local_eps ||
             M(fabs)(vcx0 - nwx) + M(fabs)(vcy0 - nwy) + M(fabs)(vcz0 - nwz) > local_eps) && err0 > local_eps) {
            err1 = (nwx * vcx1 + nwy * vcy1 + nwz * vcz1) - M(sqrt)(vcx1 * vcx1 + vcy1 * vcy1 + vcz1 * vcz1) * ct;
            tmp = R(1.) / (err1 - err0);
            vcx2 = tmp * (vcx0 * err1 - vcx1 * err0); vcy2 = tmp * (vcy0 * err1 - vcy1 * err0); vcz2 = tmp * (vcz0 * err1 - vcz1 * err0);
            err2 = (nwx * vcx2 + nwy * vcy2 + nwz * vcz2) - M(sqrt)(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2) * ct;
            if (err2 > local_eps) {
                for (int it = 2; it <= 4; it++) {
                    vcx0 = vcx1; vcy0 = vcy1; vcz0 = vcz1;
                    vcx1 = vcx2; vcy1 = vcy2; vcz1 = vcz2;
                    err0 = (nwx * vcx0 + nwy * vcy0 + nwz * vcz0) - M(sqrt)(vcx0 * vcx0 + vcy0 * vcy0 + vcz0 * vcz0) * ct;
                    err1 = (nwx * vcx1 + nwy * vcy1 + nwz * vcz1) - M(sqrt)(vcx1 * vcx1 + vcy1 * vcy1 + vcz1 * vcz1) * ct;
                    tmp = R(1.) / (err1 - err0);
                    if (isinf(tmp)) break;
                    vcx2 = tmp * (vcx0 * err1 - vcx1 * err0); vcy2 = tmp * (vcy0 * err1 - vcy1 * err0); vcz2 = tmp * (vcz0 * err1 - vcz1 * err0);
                    err2 = (nwx * vcx2 + nwy * vcy2 + nwz * vcz2) - M(sqrt)(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2) * ct;
                }
            }
            tmp = R(1.) / ((R(1e-30)) + M(sqrt)(vcx2 * vcx2 + vcy2 * vcy2 + vcz2 * vcz2));
            c->cn_x[n] = vcx2 * tmp; c->cn_y[n] = vcy2 * tmp; c->cn_z[n] = vcz2 * tmp;
        }
    }
}

void orc_extrapolate_normal_to_solid(orc_ctx* c, i64 ilo, i64 ihi) { /* :880-906, range [0..n+1] */
    const i64 ny = c->p.ny, nz = c->p.nz;
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k <= nz + 1; k++) for (i64 j = 0; j <= ny + 1; j++) for (i64 i = ilo; i <= ihi; i++) {
        if (c->walls_type[S4(c, i, j, k)] != 2) continue;
        real sx = R(0.), sy = R(0.), sz = R(0.), wsum = R(0.);
        for (int q = 1; q < 19; q++) {
            i64 ii = i + EX[q], jj = j + EY[q], kk = k + EZ[q];
            if (c->walls_type[S4(c, ii, jj, kk)] <= 0) {
                i64 n = S2(c, ii, jj, kk);
                sx += c->cn_x[n] * wq(q); sy += c->cn_y[n] * wq(q); sz += c->cn_z[n] * wq(q); wsum += wq(q);
            }
        }
        i64 n = S2(c, i, j, k);
        c->cn_x[n] = sx / wsum; c->cn_y[n] = sy / wsum; c->cn_z[n] = sz / wsum;
    }
}

void orc_csf_curvature(orc_ctx* c, i64 ilo, i64 ihi) { /* :908-1003, range [1..n]; pow(x,2) as in :998-1001 */
    const i64 ny = c->p.ny, nz = c->p.nz;
#pragma omp parallel for schedule(static)
    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        real kxx = iso4_cn(c, c->cn_x, i, j, k, 0), kyy = iso4_cn(c, c->cn_y, i, j, k, 1), kzz = iso4_cn(c, c->cn_z, i, j, k, 2);
        real kxy = iso4_cn(c, c->cn_x, i, j, k, 1), kxz = iso4_cn(c, c->cn_x, i, j, k, 2);
        real kyx = iso4_cn(c, c->cn_y, i, j, k, 0), kyz = iso4_cn(c, c->cn_y, i, j, k, 2);
        real kzx = iso4_cn(c, c->cn_z, i, j, k, 0), kzy = iso4_cn(c, c->cn_z, i, j, k, 1);
        i64 n = S2(c, i, j, k);
        real cx = c->cn_x[n], cy = c->cn_y[n], cz = c->cn_z[n];
        c->curv[S1(c, i, j, k)] = (M(pow)(cx, 2) - R(1.)) * kxx + (M(pow)(cy, 2) - R(1.)) * kyy + (M(pow)(cz, 2) - R(1.)) * kzz +
                                  cx * cy * (kxy + kyx) + cx * cz * (kxz + kzx) + cy * cz * (kzy + kyz);
    }
}

void orc_color_gradient(orc_ctx* c) { /* call order: src/main_iteration_GPU.cu:2027-2055 */
    const i64 nx = c->p.nx;
    orc_extrapolate_phi_to_solid(c, -2, nx + 3);
    orc_normal_directions(c, -1, nx + 2);
    orc_alter_color_gradient(c, -1, nx + 2);
    orc_extrapolate_normal_to_solid(c, 0, nx + 1);
    orc_csf_curvature(c, 1, nx);
}

/* =====================================================================================================
 * collision: src/main_iteration_GPU.cu:115-345 (identical in the odd and even kernels, :453-683)
 * in:  g1[19], g2[19] = pre-collision component PDFs in natural direction order
 * out: g1, g2 = post-collision recoloured PDFs; returns phi
 * ===================================================================================================== */
static real collide_node(const orc_ctx* c, real* g1, real* g2, real cnx, real cny, real cnz, real curv, real cnorm) {
    real f[19];
    for (int q = 0; q < 19; q++) f[q] = g1[q] + g2[q];
    real rho1 = g1[0], rho2 = g2[0];
    for (int q = 1; q < 19; q++) { rho1 = rho1 + g1[q]; rho2 = rho2 + g2[q]; }
    const real phi_loc = (rho1 - rho2) / (rho1 + rho2);

    real tmp = R(0.5) * c->p.lbm_gamma * curv * cnorm;
    const real fx = tmp * cnx, fy = tmp * cny, fz = tmp * cnz + c->force_z;

    const real omega = R(1.) / (R(6.) / ((R(1.0) + phi_loc) * c->la_nui1 + (R(1.0) - phi_loc) * c->la_nui2) + R(0.5));
    real s_e, s_e2, s_q, s_pi, s_t;
    const real s_nu = omega;
    switch (c->p.mrt) { /* :157-186 */
        case 1: s_e = omega; s_e2 = omega; s_pi = omega; s_q = R(8.) * (R(2.) - omega) / (R(8.) - omega); s_t = s_q; break;
        case 3: s_e = omega; s_e2 = omega; s_pi = omega; s_q = omega; s_t = omega; break;
        case 4: s_e = omega; s_e2 = omega; s_pi = omega; s_q = (R(6.) - R(3.) * omega) / (R(3.) - omega); s_t = omega; break;
        default: s_e = R(1.19); s_e2 = R(1.4); s_pi = R(1.4); s_q = R(1.2); s_t = R(1.98); break;
    }
    /* includes/Module.h:104-110 */
    const real mrt_coef1 = R(1.) / R(19.), mrt_coef2 = R(1.) / R(2394.), mrt_coef3 = R(1.) / R(252.), mrt_coef4 = R(1.) / R(72.);
    const real mrt_e2_coef1 = R(0.), mrt_e2_coef2 = R(-475.) / R(63.), mrt_omega_xx = R(0.);

    const real den = rho1 + rho2;
    const real ux = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14] + R(0.5) * fx;
    const real uy = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18] + R(0.5) * fy;
    const real uz = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18] + R(0.5) * fz;
    const real u2 = ux * ux + uy * uy + uz * uz;

    real sum1 = f[1] + f[2] + f[3] + f[4] + f[5] + f[6];
    real sum2 = f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    real sum3 = f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    real sum4 = f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    real sum5 = f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    real sum6 = R(2.) * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6];
    real sum7 = f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - R(2.) * (f[15] + f[16] + f[17] + f[18]);
    real sum8 = f[3] + f[4] - f[5] - f[6];
    real sum9 = f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];

    real m_rho = den;
    real m_e = R(-30.) * f[0] - R(11.) * sum1 + R(8.) * sum2;
    real m_e2 = R(12.) * f[0] - R(4.) * sum1 + sum2;
    real m_jx = f[1] - f[2] + sum3;
    real m_qx = R(-4.) * (f[1] - f[2]) + sum3;
    real m_jy = f[3] - f[4] + sum4;
    real m_qy = R(-4.) * (f[3] - f[4]) + sum4;
    real m_jz = f[5] - f[6] + sum5;
    real m_qz = R(-4.) * (f[5] - f[6]) + sum5;
    real m_3pxx = sum6 + sum7;
    real m_3pixx = R(-2.) * sum6 + sum7;
    real m_pww = sum8 + sum9;
    real m_piww = R(-2.) * sum8 + sum9;
    real m_pxy = f[7] - f[8] - f[9] + f[10];
    real m_pyz = f[15] - f[16] - f[17] + f[18];
    real m_pzx = f[11] - f[12] - f[13] + f[14];
    real m_tx = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    real m_ty = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    real m_tz = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];

    /* relaxation in moment space with the forcing terms: :228-246 */
    m_e = m_e - s_e * (m_e - (R(-11.0) * den + R(19.0) * u2)) + (R(38.) - R(19.) * s_e) * (fx * ux + fy * uy + fz * uz);
    m_e2 = m_e2 - s_e2 * (m_e2 - (mrt_e2_coef1 * den + mrt_e2_coef2 * u2)) + (R(-11.) + R(5.5) * s_e2) * (fx * ux + fy * uy + fz * uz);
    m_jx = m_jx + fx;
    m_qx = m_qx - s_q * (m_qx - (R(-0.666666666666666667) * ux)) + (R(-0.666666666666666667) + R(0.333333333333333333) * s_q) * fx;
    m_jy = m_jy + fy;
    m_qy = m_qy - s_q * (m_qy - (R(-0.666666666666666667) * uy)) + (R(-0.666666666666666667) + R(0.333333333333333333) * s_q) * fy;
    m_jz = m_jz + fz;
    m_qz = m_qz - s_q * (m_qz - (R(-0.666666666666666667) * uz)) + (R(-0.666666666666666667) + R(0.333333333333333333) * s_q) * fz;
    m_3pxx = m_3pxx - s_nu * (m_3pxx - (R(3.) * ux * ux - u2)) + (R(2.) - s_nu) * (R(2.) * fx * ux - fy * uy - fz * uz);
    m_3pixx = m_3pixx - s_pi * (m_3pixx - mrt_omega_xx * (R(3.) * ux * ux - u2)) + (R(1.) - R(0.5) * s_pi) * (R(-2.) * fx * ux + fy * uy + fz * uz);
    m_pww = m_pww - s_nu * (m_pww - (uy * uy - uz * uz)) + (R(2.) - s_nu) * (fy * uy - fz * uz);
    m_piww = m_piww - s_pi * (m_piww - mrt_omega_xx * (uy * uy - uz * uz)) + (R(1.) - R(0.5) * s_pi) * (-fy * uy + fz * uz);
    m_pxy = m_pxy - s_nu * (m_pxy - (ux * uy)) + (R(1.) - R(0.5) * s_nu) * (fx * uy + fy * ux);
    m_pyz = m_pyz - s_nu * (m_pyz - (uy * uz)) + (R(1.) - R(0.5) * s_nu) * (fy * uz + fz * uy);
    m_pzx = m_pzx - s_nu * (m_pzx - (ux * uz)) + (R(1.) - R(0.5) * s_nu) * (fx * uz + fz * ux);
    m_tx = m_tx - s_t * (m_tx);
    m_ty = m_ty - s_t * (m_ty);
    m_tz = m_tz - s_t * (m_tz);

    /* back to PDFs: :250-297 */
    m_rho = mrt_coef1 * m_rho;
    m_e = mrt_coef2 * m_e;
    m_e2 = mrt_coef3 * m_e2;
    m_jx = R(0.1) * m_jx; m_qx = R(0.025) * m_qx;
    m_jy = R(0.1) * m_jy; m_qy = R(0.025) * m_qy;
    m_jz = R(0.1) * m_jz; m_qz = R(0.025) * m_qz;
    m_3pxx = R(2.) * mrt_coef4 * m_3pxx;
    m_3pixx = mrt_coef4 * m_3pixx;
    m_pww = R(6.) * mrt_coef4 * m_pww;
    m_piww = R(3.) * mrt_coef4 * m_piww;
    m_pxy = R(0.25) * m_pxy; m_pyz = R(0.25) * m_pyz; m_pzx = R(0.25) * m_pzx;
    m_tx = R(0.125) * m_tx; m_ty = R(0.125) * m_ty; m_tz = R(0.125) * m_tz;
    sum1 = m_rho - R(11.) * m_e - R(4.) * m_e2;
    sum2 = R(2.) * m_3pxx - R(4.) * m_3pixx;
    sum3 = m_pww - R(2.) * m_piww;
    sum4 = m_rho + R(8.) * m_e + m_e2;
    sum5 = m_jx + m_qx;
    sum6 = m_jy + m_qy;
    sum7 = m_jz + m_qz;
    sum8 = m_3pxx + m_3pixx;
    sum9 = m_pww + m_piww;

    f[0] = m_rho - R(30.) * m_e + R(12.) * m_e2;
    f[1] = sum1 + m_jx - R(4.) * m_qx + sum2;
    f[2] = sum1 - m_jx + R(4.) * m_qx + sum2;
    f[3] = sum1 + m_jy - R(4.) * m_qy - R(0.5) * sum2 + sum3;
    f[4] = sum1 - m_jy + R(4.) * m_qy - R(0.5) * sum2 + sum3;
    f[5] = sum1 + m_jz - R(4.) * m_qz - R(0.5) * sum2 - sum3;
    f[6] = sum1 - m_jz + R(4.) * m_qz - R(0.5) * sum2 - sum3;
    f[7] = sum4 + sum5 + sum6 + sum8 + sum9 + m_pxy + m_tx - m_ty;
    f[8] = sum4 - sum5 + sum6 + sum8 + sum9 - m_pxy - m_tx - m_ty;
    f[9] = sum4 + sum5 - sum6 + sum8 + sum9 - m_pxy + m_tx + m_ty;
    f[10] = sum4 - sum5 - sum6 + sum8 + sum9 + m_pxy - m_tx + m_ty;
    f[11] = sum4 + sum5 + sum7 + sum8 - sum9 + m_pzx - m_tx + m_tz;
    f[12] = sum4 - sum5 + sum7 + sum8 - sum9 - m_pzx + m_tx + m_tz;
    f[13] = sum4 + sum5 - sum7 + sum8 - sum9 - m_pzx - m_tx - m_tz;
    f[14] = sum4 - sum5 - sum7 + sum8 - sum9 + m_pzx + m_tx - m_tz;
    f[15] = sum4 + sum6 + sum7 - sum8 * R(2.) + m_pyz + m_ty - m_tz;
    f[16] = sum4 - sum6 + sum7 - sum8 * R(2.) - m_pyz - m_ty - m_tz;
    f[17] = sum4 + sum6 - sum7 - sum8 * R(2.) - m_pyz + m_ty + m_tz;
    f[18] = sum4 - sum6 - sum7 - sum8 * R(2.) + m_pyz - m_ty + m_tz;

    /* recolouring (R-K): :302-345 */
    const real tmp1 = rho1 / den;
    g1[0] = tmp1 * f[0];
    g2[0] = f[0] * (R(1.) - tmp1);
    tmp = rho1 * rho2 * c->p.lbm_beta / den;
    const real RK_weight2 = R(1.) / M(sqrt)(R(2.)) / R(36.); /* includes/Fluid_multiphase.h:32 */
    /* e.cn in the reference's operand order: first listed component signed, second added or subtracted */
    const real ecn[19] = {R(0.), (cnx), (-cnx), (cny), (-cny), (cnz), (-cnz),
                          (cnx + cny), (-cnx + cny), (cnx - cny), (-cnx - cny),
                          (cnx + cnz), (-cnx + cnz), (cnx - cnz), (-cnx - cnz),
                          (cny + cnz), (-cny + cnz), (cny - cnz), (-cny - cnz)};
    for (int q = 1; q < 7; q++) g1[q] = tmp1 * f[q] + W1 * tmp * ecn[q];
    for (int q = 7; q < 19; q++) g1[q] = tmp1 * f[q] + RK_weight2 * tmp * ecn[q];
    for (int q = 1; q < 19; q++) g2[q] = f[q] - g1[q];
    return phi_loc;
}

/* kernel_odd_color_GPU (:56-388) / kernel_even_color_GPU (:395-726), x restricted to [ilo,ihi] */
void orc_collide(orc_ctx* c, int odd, i64 ilo, i64 ihi) {
    const i64 ny = c->p.ny, nz = c->p.nz;
#pragma omp parallel for schedule(static)
    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        if (c->walls[S2(c, i, j, k)] != 0) continue;
        real g1[19], g2[19];
        if (odd) { /* pull g_q from x - e_q, slot q */
            for (int q = 0; q < 19; q++) {
                g1[q] = c->pdf[F1(c, i - EX[q], j - EY[q], k - EZ[q], q, 0)];
                g2[q] = c->pdf[F1(c, i - EX[q], j - EY[q], k - EZ[q], q, 1)];
            }
        } else { /* local slot opc(q) holds f_q */
            for (int q = 0; q < 19; q++) {
                g1[q] = c->pdf[F1(c, i, j, k, OPC[q], 0)];
                g2[q] = c->pdf[F1(c, i, j, k, OPC[q], 1)];
            }
        }
        i64 n2 = S2(c, i, j, k);
        real ph = collide_node(c, g1, g2, c->cn_x[n2], c->cn_y[n2], c->cn_z[n2], c->curv[S1(c, i, j, k)], c->c_norm[n2]);
        c->phi[S4(c, i, j, k)] = ph;
        if (odd) { /* push g_q* to x + e_q, slot opc(q) */
            for (int q = 0; q < 19; q++) {
                c->pdf[F1(c, i + EX[q], j + EY[q], k + EZ[q], OPC[q], 0)] = g1[q];
                c->pdf[F1(c, i + EX[q], j + EY[q], k + EZ[q], OPC[q], 1)] = g2[q];
            }
        } else {
            for (int q = 0; q < 19; q++) { c->pdf[F1(c, i, j, k, q, 0)] = g1[q]; c->pdf[F1(c, i, j, k, q, 1)] = g2[q]; }
        }
    }
}

/* =====================================================================================================
 * inlet / outlet boundary kernels: src/main_iteration_GPU.cu:1009-1524.  x restricted to [ilo,ihi].
 * The five unknown directions at the inlet are q_in = {5,11,12,15,16} (ez=+1); at the outlet their opposites
 * q_out = {6,14,13,18,17}.  "before odd" variants (run after an even step) write the ghost plane the odd pull
 * will read; "after odd" variants patch the swapped slots at k=1 / k=nz.
 * ===================================================================================================== */
static const int QIN[5] = {5, 11, 12, 15, 16};
#define BLEND(newv, oldv, wi) ((newv) * (1 - (wi)) + (oldv) * (wi))

static void inlet_phi(orc_ctx* c, i64 i, i64 j, int wi) { /* :1021-1024 */
    c->phi[S4(c, i, j, 0)] = c->phi_inlet * (1 - wi) + c->phi[S4(c, i, j, 0)] * wi;
    c->phi[S4(c, i, j, -1)] = c->phi[S4(c, i, j, 0)];
    c->phi[S4(c, i, j, -2)] = c->phi[S4(c, i, j, 0)];
    c->phi[S4(c, i, j, -3)] = c->phi[S4(c, i, j, 0)];
}

void orc_inlet_velocity(orc_ctx* c, int after_odd, i64 ilo, i64 ihi) { /* :1009-1077 */
    const i64 ny = c->p.ny;
    for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        int wi = c->walls[S2(c, i, j, 1)];
        inlet_phi(c, i, j, wi);
        real tmp2 = c->W_in[S1(c, i, j, 0)] * c->relaxation;
        real tmp1 = tmp2 * c->p.sa_inject;
        tmp2 = tmp2 - tmp1;
        for (int g = 0; g < 2; g++) {
            real t = g == 0 ? tmp1 : tmp2;
            for (int n = 0; n < 5; n++) {
                int q = QIN[n], o = OPC[q];
                real wgt = n == 0 ? W1 : W2;
                if (!after_odd) { /* ghost (x-e_q, k=0) slot q  <-  (x, k=1) slot opc(q) + 6 w flux */
                    i64 dst = F1(c, i - EX[q], j - EY[q], 0, q, g);
                    c->pdf[dst] = BLEND(c->pdf[F1(c, i, j, 1, o, g)] + R(6.0) * wgt * t, c->pdf[dst], wi);
                } else { /* (x, k=1) slot opc(q)  <-  ghost (x-e_q, k=0) slot q + 6 w flux */
                    i64 dst = F1(c, i, j, 1, o, g);
                    c->pdf[dst] = BLEND(c->pdf[F1(c, i - EX[q], j - EY[q], 0, q, g)] + R(6.0) * wgt * t, c->pdf[dst], wi);
                }
            }
        }
    }
}

/* Zou-He helper: in-plane density sum and transverse corrections for component g at plane k.
 * before odd: the "pull" view (slot q at x-e_q); after odd: the swapped local view (slot opc(q) at x). */
static void zh_inplane(const orc_ctx* c, int after_odd, i64 i, i64 j, i64 k, int g, real v[11]) {
    /* v[q] for q in {0,1,2,3,4,7,8,9,10}: value of f_q arriving at (i,j,k) */
    static const int QS[9] = {0, 1, 2, 3, 4, 7, 8, 9, 10};
    for (int n = 0; n < 9; n++) {
        int q = QS[n];
        v[q] = after_odd ? c->pdf[F1(c, i, j, k, OPC[q], g)] : c->pdf[F1(c, i - EX[q], j - EY[q], k, q, g)];
    }
}

void orc_inlet_pressure(orc_ctx* c, int after_odd, i64 ilo, i64 ihi) { /* :1085-1241 */
    const i64 ny = c->p.ny;
    for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        int wi = c->walls[S2(c, i, j, 1)];
        inlet_phi(c, i, j, wi);
        real tmpRho2 = c->rho_in;
        real tmpRho1 = c->rho_in * c->p.sa_inject;
        tmpRho2 = tmpRho2 - tmpRho1;
        for (int g = 0; g < 2; g++) {
            real v[11];
            zh_inplane(c, after_odd, i, j, 1, g, v);
            /* known outgoing (ez=-1) values f_opc(q), q in QIN, in the reference's order 6,14,13,18,17 */
            real out[5];
            for (int n = 0; n < 5; n++) {
                int q = QIN[n], o = OPC[q];
                out[n] = after_odd ? c->pdf[F1(c, i, j, 1, q, g)]                       /* swapped: slot q holds f_o */
                                   : c->pdf[F1(c, i - EX[o], j - EY[o], 2, o, g)];      /* pull f_o from k=2 */
            }
            real tr = g == 0 ? tmpRho1 : tmpRho2;
            /* operand orders differ between the two variants (:1108-1129 vs :1190-1208): the "after odd" kernel
             * lists local slots 0,2,1,4,3,8,7,10,9, i.e. arriving directions 0,1,2,3,4,9,10,7,8 */
            real inplane = after_odd ? (v[0] + v[1] + v[2] + v[3] + v[4] + v[9] + v[10] + v[7] + v[8])
                                     : (v[0] + v[1] + v[2] + v[3] + v[4] + v[7] + v[8] + v[9] + v[10]);
            real t = (tr - (inplane + R(2.) * (out[0] + out[1] + out[2] + out[3] + out[4]))) * c->relaxation;
            real tnx = after_odd ? R(0.5) * (v[1] + v[9] + v[7] - (v[2] + v[10] + v[8]))
                                 : R(0.5) * (v[1] + v[7] + v[9] - (v[2] + v[8] + v[10]));
            real tny = after_odd ? R(0.5) * (v[3] + v[8] + v[7] - (v[4] + v[9] + v[10]))
                                 : R(0.5) * (v[3] + v[7] + v[8] - (v[4] + v[10] + v[9]));
            /* corrections per unknown q in QIN: 5:+t/3 ; 11: +t/6 - tnx ; 12: +t/6 + tnx ; 15: +t/6 - tny ; 16: +t/6 + tny */
            for (int n = 0; n < 5; n++) {
                int q = QIN[n], o = OPC[q];
                real val;
                if (n == 0) val = out[n] + R(0.333333333333333333) * t;
                else {
                    real corr = (q == 11) ? -tnx : (q == 12) ? tnx : (q == 15) ? -tny : tny;
                    val = out[n] + R(0.166666666666666667) * t + corr;
                }
                i64 dst = after_odd ? F1(c, i, j, 1, o, g) : F1(c, i - EX[q], j - EY[q], 0, q, g);
                c->pdf[dst] = BLEND(val, c->pdf[dst], wi);
            }
        }
    }
}

void orc_outlet_convective(orc_ctx* c, int after_odd, i64 ilo, i64 ihi) { /* :1246-1357 */
    const i64 ny = c->p.ny, nz = c->p.nz;
    for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        real u_convec = c->uin_avg;
        real temp = R(1.) / (R(1.) + u_convec);
        int wi = c->walls[S2(c, i, j, nz)];
        c->phi[S4(c, i, j, nz + 1)] = ((c->phi_convec[S1(c, i, j, 0)] + u_convec * c->phi[S4(c, i, j, nz)]) * temp) * (1 - wi) + c->phi[S4(c, i, j, nz + 1)] * wi;
        c->phi_convec[S1(c, i, j, 0)] = c->phi[S4(c, i, j, nz + 1)];
        c->phi[S4(c, i, j, nz + 2)] = c->phi[S4(c, i, j, nz + 1)];
        c->phi[S4(c, i, j, nz + 3)] = c->phi[S4(c, i, j, nz + 1)];
        c->phi[S4(c, i, j, nz + 4)] = c->phi[S4(c, i, j, nz + 1)];
        for (int g = 0; g < 2; g++) {
            real* buf = g == 0 ? c->f_convec : c->g_convec;
            for (int n = 0; n < 5; n++) {
                int o = OPC[QIN[n]]; /* unknown incoming directions at the outlet: ez = -1 */
                i64 dst, inner;
                if (!after_odd) { dst = F1(c, i - EX[o], j - EY[o], nz + 1, o, g); inner = F1(c, i - EX[o], j - EY[o], nz, o, g); }
                else { dst = F1(c, i, j, nz, OPC[o], g); inner = F1(c, i, j, nz - 1, OPC[o], g); }
                c->pdf[dst] = ((buf[CNV(c, i, j, o)] + u_convec * c->pdf[inner]) * temp) * (1 - wi) + c->pdf[dst] * wi;
            }
            for (int n = 0; n < 5; n++) {
                int o = OPC[QIN[n]];
                i64 dst = !after_odd ? F1(c, i - EX[o], j - EY[o], nz + 1, o, g) : F1(c, i, j, nz, OPC[o], g);
                buf[CNV(c, i, j, o)] = c->pdf[dst];
            }
        }
    }
}

void orc_outlet_pressure(orc_ctx* c, int after_odd, i64 ilo, i64 ihi) { /* :1363-1524 */
    const i64 ny = c->p.ny, nz = c->p.nz;
    for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        int wi = c->walls[S2(c, i, j, nz)];
        for (int d = 1; d <= 4; d++) c->phi[S4(c, i, j, nz + d)] = c->phi[S4(c, i, j, nz)];
        real v[2][11], out[2][5];
        for (int g = 0; g < 2; g++) {
            zh_inplane(c, after_odd, i, j, nz, g, v[g]);
            for (int n = 0; n < 5; n++) { /* known outgoing (ez=+1) f_q, q in QIN: pulled from k=nz-1, or swapped slot opc(q) */
                int q = QIN[n];
                out[g][n] = after_odd ? c->pdf[F1(c, i, j, nz, OPC[q], g)] : c->pdf[F1(c, i - EX[q], j - EY[q], nz - 1, q, g)];
            }
        }
        /* :1380-1408 (before odd) / :1469-1497 (after odd: local slots 0,2,1,4,3,8,7,10,9 = directions 0,1,2,3,4,9,10,7,8) */
        real tmp1;
        if (!after_odd)
            tmp1 = (v[0][0] + v[0][1] + v[0][2] + v[0][3] + v[0][4] + v[0][7] + v[0][8] + v[0][9] + v[0][10] +
                    R(2.) * (out[0][0] + out[0][1] + out[0][2] + out[0][3] + out[0][4]) +
                    v[1][0] + v[1][1] + v[1][2] + v[1][3] + v[1][4] + v[1][7] + v[1][8] + v[1][9] + v[1][10] +
                    R(2.) * (out[1][0] + out[1][1] + out[1][2] + out[1][3] + out[1][4])) - c->rho_out;
        else
            tmp1 = (v[0][0] + v[0][1] + v[0][2] + v[0][3] + v[0][4] + v[0][9] + v[0][10] + v[0][7] + v[0][8] +
                    R(2.) * (out[0][0] + out[0][1] + out[0][2] + out[0][3] + out[0][4]) +
                    v[1][0] + v[1][1] + v[1][2] + v[1][3] + v[1][4] + v[1][9] + v[1][10] + v[1][7] + v[1][8] +
                    R(2.) * (out[1][0] + out[1][1] + out[1][2] + out[1][3] + out[1][4])) - c->rho_out;
        real tmp2 = tmp1 * R(0.5) * (R(1.) - c->phi[S4(c, i, j, nz)]);
        tmp1 = tmp1 - tmp2;
        for (int g = 0; g < 2; g++) {
            const real* vv = v[g];
            real t = g == 0 ? tmp1 : tmp2;
            /* :1414-1419 / :1502-1505 */
            real tnx = after_odd ? R(0.5) * (vv[1] + vv[9] + vv[7] - (vv[2] + vv[10] + vv[8]))
                                 : R(0.5) * (vv[1] + vv[7] + vv[9] - (vv[2] + vv[8] + vv[10]));
            real tny = R(0.5) * (vv[3] + vv[7] + vv[8] - (vv[4] + vv[10] + vv[9]));
            /* unknown o = opc(q): 6: -t/3 ; 13 (opc 12): -t/6 - tnx ; 14 (opc 11): -t/6 + tnx ; 17 (opc 16): -t/6 - tny ; 18 (opc 15): -t/6 + tny */
            for (int n = 0; n < 5; n++) {
                int q = QIN[n], o = OPC[q];
                real val;
                if (n == 0) val = out[g][n] - R(0.333333333333333333) * t;
                else {
                    real corr = (o == 13) ? -tnx : (o == 14) ? tnx : (o == 17) ? -tny : tny;
                    val = out[g][n] - R(0.166666666666666667) * t + corr;
                }
                i64 dst = after_odd ? F1(c, i, j, nz, q, g) : F1(c, i - EX[o], j - EY[o], nz + 1, o, g);
                c->pdf[dst] = BLEND(val, c->pdf[dst], wi);
            }
        }
    }
}

/* =====================================================================================================
 * periodic kernels: src/main_iteration_GPU.cu:1529-1733
 * ===================================================================================================== */
/* direction sets by the sign of one velocity component */
static int comp(int q, int axis) { return axis == 1 ? EY[q] : EZ[q]; }

void orc_periodic_pdf(orc_ctx* c, int axis /*1=y,2=z*/, int odd, i64 ilo, i64 ihi) { /* :1529-1591, :1610-1672 */
    const i64 ny = c->p.ny, nz = c->p.nz;
    const i64 n = axis == 1 ? ny : nz, mlim = axis == 1 ? nz : ny;
    for (i64 m = 1; m <= mlim; m++) for (i64 i = ilo; i <= ihi; i++) for (int g = 0; g < 2; g++) for (int q = 1; q < 19; q++) {
        int s = comp(q, axis);
        if (s == 0) continue;
        /* even: slot q with e=-1 at layer 1 -> ghost n+1 ; slot q with e=+1 at layer n -> ghost 0 ; odd: the reverse copies */
        i64 a = s < 0 ? 1 : n, b = s < 0 ? n + 1 : 0;
        i64 src = odd ? b : a, dst = odd ? a : b;
        if (axis == 1) c->pdf[F1(c, i, dst, m, q, g)] = c->pdf[F1(c, i, src, m, q, g)];
        else c->pdf[F1(c, i, m, dst, q, g)] = c->pdf[F1(c, i, m, src, q, g)];
    }
}

void orc_periodic_pdf_edges(orc_ctx* c, int odd, i64 ilo, i64 ihi) { /* :1593-1608, :1674-1689 */
    const i64 ny = c->p.ny, nz = c->p.nz;
    static const int QE[4] = {18, 16, 17, 15};
    for (i64 i = ilo; i <= ihi; i++) for (int g = 0; g < 2; g++) for (int n = 0; n < 4; n++) {
        int q = QE[n];
        i64 ja = EY[q] < 0 ? 1 : ny, jb = EY[q] < 0 ? ny + 1 : 0;
        i64 ka = EZ[q] < 0 ? 1 : nz, kb = EZ[q] < 0 ? nz + 1 : 0;
        if (!odd) c->pdf[F1(c, i, jb, kb, q, g)] = c->pdf[F1(c, i, ja, ka, q, g)];
        else c->pdf[F1(c, i, ja, ka, q, g)] = c->pdf[F1(c, i, jb, kb, q, g)];
    }
}

void orc_periodic_phi(orc_ctx* c, int which /*1=y,2=z,3=zy edges*/, i64 ilo, i64 ihi) { /* :1691-1733, overlap_phi = 4 */
    const i64 ny = c->p.ny, nz = c->p.nz;
    const int ov = 4;
    if (which == 2) {
        for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) for (int k = 1; k <= ov; k++) {
            c->phi[S4(c, i, j, k + nz)] = c->phi[S4(c, i, j, k)];
            c->phi[S4(c, i, j, k - ov)] = c->phi[S4(c, i, j, nz + k - ov)];
        }
    } else if (which == 1) {
        for (i64 k = 1; k <= nz; k++) for (i64 i = ilo; i <= ihi; i++) for (int j = 1; j <= ov; j++) {
            c->phi[S4(c, i, j + ny, k)] = c->phi[S4(c, i, j, k)];
            c->phi[S4(c, i, j - ov, k)] = c->phi[S4(c, i, ny + j - ov, k)];
        }
    } else {
        for (i64 i = ilo; i <= ihi; i++) for (int k = 1; k <= ov; k++) for (int j = 1; j <= ov; j++) {
            c->phi[S4(c, i, j - ov, k - ov)] = c->phi[S4(c, i, ny + j - ov, nz + k - ov)];
            c->phi[S4(c, i, j + ny, k - ov)] = c->phi[S4(c, i, j, nz + k - ov)];
            c->phi[S4(c, i, j + ny, k + nz)] = c->phi[S4(c, i, j, k)];
            c->phi[S4(c, i, j - ov, k + nz)] = c->phi[S4(c, i, ny + j - ov, k)];
        }
    }
}

/* =====================================================================================================
 * porous plate: src/main_iteration_GPU.cu:1744-1882
 * blocked component: bounce-back across the plate plane zp; the other component: pass-through copies
 * ===================================================================================================== */
void orc_porous_plate(orc_ctx* c, int after_odd, i64 ilo, i64 ihi) {
    const i64 ny = c->p.ny, nz = c->p.nz;
    const int zp = c->p.Z_porous_plate, cmd = c->p.porous_plate_cmd;
    if (!(zp >= 1 && zp <= nz) || (cmd != 1 && cmd != 2)) return;
    const int gb = cmd == 1 ? 0 : 1, gp = 1 - gb; /* blocked / passing component */
    for (i64 j = 1; j <= ny; j++) for (i64 i = ilo; i <= ihi; i++) {
        for (int n = 0; n < 5; n++) {
            int q = QIN[n], o = OPC[q]; /* q: ez=+1, o: ez=-1 */
            if (!after_odd) {
                /* :1758-1768: slot o at (x-e_o, zp) <- slot q at (x, zp-1) ; slot q at (x-e_q, zp) <- slot o at (x, zp+1) */
                c->pdf[F1(c, i - EX[o], j - EY[o], zp, o, gb)] = c->pdf[F1(c, i, j, zp - 1, q, gb)];
                c->pdf[F1(c, i - EX[q], j - EY[q], zp, q, gb)] = c->pdf[F1(c, i, j, zp + 1, o, gb)];
            } else {
                /* :1828-1838 */
                c->pdf[F1(c, i, j, zp - 1, q, gb)] = c->pdf[F1(c, i - EX[o], j - EY[o], zp, o, gb)];
                c->pdf[F1(c, i, j, zp + 1, o, gb)] = c->pdf[F1(c, i - EX[q], j - EY[q], zp, q, gb)];
            }
        }
        for (int n = 0; n < 5; n++) {
            int q = QIN[n], o = OPC[q];
            if (!after_odd) { /* :1770-1780 */
                c->pdf[F1(c, i, j, zp, o, gp)] = c->pdf[F1(c, i, j, zp + 1, o, gp)];
                c->pdf[F1(c, i, j, zp, q, gp)] = c->pdf[F1(c, i, j, zp - 1, q, gp)];
            } else { /* :1840-1850 */
                c->pdf[F1(c, i, j, zp - 1, q, gp)] = c->pdf[F1(c, i, j, zp, q, gp)];
                c->pdf[F1(c, i, j, zp + 1, o, gp)] = c->pdf[F1(c, i, j, zp, o, gp)];
            }
        }
    }
}

/* =====================================================================================================
 * one time step: main_iteration_kernel_GPU, src/main_iteration_GPU.cu:1890-2055
 * ===================================================================================================== */
void orc_step(orc_ctx* c, int ntime) {
    const i64 nx = c->p.nx;
    const int odd = (ntime % 2) != 0;
    orc_collide(c, odd, 1, nx);
    if (c->p.kper) { orc_periodic_pdf(c, 2, odd, 1, nx); orc_periodic_phi(c, 2, 1, nx); }
    if (c->p.jper) { orc_periodic_pdf(c, 1, odd, 1, nx); orc_periodic_phi(c, 1, 1, nx); }
    if (c->p.jper && c->p.kper) { orc_periodic_pdf_edges(c, odd, 1, nx); orc_periodic_phi(c, 3, 1, nx); }
    if (c->p.kper == 0 && c->p.wall_z_min == 0 && c->p.wall_z_max == 0) {
        if (c->p.inlet_BC == 1) orc_inlet_velocity(c, odd, 1, nx);
        else if (c->p.inlet_BC == 2) orc_inlet_pressure(c, odd, 1, nx);
        if (c->p.outlet_BC == 1) orc_outlet_convective(c, odd, 1, nx);
        else if (c->p.outlet_BC == 2) orc_outlet_pressure(c, odd, 1, nx);
    }
    if (c->p.porous_plate_cmd != 0) orc_porous_plate(c, odd, 1, nx);
    orc_color_gradient(c);
}

void orc_run(orc_ctx* c, int ntime_first, int nsteps) {
    for (int n = 0; n < nsteps; n++) orc_step(c, ntime_first + n);
}

/* =====================================================================================================
 * monitor: src/Misc.cpp:222-274 (compute_macro_vars) + src/Monitor.cpp:17-171 (sums), sequential k,j,i order
 * out[0..9] = saturation, saturation_full_domain, vol1_sum, vol2_sum, mass1_sum, mass2_sum, ca, umax_global,
 *             kinetic_energy[0], kinetic_energy[1];  prof (7*nz) = fl1, fl2, pre, mass1, mass2, vol1, vol2 per slice
 * (only meaningful after an even step: PDFs must sit in their natural slots)
 * ===================================================================================================== */
void orc_monitor(orc_ctx* c, double* out, double* prof) {
    const i64 nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        int wi = c->walls[S2(c, i, j, k)];
        real ft[19];
        for (int q = 0; q < 19; q++) ft[q] = c->pdf[F1(c, i, j, k, q, 0)] + c->pdf[F1(c, i, j, k, q, 1)];
        real s = ft[0];
        for (int q = 1; q < 19; q++) s = s + ft[q];
        c->rho[S1(c, i, j, k)] = s * (1 - wi);
        real tmp = R(0.5) * c->p.lbm_gamma * c->curv[S1(c, i, j, k)] * c->c_norm[S2(c, i, j, k)];
        real fx = tmp * c->cn_x[S2(c, i, j, k)], fy = tmp * c->cn_y[S2(c, i, j, k)], fz = tmp * c->cn_z[S2(c, i, j, k)] + c->force_z;
        c->u[S1(c, i, j, k)] = (ft[1] - ft[2] + ft[7] - ft[8] + ft[9] - ft[10] + ft[11] - ft[12] + ft[13] - ft[14] - R(0.5) * fx) * (1 - wi);
        c->v[S1(c, i, j, k)] = (ft[3] - ft[4] + ft[7] + ft[8] - ft[9] - ft[10] + ft[15] - ft[16] + ft[17] - ft[18] - R(0.5) * fy) * (1 - wi);
        c->w[S1(c, i, j, k)] = (ft[5] - ft[6] + ft[11] + ft[12] - ft[13] - ft[14] + ft[15] + ft[16] - ft[17] - ft[18] - R(0.5) * fz) * (1 - wi);
    }
#define PHI_H(i, j, k, wi) (R(0.) * (wi) + c->phi[S4(c, i, j, k)] * (1 - (wi))) /* host-side zeroing in solids, src/Misc.cpp:269 */
    real umax = R(0.), usq1 = R(0.), usq2 = R(0.);
    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        int wi = c->walls[S2(c, i, j, k)];
        real uu = c->u[S1(c, i, j, k)], vv = c->v[S1(c, i, j, k)], ww = c->w[S1(c, i, j, k)];
        real t = (uu * uu + vv * vv + ww * ww) * (1 - wi);
        if (umax < t) umax = t;
        real ph = PHI_H(i, j, k, wi);
        if (ph > R(0.999)) usq1 = usq1 + t;
        else if (ph < R(-0.999)) usq2 = usq2 + t;
    }
    real* P = (real*)calloc(7 * nz, sizeof(real));
    for (i64 k = 1; k <= nz; k++) {
        real t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, prek = 0;
        for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
            int wi = c->walls[S2(c, i, j, k)];
            real ph = PHI_H(i, j, k, wi), rr = c->rho[S1(c, i, j, k)], ww = c->w[S1(c, i, j, k)];
            t3 = t3 + R(0.5) * (R(1.) + ph) * (1 - wi);
            t4 = t4 + R(0.5) * (R(1.) - ph) * (1 - wi);
            t5 = t5 + rr * R(0.5) * (R(1.) + ph) * (1 - wi);
            t6 = t6 + rr * R(0.5) * (R(1.) - ph) * (1 - wi);
            t1 = t1 + ww * R(0.5) * (R(1.) + ph) * (1 - wi);
            t2 = t2 + ww * R(0.5) * (R(1.) - ph) * (1 - wi);
            prek = prek + rr * (1 - wi);
        }
        P[0 * nz + k - 1] = t1; P[1 * nz + k - 1] = t2; P[2 * nz + k - 1] = prek;
        P[3 * nz + k - 1] = t5; P[4 * nz + k - 1] = t6; P[5 * nz + k - 1] = t3; P[6 * nz + k - 1] = t4;
    }
    real m1 = 0, m2 = 0, v1 = 0, v2 = 0;
    for (i64 k = c->p.n_exclude_inlet + 1; k <= nz - c->p.n_exclude_outlet; k++) {
        m1 = m1 + P[3 * nz + k - 1]; m2 = m2 + P[4 * nz + k - 1]; v1 = v1 + P[5 * nz + k - 1]; v2 = v2 + P[6 * nz + k - 1];
    }
    real t3 = 0, t4 = 0;
    for (i64 k = 1; k <= nz; k++) { t3 = t3 + P[5 * nz + k - 1]; t4 = t4 + P[6 * nz + k - 1]; }
    real fl1_avg = 0, fl2_avg = 0;
    for (i64 k = c->p.n_exclude_inlet + 1; k <= nz - c->p.n_exclude_outlet; k++) { fl1_avg = fl1_avg + P[0 * nz + k - 1]; fl2_avg = fl2_avg + P[1 * nz + k - 1]; }
    fl1_avg = fl1_avg / (real)(nz - c->p.n_exclude_outlet - c->p.n_exclude_inlet);
    fl2_avg = fl2_avg / (real)(nz - c->p.n_exclude_outlet - c->p.n_exclude_inlet);
    real fl_avg = fl1_avg + fl2_avg;
    real tt = fl_avg / c->A_xy;
    real ca = tt * c->p.la_nu1 / c->p.lbm_gamma;
    out[0] = (double)(v1 / (v1 + v2));
    out[1] = (double)(t3 / (t3 + t4));
    out[2] = v1; out[3] = v2; out[4] = m1; out[5] = m2; out[6] = ca; out[7] = (double)M(sqrt)(umax);
    out[8] = (double)(R(0.5) * usq1); out[9] = (double)(R(0.5) * usq2);
    if (prof) for (i64 n = 0; n < 7 * nz; n++) prof[n] = (double)P[n];
    free(P);
}

/* cal_saturation: src/Monitor.cpp:472-495 */
double orc_cal_saturation(orc_ctx* c) {
    const i64 nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    real v1 = 0, v2 = 0;
    for (i64 k = 1; k <= nz; k++) for (i64 j = 1; j <= ny; j++) for (i64 i = 1; i <= nx; i++) {
        int wi = c->walls[S2(c, i, j, k)];
        v1 = v1 + R(0.5) * (R(1.) + c->phi[S4(c, i, j, k)]) * (1 - wi);
        v2 = v2 + R(0.5) * (R(1.) - c->phi[S4(c, i, j, k)]) * (1 - wi);
    }
    return (double)(v1 / (v1 + v2 + c->eps));
}
