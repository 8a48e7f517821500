/* oracle/ul_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the
 * product path (wholecellfe_b200/, libwcfe.so).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may load it, and only as the checker / CPU baseline.
 *
 * What it is: a plain-C, single-threaded restatement of the reference's explicit updated-Lagrangian
 * step -- every function main.cpp reaches in wholeCellFE_{k1,k31,kMD}/cellBP/UL_growth_explicit.cpp and
 * the sequencing of cellBP/main.cpp:168-418 with El Topo switched off (perform_improvement 0,
 * collision_safety 0: integrate() == positions <- newpositions, actual_dt == dt).  Each function cites
 * the reference lines it follows (k1 tree unless noted).  Arithmetic is written in the reference's
 * operation order and the file is compiled with -ffp-contract=off, so that the discrete decisions
 * (point-in-triangle, boundary orientation test, Poisson firing, anchor-reset branch, clamps) are taken
 * on bit-identical operands.
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks this file against oracle/_ref (the unmodified reference
 * translation units compiled by oracle/Makefile) where that exists, and tests/test_oracle_golden.py
 * checks it against tests/golden/ (npz files), vectors generated from oracle/_ref by tests/golden/make_golden.py.
 * The reference itself has no tests / golden vectors for this path.
 *
 * Defined behaviour where the reference has UB:
 *   q1  srand(time(NULL)) per step (UL:1460) is dropped; the stream is glibc rand() after srand(seed).
 *   q2  updateFA leaves subs_u uninitialised when no triangle contains the anchor (UL:2018): defined 0.
 *   --  boundary neighbour j1/j2 uninitialised when not found (UL:1474-1475): defined 0.
 */
#include "ul_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* parameters                                                                                       */
/* ------------------------------------------------------------------------------------------------ */

/* kMD/cellBP/UL_growth_explicit.cpp:1208,1211,1214 (identical at :1383-1389 and :2036-2042) */
static const double MD_PX[6] = {0., 2.51250583e-03, 4.40265201e-03, 5.91408228e-03, 11.76817302e-03, 17.001e-03};
static const double MD_PY[6] = {-7.15514497, 71.80501308, 356.62420075, 356.96656556, 126.21811585, 528.40481124};
static const double MD_K[6]  = {31426.8556, 99485.86235, 67080.67086, -7908.39043, 21978.66923, 76858.39733};

void ulo_default_params(ulo_params *p, int variant) {
    memset(p, 0, sizeof *p);
    p->law = variant == 2 ? 1 : 0;
    p->kint_force = variant == 0 ? 1000.0 : 31000.0;                      /* UL:458,559 per variant */
    p->kint_bond = variant == 0 ? 1000.0 : (variant == 1 ? 31000.0 : -1.0);   /* main.cpp:350-354 */
    p->subs_scale = 0.0;                                                  /* main.cpp:299 */
    p->mu_cell = 1000; p->mu_subs = 1000000;                              /* UL:724,727 */
    p->t_act_rate = 200.0; p->t_act_max = 400; p->t_act_t1 = 2;           /* UL:744-756 */
    p->thick = 1; p->kappa = 0.5;                                         /* UL:767,871 */
    p->damp = 0.001; p->rho = 1.0; p->drag = 0.001;                       /* UL:596,597,569 */
    p->rho_i_max = 100;                                                   /* UL:1231 */
    p->k_bend = 20.;                                                      /* UL:1632 */
    p->k_area = 1; p->area_0 = 10 * 10 * 3.1416;                          /* UL:1755-1756 */
    p->ap_rate = 10; p->ap_quantum = 5; p->ap_density = 100;              /* UL:1528,1539,1530 */
    p->k_on = 1; p->ka = 0.004; p->kb = 10; p->Fa = 15; p->Fb = 15;       /* UL:2079-2083 */
    p->pin_limit = 49.00; p->t_i = 3; p->t_f = 8;                         /* UL:298-301 */
}

/* ------------------------------------------------------------------------------------------------ */
/* glibc rand(): TYPE_3 additive feedback generator x^31 + x^3 + 1 (glibc stdlib/random_r.c).         */
/* The reference draws rand()/(RAND_MAX+1.) once per boundary vertex per step (UL:1525) after         */
/* srand(1) (main.cpp:709).                                                                          */
/* ------------------------------------------------------------------------------------------------ */
void ulo_srand(ulo_rng *g, unsigned int seed) {
    int i;
    if (seed == 0) seed = 1;
    g->r[0] = seed;
    for (i = 1; i < 31; i++) {
        /* 16807 * r[i-1] % 2147483647 without overflow (Schrage), as srandom_r does */
        long hi = (int32_t)g->r[i - 1] / 127773, lo = (int32_t)g->r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        g->r[i] = (uint32_t)word;
    }
    g->f = 3; g->b = 0;
    for (i = 0; i < 310; i++) (void)ulo_rand(g);
}
int ulo_rand(ulo_rng *g) {
    uint32_t v;
    g->r[g->f] += g->r[g->b];
    v = g->r[g->f] >> 1;
    g->f = (g->f + 1) % 31; g->b = (g->b + 1) % 31;
    return (int)v;
}

/* ------------------------------------------------------------------------------------------------ */
/* mesh incidence (eltopo3d/nondestructivetrimesh.cpp:138-194 add_triangle, :527-543 boundary flags)  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { long long key; } edgekey;
static int cmp_ll(const void *a, const void *b) {
    long long x = *(const long long *)a, y = *(const long long *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
int ulo_mesh_build(ulo_mesh *m, int nn, int ne, const int *tri) {
    int e, j, v;
    long long *edges;
    m->nn = nn; m->ne = ne;
    m->tri = (int *)malloc(sizeof(int) * 3 * (size_t)(ne > 0 ? ne : 1));
    memcpy(m->tri, tri, sizeof(int) * 3 * (size_t)ne);
    m->ring_ptr = (int *)calloc((size_t)nn + 1, sizeof(int));
    for (e = 0; e < ne; e++) for (j = 0; j < 3; j++) m->ring_ptr[tri[3 * e + j] + 1]++;
    for (v = 0; v < nn; v++) m->ring_ptr[v + 1] += m->ring_ptr[v];
    m->ring = (int *)malloc(sizeof(int) * (size_t)(3 * ne > 0 ? 3 * ne : 1));
    {
        int *fill = (int *)calloc((size_t)nn + 1, sizeof(int));
        for (e = 0; e < ne; e++) for (j = 0; j < 3; j++) { v = tri[3 * e + j]; m->ring[m->ring_ptr[v] + fill[v]++] = e; }
        free(fill);
    }
    /* boundary vertex = endpoint of an edge with an odd number of incident triangles */
    m->is_bnd = (unsigned char *)calloc((size_t)(nn > 0 ? nn : 1), 1);
    edges = (long long *)malloc(sizeof(long long) * (size_t)(3 * ne > 0 ? 3 * ne : 1));
    for (e = 0; e < ne; e++) for (j = 0; j < 3; j++) {
        long long a = tri[3 * e + j], b = tri[3 * e + (j + 1) % 3], t;
        if (a > b) { t = a; a = b; b = t; }
        edges[3 * e + j] = a * (long long)nn + b;
    }
    qsort(edges, (size_t)3 * ne, sizeof(long long), cmp_ll);
    for (e = 0; e < 3 * ne;) {
        int k = e;
        while (k < 3 * ne && edges[k] == edges[e]) k++;
        if ((k - e) % 2 == 1) { m->is_bnd[edges[e] / nn] = 1; m->is_bnd[edges[e] % nn] = 1; }
        e = k;
    }
    free(edges);
    return 0;
}
void ulo_mesh_free(ulo_mesh *m) { free(m->tri); free(m->ring_ptr); free(m->ring); free(m->is_bnd); memset(m, 0, sizeof *m); }

/* ------------------------------------------------------------------------------------------------ */
/* small dense algebra in the order Eigen 3.3.4 evaluates it for fixed 3x3 (Eigen/src/LU/InverseImpl.h, */
/* Determinant.h); matrices are row-major double[9]                                                  */
/* ------------------------------------------------------------------------------------------------ */
#define M(a, i, j) ((a)[3 * (i) + (j)])
static void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double *a) { return sqrt(a[0] * a[0] + (a[1] * a[1] + a[2] * a[2])); }
static double cof(const double *m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(m, i1, j1) * M(m, i2, j2) - M(m, i1, j2) * M(m, i2, j1);
}
static void inv3(const double *m, double *r) {
    double c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
    double det = c0 * M(m, 0, 0) + (c1 * M(m, 1, 0) + c2 * M(m, 2, 0));
    double invdet = 1.0 / det;
    int i, j;
    M(r, 0, 0) = c0 * invdet; M(r, 0, 1) = c1 * invdet; M(r, 0, 2) = c2 * invdet;
    for (i = 1; i < 3; i++) for (j = 0; j < 3; j++) M(r, i, j) = cof(m, j, i) * invdet;
}
static double det3h(const double *m, int a, int b, int c) { return M(m, 0, a) * (M(m, 1, b) * M(m, 2, c) - M(m, 1, c) * M(m, 2, b)); }
static double det3(const double *m) { return det3h(m, 0, 1, 2) - det3h(m, 1, 0, 2) + det3h(m, 2, 0, 1); }
static void matmul3(const double *a, const double *b, double *c) {
    int i, j;
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(c, i, j) = M(a, i, 0) * M(b, 0, j) + M(a, i, 1) * M(b, 1, j) + M(a, i, 2) * M(b, 2, j);
}
static void transpose3(const double *a, double *t) { int i, j; for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(t, i, j) = M(a, j, i); }

/* ------------------------------------------------------------------------------------------------ */
/* evalElementRe: cell overload UL:710-927 (mat 1), substrate overload UL:988-1174 (mat 2).           */
/* The dead diffusion residuals (UL:915-925, assignments commented out at :617-618) are not computed.  */
/* ------------------------------------------------------------------------------------------------ */
void ulo_eval_element(const ulo_params *p, int mat, const double *X9, const double *x9, const double *ea18,
                      double dt, const double *centroid3, double time, double *Re9, double *Re_ea18, double *MMe3) {
    double thetag_dot = 0.0;                                   /* UL:714 */
    double mu = mat == 1 ? p->mu_cell : p->mu_subs;            /* UL:721-728 */
    double t_act = 0;                                          /* UL:742-764 */
    double thick = p->thick;
    double G1[3], G2[3], G1xG2[3], N[3], nG, g1[3], g2[3], gx[3], n[3], ng;
    double eav[6], ea_t[9], aux[9], be_t[9], NdN[9], be_t_s[9], G[9], Ginv[9], G1inv[3], G2inv[3];
    double F[9], Ft[9], C[9], Cinv[9], Cinv_s[9], Spas[9], Sact[9], S[9], tmp[9], be[9], beinv[9], ea_s[9], ea_s_vec[6];
    double theta_e_t, theta_g, theta, theta_e, lamda_N, pres;
    int i, j, k, q, pp;
    if (mat == 1) t_act = time < p->t_act_t1 ? time * p->t_act_rate : p->t_act_max;
    for (k = 0; k < 3; k++) {                                  /* UL:770-771 */
        G1[k] = -1.0 * X9[k] + 1.0 * X9[3 + k];
        G2[k] = -1.0 * X9[k] + 1.0 * X9[6 + k];
    }
    cross3(G1, G2, G1xG2); nG = norm3(G1xG2);
    for (k = 0; k < 3; k++) N[k] = G1xG2[k] / nG;              /* UL:773 */
    for (q = 0; q < 6; q++) eav[q] = 1.0 / 3.0 * ea18[q] + 1.0 / 3.0 * ea18[6 + q] + 1.0 / 3.0 * ea18[12 + q];   /* UL:781 */
    ea_t[0] = eav[0]; ea_t[1] = eav[3]; ea_t[2] = eav[4];      /* UL:784-786 */
    ea_t[3] = eav[3]; ea_t[4] = eav[1]; ea_t[5] = eav[5];
    ea_t[6] = eav[4]; ea_t[7] = eav[5]; ea_t[8] = eav[2];
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(aux, i, j) = (i == j ? 1.0 : 0.0) - 2.0 * M(ea_t, i, j);   /* UL:790 */
    inv3(aux, be_t);                                           /* UL:791 */
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { M(NdN, i, j) = N[i] * N[j]; M(be_t_s, i, j) = M(be_t, i, j) - M(NdN, i, j); }   /* UL:797-798 */
    for (k = 0; k < 3; k++) {                                  /* UL:805-806 */
        g1[k] = -1.0 * x9[k] + 1.0 * x9[3 + k];
        g2[k] = -1.0 * x9[k] + 1.0 * x9[6 + k];
    }
    cross3(g1, g2, gx); ng = norm3(gx);
    for (k = 0; k < 3; k++) n[k] = gx[k] / ng;                 /* UL:808 */
    for (k = 0; k < 3; k++) { M(G, k, 0) = G1[k]; M(G, k, 1) = G2[k]; M(G, k, 2) = N[k]; }   /* UL:812-815 */
    inv3(G, Ginv);                                             /* UL:821 */
    for (k = 0; k < 3; k++) { G1inv[k] = M(Ginv, 0, k); G2inv[k] = M(Ginv, 1, k); }
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(F, i, j) = g1[i] * G1inv[j] + g2[i] * G2inv[j] + n[i] * N[j];   /* UL:826 */
    transpose3(F, Ft); matmul3(Ft, F, C); inv3(C, Cinv);       /* UL:829-830 */
    for (i = 0; i < 9; i++) Cinv_s[i] = Cinv[i] - NdN[i];      /* UL:831 */
    theta_e_t = sqrt(det3(be_t));                              /* UL:845 */
    theta_g = 1.0 + thetag_dot * dt;                           /* UL:846 */
    theta = det3(F);                                           /* UL:848 */
    theta_e = theta * theta_e_t / theta_g;                     /* UL:850 */
    lamda_N = 1.0 / theta_e;                                   /* UL:860 */
    pres = -mu * (lamda_N * lamda_N);                          /* UL:861 */
    for (i = 0; i < 9; i++) Spas[i] = (mu / theta_g) * be_t_s[i] + (theta_g * pres) * Cinv_s[i];   /* UL:864 */
    if (mat == 1) {                                            /* UL:871-877; kappa = 0.5 multiplies the a0 term by 0 (quirk q5) */
        double cen[3], a0[3], na;
        for (k = 0; k < 3; k++) { cen[k] = 1. / 3. * ((x9[k] + x9[3 + k]) + x9[6 + k]); a0[k] = cen[k] - centroid3[k]; }
        na = norm3(a0);
        for (k = 0; k < 3; k++) a0[k] = a0[k] / na;
        for (i = 0; i < 3; i++) for (j = 0; j < 3; j++)
            M(Sact, i, j) = t_act * (p->kappa * (i == j ? 1.0 : 0.0) + ((1 - 2 * p->kappa) * a0[i]) * a0[j]);
    } else {
        for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(Sact, i, j) = t_act * (i == j ? 1.0 : 0.0);   /* UL:1136 */
    }
    for (i = 0; i < 9; i++) S[i] = Sact[i] + Spas[i];          /* UL:878 */
    matmul3(F, be_t, tmp); matmul3(tmp, Ft, be);               /* UL:883 */
    for (i = 0; i < 9; i++) be[i] = (1.0 / (theta_g * theta_g)) * be[i];
    inv3(be, beinv);
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) M(ea_s, i, j) = 0.5 * ((i == j ? 1.0 : 0.0) - M(beinv, i, j));   /* UL:884 */
    ea_s_vec[0] = M(ea_s, 0, 0); ea_s_vec[1] = M(ea_s, 1, 1); ea_s_vec[2] = M(ea_s, 2, 2);        /* UL:886 */
    ea_s_vec[3] = M(ea_s, 0, 1); ea_s_vec[4] = M(ea_s, 0, 2); ea_s_vec[5] = M(ea_s, 1, 2);
    for (i = 0; i < 3; i++) {                                  /* UL:892-913 */
        for (pp = 0; pp < 3; pp++) {
            double temp[9], varg1[3], varg2[3], varF[9], varFt[9], a[9], b[9], varE[9], vv[6], sv[6], dot;
            memset(temp, 0, sizeof temp);
            M(temp, pp, i) = 1.0;
            for (k = 0; k < 3; k++) { varg1[k] = -1.0 * M(temp, k, 0) + M(temp, k, 1); varg2[k] = -1.0 * M(temp, k, 0) + M(temp, k, 2); }
            for (j = 0; j < 3; j++) for (k = 0; k < 3; k++) M(varF, j, k) = varg1[j] * G1inv[k] + varg2[j] * G2inv[k];   /* UL:899 */
            transpose3(varF, varFt); matmul3(varFt, F, a); matmul3(Ft, varF, b);
            for (k = 0; k < 9; k++) varE[k] = 0.5 * (a[k] + b[k]);                                  /* UL:900-901 */
            vv[0] = M(varE, 0, 0); vv[1] = M(varE, 1, 1); vv[2] = M(varE, 2, 2);
            vv[3] = 2 * M(varE, 0, 1); vv[4] = 2 * M(varE, 0, 2); vv[5] = 2 * M(varE, 1, 2);         /* UL:904 */
            sv[0] = M(S, 0, 0); sv[1] = M(S, 1, 1); sv[2] = M(S, 2, 2); sv[3] = M(S, 0, 1); sv[4] = M(S, 0, 2); sv[5] = M(S, 1, 2);
            dot = 0; for (k = 0; k < 6; k++) dot += sv[k] * vv[k];
            Re9[i * 3 + pp] = dot * nG / 2.0 * thick;                                               /* UL:907 */
        }
        for (q = 0; q < 6; q++) Re_ea18[i * 6 + q] = (1. / 3.) * (ea_s_vec[q] - ea18[6 * i + q]) * nG / 2.0 * thick;   /* UL:911 */
        MMe3[i] = (1. / 3.) * nG / 2.0 * thick;                                                     /* UL:913 */
    }
}

/* calculateRes: cell UL:632-707, substrate UL:930-985.  Serial scatter-add in element order. */
void ulo_calculate_res(const ulo_params *p, int mat, const ulo_mesh *m, const double *X, const double *x, const double *ea,
                       double dt, const double *centroid3, double time, double *Res, double *Res_ea, double *MM) {
    int e, i, k;
    memset(Res, 0, sizeof(double) * 3 * (size_t)m->nn);
    memset(Res_ea, 0, sizeof(double) * 6 * (size_t)m->nn);
    memset(MM, 0, sizeof(double) * (size_t)m->nn);
    for (e = 0; e < m->ne; e++) {
        double X9[9], x9[9], ea18[18], Re[9], Rea[18], MMe[3];
        const int *t = m->tri + 3 * e;
        for (i = 0; i < 3; i++) {
            for (k = 0; k < 3; k++) { X9[3 * i + k] = X[3 * t[i] + k]; x9[3 * i + k] = x[3 * t[i] + k]; }
            for (k = 0; k < 6; k++) ea18[6 * i + k] = ea[6 * t[i] + k];
        }
        ulo_eval_element(p, mat, X9, x9, ea18, dt, centroid3, time, Re, Rea, MMe);
        for (i = 0; i < 3; i++) {
            for (k = 0; k < 3; k++) Res[t[i] * 3 + k] += Re[i * 3 + k];
            for (k = 0; k < 6; k++) Res_ea[t[i] * 6 + k] += Rea[i * 6 + k];
            MM[t[i]] += MMe[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Dormant physics: functions the reference defines and whose calls are           */
/* commented out -- restated so that the device versions (wcfe_dormant_field) have a checker.          */
/* ------------------------------------------------------------------------------------------------ */
/* Diffusion residuals of the two integrin species: element part evalElementRe UL:833-842, :915-925 (D_A = 1, D_C = 1e-10,
 * shape-function gradients from the rows of [g1 g2 n]^-1 on the predicted coordinates), serial scatter-add in element order
 * calculateRes UL:702-704.  Their consumers are commented out (UL:617-618). */
void ulo_diffusion_res(const ulo_params *p, const ulo_mesh *m, const double *X, const double *x, const double *cA, const double *cC,
                       double dt, double *Res_Adif, double *Res_Cdif) {
    const double D_A = 1., D_C = 1e-10;                        /* UL:919-920 */
    int e, i, k;
    memset(Res_Adif, 0, sizeof(double) * (size_t)m->nn);
    memset(Res_Cdif, 0, sizeof(double) * (size_t)m->nn);
    for (e = 0; e < m->ne; e++) {
        const int *t = m->tri + 3 * e;
        double G1[3], G2[3], G1xG2[3], nG, g1[3], g2[3], gx[3], ng, n[3], g[9], ginv[9], g1inv[3], g2inv[3], gradA[3], gradC[3], c_A, c_C;
        for (k = 0; k < 3; k++) {
            G1[k] = -1.0 * X[3 * t[0] + k] + 1.0 * X[3 * t[1] + k]; G2[k] = -1.0 * X[3 * t[0] + k] + 1.0 * X[3 * t[2] + k];   /* UL:770-771 */
            g1[k] = -1.0 * x[3 * t[0] + k] + 1.0 * x[3 * t[1] + k]; g2[k] = -1.0 * x[3 * t[0] + k] + 1.0 * x[3 * t[2] + k];   /* UL:805-806 */
        }
        cross3(G1, G2, G1xG2); nG = norm3(G1xG2);
        cross3(g1, g2, gx); ng = norm3(gx);
        for (k = 0; k < 3; k++) n[k] = gx[k] / ng;
        for (k = 0; k < 3; k++) { M(g, k, 0) = g1[k]; M(g, k, 1) = g2[k]; M(g, k, 2) = n[k]; }                                /* UL:816-819 */
        inv3(g, ginv);                                                                                                          /* UL:834 */
        for (k = 0; k < 3; k++) { g1inv[k] = M(ginv, 0, k); g2inv[k] = M(ginv, 1, k); }
        for (k = 0; k < 3; k++) {                                                                                               /* UL:839-840 */
            gradA[k] = cA[t[0]] * g1inv[k] + cA[t[1]] * g2inv[k] + cA[t[2]] * (-g1inv[k] - g2inv[k]);
            gradC[k] = cC[t[0]] * g1inv[k] + cC[t[1]] * g2inv[k] + cC[t[2]] * (-g1inv[k] - g2inv[k]);
        }
        c_A = (1. / 3.) * (cA[t[0]] + cA[t[1]] + cA[t[2]]);                                                                     /* UL:841-842 */
        c_C = (1. / 3.) * (cC[t[0]] + cC[t[1]] + cC[t[2]]);
        for (i = 0; i < 3; i++) {
            double gN[3], dA = 0, dC = 0;
            for (k = 0; k < 3; k++) gN[k] = (i == 0 ? 1.0 : 0.0) * g1inv[k] + (i == 1 ? 1.0 : 0.0) * g2inv[k] + (i == 2 ? 1.0 : 0.0) * (-g1inv[k] - g2inv[k]);   /* UL:917 */
            for (k = 0; k < 3; k++) { dA += gradA[k] * gN[k]; dC += gradC[k] * gN[k]; }
            Res_Adif[t[i]] += (c_A * (1. / 3.) - dt * D_A * dA) * nG / 2.0 * p->thick;                                          /* UL:924 */
            Res_Cdif[t[i]] += (c_C * (1. / 3.) - dt * D_C * dC) * nG / 2.0 * p->thick;                                          /* UL:925 */
        }
    }
}
/* calculatePressureForce UL:1765-1816 (the bulge test's follower load; call commented out at UL:577): -pres (MM/thick) n with the
 * area-weighted vertex normal on x, pres a trapezoid in time, thick the function's own literal 0.1 */
void ulo_force_pressure(const ulo_mesh *m, const double *x, const double *MM, double t_tot, double *F) {
    double pres = 2000.0, thick = 0.1;
    int ni, r, k;
    if (t_tot < 1) pres = 2000.0 * t_tot;
    if (t_tot > 2) pres = 2000.0 * (3 - t_tot);
    if (t_tot > 3) pres = 0.0;
    for (ni = 0; ni < m->nn; ni++) {
        double normal[3] = {0, 0, 0}, Asum = 0.0, nn_;
        for (r = m->ring_ptr[ni]; r < m->ring_ptr[ni + 1]; r++) {
            const int *t = m->tri + 3 * m->ring[r];
            double E1[3], E2[3], nt[3], a, l;
            for (k = 0; k < 3; k++) { E1[k] = x[3 * t[1] + k] - x[3 * t[0] + k]; E2[k] = x[3 * t[2] + k] - x[3 * t[0] + k]; }
            cross3(E1, E2, nt);
            l = norm3(nt); a = 0.5 * l;
            for (k = 0; k < 3; k++) normal[k] += a * (nt[k] / l);
            Asum += a;
        }
        for (k = 0; k < 3; k++) normal[k] = normal[k] / Asum;
        nn_ = norm3(normal);
        for (k = 0; k < 3; k++) normal[k] = normal[k] / nn_;
        for (k = 0; k < 3; k++) F[3 * ni + k] = -pres * (MM[ni] / thick) * normal[k];
    }
}
/* calculateConstraintForce UL:1178-1194 (penalty towards the sphere of radius 350 about the origin; call commented out at UL:551) */
void ulo_force_sphere(int nn, const double *x, double *F) {
    const double radius = 350.0, penalty_dist = 4000.0;
    int ni, k;
    for (ni = 0; ni < nn; ni++) {
        const double *xi = x + 3 * ni;
        double dist = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
        for (k = 0; k < 3; k++) F[3 * ni + k] = penalty_dist * (radius - dist) * (xi[k] / dist);
    }
}

/* calculateAreaTot UL:1641-1667 */
void ulo_area_tot(const ulo_mesh *m, const double *x, double *out4) {
    double area_tot = 0, cen[3] = {0, 0, 0};
    int e, k;
    for (e = 0; e < m->ne; e++) {
        const double *a = x + 3 * m->tri[3 * e], *b = x + 3 * m->tri[3 * e + 1], *c = x + 3 * m->tri[3 * e + 2];
        double e1[3], e2[3], cr[3], av[3], ce[3], area;
        for (k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
        cross3(e1, e2, cr);
        for (k = 0; k < 3; k++) { av[k] = 0.5 * cr[k]; ce[k] = 1. / 3. * ((a[k] + b[k]) + c[k]); }
        area = norm3(av);
        for (k = 0; k < 3; k++) cen[k] += ce[k] * area;
        area_tot += area;
    }
    out4[0] = area_tot;
    for (k = 0; k < 3; k++) out4[1 + k] = cen[k] / area_tot;
}

/* nodal area: one third of the one-ring triangle areas on x (UL:1239-1252, repeated at :1351-1364) */
static double node_area(const ulo_mesh *m, const double *x, int ni) {
    double Asum = 0.0;
    int r, k;
    for (r = m->ring_ptr[ni]; r < m->ring_ptr[ni + 1]; r++) {
        const int *t = m->tri + 3 * m->ring[r];
        double E1[3], E2[3], nrm[3];
        for (k = 0; k < 3; k++) { E1[k] = x[3 * t[1] + k] - x[3 * t[0] + k]; E2[k] = x[3 * t[2] + k] - x[3 * t[0] + k]; }
        cross3(E1, E2, nrm);
        Asum += 0.5 * norm3(nrm) / 3;
    }
    return Asum;
}

/* the MD force law of the kMD variant: kMD UL:1246-1265 (same text at :1395-1416 and :2113-2135) */
static void md_force(const double *u, double *force_s) {
    double D = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double uD[3];
    int j, k;
    if (D == 0) { uD[0] = uD[1] = uD[2] = 0; } else { for (k = 0; k < 3; k++) uD[k] = u[k] / D; }
    force_s[0] = force_s[1] = force_s[2] = 0;       /* reference: uninitialised if no branch fires (D NaN) */
    for (j = 0; j < 5; j++) {
        if (D < MD_PX[j + 1] && D >= MD_PX[j]) {
            for (k = 0; k < 3; k++) force_s[k] = uD[k] * (MD_PY[j] + MD_K[j] * (D - MD_PX[j]));
        } else if (D >= MD_PX[5]) {
            for (k = 0; k < 3; k++) force_s[k] = uD[k] * (MD_PY[5] + MD_K[5] * (D - MD_PX[5]));
        }
    }
}

/* calculateForce_FA UL:1196-1295 (kMD: kMD UL:1198-1303) */
void ulo_force_FA(const ulo_params *p, const ulo_mesh *m, const double *x, const double *cC, const double *fa_u, double *F) {
    int ni, k;
    memset(F, 0, sizeof(double) * 3 * (size_t)m->nn);
    for (ni = 0; ni < m->nn; ni++) {
        double Asum = node_area(m, x, ni);
        const double *u = fa_u + 3 * ni;
        if (p->law == 0) {
            for (k = 0; k < 3; k++) F[3 * ni + k] = -cC[ni] * p->kint_force * u[k] * Asum * p->rho_i_max;   /* UL:1259 */
        } else {
            double fs[3];
            md_force(u, fs);
            for (k = 0; k < 3; k++) F[3 * ni + k] = -cC[ni] * fs[k] * Asum * p->rho_i_max;                  /* kMD UL:1267 */
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* point-in-triangle with the reference's sign conventions (UL:1327-1340 == UL:2042-2057).            */
/* Returns 1 when A1>=0 && A2>=0 && A3>=0 and the three sub-areas / A_tri in w.                       */
/* ------------------------------------------------------------------------------------------------ */
static int contains(const double *n1, const double *n2, const double *n3, const double *pt, double *A, double *A_tri_out) {
    double e1[3], e2[3], ep[3], nrm[3], A3n[3], A2n[3], A_tri, A3 = -1, A2 = -1, A1;
    int k;
    for (k = 0; k < 3; k++) { e1[k] = n2[k] - n1[k]; e2[k] = n3[k] - n1[k]; }
    cross3(e1, e2, nrm);
    A_tri = 0.5 * norm3(nrm);
    for (k = 0; k < 3; k++) ep[k] = pt[k] - n1[k];
    cross3(e1, ep, A3n);
    if (A3n[2] >= 0) A3 = 0.5 * norm3(A3n);
    cross3(ep, e2, A2n);
    if (A2n[2] >= 0) A2 = 0.5 * norm3(A2n);
    A1 = A_tri - A2 - A3;
    A[0] = A1; A[1] = A2; A[2] = A3; *A_tri_out = A_tri;
    return A1 >= 0 && A2 >= 0 && A3 >= 0;
}

/* uniform grid over substrate triangles: only a candidate filter; predicate and visiting order
   (ascending triangle index) are those of the brute-force loops, so results are identical */
typedef struct { int nx, ny; double x0, y0, h; int *ptr; int *items; } ugrid;
static void grid_free(ugrid *g) { if (g) { free(g->ptr); free(g->items); free(g); } }
static ugrid *grid_build(const ulo_mesh *m, const double *x) {
    ugrid *g = (ugrid *)calloc(1, sizeof *g);
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300, eps;
    int i, e, pass;
    for (i = 0; i < m->nn; i++) {
        if (x[3 * i] < xmin) xmin = x[3 * i];
        if (x[3 * i] > xmax) xmax = x[3 * i];
        if (x[3 * i + 1] < ymin) ymin = x[3 * i + 1];
        if (x[3 * i + 1] > ymax) ymax = x[3 * i + 1];
    }
    g->h = sqrt(((xmax - xmin) * (ymax - ymin)) / (m->ne > 0 ? m->ne : 1)) * 1.5;
    if (!(g->h > 0)) g->h = 1;
    eps = 1e-9 * ((xmax - xmin) + (ymax - ymin) + 1.0);
    g->x0 = xmin - 2 * eps; g->y0 = ymin - 2 * eps;
    g->nx = (int)((xmax - g->x0 + 2 * eps) / g->h) + 1; g->ny = (int)((ymax - g->y0 + 2 * eps) / g->h) + 1;
    g->ptr = (int *)calloc((size_t)g->nx * g->ny + 1, sizeof(int));
    for (pass = 0; pass < 2; pass++) {
        int *fill = pass ? (int *)calloc((size_t)g->nx * g->ny, sizeof(int)) : NULL;
        for (e = 0; e < m->ne; e++) {
            const int *t = m->tri + 3 * e;
            double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
            int ix0, ix1, iy0, iy1, ix, iy, j;
            for (j = 0; j < 3; j++) {
                double px = x[3 * t[j]], py = x[3 * t[j] + 1];
                if (px < bx0) bx0 = px;
                if (px > bx1) bx1 = px;
                if (py < by0) by0 = py;
                if (py > by1) by1 = py;
            }
            ix0 = (int)((bx0 - eps - g->x0) / g->h); ix1 = (int)((bx1 + eps - g->x0) / g->h);
            iy0 = (int)((by0 - eps - g->y0) / g->h); iy1 = (int)((by1 + eps - g->y0) / g->h);
            if (ix0 < 0) ix0 = 0;
            if (iy0 < 0) iy0 = 0;
            if (ix1 >= g->nx) ix1 = g->nx - 1;
            if (iy1 >= g->ny) iy1 = g->ny - 1;
            for (iy = iy0; iy <= iy1; iy++) for (ix = ix0; ix <= ix1; ix++) {
                int c = iy * g->nx + ix;
                if (!pass) g->ptr[c + 1]++; else g->items[g->ptr[c] + fill[c]++] = e;
            }
        }
        if (!pass) {
            int c;
            for (c = 0; c < g->nx * g->ny; c++) g->ptr[c + 1] += g->ptr[c];
            g->items = (int *)malloc(sizeof(int) * (size_t)(g->ptr[g->nx * g->ny] > 0 ? g->ptr[g->nx * g->ny] : 1));
        } else free(fill);
    }
    return g;
}
static int grid_candidates(const ugrid *g, const double *pt, const int **items) {
    int ix, iy, c;
    double fx = (pt[0] - g->x0) / g->h, fy = (pt[1] - g->y0) / g->h;
    if (!(fx >= 0) || !(fy >= 0) || fx >= g->nx || fy >= g->ny) return 0;
    ix = (int)fx; iy = (int)fy; c = iy * g->nx + ix;
    *items = g->items + g->ptr[c];
    return g->ptr[c + 1] - g->ptr[c];
}

/* calculateForce_FA_subs UL:1297-1454 (kMD: kMD UL:1305-1475).  Current substrate coordinates; every
   containing triangle contributes (no break); scatter-add order = ascending cell node, then triangle. */
void ulo_force_FA_subs(ulo_sim *s, double *F) {
    const ulo_params *p = &s->p;
    const ulo_mesh *ms = &s->subs.m, *mc = &s->cell.m;
    const double *xs = s->subs.x, *xc = s->cell.x;
    ugrid *g = NULL;
    int ni, k, c;
    memset(F, 0, sizeof(double) * 3 * (size_t)ms->nn);
    if (s->use_grid) g = grid_build(ms, xs);
    for (ni = 0; ni < mc->nn; ni++) {
        double pt[3];
        const int *cand = NULL;
        int ncand = ms->ne;
        for (k = 0; k < 3; k++) pt[k] = xc[3 * ni + k] - s->cell.fa_u[3 * ni + k];   /* UL:1303 */
        if (g) ncand = grid_candidates(g, pt, &cand);
        s->loc_force_cnt[ni] = 0; s->loc_force_first[ni] = -1;
        for (c = 0; c < ncand; c++) {
            int ei = g ? cand[c] : c;
            const int *t = ms->tri + 3 * ei;
            double A[3], A_tri;
            if (contains(xs + 3 * t[0], xs + 3 * t[1], xs + 3 * t[2], pt, A, &A_tri)) {
                const double *u = s->cell.fa_u + 3 * ni;
                double Asum = node_area(mc, xc, ni);
                double fs[3];
                int j;
                if (s->loc_force_cnt[ni]++ == 0) s->loc_force_first[ni] = ei;
                if (p->law == 1) md_force(u, fs);
                for (j = 0; j < 3; j++) for (k = 0; k < 3; k++) {
                    if (p->law == 0) F[t[j] * 3 + k] += s->cell.cC[ni] * p->kint_force * u[k] * (A[j] / A_tri) * Asum * p->rho_i_max;   /* UL:1414-1422 */
                    else F[t[j] * 3 + k] += s->cell.cC[ni] * fs[k] * (A[j] / A_tri) * Asum * p->rho_i_max;                               /* kMD UL:1457-1465 */
                }
            }
        }
    }
    grid_free(g);
}

/* boundary neighbours j1 (start) / j2 (end) of boundary vertex ni: UL:1479-1507 (== :1571-1599 == :1690-1723).
   Later matches overwrite earlier ones, exactly as the reference loop does. */
static void bnd_neighbours(const ulo_mesh *m, const double *x, int ni, double *j1, double *j2) {
    int r, enj, enk, k;
    const double *xi = x + 3 * ni;
    for (k = 0; k < 3; k++) { j1[k] = 0; j2[k] = 0; }
    for (r = m->ring_ptr[ni]; r < m->ring_ptr[ni + 1]; r++) {
        const int *t = m->tri + 3 * m->ring[r];
        for (enj = 0; enj < 3; enj++) {
            int nj = t[enj];
            if (m->is_bnd[nj] && nj != ni) {
                int nint = ni;
                double a[3], b[3], nc[3];
                for (enk = 0; enk < 3; enk++) if (t[enk] != nj && t[enk] != ni) { nint = t[enk]; break; }
                for (k = 0; k < 3; k++) { a[k] = xi[k] - x[3 * nj + k]; b[k] = x[3 * nint + k] - xi[k]; }
                cross3(a, b, nc);                                          /* UL:1497 */
                if (nc[2] > 0) { for (k = 0; k < 3; k++) j1[k] = x[3 * nj + k]; }
                else { for (k = 0; k < 3; k++) j2[k] = x[3 * nj + k]; }
            }
        }
    }
}
/* outward normal and edge lengths at a boundary vertex: UL:1512-1520 (== :1624-1629 == :1747-1752) */
static void bnd_normal(const double *xi, const double *j1, const double *j2, double *normal_i, double *total_length) {
    double d1[3], d2[3], n1[3], n2[3], me3[3] = {-0.0, -0.0, -1.0}, nn1, nn2, l1, l2, tl;
    int k;
    for (k = 0; k < 3; k++) { d1[k] = xi[k] - j1[k]; d2[k] = j2[k] - xi[k]; }
    cross3(me3, d1, n1); nn1 = norm3(n1);
    for (k = 0; k < 3; k++) n1[k] = n1[k] / nn1;
    cross3(me3, d2, n2); nn2 = norm3(n2);
    for (k = 0; k < 3; k++) n2[k] = n2[k] / nn2;
    l1 = norm3(d1); l2 = norm3(d2); tl = l1 + l2;
    for (k = 0; k < 3; k++) normal_i[k] = l1 / tl * n1[k] + l2 / tl * n2[k];
    *total_length = tl;
}

/* the two boundary neighbours of a boundary vertex of the cell as the force routines find them (UL:1479-1507), for the CPU tests of
 * the device's boundary math; returns 0 for an interior vertex */
int ulo_boundary_neighbours(ulo_sim *s, int ni, double *j1, double *j2) {
    const ulo_mesh *m = &s->cell.m;
    if (ni < 0 || ni >= m->nn || !m->is_bnd[ni]) return 0;
    bnd_neighbours(m, s->cell.x, ni, j1, j2);
    return 1;
}

/* calculateForce_AP UL:1456-1547; one rand() per boundary vertex in ascending vertex order */
void ulo_force_AP(ulo_sim *s, const double *x, double dt, double *F) {
    const ulo_mesh *m = &s->cell.m;
    const ulo_params *p = &s->p;
    int ni, k, ii;
    memset(F, 0, sizeof(double) * 3 * (size_t)m->nn);
    memset(s->ap_fire, 0, (size_t)m->nn);
    for (ni = 0; ni < m->nn; ni++) {
        double j1[3], j2[3], normal_i[3], tl, Pr, factin = 0;
        if (!m->is_bnd[ni]) continue;
        bnd_neighbours(m, x, ni, j1, j2);
        bnd_normal(x + 3 * ni, j1, j2, normal_i, &tl);
        Pr = ((double)ulo_rand(&s->rng) / ((double)2147483647 + 1.));      /* UL:1525 */
        for (ii = 0; ii < p->ap_density * tl / 2; ii++) {                  /* UL:1530 */
            if (Pr < p->ap_rate * dt * exp(-p->ap_rate * dt)) { factin += p->ap_quantum; s->ap_fire[ni] = 1; }
        }
        for (k = 0; k < 3; k++) F[3 * ni + k] = factin * normal_i[k];
    }
}

/* calculateBendingForce2D UL:1550-1639 */
void ulo_force_bending(const ulo_mesh *m, const double *x, double k_bend, double *F) {
    int ni, k;
    memset(F, 0, sizeof(double) * 3 * (size_t)m->nn);
    for (ni = 0; ni < m->nn; ni++) {
        double j1[3], j2[3], normal_i[3], tl, ed1[3], ed2[3], n1, n2, costurn, e3s[3], turna, kap;
        const double *xi = x + 3 * ni;
        if (!m->is_bnd[ni]) continue;
        bnd_neighbours(m, x, ni, j1, j2);
        for (k = 0; k < 3; k++) { ed1[k] = xi[k] - j1[k]; ed2[k] = j2[k] - xi[k]; }
        n1 = norm3(ed1); n2 = norm3(ed2);
        for (k = 0; k < 3; k++) { ed1[k] = ed1[k] / n1; ed2[k] = ed2[k] / n2; }
        costurn = ed1[0] * ed2[0] + (ed1[1] * ed2[1] + ed1[2] * ed2[2]);
        cross3(ed1, ed2, e3s);
        turna = e3s[2] > 0 ? costurn - 1 : 1 - costurn;                    /* UL:1611-1615 */
        bnd_normal(xi, j1, j2, normal_i, &tl);
        kap = turna * 2 / tl;                                              /* UL:1620 */
        for (k = 0; k < 3; k++) F[3 * ni + k] = k_bend * kap * normal_i[k];
    }
}

/* calculateAreaConstraintForce UL:1670-1763 */
void ulo_force_area(const ulo_params *p, const ulo_mesh *m, const double *x, double area_tot, double *F) {
    int ni, k;
    for (ni = 0; ni < m->nn; ni++) {
        double j1[3], j2[3], normal_i[3], tl;
        if (!m->is_bnd[ni]) continue;
        bnd_neighbours(m, x, ni, j1, j2);
        bnd_normal(x + 3 * ni, j1, j2, normal_i, &tl);
        for (k = 0; k < 3; k++) F[3 * ni + k] = -p->k_area * (area_tot - p->area_0) / 2 * tl * normal_i[k];   /* UL:1757 */
    }
}

/* updateFA (11-argument overload) UL:2005-2114 (kMD: kMD UL:2029-2174) */
void ulo_update_FA(ulo_sim *s, double dt) {
    const ulo_params *p = &s->p;
    const ulo_mesh *ms = &s->subs.m;
    const double *Xs = s->subs.X, *xs = s->subs.x;
    ugrid *g = NULL;
    int ni, k, c, nn = s->cell.m.nn;
    if (s->use_grid) g = grid_build(ms, Xs);
    for (ni = 0; ni < nn; ni++) {
        double cell_u[3], subs_u[3] = {0, 0, 0}, pt[3], *fa = s->cell.fa_u + 3 * ni, force_FA, k_off, c_old, surv;
        const int *cand = NULL;
        int ncand = ms->ne;
        for (k = 0; k < 3; k++) { cell_u[k] = s->cell.x[3 * ni + k] - s->cell.X[3 * ni + k]; pt[k] = s->cell.X[3 * ni + k] - fa[k]; }   /* UL:2016,2020 */
        if (g) ncand = grid_candidates(g, pt, &cand);
        s->loc_bond[ni] = -1;
        for (c = 0; c < ncand; c++) {                                       /* UL:2026-2065: reference coords, last match wins */
            int ei = g ? cand[c] : c;
            const int *t = ms->tri + 3 * ei;
            double A[3], A_tri;
            if (contains(Xs + 3 * t[0], Xs + 3 * t[1], Xs + 3 * t[2], pt, A, &A_tri)) {
                for (k = 0; k < 3; k++)
                    subs_u[k] = (A[0] / A_tri) * (xs[3 * t[0] + k] - Xs[3 * t[0] + k]) + (A[1] / A_tri) * (xs[3 * t[1] + k] - Xs[3 * t[1] + k])
                              + (A[2] / A_tri) * (xs[3 * t[2] + k] - Xs[3 * t[2] + k]);           /* UL:2063 */
                s->loc_bond[ni] = ei;
            }
        }
        for (k = 0; k < 3; k++) fa[k] = fa[k] + cell_u[k] - subs_u[k];      /* UL:2072 */
        if (p->law == 0) {
            force_FA = p->kint_bond * sqrt(fa[0] * fa[0] + fa[1] * fa[1]);  /* UL:2077 (2-D norm, quirk q4) */
        } else {
            double fs[3];
            md_force(fa, fs);
            force_FA = sqrt(fs[0] * fs[0] + fs[1] * fs[1] + fs[2] * fs[2]); /* kMD UL:2137 */
        }
        k_off = p->ka * exp(force_FA / p->Fa) + p->kb * exp(-force_FA / p->Fb);   /* UL:2084 */
        c_old = s->cell.cC[ni];
        s->cell.cC[ni] = s->cell.cC[ni] + dt * p->k_on * (1 - s->cell.cC[ni]) - dt * k_off * s->cell.cC[ni];   /* UL:2088 */
        s->cc_clamp[ni] = 0;
        if (s->cell.cC[ni] < 0) { s->cell.cC[ni] = 0; s->cc_clamp[ni] = 1; }        /* UL:2095 */
        if (s->cell.cA[ni] < 0) s->cell.cA[ni] = 0;                                 /* UL:2096 */
        surv = c_old - dt * k_off * c_old;
        if (surv > 0 && s->cell.cC[ni] > 0) {                                       /* UL:2107-2111 */
            for (k = 0; k < 3; k++) fa[k] = fa[k] * surv / s->cell.cC[ni];
            s->fa_reset[ni] = 0;
        } else {
            for (k = 0; k < 3; k++) fa[k] = 0.0 * fa[k];
            s->fa_reset[ni] = 1;
        }
    }
    grid_free(g);
}

/* driverVelocities_Explicit UL:103-124 */
void ulo_driver_explicit(int nn, const double *V, const double *A, double dt, double *Vh) {
    int ni, k;
    for (ni = 0; ni < nn; ni++) {
        for (k = 0; k < 3; k++) Vh[3 * ni + k] = V[3 * ni + k] + (dt / 2.0) * A[3 * ni + k];
        Vh[3 * ni + 2] = 0.0;
    }
}
/* driverVelocities_Biaxial UL:275-345 (Vbiax_x = Vbiax_y = 0) */
void ulo_driver_biaxial(const ulo_params *p, int nn, const double *X, const double *V, const double *A, double dt, double t_tot, double *Vh) {
    double Vbx = 0.0, Vby = 0.0, t_i = p->t_i, t_f = p->t_f, L = p->pin_limit;
    int ni, k;
    for (ni = 0; ni < nn; ni++) {
        const double *Xi = X + 3 * ni;
        double *v = Vh + 3 * ni;
        for (k = 0; k < 3; k++) v[k] = V[3 * ni + k] + (dt / 2.0) * A[3 * ni + k];
        v[2] = 0.0;
        if (t_tot < t_i) {
            if (Xi[0] < -L) v[0] = 0.0;
            if (Xi[0] > L) v[0] = 0.0;
            if (Xi[1] > L) v[1] = 0.0;
            if (Xi[1] < -L) v[1] = 0.0;
        }
        if (t_tot > t_i && t_tot < t_f) {
            if (Xi[0] < -L - (t_tot - t_i) * Vbx) v[0] = -Vbx;
            if (Xi[0] > L + (t_tot - t_i) * Vbx) v[0] = Vbx;
            if (Xi[1] > L + (t_tot - t_i) * Vby) v[1] = Vby;
            if (Xi[1] < -L - (t_tot - t_i) * Vby) v[1] = -Vby;
        }
        if (t_tot > t_f) {
            if (Xi[0] < -L - (t_f - t_i) * Vbx) v[0] = 0.0;
            if (Xi[0] > L + (t_f - t_i) * Vbx) v[0] = 0.0;
            if (Xi[1] > L + (t_f - t_i) * Vby) v[1] = 0.0;
            if (Xi[1] < -L - (t_f - t_i) * Vby) v[1] = 0.0;
        }
    }
}

/* The drivers the reference defines and main.cpp leaves commented out (main.cpp:217-220, :291-295); identical in the three variants.
 * kind: 2 driverVelocities_Uniaxial UL:125-168, 3 driverVelocities_OffBiaxial UL:214-273, 4 driverVelocities_BulgeTest UL:170-212;
 * 0 / 1 dispatch to the two drivers main.cpp calls.  Thresholds are the reference's literals. */
int ulo_driver_kind(const ulo_params *p, int kind, int nn, const double *X, const double *V, const double *A, double dt, double t_tot, double *Vh) {
    int ni, k;
    if (kind == 0) { ulo_driver_explicit(nn, V, A, dt, Vh); return 0; }
    if (kind == 1) { ulo_driver_biaxial(p, nn, X, V, A, dt, t_tot, Vh); return 0; }
    if (kind < 0 || kind > 4) return -1;
    for (ni = 0; ni < nn; ni++) {
        const double *Xi = X + 3 * ni;
        double *v = Vh + 3 * ni;
        for (k = 0; k < 3; k++) v[k] = V[3 * ni + k] + (dt / 2.0) * A[3 * ni + k];
        if (kind == 2) {                                                        /* UL:141-164 */
            v[2] = 0.0;
            if (Xi[0] < -0.0499) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
            if (Xi[0] > 0.9499 + 0.1 * t_tot && t_tot <= 2.0) { v[0] = 0.1; v[1] = 0.0; v[2] = 0.0; }
            if (Xi[0] > 0.9499 + 2.00 * 0.1 && t_tot >= 2.0) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
        } else if (kind == 3) {                                                 /* UL:230-269 */
            v[2] = 0.0;
            if (Xi[0] < -0.04) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
            if (Xi[0] > 0.94 + 0.1 * t_tot && t_tot < 1.0) { v[0] = 0.1; v[1] = 0.0; v[2] = 0.0; }
            if (Xi[0] > 0.94 + 0.99 * 0.1 && t_tot >= 1.0) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
            if (Xi[1] > 0.99) { v[1] = 0.0; v[2] = 0.0; }
            if (Xi[1] < 0.01) { v[1] = 0.0; v[2] = 0.0; }
        } else {                                                                /* UL:184-209: the normal component is kept */
            const int mid = Xi[2] < 0.001 && Xi[2] > -0.001;
            if ((Xi[0] < 0.001 && mid) || (Xi[0] > 0.999 && mid) || (Xi[1] < 0.001 && mid) || (Xi[1] > 0.999 && mid)) { v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; }
        }
    }
    return 0;
}

/* nodal update shared by both updateForces_Explicit overloads: UL:601-628 / UL:478-503 */
static void nodal_update(const ulo_params *p, int nn, const double *f, const double *MM, const double *Res_ea,
                         double *A, double *V, const double *Vh, double *ea, double dt, double dt_real) {
    int ni, k;
    for (ni = 0; ni < nn; ni++) {
        for (k = 0; k < 3; k++) A[3 * ni + k] = (f[3 * ni + k] - p->damp * MM[ni] * p->rho * V[3 * ni + k]) / (MM[ni] * p->rho);
        for (k = 0; k < 6; k++) ea[6 * ni + k] += Res_ea[6 * ni + k] / MM[ni];
    }
    for (ni = 0; ni < nn; ni++) for (k = 0; k < 3; k++) {
        if (dt_real > dt / 2.0) V[3 * ni + k] = Vh[3 * ni + k] + (dt_real - dt / 2.0) * A[3 * ni + k];
        else V[3 * ni + k] = V[3 * ni + k] + dt_real * A[3 * ni + k];
    }
}
static double norm_n(const double *a, int n) { double s = 0; int i; for (i = 0; i < n; i++) s += a[i] * a[i]; return sqrt(s); }

/* updateForces_Explicit, cell overload UL:508-630 */
static void update_forces_cell(ulo_sim *s, double dt, double dt_real, double t_tot) {
    ulo_surface *c = &s->cell;
    const ulo_params *p = &s->p;
    int nn = c->m.nn, i;
    double area[4], *Res, *Rea, *MM, *F_FA, *F_AP, *F_vd, *F_b, *F_a, *f;
    Res = (double *)malloc(sizeof(double) * 3 * nn); Rea = (double *)malloc(sizeof(double) * 6 * nn); MM = (double *)malloc(sizeof(double) * nn);
    F_FA = (double *)malloc(sizeof(double) * 3 * nn); F_AP = (double *)malloc(sizeof(double) * 3 * nn); F_vd = (double *)malloc(sizeof(double) * 3 * nn);
    F_b = (double *)malloc(sizeof(double) * 3 * nn); F_a = (double *)calloc(3 * (size_t)nn, sizeof(double)); f = (double *)malloc(sizeof(double) * 3 * nn);
    ulo_area_tot(&c->m, c->x, area);                                                    /* UL:531 */
    ulo_calculate_res(p, 1, &c->m, c->X, c->x, c->ea, dt_real, area + 1, t_tot, Res, Rea, MM);   /* UL:543 */
    ulo_force_FA(p, &c->m, c->x, c->cC, c->fa_u, F_FA);                                 /* UL:562 */
    ulo_force_AP(s, c->x, dt_real, F_AP);                                               /* UL:565 */
    for (i = 0; i < 3 * nn; i++) F_vd[i] = -p->drag * c->Vh[i];                         /* UL:568-574 */
    ulo_force_bending(&c->m, c->x, p->k_bend, F_b);                                     /* UL:584 */
    ulo_force_area(p, &c->m, c->x, area[0], F_a);                                       /* UL:589 */
    for (i = 0; i < 3 * nn; i++) f[i] = F_vd[i] + F_b[i] + F_AP[i] + F_FA[i] + F_a[i] - Res[i];   /* UL:592 */
    for (i = 0; i < 4; i++) s->diag[i] = area[i];
    s->diag[4] = norm_n(Res, 3 * nn);
    memcpy(s->sf[0], Res, sizeof(double) * 3 * nn); memcpy(s->sf[1], F_FA, sizeof(double) * 3 * nn); memcpy(s->sf[2], F_AP, sizeof(double) * 3 * nn);
    memcpy(s->sf[3], F_b, sizeof(double) * 3 * nn); memcpy(s->sf[4], F_a, sizeof(double) * 3 * nn);
    memcpy(s->s_mm[0], MM, sizeof(double) * nn); memcpy(s->s_rea[0], Rea, sizeof(double) * 6 * nn);
    nodal_update(p, nn, f, MM, Rea, c->A, c->V, c->Vh, c->ea, dt, dt_real);
    free(Res); free(Rea); free(MM); free(F_FA); free(F_AP); free(F_vd); free(F_b); free(F_a); free(f);
}
/* updateForces_Explicit, substrate overload UL:429-504 */
static void update_forces_subs(ulo_sim *s, double dt, double dt_real, double t_tot) {
    ulo_surface *b = &s->subs;
    const ulo_params *p = &s->p;
    int nn = b->m.nn, i;
    double zero3[3] = {0, 0, 0};
    double *Res = (double *)malloc(sizeof(double) * 3 * nn), *Rea = (double *)malloc(sizeof(double) * 6 * nn), *MM = (double *)malloc(sizeof(double) * nn);
    double *F_FA = (double *)malloc(sizeof(double) * 3 * nn), *f = (double *)malloc(sizeof(double) * 3 * nn);
    ulo_calculate_res(p, 2, &b->m, b->X, b->x, b->ea, dt_real, zero3, t_tot, Res, Rea, MM);   /* UL:452 */
    ulo_force_FA_subs(s, F_FA);                                                         /* UL:467 */
    for (i = 0; i < 3 * nn; i++) f[i] = F_FA[i] - Res[i];                               /* UL:471 */
    s->diag[5] = norm_n(Res, 3 * nn);
    memcpy(s->sf[5], Res, sizeof(double) * 3 * nn); memcpy(s->sf[6], F_FA, sizeof(double) * 3 * nn);
    memcpy(s->s_mm[1], MM, sizeof(double) * nn); memcpy(s->s_rea[1], Rea, sizeof(double) * 6 * nn);
    nodal_update(p, nn, f, MM, Rea, b->A, b->V, b->Vh, b->ea, dt, dt_real);
    free(Res); free(Rea); free(MM); free(F_FA); free(f);
}

/* ------------------------------------------------------------------------------------------------ */
/* simulation object and one step (main.cpp:168-418, El Topo off)                                    */
/* ------------------------------------------------------------------------------------------------ */
static void surf_init(ulo_surface *s, int nn, const double *xyz, int ne, const int *tri) {
    size_t n = (size_t)(nn > 0 ? nn : 1);
    ulo_mesh_build(&s->m, nn, ne, tri);
    s->x = (double *)malloc(sizeof(double) * 3 * n); memcpy(s->x, xyz, sizeof(double) * 3 * (size_t)nn);
    s->X = (double *)malloc(sizeof(double) * 3 * n); memcpy(s->X, xyz, sizeof(double) * 3 * (size_t)nn);
    s->V = (double *)calloc(3 * n, sizeof(double)); s->Vh = (double *)calloc(3 * n, sizeof(double)); s->A = (double *)calloc(3 * n, sizeof(double));
    s->ea = (double *)calloc(6 * n, sizeof(double)); s->fa_u = (double *)calloc(3 * n, sizeof(double));   /* scriptinit.cpp:567-580 zero init */
    s->cA = (double *)calloc(n, sizeof(double)); s->cB = (double *)calloc(n, sizeof(double)); s->cC = (double *)calloc(n, sizeof(double));
}
static void surf_free(ulo_surface *s) {
    ulo_mesh_free(&s->m);
    free(s->x); free(s->X); free(s->V); free(s->Vh); free(s->A); free(s->ea); free(s->fa_u); free(s->cA); free(s->cB); free(s->cC);
}
ulo_sim *ulo_create(int variant, int nn_c, const double *xyz_c, int ne_c, const int *tri_c,
                    int nn_s, const double *xyz_s, int ne_s, const int *tri_s) {
    ulo_sim *s = (ulo_sim *)calloc(1, sizeof *s);
    size_t n = (size_t)(nn_c > 0 ? nn_c : 1);
    int i;
    ulo_default_params(&s->p, variant);
    surf_init(&s->cell, nn_c, xyz_c, ne_c, tri_c);
    surf_init(&s->subs, nn_s, xyz_s, ne_s, tri_s);
    for (i = 0; i < nn_c; i++) { s->cell.cA[i] = 100; s->cell.cC[i] = 0; }      /* main.cpp:671-674 */
    ulo_srand(&s->rng, 1);                                                      /* main.cpp:709 */
    s->loc_bond = (int *)calloc(n, sizeof(int)); s->loc_force_cnt = (int *)calloc(n, sizeof(int)); s->loc_force_first = (int *)calloc(n, sizeof(int));
    s->ap_fire = (unsigned char *)calloc(n, 1); s->fa_reset = (unsigned char *)calloc(n, 1); s->cc_clamp = (unsigned char *)calloc(n, 1);
    for (i = 0; i < 7; i++) s->sf[i] = (double *)calloc(3 * (size_t)((i < 5 ? nn_c : nn_s) > 0 ? (i < 5 ? nn_c : nn_s) : 1), sizeof(double));
    s->s_mm[0] = (double *)calloc(n, sizeof(double)); s->s_rea[0] = (double *)calloc(6 * n, sizeof(double));
    s->s_mm[1] = (double *)calloc((size_t)(nn_s > 0 ? nn_s : 1), sizeof(double)); s->s_rea[1] = (double *)calloc(6 * (size_t)(nn_s > 0 ? nn_s : 1), sizeof(double));
    return s;
}
void ulo_destroy(ulo_sim *s) {
    if (!s) return;
    surf_free(&s->cell); surf_free(&s->subs);
    free(s->loc_bond); free(s->loc_force_cnt); free(s->loc_force_first); free(s->ap_fire); free(s->fa_reset); free(s->cc_clamp);
    { int i; for (i = 0; i < 7; i++) free(s->sf[i]); }
    free(s->s_mm[0]); free(s->s_mm[1]); free(s->s_rea[0]); free(s->s_rea[1]);
    free(s);
}
ulo_params *ulo_params_of(ulo_sim *s) { return &s->p; }
void ulo_set_use_grid(ulo_sim *s, int on) { s->use_grid = on; }
void ulo_seed(ulo_sim *s, unsigned int seed) { ulo_srand(&s->rng, seed); }
void ulo_rng_discard(ulo_sim *s, long n) { long i; for (i = 0; i < n; i++) (void)ulo_rand(&s->rng); }

static double *field_ptr(ulo_surface *s, int field, int *w) {
    switch (field) {
    case 0: *w = 3; return s->x;
    case 1: *w = 3; return s->V;
    case 2: *w = 3; return s->A;
    case 3: *w = 6; return s->ea;
    case 4: *w = 1; return s->cA;
    case 5: *w = 1; return s->cB;
    case 6: *w = 1; return s->cC;
    case 7: *w = 3; return s->fa_u;
    case 8: *w = 3; return s->X;
    case 9: *w = 3; return s->Vh;
    }
    return NULL;
}
int ulo_get(ulo_sim *s, int surface, int field, double *out) {
    ulo_surface *sf = surface == 0 ? &s->cell : &s->subs;
    int w; double *p = field_ptr(sf, field, &w);
    if (!p) return -1;
    memcpy(out, p, sizeof(double) * (size_t)w * sf->m.nn);
    return 0;
}
int ulo_set(ulo_sim *s, int surface, int field, const double *in) {
    ulo_surface *sf = surface == 0 ? &s->cell : &s->subs;
    int w; double *p = field_ptr(sf, field, &w);
    if (!p) return -1;
    memcpy(p, in, sizeof(double) * (size_t)w * sf->m.nn);
    if (field == 0) memcpy(sf->X, in, sizeof(double) * 3 * (size_t)sf->m.nn);
    return 0;
}
int ulo_get_discrete(ulo_sim *s, int which, int *out) {
    int i, nn = s->cell.m.nn;
    for (i = 0; i < nn; i++) {
        switch (which) {
        case 0: out[i] = s->loc_bond[i]; break;
        case 1: out[i] = s->loc_force_cnt[i]; break;
        case 2: out[i] = s->loc_force_first[i]; break;
        case 3: out[i] = s->ap_fire[i]; break;
        case 4: out[i] = s->fa_reset[i]; break;
        case 5: out[i] = s->cc_clamp[i]; break;
        default: return -1;
        }
    }
    return 0;
}
void ulo_get_diag(ulo_sim *s, double *out8) { memcpy(out8, s->diag, sizeof s->diag); }

void ulo_step(ulo_sim *s, double dt, double t_curr) {
    ulo_surface *c = &s->cell, *b = &s->subs;
    int i, nc = c->m.nn, ns = b->m.nn;
    ulo_driver_explicit(nc, c->V, c->A, dt, c->Vh);                                     /* main.cpp:224 */
    memcpy(c->X, c->x, sizeof(double) * 3 * (size_t)nc);                                /* main.cpp:242 */
    for (i = 0; i < 3 * nc; i++) c->x[i] = c->X[i] + dt * c->Vh[i];                     /* main.cpp:227-231, integrate() */
    ulo_driver_biaxial(&s->p, ns, b->x, b->V, b->A, dt, t_curr, b->Vh);                 /* main.cpp:293 */
    memcpy(b->X, b->x, sizeof(double) * 3 * (size_t)ns);                                /* main.cpp:311 */
    for (i = 0; i < 3 * ns; i++) b->x[i] = b->X[i] + s->p.subs_scale * dt * b->Vh[i];   /* main.cpp:298-302 */
    ulo_update_FA(s, dt);                                                               /* main.cpp:354 */
    update_forces_cell(s, dt, dt, t_curr);                                              /* main.cpp:360 */
    update_forces_subs(s, dt, dt, t_curr);                                              /* main.cpp:361 */
}

int ulo_force_field(ulo_sim *s, int which, double dt, double t_curr, double *out, double *mm, double *res_ea, double *aux4) {
    ulo_surface *c = &s->cell, *b = &s->subs;
    double area[4], zero3[3] = {0, 0, 0};
    int nc = c->m.nn, ns = b->m.nn;
    ulo_area_tot(&c->m, c->x, area);
    if (aux4) memcpy(aux4, area, sizeof area);
    if (which == 0 || which == 5) {
        ulo_surface *sf = which == 0 ? c : b;
        int nn = which == 0 ? nc : ns;
        double *MM = (double *)malloc(sizeof(double) * nn), *Rea = (double *)malloc(sizeof(double) * 6 * nn);
        ulo_calculate_res(&s->p, which == 0 ? 1 : 2, &sf->m, sf->X, sf->x, sf->ea, dt, which == 0 ? area + 1 : zero3, t_curr, out, Rea, MM);
        if (mm) memcpy(mm, MM, sizeof(double) * nn);
        if (res_ea) memcpy(res_ea, Rea, sizeof(double) * 6 * nn);
        free(MM); free(Rea);
        return 0;
    }
    if (which == 7 || which == 8) {                             /* Res_Adif / Res_Cdif: out[nn] */
        double *RA = (double *)malloc(sizeof(double) * nc), *RC = (double *)malloc(sizeof(double) * nc);
        ulo_diffusion_res(&s->p, &c->m, c->X, c->x, c->cA, c->cC, dt, RA, RC);
        memcpy(out, which == 7 ? RA : RC, sizeof(double) * nc);
        free(RA); free(RC);
        return 0;
    }
    if (which == 9) {                                           /* pressure, with the lumped mass of calculateRes(cell) */
        double *MM = (double *)malloc(sizeof(double) * nc), *Rea = (double *)malloc(sizeof(double) * 6 * nc), *Res = (double *)malloc(sizeof(double) * 3 * nc);
        ulo_calculate_res(&s->p, 1, &c->m, c->X, c->x, c->ea, dt, area + 1, t_curr, Res, Rea, MM);
        ulo_force_pressure(&c->m, c->x, MM, t_curr, out);
        free(MM); free(Rea); free(Res);
        return 0;
    }
    if (which == 10) { ulo_force_sphere(nc, c->x, out); return 0; }
    switch (which) {
    case 1: ulo_force_FA(&s->p, &c->m, c->x, c->cC, c->fa_u, out); return 0;
    case 2: ulo_force_AP(s, c->x, dt, out); return 0;
    case 3: ulo_force_bending(&c->m, c->x, s->p.k_bend, out); return 0;
    case 4: memset(out, 0, sizeof(double) * 3 * (size_t)nc); ulo_force_area(&s->p, &c->m, c->x, area[0], out); return 0;
    case 6: ulo_force_FA_subs(s, out); return 0;
    }
    return -1;
}

int ulo_step_force_field(ulo_sim *s, int which, double *out, double *mm, double *res_ea) {
    int surf, nn;
    if (which < 0 || which > 6) return -1;
    surf = which >= 5; nn = surf ? s->subs.m.nn : s->cell.m.nn;
    memcpy(out, s->sf[which], sizeof(double) * 3 * (size_t)nn);
    if (mm) memcpy(mm, s->s_mm[surf], sizeof(double) * (size_t)nn);
    if (res_ea) memcpy(res_ea, s->s_rea[surf], sizeof(double) * 6 * (size_t)nn);
    return 0;
}
