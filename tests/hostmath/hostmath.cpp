// tests/hostmath/hostmath.cpp -- TEST build of wholecellfe_b200/csrc/wcfe_math.cuh for a CPU-only box.
// Lets `pytest -m "not gpu"` check the kernels' per-element / per-node math against the oracle without a
// GPU.  This is not a fallback: nothing in the product loads this library.
#include "../../wholecellfe_b200/csrc/wcfe_math.cuh"
#include <cstring>
using namespace wcfe;
extern "C" {
void hm_elem_eval(const double *X9, const double *x9, const double *e6, double mu, double tk, double theta_g, double thick,
                  double *r1, double *r2, double *eas, double *A0, double *acur) {
    double X[3][3], x[3][3];
    memcpy(X, X9, sizeof X); memcpy(x, x9, sizeof x);
    ElemOut o; elem_eval(X, x, e6, mu, tk, theta_g, thick, o);
    memcpy(r1, o.r1, 24); memcpy(r2, o.r2, 24); memcpy(eas, o.eas, 48); *A0 = o.A0; *acur = o.a_cur;
}
int hm_point_in_triangle(const double *n1, const double *n2, const double *n3, const double *pt, double *A, double *Atri) {
    return point_in_triangle(n1, n2, n3, pt, A, *Atri) ? 1 : 0;
}
int hm_point_in_triangle_2d(const double *n1, const double *n2, const double *n3, const double *pt, double *A, double *Atri) {
    return point_in_triangle_2d(n1[0], n1[1], n2[0], n2[1], n3[0], n3[1], pt[0], pt[1], A, *Atri) ? 1 : 0;
}
unsigned int hm_philox(unsigned int seed, unsigned int step, unsigned int ord) { return philox_u32(seed, step, ord); }
}
extern "C" {
static Consts mk(int law, double kf, double kb, const double *px, const double *py, const double *kk) {
    Consts c; memset(&c, 0, sizeof c);
    c.law = law; c.kint_force = kf; c.kint_bond = kb; c.k_on = 1; c.ka = 0.004; c.kb = 10; c.Fa = 15; c.Fb = 15;
    for (int i = 0; i < 6; i++) { c.md_px[i] = px[i]; c.md_py[i] = py[i]; c.md_k[i] = kk[i]; }
    return c;
}
void hm_integrin_force(int law, double kf, const double *px, const double *py, const double *kk, const double *u, double *f) {
    Consts c = mk(law, kf, kf, px, py, kk); integrin_force(c, u, f);
}
int hm_bond_update(int law, double kf, double kb, const double *px, const double *py, const double *kk, double dt,
                   const double *cell_u, const double *subs_u, double *fa, double *cC) {
    Consts c = mk(law, kf, kb, px, py, kk); return bond_update(c, dt, cell_u, subs_u, fa, *cC);
}
}
extern "C" void hm_elem_eval_planar(const double *X6, const double *x6, const double *e3, double mu, double tk, double thick,
                                    double *r1, double *r2, double *eas, double *A0, double *acur) {
    double X[3][2], x[3][2];
    memcpy(X, X6, sizeof X); memcpy(x, x6, sizeof x);
    ElemOut2 o; elem_eval_planar(X, x, e3, mu, tk, thick, o);
    memcpy(r1, o.r1, 16); memcpy(r2, o.r2, 16); memcpy(eas, o.eas, 24); *A0 = o.A0; *acur = o.a_cur;
}

extern "C" void hm_elem_eval_frozen(const double *X6, const double *e3, double mu, double thick, double *r1, double *r2, double *A0) {
    double X[3][2]; memcpy(X, X6, sizeof X);
    ElemOutFrozen o; elem_eval_frozen(X, e3, mu, thick, o);
    memcpy(r1, o.r1, 16); memcpy(r2, o.r2, 16); *A0 = o.A0;
}

// boundary of the cell: outward normal, total edge length, bending curvature, actin loop trips (UL:1512-1520, :1604-1620, :1530)
extern "C" void hm_boundary(const double *xi, const double *j1, const double *j2, double density, double *nrm, double *tl, double *kappa, double *trips) {
    bnd_normal(xi, j1, j2, nrm, *tl);
    *kappa = bnd_curvature(xi, j1, j2, *tl);
    *trips = ap_trips(density, *tl);
}

// nodal update (UL:601-628): the three forms the kernels use -- correctly rounded quotient, reciprocal of M rho shared by a node's
// components, reciprocal of M supplied
extern "C" void hm_node_update(double f, double MM, double damp, double rho, double dt, double dt_real, double Vh, double V, double A, double *out6) {
    double v = V, a = A;
    node_update(f, MM, damp, rho, dt, dt_real, Vh, v, a); out6[0] = v; out6[1] = a;
    v = V; a = A;
    node_update_q(f, MM, recip_of(MM * rho), damp, rho, dt, dt_real, Vh, v, a); out6[2] = v; out6[3] = a;
    v = V; a = A;
    node_update_r(f, MM, 1.0 / MM, damp, rho, dt, dt_real, Vh, v, a); out6[4] = v; out6[5] = a;
}

// ---- host logic of the patch builder: row numbering of a substrate patch's node table (csrc/wcfe_rows.h)
#include "../../wholecellfe_b200/csrc/wcfe_rows.h"
extern "C" long long hm_ring_table_rows(const unsigned short *jk, int njk, const unsigned short *rptr, int P, int NL, unsigned short *row_of, long long *wavefronts_identity) {
    std::vector<unsigned short> vjk(jk, jk + njk), vr(rptr, rptr + P + 1), id((size_t)NL);
    for (int l = 0; l < NL; l++) id[l] = (unsigned short)l;
    const std::vector<unsigned short> r = wcfe::ring_table_rows(vjk, vr, P, NL);
    for (int l = 0; l < NL; l++) row_of[l] = r[l];
    *wavefronts_identity = wcfe::ring_table_wavefronts(vjk, vr, P, NL, id);
    return wcfe::ring_table_wavefronts(vjk, vr, P, NL, r);
}
