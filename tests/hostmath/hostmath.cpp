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
unsigned int hm_philox(unsigned int seed, unsigned int step, unsigned int ord) { return philox_u32(seed, step, ord); }
}
