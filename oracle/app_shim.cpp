// app_shim.cpp -- TEST INFRASTRUCTURE.  What the UNMODIFIED reference application (cellBP/main.cpp + El Topo, compiled where it
// lies by `make -C oracle app`) needs from outside its own tree, so that it links and runs deterministically in a container
// without BLAS/LAPACK:
//   * the five Fortran symbols El Topo pulls in (common/blas_wrapper.h:47-59, common/lapack_wrapper.h:154,187); none of them is on
//     the UL path, they serve El Topo's implicit smoothing / eigen-classification of vertices;
//   * --wrap=srand: the first call (main.cpp:709 `srand(1)`) is honoured, the per-step `srand(time(NULL))` of
//     UL_growth_explicit.cpp:1460 is dropped;
//   * --wrap=sprintf / __sprintf_chk: main.cpp:528 prints g_output_path into itself (overlapping buffers, undefined behaviour;
//     with glibc 2.39 the run directory lands in "/run0001", quirk q10).  The wrapper formats into a temporary first, which gives
//     the result the author obviously intended.
// Nothing here is part of the product; nothing here is copied from the reference.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <Eigen/Dense>

extern "C" {

double ddot_(const int *n, const double *x, const int *incx, const double *y, const int *incy) {
    double s = 0.0;
    for (int i = 0; i < *n; i++) s += x[(size_t)i * *incx] * y[(size_t)i * *incy];
    return s;
}
double dnrm2_(const int *n, const double *x, const int *incx) {
    double s = 0.0;
    for (int i = 0; i < *n; i++) s += x[(size_t)i * *incx] * x[(size_t)i * *incx];
    return std::sqrt(s);
}
int idamax_(const int *n, const double *x, const int *incx) {      // 1-based, first maximum of |x|
    if (*n < 1) return 0;
    int best = 0;
    for (int i = 1; i < *n; i++) if (std::fabs(x[(size_t)i * *incx]) > std::fabs(x[(size_t)best * *incx])) best = i;
    return best + 1;
}
void dcopy_(const int *n, const double *x, const int *incx, double *y, const int *incy) {
    for (int i = 0; i < *n; i++) y[(size_t)i * *incy] = x[(size_t)i * *incx];
}
void daxpy_(const int *n, const double *alpha, const double *x, const int *incx, double *y, const int *incy) {
    for (int i = 0; i < *n; i++) y[(size_t)i * *incy] += *alpha * x[(size_t)i * *incx];
}
void dscal_(const int *n, const double *alpha, double *x, const int *incx) {
    for (int i = 0; i < *n; i++) x[(size_t)i * *incx] *= *alpha;
}
// symmetric eigenproblem, jobz = 'V', column-major a (either triangle is fine: the callers pass full symmetric matrices),
// eigenvalues ascending in w, eigenvectors in the columns of a
int dsyev_(char *jobz, char *uplo, int *n, double *a, int *lda, double *w, double *work, int *lwork, int *info) {
    (void)uplo; (void)work;
    if (*lwork == -1) { work[0] = 3.0 * *n; *info = 0; return 0; }
    Eigen::MatrixXd A(*n, *n);
    for (int j = 0; j < *n; j++) for (int i = 0; i < *n; i++) A(i, j) = a[(size_t)j * *lda + i];
    for (int j = 0; j < *n; j++) for (int i = 0; i < j; i++) A(i, j) = A(j, i);          // uplo = 'L': the lower triangle is the data
    Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> es(A, (*jobz == 'V' || *jobz == 'v') ? Eigen::ComputeEigenvectors : Eigen::EigenvaluesOnly);
    if (es.info() != Eigen::Success) { *info = 1; return 0; }
    for (int i = 0; i < *n; i++) w[i] = es.eigenvalues()(i);
    if (*jobz == 'V' || *jobz == 'v') for (int j = 0; j < *n; j++) for (int i = 0; i < *n; i++) a[(size_t)j * *lda + i] = es.eigenvectors()(i, j);
    *info = 0;
    return 0;
}

void __real_srand(unsigned int);
void __wrap_srand(unsigned int seed) {
    static bool first = true;
    if (first) { first = false; __real_srand(seed); }
}

int __wrap_sprintf(char *dst, const char *fmt, ...) {
    va_list ap, ap2;
    va_start(ap, fmt);
    va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, fmt, ap);
    va_end(ap);
    std::vector<char> tmp((size_t)n + 1);
    vsnprintf(tmp.data(), tmp.size(), fmt, ap2);
    va_end(ap2);
    memcpy(dst, tmp.data(), (size_t)n + 1);
    return n;
}
int __wrap___sprintf_chk(char *dst, int flag, size_t dstlen, const char *fmt, ...) {
    (void)flag;
    va_list ap, ap2;
    va_start(ap, fmt);
    va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, fmt, ap);
    va_end(ap);
    std::vector<char> tmp((size_t)n + 1);
    vsnprintf(tmp.data(), tmp.size(), fmt, ap2);
    va_end(ap2);
    if ((size_t)n + 1 > dstlen) { fprintf(stderr, "app_shim: sprintf would overflow its destination\n"); abort(); }
    memcpy(dst, tmp.data(), (size_t)n + 1);
    return n;
}

}   // extern "C"
