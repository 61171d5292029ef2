/* TEST INFRASTRUCTURE (oracle build only) -- never part of the product library.
 *
 * Minimal CBLAS prototype shim. The image has an LP64 OpenBLAS shared object
 * (opencv_python_headless.libs/libopenblasp-*.so) but no <cblas.h>.  The
 * reference includes <cblas.h> through include/cblas_ct.h:10; this header
 * declares the standard CBLAS entry points that the reference sources call
 * (see `grep -o 'cblas_[a-z_]*' src` in the reference tree).
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int blasint;

typedef enum CBLAS_ORDER     { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113, CblasConjNoTrans = 114 } CBLAS_TRANSPOSE;
typedef CBLAS_ORDER CBLAS_LAYOUT;

/* level 1 */
float  cblas_snrm2 (blasint n, const float*  x, blasint incx);
double cblas_dnrm2 (blasint n, const double* x, blasint incx);
float  cblas_scnrm2(blasint n, const void*   x, blasint incx);
double cblas_dznrm2(blasint n, const void*   x, blasint incx);

void cblas_sscal (blasint n, float  alpha, float*  x, blasint incx);
void cblas_dscal (blasint n, double alpha, double* x, blasint incx);
void cblas_cscal (blasint n, const void* alpha, void* x, blasint incx);
void cblas_zscal (blasint n, const void* alpha, void* x, blasint incx);
void cblas_csscal(blasint n, float  alpha, void* x, blasint incx);
void cblas_zdscal(blasint n, double alpha, void* x, blasint incx);

void cblas_saxpy(blasint n, float  alpha, const float*  x, blasint incx, float*  y, blasint incy);
void cblas_daxpy(blasint n, double alpha, const double* x, blasint incx, double* y, blasint incy);
void cblas_caxpy(blasint n, const void* alpha, const void* x, blasint incx, void* y, blasint incy);
void cblas_zaxpy(blasint n, const void* alpha, const void* x, blasint incx, void* y, blasint incy);

float  cblas_sdot(blasint n, const float*  x, blasint incx, const float*  y, blasint incy);
double cblas_ddot(blasint n, const double* x, blasint incx, const double* y, blasint incy);
void cblas_cdotc_sub(blasint n, const void* x, blasint incx, const void* y, blasint incy, void* ret);
void cblas_zdotc_sub(blasint n, const void* x, blasint incx, const void* y, blasint incy, void* ret);
void cblas_cdotu_sub(blasint n, const void* x, blasint incx, const void* y, blasint incy, void* ret);
void cblas_zdotu_sub(blasint n, const void* x, blasint incx, const void* y, blasint incy, void* ret);

/* level 2 */
void cblas_sger (CBLAS_ORDER order, blasint m, blasint n, float  alpha, const float*  x, blasint incx, const float*  y, blasint incy, float*  a, blasint lda);
void cblas_dger (CBLAS_ORDER order, blasint m, blasint n, double alpha, const double* x, blasint incx, const double* y, blasint incy, double* a, blasint lda);
void cblas_cgeru(CBLAS_ORDER order, blasint m, blasint n, const void* alpha, const void* x, blasint incx, const void* y, blasint incy, void* a, blasint lda);
void cblas_zgeru(CBLAS_ORDER order, blasint m, blasint n, const void* alpha, const void* x, blasint incx, const void* y, blasint incy, void* a, blasint lda);
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, blasint m, blasint n, float  alpha, const float*  a, blasint lda, const float*  x, blasint incx, float  beta, float*  y, blasint incy);
void cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, blasint m, blasint n, double alpha, const double* a, blasint lda, const double* x, blasint incx, double beta, double* y, blasint incy);
void cblas_cgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, blasint m, blasint n, const void* alpha, const void* a, blasint lda, const void* x, blasint incx, const void* beta, void* y, blasint incy);
void cblas_zgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, blasint m, blasint n, const void* alpha, const void* a, blasint lda, const void* x, blasint incx, const void* beta, void* y, blasint incy);

/* level 3 */
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k, float  alpha, const float*  a, blasint lda, const float*  b, blasint ldb, float  beta, float*  c, blasint ldc);
void cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k, double alpha, const double* a, blasint lda, const double* b, blasint ldb, double beta, double* c, blasint ldc);
void cblas_cgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k, const void* alpha, const void* a, blasint lda, const void* b, blasint ldb, const void* beta, void* c, blasint ldc);
void cblas_zgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, blasint m, blasint n, blasint k, const void* alpha, const void* a, blasint lda, const void* b, blasint ldb, const void* beta, void* c, blasint ldc);

/* OpenBLAS threading control */
void openblas_set_num_threads(int num_threads);
int  openblas_get_num_threads(void);

#ifdef __cplusplus
}
#endif

#endif
