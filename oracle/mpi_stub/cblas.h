/* oracle/mpi_stub/cblas.h -- TEST INFRASTRUCTURE: hand-declared prototypes of the four CBLAS entry points the
 * reference LU path calls (conflux_opt.hpp:1347,1539,1628); resolved against the OpenBLAS 0.3.15 shipped
 * inside the opencv wheel (see build_ref.sh).  Enum values are the standard CBLAS ones. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
typedef CBLAS_ORDER CBLAS_LAYOUT;
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, double alpha, const double* A,
                 int lda, const double* B, int ldb, double beta, double* C, int ldc);
void cblas_dtrsm(CBLAS_ORDER, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int M, int N, double alpha,
                 const double* A, int lda, double* B, int ldb);
void openblas_set_num_threads(int);
#ifdef __cplusplus
}
#endif
