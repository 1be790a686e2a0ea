/* oracle/mpi_stub/lapacke.h -- TEST INFRASTRUCTURE: prototype of the one LAPACKE routine on the reference LU
 * path (conflux_opt.hpp:158). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv);
#ifdef __cplusplus
}
#endif
