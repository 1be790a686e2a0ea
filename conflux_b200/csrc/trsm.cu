// conflux_b200/csrc/trsm.cu -- the two triangular solves of the LU step as blocked GEMMs on the FP64 tensor path.
//
// Replaces, relative to /root/reference/src/conflux/lu/conflux_opt.hpp:
//   :1347-1358  cblas_dtrsm(Right, Upper, NoTrans, NonUnit)  A10 <- A10 * U00^-1     -> trsm_right_upper_T
//   :1539-1550  cblas_dtrsm(Left,  Lower, NoTrans, Unit)     A01 <- L00^-1 * A01     -> trsm_left_lower_unit
// A00 = L00\U00 is v x v (<= 2 MB).  Its nb x nb diagonal blocks are inverted once per step by a small kernel
// (one CTA per block, substitution in shared memory); the solve is then a right-looking sweep of GEMMs
// (multiply by the inverse block, rank-nb update of the remaining block rows) that all run on gemm.cu's DMMA
// kernel with the panels kept K-major (transposed L panel, row-major U panel).
#include "common.cuh"
#include "kernels.h"

namespace cflx {
namespace {
// grid = (nblk, 2): y == 0 -> Uinv[j] = inv(U_jj) row-major; y == 1 -> LinvT[j] = inv(L_jj)^T row-major
__global__ void diag_inverse_kernel(const double* __restrict__ A00, int v, int nb, double* __restrict__ Uinv,
                                    double* __restrict__ LinvT) {
    extern __shared__ double sm[];
    double* S = sm;                  // [nb][nb+1] the diagonal block
    double* X = sm + nb * (nb + 1);  // [nb][nb+1] result (column c owned by thread c)
    const int j = blockIdx.x, c = threadIdx.x;
    const double* blk = A00 + (size_t)(j * nb) * v + j * nb;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) S[(e / nb) * (nb + 1) + e % nb] = blk[(size_t)(e / nb) * v + e % nb];
    __syncthreads();
    if (c < nb) {
        if (blockIdx.y == 0) {  // U X = I, column c by back substitution
            for (int r = nb - 1; r > c; --r) X[r * (nb + 1) + c] = 0.0;
            X[c * (nb + 1) + c] = 1.0 / S[c * (nb + 1) + c];
            for (int r = c - 1; r >= 0; --r) {
                double s = 0.0;
                for (int t = r + 1; t <= c; ++t) s += S[r * (nb + 1) + t] * X[t * (nb + 1) + c];
                X[r * (nb + 1) + c] = -s / S[r * (nb + 1) + r];
            }
        } else {  // L Y = I (unit diagonal), column c by forward substitution
            for (int r = 0; r < c; ++r) X[r * (nb + 1) + c] = 0.0;
            X[c * (nb + 1) + c] = 1.0;
            for (int r = c + 1; r < nb; ++r) {
                double s = 0.0;
                for (int t = c; t < r; ++t) s += S[r * (nb + 1) + t] * X[t * (nb + 1) + c];
                X[r * (nb + 1) + c] = -s;
            }
        }
    }
    __syncthreads();
    double* out = (blockIdx.y == 0 ? Uinv : LinvT) + (size_t)j * nb * nb;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int r = e / nb, cc = e % nb;
        out[e] = blockIdx.y == 0 ? X[r * (nb + 1) + cc] : X[cc * (nb + 1) + r];
    }
}
}  // namespace

int launch_diag_inverses(const double* A00, int v, int nb, double* Uinv, double* LinvT, cudaStream_t stream) {
    const int nblk = v / nb;
    const size_t smem = 2 * (size_t)nb * (nb + 1) * sizeof(double);
    static PerDeviceMax cfg;
    if (smem > 48 * 1024 && cfg.raise(smem))
        CFLX_CUDA(cudaFuncSetAttribute(diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = ((nb + 31) / 32) * 32;
    diag_inverse_kernel<<<dim3(nblk, 2), threads < 64 ? 64 : threads, smem, stream>>>(A00, v, nb, Uinv, LinvT);
    CFLX_CUDA(cudaGetLastError());
    return CFLX_OK;
}

// X * U00 = P  <=>  U00^T X^T = P^T.  PT/LT are the transposed panels [v][ld]; sweep over block rows j:
//   LT_j = inv(U_jj)^T * PT_j ;  PT_i -= U_ji^T * LT_j  (i > j)
int trsm_right_upper_T(const double* A00, const double* Uinv, int v, int nb, double* PT, double* LT, int64_t ld, int n,
                       cudaStream_t stream) {
    if (n <= 0) return CFLX_OK;
    const int nblk = v / nb;
    const int N = (n + 1) & ~1;
    for (int j = 0; j < nblk; ++j) {
        GemmArgs g{};
        g.M = nb; g.N = N; g.K = nb;
        g.AT = Uinv + (size_t)j * nb * nb; g.ldat = nb;              // AT[k][m] = inv(U_jj)[k][m]
        g.B = PT + (int64_t)j * nb * ld; g.ldb = ld;
        g.C = nullptr; g.ldc = ld;
        g.D = LT + (int64_t)j * nb * ld; g.ldd = ld;
        g.alpha = 1.0; g.beta = 0.0;
        CFLX_TRY(launch_gemm_tn(g, stream));
        const int rest = v - (j + 1) * nb;
        if (rest > 0) {
            GemmArgs u{};
            u.M = rest; u.N = N; u.K = nb;
            u.AT = A00 + (size_t)(j * nb) * v + (j + 1) * nb; u.ldat = v;  // AT[k][m] = U[j*nb+k][(j+1)*nb+m]
            u.B = LT + (int64_t)j * nb * ld; u.ldb = ld;
            u.C = PT + (int64_t)(j + 1) * nb * ld; u.ldc = ld;
            u.D = PT + (int64_t)(j + 1) * nb * ld; u.ldd = ld;
            u.alpha = -1.0; u.beta = 1.0;
            CFLX_TRY(launch_gemm_tn(u, stream));
        }
    }
    return CFLX_OK;
}

// L00 * X = R (unit lower).  R/U are [v][ld]:  U_j = inv(L_jj) * R_j ;  R_i -= L_ij * U_j  (i > j)
int trsm_left_lower_unit(const double* A00T, const double* LinvT, int v, int nb, double* R, double* U, int64_t ld, int n,
                         cudaStream_t stream) {
    if (n <= 0) return CFLX_OK;
    const int nblk = v / nb;
    for (int j = 0; j < nblk; ++j) {
        GemmArgs g{};
        g.M = nb; g.N = n; g.K = nb;
        g.AT = LinvT + (size_t)j * nb * nb; g.ldat = nb;             // AT[k][m] = inv(L_jj)[m][k]
        g.B = R + (int64_t)j * nb * ld; g.ldb = ld;
        g.C = nullptr; g.ldc = ld;
        g.D = U + (int64_t)j * nb * ld; g.ldd = ld;
        g.alpha = 1.0; g.beta = 0.0;
        CFLX_TRY(launch_gemm_tn(g, stream));
        const int rest = v - (j + 1) * nb;
        if (rest > 0) {
            GemmArgs u{};
            u.M = rest; u.N = n; u.K = nb;
            u.AT = A00T + (size_t)(j * nb) * v + (j + 1) * nb; u.ldat = v;  // AT[k][m] = L[(j+1)*nb+m][j*nb+k]
            u.B = U + (int64_t)j * nb * ld; u.ldb = ld;
            u.C = R + (int64_t)(j + 1) * nb * ld; u.ldc = ld;
            u.D = R + (int64_t)(j + 1) * nb * ld; u.ldd = ld;
            u.alpha = -1.0; u.beta = 1.0;
            CFLX_TRY(launch_gemm_tn(u, stream));
        }
    }
    return CFLX_OK;
}

}  // namespace cflx
