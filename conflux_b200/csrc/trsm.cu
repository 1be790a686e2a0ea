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
// grid = (nblk, 2): y == 0 -> Uinv[j] = inv(U_jj) row-major; y == 1 -> LinvT[j] = inv(L_jj)^T row-major.
// One WARP per column of the inverse: the column lives in registers spread over the lanes (lane l holds entries l and
// l + 32), every substitution step is a two-term partial dot product per lane + a warp reduction, so a 64 x 64 block
// takes 64 steps of ~100 cycles per column instead of a 2000-FMA serial chain per thread.
template <int NB>
__global__ void __launch_bounds__(1024) diag_inverse_kernel(const double* __restrict__ A00, int v, double* __restrict__ Uinv,
                                                            double* __restrict__ LinvT) {
    static_assert(NB <= 64, "two entries per lane");
    __shared__ double S[NB][NB + 1];
    const int j = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double* blk = A00 + (size_t)(j * NB) * v + j * NB;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) S[e / NB][e % NB] = blk[(size_t)(e / NB) * v + e % NB];
    __syncthreads();
    double* out = (blockIdx.y == 0 ? Uinv : LinvT) + (size_t)j * NB * NB;
    const int t0 = lane, t1 = lane + 32;
    for (int c = warp; c < NB; c += nwarps) {
        double x0 = 0.0, x1 = 0.0;  // entries t0 and t1 of column c
        if (blockIdx.y == 0) {  // U X = I: x[c] = 1/U[c][c]; x[r] = -(sum_{t=r+1..c} U[r][t] x[t]) / U[r][r], r < c
            const double xc = 1.0 / S[c][c];
            if (t0 == c) x0 = xc;
            if (t1 == c) x1 = xc;
            for (int r = c - 1; r >= 0; --r) {
                double s = 0.0;
                if (t0 > r && t0 <= c) s = S[r][t0] * x0;
                if (t1 > r && t1 <= c && t1 < NB) s = fma(S[r][t1], x1, s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const double xr = -s / S[r][r];
                if (t0 == r) x0 = xr;
                if (t1 == r) x1 = xr;
            }
            if (t0 < NB) out[(size_t)t0 * NB + c] = x0;  // Uinv[r][c]
            if (t1 < NB) out[(size_t)t1 * NB + c] = x1;
        } else {  // L Y = I (unit diagonal): y[c] = 1; y[r] = -sum_{t=c..r-1} L[r][t] y[t], r > c
            if (t0 == c) x0 = 1.0;
            if (t1 == c) x1 = 1.0;
            for (int r = c + 1; r < NB; ++r) {
                double s = 0.0;
                if (t0 >= c && t0 < r) s = S[r][t0] * x0;
                if (t1 >= c && t1 < r) s = fma(S[r][t1], x1, s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (t0 == r) x0 = -s;
                if (t1 == r) x1 = -s;
            }
            if (t0 < NB) out[(size_t)c * NB + t0] = x0;  // LinvT[c][r] = Linv[r][c]
            if (t1 < NB) out[(size_t)c * NB + t1] = x1;
        }
    }
}

template <int NB>
int launch_diag_nb(const double* A00, int v, double* Uinv, double* LinvT, cudaStream_t stream) {
    const int warps = NB < 32 ? NB : 32;  // one warp per column (two columns per warp at NB = 64)
    diag_inverse_kernel<NB><<<dim3(v / NB, 2), 32 * warps, 0, stream>>>(A00, v, Uinv, LinvT);
    CFLX_CUDA(cudaGetLastError());
    return CFLX_OK;
}
}  // namespace

int launch_diag_inverses(const double* A00, int v, int nb, double* Uinv, double* LinvT, cudaStream_t stream) {
    switch (nb) {
        case 64: return launch_diag_nb<64>(A00, v, Uinv, LinvT, stream);
        case 32: return launch_diag_nb<32>(A00, v, Uinv, LinvT, stream);
        case 16: return launch_diag_nb<16>(A00, v, Uinv, LinvT, stream);
        case 8: return launch_diag_nb<8>(A00, v, Uinv, LinvT, stream);
        case 4: return launch_diag_nb<4>(A00, v, Uinv, LinvT, stream);
        default:
            set_last_error("diag_inverses: unsupported block size %d", nb);
            return CFLX_ERR_UNSUPPORTED;
    }
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
    // (callers solve a column window by offsetting R and U: the columns are independent right-hand sides)
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
