/*
 * oracle/lu_oracle.c -- CPU RESTATEMENT ORACLE of CONFLUX's LU hot path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it; the product library
 * (conflux_b200/csrc) never links or executes it.
 *
 * Plain C, no BLAS: every rank of the Px x Py x Pz grid is simulated sequentially in one process, in the
 * step order of the reference's conflux::LU_rep<T> (all citations relative to /root/reference):
 *   src/conflux/lu/conflux_opt.hpp:535-1803  main loop (steps 0-6 + validation stores)
 *   src/conflux/lu/conflux_opt.hpp:143-166   LUP            -> getrf_perm()
 *   src/conflux/lu/conflux_opt.hpp:176-218   push_pivots_up -> push_rows()
 *   src/conflux/lu/conflux_opt.hpp:220-336   tournament_rounds
 *   src/conflux/lu/conflux_opt.cpp:55-148    flipbit, butterfly_pair, g2lnoTile, analyze_pivots
 *   src/conflux/lu/utils.hpp:85-116          inverse_permute_rows
 *   src/conflux/lu/lu_params.hpp:49-82,364-375  sizes, InitMatrix generator (mt19937_64(42+rank), 5+U[0,1))
 *   src/conflux/lu/layout.cpp:95-123         tile layout (tile (gi,gj) -> rank (gi%Px, gj%Py), local (gi/Px, gj/Py))
 * The arithmetic the reference delegates to an un-pinned BLAS/LAPACK (LAPACKE_dgetrf, cblas_dtrsm x2,
 * cblas_dgemm; conflux_opt.hpp:158,1347,1539,1628) is restated from the published LAPACK/BLAS definitions:
 * dgetf2-style right-looking partial pivoting (idamax = first maximal |a|, reciprocal scaling), forward
 * substitution TRSMs, plain GEMM.  Bitwise parity with a particular BLAS is therefore NOT claimed; the pinned
 * contract is (1) the pivot sequence (exact), (2) L\U within 1e-10*||A||, (3) ||PA-LU||/||A|| -- all checked
 * against the real reference built by oracle/build_ref.sh (tests/test_oracle_vs_reference.py) and against the
 * committed golden vectors in tests/golden/.
 * Deviation kept from the survey: when the tournament has zero rounds (Px == 1) A00 is taken from the local
 * LUP result (the reference leaves A00 zero there and returns NaN).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ sizes (lu_params.hpp:67-82) */
typedef struct {
    int N, v, Px, Py, Pz, M, Ml, Nl, Nt, nlayr, P;
} odims;

static int iceil(int a, int b) { return (a + b - 1) / b; }

int oracle_dims(int N, int v, int Px, int Py, int Pz, int* out /* M,N,Ml,Nl,Nt,nlayr */) {
    int tx = iceil(N, v * Px), ty = iceil(N, v * Py);
    int M = v * Px * tx, NN = v * Py * ty;
    out[0] = M;
    out[1] = NN;
    out[2] = iceil(M / v, Px) * v;
    out[3] = iceil(NN / v, Py) * v;
    out[4] = NN / v;
    out[5] = (v + Pz - 1) / Pz;
    return 0;
}

/* ------------------------------------------------------------------ std::mt19937_64 restated */
typedef struct {
    uint64_t mt[312];
    int idx;
} mt64;
static void mt64_seed(mt64* s, uint64_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 312; ++i) s->mt[i] = 6364136223846793005ULL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 62)) + (uint64_t)i;
    s->idx = 312;
}
static uint64_t mt64_next(mt64* s) {
    if (s->idx >= 312) {
        for (int i = 0; i < 312; ++i) {
            uint64_t x = (s->mt[i] & 0xFFFFFFFF80000000ULL) | (s->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
            uint64_t xa = x >> 1;
            if (x & 1ULL) xa ^= 0xB5026F5AA96619E9ULL;
            s->mt[i] = s->mt[(i + 156) % 312] ^ xa;
        }
        s->idx = 0;
    }
    uint64_t y = s->mt[s->idx++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= (y >> 43);
    return y;
}
/* libstdc++ uniform_real_distribution<double>()(mt19937_64): generate_canonical<double,53> draws ONE 64-bit
 * word: double(u) / 2^64, clamped below 1 (lu_params.hpp:366-371 binds dist(eng)). */
static double mt64_uniform(mt64* s) {
    double r = (double)mt64_next(s) * (1.0 / 18446744073709551616.0);
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* InitMatrix, random branch (lu_params.hpp:364-375): layer-0 rank r draws its tiles (lti outer, ltj inner,
 * row-major inside a tile; grid_layout.hpp:68-92) from mt19937_64(seed + r); other layers are zero. */
int oracle_init_matrix(int N, int v, int Px, int Py, int Pz, int seed, double* A_all) {
    int d[6];
    oracle_dims(N, v, Px, Py, Pz, d);
    int Ml = d[2], Nl = d[3];
    size_t loc = (size_t)Ml * Nl;
    memset(A_all, 0, sizeof(double) * loc * Px * Py * Pz);
    for (int pi = 0; pi < Px; ++pi)
        for (int pj = 0; pj < Py; ++pj) {
            int rank = (pi * Py + pj) * Pz;
            double* A = A_all + (size_t)rank * loc;
            mt64 g;
            mt64_seed(&g, (uint64_t)(seed + rank));
            for (int lti = 0; lti < Ml / v; ++lti)
                for (int ltj = 0; ltj < Nl / v; ++ltj)
                    for (int li = 0; li < v; ++li)
                        for (int lj = 0; lj < v; ++lj) A[(size_t)(lti * v + li) * Nl + ltj * v + lj] = 5.0 + mt64_uniform(&g);
        }
    return 0;
}

/* ------------------------------------------------------------------ small helpers */
int oracle_flipbit(int n, int k) { return n ^ (1 << k); } /* conflux_opt.cpp:55 */

int oracle_butterfly_pair(int pi, int r, int Px) { /* conflux_opt.cpp:59-72 */
    int src = oracle_flipbit(pi, r);
    if (src >= Px) {
        if (r == 0) src = pi;
        else {
            src = oracle_flipbit(src, r - 1);
            if (src >= Px) src = Px - 1;
        }
    }
    return src;
}

/* LUP (conflux_opt.hpp:143-166): partial-pivot LU of the n x v values in a (ld lda, row-major, IN PLACE) and
 * perm[0..max(2v,n)) = identity with LAPACK's sequential interchanges applied.  dgetf2 semantics: pivot =
 * first maximal |a(i,j)|, rows swapped, column scaled by the reciprocal, rank-1 update. */
void oracle_getrf_perm(int n, int v, double* a, int lda, int* perm) {
    int m = n > 2 * v ? n : 2 * v;
    for (int i = 0; i < m; ++i) perm[i] = i;
    int steps = n < v ? n : v;
    for (int j = 0; j < steps; ++j) {
        int p = j;
        double best = fabs(a[(size_t)j * lda + j]);
        for (int i = j + 1; i < n; ++i) {
            double x = fabs(a[(size_t)i * lda + j]);
            if (x > best) { best = x; p = i; }
        }
        if (p != j) {
            for (int c = 0; c < v; ++c) {
                double t = a[(size_t)j * lda + c];
                a[(size_t)j * lda + c] = a[(size_t)p * lda + c];
                a[(size_t)p * lda + c] = t;
            }
            int t = perm[j]; perm[j] = perm[p]; perm[p] = t;
        }
        double piv = a[(size_t)j * lda + j];
        if (piv != 0.0) {
            double rinv = 1.0 / piv;
            for (int i = j + 1; i < n; ++i) {
                double* row = a + (size_t)i * lda;
                double l = row[j] * rinv;
                row[j] = l;
                const double* prow = a + (size_t)j * lda;
                for (int c = j + 1; c < v; ++c) row[c] -= l * prow[c];
            }
        }
    }
}

/* inverse_permute_rows (utils.hpp:85-116,119-138), row-major: out[i, :] = in[perm[i], in_cols-out_cols : ] */
void oracle_inverse_permute_rows(const double* in, double* out, int in_cols, int out_rows, int out_cols,
                                 const int* perm) {
    int col = in_cols - out_cols;
    for (int i = 0; i < out_rows; ++i)
        memcpy(out + (size_t)i * out_cols, in + (size_t)perm[i] * in_cols + col, sizeof(double) * out_cols);
}
/* permute_rows (utils.hpp:49-79), row-major: out[perm[i], :] = in[i, in_cols-out_cols:] for i < in_rows */
void oracle_permute_rows(const double* in, double* out, int in_rows, int in_cols, int out_cols, const int* perm) {
    int col = in_cols - out_cols;
    for (int i = 0; i < in_rows; ++i)
        memcpy(out + (size_t)perm[i] * out_cols, in + (size_t)i * in_cols + col, sizeof(double) * out_cols);
}

/* analyze_pivots (conflux_opt.cpp:100-148): curPivots = {npiv, rows...}.  Returns #early (== #late). */
int oracle_analyze_pivots(int fnpr, int n_rows, const int* curPivots, int* early, int* late) {
    if (fnpr >= n_rows) return 0;
    int npiv = curPivots[0];
    char* is_piv = (char*)calloc((size_t)n_rows, 1);
    for (int i = 0; i < npiv; ++i) is_piv[curPivots[i + 1]] = 1;
    int ne = 0, nl = 0;
    int lim = fnpr + npiv < n_rows ? fnpr + npiv : n_rows;
    for (int i = fnpr; i < lim; ++i)
        if (!is_piv[i]) early[ne++] = i;
    for (int i = fnpr + npiv; i < n_rows; ++i)
        if (is_piv[i]) late[nl++] = i;
    free(is_piv);
    if (ne != nl) { fprintf(stderr, "[oracle] analyze_pivots: %d early vs %d late\n", ne, nl); abort(); }
    return ne;
}

/* push_pivots_up (conflux_opt.hpp:176-218) on a row-major n_rows x n_cols array of `es`-byte elements. */
static void push_rows(void* in_, void* tmp_, int n_rows, int n_cols, size_t es, const int* curPivots, int fnpr,
                      const int* early, const int* late, int nel) {
    if (n_rows == 0 || n_cols == 0 || fnpr >= n_rows) return;
    char* in = (char*)in_;
    char* tmp = (char*)tmp_;
    size_t rb = (size_t)n_cols * es;
    int npiv = curPivots[0];
    for (int i = 0; i < npiv; ++i) memcpy(tmp + i * rb, in + (size_t)curPivots[i + 1] * rb, rb);
    for (int i = 0; i < nel; ++i) memcpy(in + (size_t)late[i] * rb, in + (size_t)early[i] * rb, rb);
    for (int i = 0; i < npiv; ++i) memcpy(in + (size_t)(fnpr + i) * rb, tmp + i * rb, rb);
}
void oracle_push_pivots_up(double* inout, int n_rows, int n_cols, const int* curPivots, int fnpr) {
    int npiv = curPivots[0];
    int* early = (int*)malloc(sizeof(int) * (npiv + 1));
    int* late = (int*)malloc(sizeof(int) * (npiv + 1));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(npiv + 1) * n_cols);
    int nel = oracle_analyze_pivots(fnpr, n_rows, curPivots, early, late);
    push_rows(inout, tmp, n_rows, n_cols, sizeof(double), curPivots, fnpr, early, late, nel);
    free(early); free(late); free(tmp);
}

/* g2lnoTile (conflux_opt.cpp:74-98): owner of global row g is (g / v) % Px */
void oracle_g2l_owner(const int* grows, int size, int Px, int v, int* owner) {
    for (int i = 0; i < size; ++i) owner[i] = (grows[i] / v) % Px;
}

/* cblas_dtrsm(RowMajor, Right, Upper, NoTrans, NonUnit): X * U = B, in place on B (m x v, ld ldb) */
static void trsm_right_upper(int m, int v, const double* U, int ldu, double* B, int ldb) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        double* x = B + (size_t)i * ldb;
        for (int j = 0; j < v; ++j) {
            double s = x[j];
            for (int t = 0; t < j; ++t) s -= x[t] * U[(size_t)t * ldu + j];
            x[j] = s / U[(size_t)j * ldu + j];
        }
    }
}
/* cblas_dtrsm(RowMajor, Left, Lower, NoTrans, Unit): L * X = B, in place on B (v x n, ld ldb) */
static void trsm_left_lower_unit(int v, int n, const double* L, int ldl, double* B, int ldb) {
    for (int i = 1; i < v; ++i) {
        double* xi = B + (size_t)i * ldb;
        for (int t = 0; t < i; ++t) {
            double l = L[(size_t)i * ldl + t];
            const double* xt = B + (size_t)t * ldb;
            for (int c = 0; c < n; ++c) xi[c] -= l * xt[c];
        }
    }
}
/* C (m x n, ldc) -= A (m x k, lda) * B (k x n, ldb) */
static void gemm_minus(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        double* c = C + (size_t)i * ldc;
        for (int t = 0; t < k; ++t) {
            double a = A[(size_t)i * lda + t];
            const double* b = B + (size_t)t * ldb;
            for (int j = 0; j < n; ++j) c[j] -= a * b[j];
        }
    }
}

/* ------------------------------------------------------------------ the LU itself */
typedef struct {
    double *A11, *A10, *A10res, *A01, *A01Tmp, *A10Rcv, *A01Rcv, *A00, *C;
    double *cand, *candPerm, *pivotBuff, *rowTmp;
    int *gri, *griTmp, *igri, *perm, *curPivots, *curPivOrder, *early, *late, *gpivots;
    int fnpr, nact;
} rstate;

static int RK(const odims* d, int pi, int pj, int pk) { return (pi * d->Py + pj) * d->Pz + pk; }

/* Runs the factorisation.  A_all: P blocks of Ml*Nl doubles in rank order (layers pk != 0 must be zero,
 * lu_params.hpp:149-155).  C_all (P blocks, may be NULL): L\U of PA in the conflux layout exactly as the
 * validation build leaves it in `C` (conflux_opt.hpp:1660-1771).  perm_out[M] = pivotIndsBuff
 * (conflux_opt.hpp:910,1822).  Returns 0, or -1 on unsupported arguments. */
int oracle_lu(int N, int v, int Px, int Py, int Pz, const double* A_all, double* C_all, int* perm_out) {
    odims d;
    int dd[6];
    oracle_dims(N, v, Px, Py, Pz, dd);
    d.N = dd[1]; d.v = v; d.Px = Px; d.Py = Py; d.Pz = Pz; d.M = dd[0]; d.Ml = dd[2]; d.Nl = dd[3];
    d.Nt = dd[4]; d.nlayr = dd[5]; d.P = Px * Py * Pz;
    if (Px != Py || v % Pz != 0) return -1; /* reference assumptions, SURVEY.md fact 6 */
    const int Ml = d.Ml, Nl = d.Nl, M = d.M, nlayr = d.nlayr, P = d.P;
    const size_t loc = (size_t)Ml * Nl;
    const int candRows = Ml > 2 * v ? Ml : 2 * v;
    rstate* R = (rstate*)calloc((size_t)P, sizeof(rstate));
    int* pivotIndsBuff = (int*)calloc((size_t)M, sizeof(int));
    for (int r = 0; r < P; ++r) {
        rstate* s = &R[r];
        s->A11 = (double*)malloc(sizeof(double) * loc);
        memcpy(s->A11, A_all + (size_t)r * loc, sizeof(double) * loc);
        s->A10 = (double*)calloc((size_t)Ml * v, sizeof(double));
        s->A10res = (double*)calloc(loc, sizeof(double));
        s->A01 = (double*)calloc((size_t)v * Nl, sizeof(double));
        s->A01Tmp = (double*)calloc((size_t)v * (Nl + 1), sizeof(double));
        s->A10Rcv = (double*)calloc((size_t)Ml * nlayr, sizeof(double));
        s->A01Rcv = (double*)calloc((size_t)nlayr * Nl, sizeof(double));
        s->A00 = (double*)calloc((size_t)v * v, sizeof(double));
        s->C = (double*)calloc(loc, sizeof(double));
        s->cand = (double*)calloc((size_t)candRows * (v + 1), sizeof(double));
        s->candPerm = (double*)calloc((size_t)candRows * (v + 1), sizeof(double));
        s->pivotBuff = (double*)calloc((size_t)candRows * v, sizeof(double));
        s->rowTmp = (double*)malloc(sizeof(double) * (size_t)v * Nl);
        s->gri = (int*)malloc(sizeof(int) * Ml);
        s->griTmp = (int*)malloc(sizeof(int) * Ml);
        s->igri = (int*)malloc(sizeof(int) * M);
        s->perm = (int*)malloc(sizeof(int) * candRows);
        s->curPivots = (int*)calloc((size_t)v + 1, sizeof(int));
        s->curPivOrder = (int*)calloc((size_t)v, sizeof(int));
        s->early = (int*)malloc(sizeof(int) * (v + 1));
        s->late = (int*)malloc(sizeof(int) * (v + 1));
        s->gpivots = (int*)calloc((size_t)v, sizeof(int));
        int pi = r / (Py * Pz);
        for (int i = 0; i < M; ++i) s->igri[i] = -1;
        for (int i = 0; i < Ml; ++i) { /* conflux_opt.hpp:430-440 */
            s->gri[i] = (i % v) + ((i / v) * Px + pi) * v;
            s->igri[s->gri[i]] = i;
        }
        s->fnpr = 0;
        s->nact = Ml;
    }
    double* snap = (double*)malloc(sizeof(double) * (size_t)Px * 2 * v * (v + 1));

    for (int k = 0; k < d.Nt; ++k) {
        const int loff = (k / Py) * v;
        const int pjk = k % Py, pik = k % Px;
        const int ncol = Nl - loff;
        /* step0_padding (conflux_opt.hpp:604-613) */
        for (int r = 0; r < P; ++r) {
            rstate* s = &R[r];
            if (s->nact < v) {
                int st = (s->nact > 0 ? s->nact : 0) * (v + 1), en = v * (v + 1);
                memset(s->cand + st, 0, sizeof(double) * (en - st));
                memset(s->candPerm + st, 0, sizeof(double) * (en - st));
            }
        }
        /* step 0: panel extract + reduce over layers onto pk = 0 (conflux_opt.hpp:618-648) */
        for (int pi = 0; pi < Px; ++pi) {
            rstate* root = &R[RK(&d, pi, pjk, 0)];
            for (int pk = 0; pk < Pz; ++pk) {
                rstate* s = &R[RK(&d, pi, pjk, pk)];
                for (int i = s->fnpr; i < Ml; ++i)
                    memcpy(s->A10 + (size_t)i * v, s->A11 + (size_t)i * Nl + loff, sizeof(double) * v);
            }
            for (int pk = 1; pk < Pz; ++pk) {
                rstate* s = &R[RK(&d, pi, pjk, pk)];
                for (size_t e = (size_t)root->fnpr * v; e < (size_t)Ml * v; ++e) root->A10[e] += s->A10[e];
            }
        }
        /* step 1: local LUP + tournament on column pj = k % Py, layer 0 (conflux_opt.hpp:693-816) */
        int nR = 0;
        while ((1 << nR) < Px) ++nR; /* ceil(log2(Px)) */
        for (int pi = 0; pi < Px; ++pi) {
            rstate* s = &R[RK(&d, pi, pjk, 0)];
            int n = s->nact;
            for (int i = 0; i < n; ++i) {
                s->cand[(size_t)i * (v + 1)] = (double)s->gri[s->fnpr + i];
                memcpy(s->cand + (size_t)i * (v + 1) + 1, s->A10 + (size_t)(s->fnpr + i) * v, sizeof(double) * v);
            }
            for (int i = 0; i < n; ++i)
                memcpy(s->pivotBuff + (size_t)i * v, s->cand + (size_t)i * (v + 1) + 1, sizeof(double) * v);
            oracle_getrf_perm(n, v, s->pivotBuff, v, s->perm);
            int src_pi = oracle_flipbit(pi, 0);
            if (src_pi > Px - 1) src_pi = Px - 1;
            size_t off = (src_pi < pi) ? (size_t)v * (v + 1) : 0;
            oracle_inverse_permute_rows(s->cand, s->candPerm + off, v + 1, v, v + 1, s->perm);
            double* t = s->cand; s->cand = s->candPerm; s->candPerm = t;
            if (nR == 0) /* Px == 1 fix (see header) */
                for (int i = 0; i < v; ++i) memcpy(s->A00 + (size_t)i * v, s->pivotBuff + (size_t)i * v, sizeof(double) * v);
        }
        for (int r = 0; r < nR; ++r) { /* tournament_rounds (conflux_opt.hpp:242-335) */
            const size_t half = (size_t)v * (v + 1);
            for (int pi = 0; pi < Px; ++pi)
                memcpy(snap + (size_t)pi * 2 * half, R[RK(&d, pi, pjk, 0)].cand, sizeof(double) * 2 * half);
            for (int pi = 0; pi < Px; ++pi) {
                rstate* s = &R[RK(&d, pi, pjk, 0)];
                int src = oracle_butterfly_pair(pi, r, Px);
                size_t send_off = 0, recv_off = half;
                if (src < pi) { send_off = half; recv_off = 0; }
                (void)send_off;
                /* what `src` sent us: its Sendrecv half if we are its partner, else its Isend of the lower half */
                size_t src_off;
                if (oracle_butterfly_pair(src, r, Px) == pi) src_off = (pi < src) ? half : 0;
                else src_off = half;
                memcpy(s->cand + recv_off, snap + (size_t)src * 2 * half + src_off, sizeof(double) * half);
                for (int i = 0; i < 2 * v; ++i)
                    memcpy(s->pivotBuff + (size_t)i * v, s->cand + (size_t)i * (v + 1) + 1, sizeof(double) * v);
                oracle_getrf_perm(2 * v, v, s->pivotBuff, v, s->perm);
                size_t off = 0;
                if (r != nR - 1) {
                    int nsrc = oracle_butterfly_pair(pi, r + 1, Px);
                    if (nsrc < pi) off = half;
                }
                oracle_inverse_permute_rows(s->cand, s->candPerm + off, v + 1, v, v + 1, s->perm);
                double* t = s->cand; s->cand = s->candPerm; s->candPerm = t;
                if (r == nR - 1)
                    for (int i = 0; i < v; ++i) memcpy(s->A00 + (size_t)i * v, s->pivotBuff + (size_t)i * v, sizeof(double) * v);
            }
        }
        for (int pi = 0; pi < Px; ++pi) { /* gpivots = column 0 of the winners (conflux_opt.hpp:810-815) */
            rstate* s = &R[RK(&d, pi, pjk, 0)];
            for (int i = 0; i < v; ++i) s->gpivots[i] = (int)s->cand[(size_t)i * (v + 1)];
        }
        /* A00: (pi, k%Py, 0) -> (k%Px, pi, 0) (conflux_opt.hpp:818-850) */
        for (int pi = 0; pi < Px && pi < Py; ++pi) {
            int from = RK(&d, pi, pjk, 0), to = RK(&d, pik, pi, 0);
            if (from != to) memcpy(R[to].A00, R[from].A00, sizeof(double) * v * v);
        }
        /* gpivots bcast over jk_comm (conflux_opt.hpp:872) */
        for (int pi = 0; pi < Px; ++pi) {
            const int* g = R[RK(&d, pi, pjk, 0)].gpivots;
            for (int pj = 0; pj < Py; ++pj)
                for (int pk = 0; pk < Pz; ++pk) {
                    rstate* s = &R[RK(&d, pi, pj, pk)];
                    if (s->gpivots != g) memcpy(s->gpivots, g, sizeof(int) * v);
                }
        }
        memcpy(pivotIndsBuff + (size_t)k * v, R[RK(&d, 0, pjk, 0)].gpivots, sizeof(int) * v);

        /* step 2: localise pivots, push them up, extract pivot rows (conflux_opt.hpp:876-1173) */
        for (int r = 0; r < P; ++r) {
            rstate* s = &R[r];
            int pi = r / (Py * Pz);
            int np = 0;
            for (int i = 0; i < v; ++i) {
                int g = s->gpivots[i];
                if ((g / v) % Px == pi) {
                    s->curPivots[1 + np] = g;
                    s->curPivOrder[np] = i;
                    ++np;
                }
            }
            s->curPivots[0] = np;
            for (int i = 0; i < np; ++i) {
                int row = s->igri[s->curPivots[i + 1]];
                if (row < s->fnpr || row >= Ml) {
                    fprintf(stderr, "[oracle] step %d rank %d: pivot row %d outside the active range\n", k, r, row);
                    abort();
                }
                s->curPivots[i + 1] = row;
            }
            int nel = oracle_analyze_pivots(s->fnpr, Ml, s->curPivots, s->early, s->late);
            push_rows(s->A11, s->rowTmp, Ml, Nl, sizeof(double), s->curPivots, s->fnpr, s->early, s->late, nel);
            push_rows(s->A10res, s->rowTmp, Ml, Nl, sizeof(double), s->curPivots, s->fnpr, s->early, s->late, nel);
            push_rows(s->A10, s->rowTmp, Ml, v, sizeof(double), s->curPivots, s->fnpr, s->early, s->late, nel);
            push_rows(s->gri, s->griTmp, Ml, 1, sizeof(int), s->curPivots, s->fnpr, s->early, s->late, nel);
            s->fnpr += np;
            s->nact -= np;
            for (int i = 0; i < Ml; ++i) s->igri[s->gri[i]] = i;
            for (int i = 0; i < np; ++i) {
                int prow = s->fnpr - np + i;
                s->A01Tmp[(size_t)i * (ncol + 1)] = 0;
                memcpy(s->A01Tmp + (size_t)i * (ncol + 1) + 1, s->A11 + (size_t)prow * Nl + loff, sizeof(double) * ncol);
            }
        }
        for (int pi = 0; pi < Px; ++pi) /* reduce pivot rows over layers (conflux_opt.hpp:1164-1173) */
            for (int pj = 0; pj < Py; ++pj) {
                rstate* root = &R[RK(&d, pi, pj, 0)];
                size_t cnt = (size_t)root->curPivots[0] * (ncol + 1);
                for (int pk = 1; pk < Pz; ++pk) {
                    rstate* s = &R[RK(&d, pi, pj, pk)];
                    for (size_t e = 0; e < cnt; ++e) root->A01Tmp[e] += s->A01Tmp[e];
                }
            }
        /* step 3: gather pivot rows on row pi = k % Px in tournament order (conflux_opt.hpp:1191-1260,1454-1512) */
        for (int pj = 0; pj < Py; ++pj) {
            rstate* root = &R[RK(&d, pik, pj, 0)];
            for (int pi = 0; pi < Px; ++pi) {
                rstate* s = &R[RK(&d, pi, pj, 0)];
                for (int i = 0; i < s->curPivots[0]; ++i)
                    memcpy(root->A01 + (size_t)s->curPivOrder[i] * ncol, s->A01Tmp + (size_t)i * (ncol + 1) + 1,
                           sizeof(double) * ncol);
            }
        }
        /* step 4: A10 <- A10 * U00^-1 on the panel column, slabs to every (pj', pk') (conflux_opt.hpp:1329-1434) */
        for (int pi = 0; pi < Px; ++pi) {
            rstate* s = &R[RK(&d, pi, pjk, 0)];
            trsm_right_upper(s->nact, v, s->A00, v, s->A10 + (size_t)s->fnpr * v, v);
            for (int pj = 0; pj < Py; ++pj)
                for (int pk = 0; pk < Pz; ++pk) {
                    rstate* t = &R[RK(&d, pi, pj, pk)];
                    for (int i = 0; i < s->nact; ++i)
                        memcpy(t->A10Rcv + (size_t)i * nlayr, s->A10 + (size_t)(s->fnpr + i) * v + pk * nlayr,
                               sizeof(double) * nlayr);
                }
        }
        /* step 5: A01 <- L00^-1 * A01 on the pivot row, slabs to every (pi', pk') (conflux_opt.hpp:1522-1593) */
        for (int pj = 0; pj < Py; ++pj) {
            rstate* s = &R[RK(&d, pik, pj, 0)];
            trsm_left_lower_unit(v, ncol, s->A00, v, s->A01, ncol);
            for (int pi = 0; pi < Px; ++pi)
                for (int pk = 0; pk < Pz; ++pk) {
                    rstate* t = &R[RK(&d, pi, pj, pk)];
                    memcpy(t->A01Rcv, s->A01 + (size_t)pk * nlayr * ncol, sizeof(double) * (size_t)nlayr * ncol);
                }
        }
        /* step 6: trailing update on every rank and layer (conflux_opt.hpp:1628-1632) */
        for (int r = 0; r < P; ++r) {
            rstate* s = &R[r];
            gemm_minus(s->nact, ncol, nlayr, s->A10Rcv, nlayr, s->A01Rcv, ncol, s->A11 + (size_t)s->fnpr * Nl + loff, Nl);
        }
        /* validation stores (conflux_opt.hpp:1660-1771) */
        {
            const int locK = k / Py;
            if (k > 0)
                for (int pi = 0; pi < Px; ++pi)
                    for (int pj = 0; pj < Py; ++pj) {
                        rstate* s = &R[RK(&d, pi, pj, 0)];
                        rstate* dst = &R[RK(&d, pik, pj, 0)];
                        int np = s->curPivots[0];
                        for (int ii = 0; ii < np; ++ii) {
                            int i = s->curPivOrder[ii];
                            size_t src_off = (size_t)(ii + s->fnpr - np) * Nl;
                            size_t dst_off = (size_t)(i + locK * v) * Nl;
                            int cnt = (pik > pj) ? (locK + 1) * v : locK * v;
                            memcpy(dst->C + dst_off, s->A10res + src_off, sizeof(double) * cnt);
                        }
                    }
            for (int pj = 0; pj < Py; ++pj) {
                rstate* s = &R[RK(&d, pik, pj, 0)];
                if (k < d.Nt - 1) {
                    size_t rowOff = (size_t)Nl * v * locK;
                    if (pik > pj) {
                        int colOff = v * (locK + 1);
                        for (int i = 0; i < v; ++i)
                            memcpy(s->C + rowOff + (size_t)i * Nl + colOff, s->A01 + (size_t)i * ncol + v,
                                   sizeof(double) * (Nl - loff - v));
                    } else {
                        int colOff = v * locK;
                        for (int i = 0; i < v; ++i)
                            memcpy(s->C + rowOff + (size_t)i * Nl + colOff, s->A01 + (size_t)i * ncol, sizeof(double) * ncol);
                    }
                }
                if (pj == pjk) {
                    size_t rowOff = (size_t)Nl * v * locK;
                    int colOff = v * locK;
                    for (int i = 0; i < v; ++i)
                        memcpy(s->C + rowOff + (size_t)i * Nl + colOff, s->A00 + (size_t)i * v, sizeof(double) * v);
                }
            }
            for (int pi = 0; pi < Px; ++pi) { /* L rows of the still-active (and just promoted) rows */
                rstate* s = &R[RK(&d, pi, pjk, 0)];
                for (int i = s->fnpr - s->curPivots[0]; i < Ml; ++i)
                    memcpy(s->A10res + (size_t)i * Nl + loff, s->A10 + (size_t)i * v, sizeof(double) * v);
            }
        }
    }
    if (perm_out) memcpy(perm_out, pivotIndsBuff, sizeof(int) * M);
    for (int r = 0; r < P; ++r) {
        rstate* s = &R[r];
        if (C_all) memcpy(C_all + (size_t)r * loc, s->C, sizeof(double) * loc);
        free(s->A11); free(s->A10); free(s->A10res); free(s->A01); free(s->A01Tmp); free(s->A10Rcv); free(s->A01Rcv);
        free(s->A00); free(s->C); free(s->cand); free(s->candPerm); free(s->pivotBuff); free(s->rowTmp); free(s->gri);
        free(s->griTmp); free(s->igri); free(s->perm); free(s->curPivots); free(s->curPivOrder); free(s->early);
        free(s->late); free(s->gpivots);
    }
    free(R); free(pivotIndsBuff); free(snap);
    return 0;
}
