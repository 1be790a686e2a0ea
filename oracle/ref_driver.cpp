/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 * C entry points around the UNMODIFIED reference (compiled from /root/reference by build_ref.sh with
 * -DCONFLUX_WITH_VALIDATION so that LU_rep fills C and permutation, conflux_opt.hpp:408-424,1660-1802,1821).
 * Ranks are threads of this process (mpi_stub/).  Used to (a) pin oracle/lu_oracle.c, (b) generate
 * tests/golden/*, (c) time the reference's CPU path for bench.py --impl reference.
 */
#include <conflux/lu/conflux_opt.hpp>
#include <conflux/lu/utils.hpp>

#include <atomic>
#include <chrono>
#include <cstring>
#include <vector>

namespace {
struct RunArgs {
    int N, v, Px, Py, Pz, n_rep;
    double* A_all;
    double* C_all;
    int* perm;
    double* ms;
    int* dims_out;
};

void rank_main(int rank, void* p) {
    auto* a = (RunArgs*)p;
    conflux::lu_params<double> params(a->N, a->N, a->v, a->Px, a->Py, a->Pz, MPI_COMM_WORLD);
    const size_t loc = (size_t)params.Ml * params.Nl;
    if (rank == 0 && a->dims_out) {
        int d[] = {params.M, params.N, params.Ml, params.Nl, params.Nt, params.nlayr};
        std::memcpy(a->dims_out, d, sizeof(d));
    }
    std::vector<double> C(loc, 0.0);
    std::vector<int> piv(params.M, -1);
    double best = 1e300;
    for (int i = 0; i < a->n_rep; ++i) {
        params.InitMatrix();
        std::fill(C.begin(), C.end(), 0.0);
        auto t0 = std::chrono::high_resolution_clock::now();
        conflux::LU_rep<double>(params, C.data(), piv.data());
        auto t1 = std::chrono::high_resolution_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (i > 0 || a->n_rep == 1) best = std::min(best, ms);
    }
    // rank in lu_comm == world rank (stub cart keeps order); coords row-major (pi,pj,pk)
    if (a->A_all) std::memcpy(a->A_all + (size_t)params.rank * loc, params.data.data(), loc * sizeof(double));
    if (a->C_all) std::memcpy(a->C_all + (size_t)params.rank * loc, C.data(), loc * sizeof(double));
    if (params.rank == 0) {
        if (a->perm) std::memcpy(a->perm, piv.data(), sizeof(int) * params.M);
        if (a->ms) *a->ms = best;
    }
    MPI_Barrier(params.lu_comm);
}
}  // namespace

namespace {
// bench.py --impl reference: ONE lu_params object (= one process, thread pools warm) runs n_warm untimed and then up
// to n_rep timed factorisations, InitMatrix before each like the miniapp (conflux_miniapp.cpp:138-149), and stops
// early once budget_s of wall time are spent (at least one timed rep is always done).  Reports what LU_rep itself
// returns (main loop between two barriers, conflux_opt.hpp:531-532,1805-1807) and the wall time around the call.
struct BenchArgs {
    int N, v, Px, Py, Pz, n_warm, n_rep;
    double budget_s;
    double* ms_inner;  // [n_rep]
    double* ms_outer;  // [n_rep]
    int done;
    std::atomic<int> stop;
    std::chrono::steady_clock::time_point t_start;
};
void bench_main(int rank, void* p) {
    auto* a = (BenchArgs*)p;
    conflux::lu_params<double> params(a->N, a->N, a->v, a->Px, a->Py, a->Pz, MPI_COMM_WORLD);
    std::vector<double> C((size_t)params.Ml * params.Nl, 0.0);
    std::vector<int> piv(params.M, -1);
    for (int i = 0; i < a->n_warm + a->n_rep; ++i) {
        params.InitMatrix();
        MPI_Barrier(params.lu_comm);
        auto t0 = std::chrono::steady_clock::now();
        std::size_t inner = conflux::LU_rep<double>(params, C.data(), piv.data());
        MPI_Barrier(params.lu_comm);
        auto t1 = std::chrono::steady_clock::now();
        if (params.rank == 0) {
            if (i >= a->n_warm) {
                a->ms_inner[i - a->n_warm] = (double)inner;
                a->ms_outer[i - a->n_warm] = std::chrono::duration<double, std::milli>(t1 - t0).count();
                a->done = i - a->n_warm + 1;
            }
            const double spent = std::chrono::duration<double>(t1 - a->t_start).count();
            const double last = std::chrono::duration<double>(t1 - t0).count();
            // stop when the next repetition would not fit the budget any more
            if (i >= a->n_warm && spent + last > a->budget_s) a->stop.store(1);
        }
        MPI_Barrier(params.lu_comm);
        if (a->stop.load()) break;
    }
    MPI_Barrier(params.lu_comm);
}
}  // namespace

extern "C" {

int ref_lu_bench(int N, int v, int Px, int Py, int Pz, int n_warm, int n_rep, double budget_s, double* ms_inner,
                 double* ms_outer, int* done, int blas_threads) {
    BenchArgs a{N, v, Px, Py, Pz, n_warm < 0 ? 0 : n_warm, n_rep < 1 ? 1 : n_rep, budget_s, ms_inner, ms_outer, 0};
    a.stop.store(0);
    a.t_start = std::chrono::steady_clock::now();
    openblas_set_num_threads(blas_threads > 0 ? blas_threads : 1);
    stub_mpi_run(Px * Py * Pz, bench_main, &a);
    *done = a.done;
    return 0;
}

// dims_out[6] = {M, N, Ml, Nl, Nt, nlayr} exactly as lu_params::initialize derives them (lu_params.hpp:67-82)
int ref_lu_dims(int N, int v, int Px, int Py, int Pz, int* dims_out) {
    int nLocalTilesx = (int)std::ceil((double)N / (v * Px));
    int nLocalTilesy = (int)std::ceil((double)N / (v * Py));
    int M = v * Px * nLocalTilesx, NN = v * Py * nLocalTilesy;
    int Nt = (int)std::ceil((double)NN / v), Mt = (int)std::ceil((double)M / v);
    dims_out[0] = M; dims_out[1] = NN;
    dims_out[2] = (int)std::ceil((double)Mt / Px) * v;
    dims_out[3] = (int)std::ceil((double)Nt / Py) * v;
    dims_out[4] = Nt; dims_out[5] = (v + Pz - 1) / Pz;
    return 0;
}

// Runs the reference LU_rep on a Px x Py x Pz grid of threads.  A_all/C_all: P blocks of Ml*Nl doubles in
// rank order (rank = (pi*Py+pj)*Pz+pk), perm: M ints, ms: best wall time of the non-warm-up reps.
int ref_lu_run(int N, int v, int Px, int Py, int Pz, int n_rep, double* A_all, double* C_all, int* perm,
               double* ms, int blas_threads) {
    RunArgs a{N, v, Px, Py, Pz, n_rep < 1 ? 1 : n_rep, A_all, C_all, perm, ms, nullptr};
    openblas_set_num_threads(blas_threads > 0 ? blas_threads : 1);
    stub_mpi_run(Px * Py * Pz, rank_main, &a);
    return 0;
}

/* ---- helper entry points used to generate tests/golden (the reference's own building blocks) ---- */
int ref_butterfly_pair(int pi, int r, int Px) { return conflux::butterfly_pair(pi, r, Px); }

// push_pivots_up on a row-major n_rows x n_cols double matrix; curPivots = {npiv, rows...} (conflux_opt.hpp:176)
void ref_push_pivots_up(double* inout, int n_rows, int n_cols, const int* curPivots, int first_non_pivot_row) {
    std::vector<double> in(inout, inout + (size_t)n_rows * n_cols), tmp((size_t)n_rows * n_cols);
    std::vector<int> cp(curPivots, curPivots + curPivots[0] + 1), early, late;
    std::vector<bool> pivots(n_rows);
    conflux::analyze_pivots(first_non_pivot_row, n_rows, cp, pivots, early, late);
    conflux::push_pivots_up<double>(in, tmp, n_rows, n_cols, conflux::order::row_major, cp, first_non_pivot_row,
                                    pivots, early, late);
    std::memcpy(inout, in.data(), in.size() * sizeof(double));
}

// LUP (conflux_opt.hpp:143): cand is n x (v+1) row-major, column 0 = row tags.  Returns perm (max(2v,n) ints)
// and the factored n x v pivotBuff.
void ref_lup(int n, int v, const double* cand, int* perm_out, double* pivotBuff_out) {
    int m = std::max(2 * v, n);
    std::vector<int> ipiv(m), perm(m);
    std::vector<double> c(cand, cand + (size_t)n * (v + 1));
    conflux::LUP<double>(n, v, v + 1, pivotBuff_out, c.data() + 1, ipiv, perm);
    std::memcpy(perm_out, perm.data(), sizeof(int) * m);
}

// inverse_permute_rows (utils.hpp:119): out[i,:] = in[perm[i], col_off:] for i < new_rows
void ref_inverse_permute_rows(const double* in, double* out, int n_rows, int n_cols, int new_rows, int new_cols,
                              const int* perm, int nperm) {
    std::vector<int> p(perm, perm + nperm);
    conflux::inverse_permute_rows<double>(const_cast<double*>(in), out, n_rows, n_cols, new_rows, new_cols,
                                          conflux::order::row_major, p);
}
void ref_permute_rows(const double* in, double* out, int n_rows, int n_cols, int new_rows, int new_cols,
                      const int* perm, int nperm) {
    std::vector<int> p(perm, perm + nperm);
    conflux::permute_rows<double>(const_cast<double*>(in), out, n_rows, n_cols, new_rows, new_cols,
                                  conflux::order::row_major, p);
}

// g2lnoTile (conflux_opt.cpp:74): owner_out[i] = owning pi of grows[i]
void ref_g2l_owner(const int* grows, int size, int Px, int v, int* owner_out) {
    std::vector<int> g(grows, grows + size);
    auto res = conflux::g2lnoTile(g, size, Px, v);
    for (auto& kv : res.second)
        for (int off : kv.second) owner_out[off] = kv.first;
}

// lu_params::InitMatrix for one rank of a grid (seeded generator + the hard-coded small matrices)
struct InitArgs { int N, v, Px, Py, Pz; double* A_all; };
static void init_main(int, void* p) {
    auto* a = (InitArgs*)p;
    conflux::lu_params<double> params(a->N, a->N, a->v, a->Px, a->Py, a->Pz, MPI_COMM_WORLD);
    size_t loc = (size_t)params.Ml * params.Nl;
    std::memcpy(a->A_all + (size_t)params.rank * loc, params.data.data(), loc * sizeof(double));
    MPI_Barrier(params.lu_comm);
}
int ref_init_matrix(int N, int v, int Px, int Py, int Pz, double* A_all) {
    InitArgs a{N, v, Px, Py, Pz, A_all};
    stub_mpi_run(Px * Py * Pz, init_main, &a);
    return 0;
}
}
