// include/conflux/lu/conflux_b200.hpp -- header-only C++ facade over the C ABI (include/conflux_b200.h) that keeps the
// reference's driver-facing names for the LU path, so that examples/conflux_miniapp.cpp reads like the reference's
// miniapp (examples/conflux_miniapp.cpp:88-167 there).  Reference interfaces mirrored:
//   conflux::lu_params<T>   src/conflux/lu/lu_params.hpp:8-459   (ctors :401-409, public fields :378-397)
//   conflux::LU_rep<T>      src/conflux/lu/conflux_opt.hpp:343-346
// MPI_Comm is replaced by conflux::comm_t (a cflx_comm*): one per rank, one GPU per rank.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../conflux_b200.h"

namespace conflux {

using comm_t = cflx_comm*;

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + cflx_last_error());
}

template <typename T>
class lu_params {
    static_assert(sizeof(T) == sizeof(double), "the B200 path is FP64 only (BASELINE.json)");

   public:
    comm_t lu_comm = nullptr;
    int rank = 0, pi = 0, pj = 0, pk = 0;
    int M = 0, N = 0, P = 0, Ml = 0, Nl = 0, Px = 0, Py = 0, Pz = 0;
    int v = 0, nlayr = 0, Mt = 0, Nt = 0, t = 0, tA11x = 0, tA11y = 0;
    int seed = 42;
    std::vector<T> data;  // local tiles, row-major Ml x Nl (conflux/COSTA tile layout, layout.cpp:95-109)
    bool use_collectives = false;
    cflx_lu* plan = nullptr;

    lu_params(int inpM, int inpN, int v_, comm_t comm) { initialize(inpM, inpN, v_, -1, -1, -1, comm); }
    lu_params(int inpM, int inpN, int v_, int Px_, int Py_, int Pz_, comm_t comm) {
        initialize(inpM, inpN, v_, Px_, Py_, Pz_, comm);
    }
    lu_params(const lu_params&) = delete;
    lu_params& operator=(const lu_params&) = delete;
    ~lu_params() { free_comms(); }

    void InitMatrix() {  // lu_params.hpp:141-376 (seeded branch)
        check(cflx_init_matrix_host(M, N, v, Px, Py, Pz, rank, seed, data.data()), "InitMatrix");
    }
    void free_comms() {
        if (plan) cflx_lu_destroy(plan);
        plan = nullptr;
    }

   private:
    void initialize(int inpM, int inpN, int v_, int Px_, int Py_, int Pz_, comm_t comm) {
        lu_comm = comm;
        check(cflx_lu_create(comm, inpM, inpN, v_, Px_, Py_, Pz_, &plan), "lu_params");
        int info[16];
        check(cflx_lu_info(plan, info), "lu_info");
        M = info[0]; N = info[1]; Ml = info[2]; Nl = info[3]; Nt = info[4]; nlayr = info[5]; P = info[6];
        Px = info[7]; Py = info[8]; Pz = info[9]; pi = info[10]; pj = info[11]; pk = info[12]; rank = info[13]; v = info[14];
        Mt = M / v; tA11x = Ml / v; tA11y = Nl / v; t = tA11y + 1;
        use_collectives = v > 1024;
        data.assign((std::size_t)Ml * Nl, T{0});
        InitMatrix();
    }
};

// Collective over gv.lu_comm; does not modify gv.data; C (>= Ml*Nl, may be null) and permutation (>= M, may be null)
// are filled as in the reference's validation build; returns the main-loop time in ms (truncated like the reference).
template <class T>
std::size_t LU_rep(lu_params<T>& gv, T* C, int* permutation) {
    double ms = 0;
    check(cflx_lu_set_local(gv.plan, gv.data.data()), "LU_rep: upload");
    check(cflx_lu_factor(gv.plan, &ms), "LU_rep: factor");
    if (C) check(cflx_lu_get_factors(gv.plan, C, permutation), "LU_rep: factors");
    else if (permutation) check(cflx_lu_get_permutation(gv.plan, permutation), "LU_rep: permutation");
    return (std::size_t)ms;
}

}  // namespace conflux
