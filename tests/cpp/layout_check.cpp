// tests/cpp/layout_check.cpp -- TEST INFRASTRUCTURE (build container only).  Compares, field by field, the COSTA
// descriptors produced by the REFERENCE's conflux_layout (src/conflux/lu/layout.cpp, both overloads; the 3-D one through
// lu_params::matrix on the thread-backed MPI stub) with the ones this repo's facade produces (layout_mine.cpp).
#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>
#include <random>
#include <tuple>

#include <conflux/lu/layout.hpp>
#include <conflux/lu/lu_params.hpp>

#include <atomic>
#include <cstdio>
#include <vector>

costa::grid_layout<double> mine_block_cyclic(double* data, int M, int N, int v, char ordering, int Px, int Py, int rank);
costa::grid_layout<double> mine_cart(double* data, int M, int N, int v, char ordering, int Px, int Py, int Pz, int pi,
                                     int pj, int pk);

static std::atomic<int> g_fail{0};
#define EXPECT(c)                                                            \
    do {                                                                     \
        if (!(c)) {                                                          \
            std::fprintf(stderr, "MISMATCH %s:%d: %s\n", __FILE__, __LINE__, #c); \
            g_fail++;                                                        \
        }                                                                    \
    } while (0)

static void same(costa::grid_layout<double>& a, costa::grid_layout<double>& b) {
    EXPECT(a.num_rows() == b.num_rows() && a.num_cols() == b.num_cols());
    EXPECT(a.num_blocks_row() == b.num_blocks_row() && a.num_blocks_col() == b.num_blocks_col());
    if (a.num_blocks_row() != b.num_blocks_row() || a.num_blocks_col() != b.num_blocks_col()) return;
    for (int i = 0; i < a.num_blocks_row(); ++i)
        for (int j = 0; j < a.num_blocks_col(); ++j) {
            EXPECT(a.grid.owner(i, j) == b.grid.owner(i, j));
            EXPECT(a.grid.rows_interval(i).start == b.grid.rows_interval(i).start);
            EXPECT(a.grid.rows_interval(i).end == b.grid.rows_interval(i).end);
            EXPECT(a.grid.cols_interval(j).start == b.grid.cols_interval(j).start);
            EXPECT(a.grid.cols_interval(j).end == b.grid.cols_interval(j).end);
        }
    EXPECT(a.blocks.num_blocks() == b.blocks.num_blocks());
    if (a.blocks.num_blocks() != b.blocks.num_blocks()) return;
    // the order in which a layout enumerates its local blocks is an internal detail (COSTA's ScaLAPACK variant walks
    // them column by column): compare them as a set keyed by the global block coordinates
    for (size_t k = 0; k < a.blocks.num_blocks(); ++k) {
        auto& x = a.blocks.get_block(k);
        bool found = false;
        for (size_t m = 0; m < b.blocks.num_blocks(); ++m) {
            auto& y = b.blocks.get_block(m);
            if (x.coordinates.row != y.coordinates.row || x.coordinates.col != y.coordinates.col) continue;
            found = true;
            EXPECT(x.data == y.data);
            EXPECT(x.stride == y.stride);
            EXPECT(x.n_rows() == y.n_rows() && x.n_cols() == y.n_cols());
        }
        EXPECT(found);
    }
}

struct Case { int N, v, Px, Py, Pz; };
static void rank_main(int, void* p) {
    auto* c = (Case*)p;
    conflux::lu_params<double> params(c->N, c->N, c->v, c->Px, c->Py, c->Pz, MPI_COMM_WORLD);
    auto mine = mine_cart(params.data.data(), params.M, params.N, params.v, 'R', params.Px, params.Py, params.Pz, params.pi,
                          params.pj, params.pk);
    same(params.matrix, mine);
    MPI_Barrier(params.lu_comm);
}

int main() {
    std::vector<double> buf(1 << 16);
    const int cases[][5] = {{16, 16, 4, 1, 1}, {64, 64, 8, 2, 2}, {96, 96, 8, 3, 3}, {48, 48, 4, 2, 2}};
    for (auto& c : cases)
        for (char ord : {'R', 'C'})
            for (int rank = 0; rank < c[3] * c[4]; ++rank) {
                auto ref = conflux::conflux_layout<double>(buf.data(), c[0], c[1], c[2], ord, c[3], c[4], rank);
                auto mine = mine_block_cyclic(buf.data(), c[0], c[1], c[2], ord, c[3], c[4], rank);
                same(ref, mine);
            }
    Case cs[] = {{64, 8, 2, 2, 1}, {64, 8, 2, 2, 2}, {96, 8, 3, 3, 1}, {64, 16, 1, 1, 2}, {100, 16, 2, 2, 1}};
    for (auto& c : cs) stub_mpi_run(c.Px * c.Py * c.Pz, rank_main, &c);
    if (g_fail.load()) {
        std::fprintf(stderr, "%d mismatches\n", g_fail.load());
        return 1;
    }
    std::puts("layouts identical");
    return 0;
}
