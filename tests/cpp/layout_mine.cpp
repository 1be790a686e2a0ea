// tests/cpp/layout_mine.cpp -- TEST INFRASTRUCTURE (build container only): this repo's conflux_layout descriptors,
// converted to real costa::grid_layout<double> objects (-DCONFLUX_B200_WITH_COSTA), for layout_check.cpp.
#include <conflux/lu/conflux_b200.hpp>

costa::grid_layout<double> mine_block_cyclic(double* data, int M, int N, int v, char ordering, int Px, int Py, int rank) {
    return conflux::conflux_layout<double>(data, M, N, v, ordering, Px, Py, rank).to_costa();
}
costa::grid_layout<double> mine_cart(double* data, int M, int N, int v, char ordering, int Px, int Py, int Pz, int pi,
                                     int pj, int pk) {
    conflux::cart_t c;
    c.dims[0] = Px; c.dims[1] = Py; c.dims[2] = Pz;
    c.coords[0] = pi; c.coords[1] = pj; c.coords[2] = pk;
    c.rank = c.cart_rank(pi, pj, pk);
    return conflux::conflux_layout<double>(data, M, N, v, ordering, c).to_costa();
}
