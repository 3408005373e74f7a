// Host build of the grid-cell helper the laserscan and wall-collision kernels use (csrc/cav_hd.cuh::cell_coord).
// TEST INFRASTRUCTURE (tests/test_geometry_native_cpu.py).
#include "cav_hd.cuh"

extern "C" void host_cell_coord(int n, const double *origin, const double *v, double cell, int negate, double *out) {
    const double inv_cell = 1.0 / cell;
    for (int i = 0; i < n; ++i) out[i] = cav::cell_coord(origin[i], v[i], cell, inv_cell, negate != 0);
}
