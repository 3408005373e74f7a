"""csrc/cav_hd.cuh::cell_coord (the world -> grid-cell mapping of k_laserscan and the wall-collision test), compiled for
the host, gives the reference's cell -- np.floor(origin -/+ v / cell), Map.py:28-29,36-37 -- for every input, including
coordinates that sit on or next to cell borders (where its fast path v * (1/cell) would round differently)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gym_collision_avoidance_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libgeometry_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "native", "geometry_host.cpp"), "-I", CSRC, "-o", out])
    return ctypes.CDLL(out)


def _run(lib, origin, v, cell, negate):
    origin = np.ascontiguousarray(np.broadcast_to(origin, v.shape), dtype=np.float64)
    out = np.zeros_like(v)
    lib.host_cell_coord(ctypes.c_int(v.size), origin.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p),
                        ctypes.c_double(cell), ctypes.c_int(int(negate)), out.ctypes.data_as(ctypes.c_void_p))
    return out


@pytest.mark.parametrize("cell", [0.1, 0.05, 0.25, 0.3])
def test_cell_coord_equals_numpy_floor_everywhere(lib, cell):
    rng = np.random.RandomState(int(cell * 1000))
    origin = (16.0 / 2.0) / cell  # Map.py:23
    vals = [rng.uniform(-12, 12, 400000)]
    k = np.arange(-130, 131, dtype=np.float64)
    grid = k * cell                                  # border coordinates as the arithmetic produces them
    vals += [grid, np.nextafter(grid, np.inf), np.nextafter(grid, -np.inf)]
    vals += [np.round(k * cell, 10), k / (1.0 / cell)]
    for ulps in (2, 3, 17, 1000):
        vals += [grid + ulps * np.spacing(grid), grid - ulps * np.spacing(grid)]
    # sample points like the laser's: pos + r * cos/sin with r on the 0.1 grid
    p = rng.uniform(-8, 8, 20000)
    r = rng.randint(0, 60, 20000) * 0.1
    th = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2, np.pi / 4], 20000)
    vals += [p + r * np.cos(th), np.round(p, 1) + r * np.cos(th)]
    v = np.ascontiguousarray(np.concatenate(vals))
    for negate in (True, False):
        want = np.floor(origin - v / cell) if negate else np.floor(origin + v / cell)
        got = _run(lib, origin, v, cell, negate)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (cell, negate, v[bad[:5]], got[bad[:5]], want[bad[:5]])
