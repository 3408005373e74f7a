"""time_to_impact (reference util.compute_time_to_impact, util.py:23-112) -- the oracle's restatement and the KERNELS'
source (csrc/cav_tti.cuh, compiled here with g++) against 6000 values computed by the reference's own Python function
(tests/golden/generator/time_to_impact.npz, minted by tools/make_tti_golden.py): approaching, receding, overlapping,
grazing, near-vertical and near-zero relative velocities."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cav_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gym_collision_avoidance_b200", "csrc")
CASES = np.load(os.path.join(HERE, "golden", "generator", "time_to_impact.npz"))["cases"]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libtti_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "native", "tti_host.cpp"), "-I", CSRC, "-o", out])
    return ctypes.CDLL(out)


def _check_f64(got, want):
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)            # sqrt of a negative discriminant, like the reference
    assert np.array_equal(np.isinf(got), np.isinf(want))  # outside the cone / not approaching
    assert np.array_equal(got == 0.0, want == 0.0)        # already overlapping
    f = np.isfinite(want)
    # NumPy evaluates x**2 through pow() and norms through BLAS dot: last-bit differences are legitimate, and the
    # quadratic's cancellation amplifies them for grazing lines
    err = np.abs(got[f] - want[f]) / np.maximum(1.0, np.abs(want[f]))
    assert np.percentile(err, 99) < 1e-12 and err.max() < 1e-6, (np.percentile(err, 99), err.max())
    assert (err == 0).mean() > 0.5


def test_oracle_time_to_impact_matches_reference_values():
    L = cav_oracle.lib()
    L.cavo_time_to_impact_f64.restype = ctypes.c_double
    L.cavo_time_to_impact_f64.argtypes = [ctypes.c_double] * 9
    got = np.array([L.cavo_time_to_impact_f64(*row[:9]) for row in CASES])
    _check_f64(got, CASES[:, 9])


def test_kernel_source_time_to_impact_matches_reference_values_and_oracle_bitwise(host):
    inp = np.ascontiguousarray(CASES[:, :9])
    got = np.zeros(len(inp))
    host.host_tti_f64(ctypes.c_int(len(inp)), inp.ctypes.data_as(ctypes.c_void_p), got.ctypes.data_as(ctypes.c_void_p))
    _check_f64(got, CASES[:, 9])
    # same expression order as the oracle: identical bits
    L = cav_oracle.lib()
    L.cavo_time_to_impact_f64.restype = ctypes.c_double
    L.cavo_time_to_impact_f64.argtypes = [ctypes.c_double] * 9
    orc = np.array([L.cavo_time_to_impact_f64(*row) for row in inp])
    assert np.array_equal(got.view(np.uint64), orc.view(np.uint64))
    # float32 instantiation == the oracle's float32 build, bit for bit
    in32 = np.ascontiguousarray(inp, dtype=np.float32)
    got32 = np.zeros(len(inp), np.float32)
    host.host_tti_f32(ctypes.c_int(len(inp)), in32.ctypes.data_as(ctypes.c_void_p), got32.ctypes.data_as(ctypes.c_void_p))
    L.cavo_time_to_impact_f32.restype = ctypes.c_float
    L.cavo_time_to_impact_f32.argtypes = [ctypes.c_float] * 9
    orc32 = np.array([L.cavo_time_to_impact_f32(*[float(x) for x in row]) for row in in32], np.float32)
    assert np.array_equal(got32.view(np.uint32), orc32.view(np.uint32))
