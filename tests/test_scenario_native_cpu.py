"""CPU coverage of the DEVICE scenario generator: cav_scenario_gen.cuh is host/device code, so the source that nvcc
compiles into k_generate_scenarios is compiled here with g++ and exercised without a GPU -- Philox4x32-10 known
answers (Random123's published vectors), the construction invariants, and the distribution against the host generator
that reproduces the reference bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from gym_collision_avoidance_b200.envs import scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gym_collision_avoidance_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libscenario_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "native", "scenario_host.cpp"), "-I", CSRC, "-o", out])
    L = ctypes.CDLL(out)
    L.host_generate_scenarios.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64] + \
        [ctypes.c_double] * 7
    L.host_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    return L


def _gen(L, S, N, seed):
    rows = np.zeros((S, N, 8))
    kind = np.zeros(S, np.int32)
    lo, hi = (4.0, 5.0) if N < 5 else (6.0, 8.0)
    L.host_generate_scenarios(rows.ctypes.data, kind.ctypes.data, S, N, seed, lo, hi, 0.5, 2.0, 0.2, 0.8,
                              scenarios.GETTING_CLOSE_RANGE)
    return rows, kind


def test_philox4x32_10_known_answers(lib):
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        o = (ctypes.c_uint32 * 4)()
        lib.host_philox_block(c, k, o)
        assert tuple(o) == want


def test_uniform_stream_is_53_bit_and_distinct_per_scenario(lib):
    a = np.zeros(4096)
    b = np.zeros(4096)
    lib.host_uniforms(7, 0, 4096, a.ctypes.data)
    lib.host_uniforms(7, 1, 4096, b.ctypes.data)
    assert (a >= 0).all() and (a < 1).all() and not np.array_equal(a, b)
    assert np.all(a * 2.0 ** 53 == np.floor(a * 2.0 ** 53))          # multiples of 2^-53
    assert abs(a.mean() - 0.5) < 0.02 and abs(a.var() - 1 / 12.0) < 0.01 and len(np.unique(a)) == 4096


def _host_kind(rows):
    swap = (rows[:, 0, 1] == 0.0) & (rows[:, 1, 1] == 0.0) & (rows[:, 0, 0] == -rows[:, 1, 0])
    antipodal = np.all(np.abs(rows[:, :, 0:2] + rows[:, :, 2:4]) < 1e-9, axis=(1, 2))
    return np.where(swap, 0, np.where(antipodal, 1, 2))


@pytest.mark.parametrize("N", [3, 10])
def test_construction_invariants(lib, N):
    rows, kind = _gen(lib, 1500, N, 5)
    assert np.isfinite(rows).all() and np.array_equal(kind, _host_kind(rows))
    r = rows[..., 5]
    assert (r >= 0.2).all() and (r <= 0.8).all() and (rows[..., 4] >= 0.5).all() and (rows[..., 4] <= 2.0).all()
    for c0 in (0, 2):
        p = rows[:, :, c0:c0 + 2]
        d = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1) + np.eye(N)[None] * 1e9
        assert (d >= r[:, :, None] + r[:, None] + scenarios.GETTING_CLOSE_RANGE - 1e-9).all()
    assert np.allclose(rows[..., 6], np.arctan2(rows[..., 3] - rows[..., 1], rows[..., 2] - rows[..., 0]), atol=1e-12)
    for s in np.nonzero(kind == 2)[0][:120]:  # random cases: no agent may simply drive straight past all earlier ones
        tc = rows[s]
        for i in range(1, N):
            assert not all(scenarios._straight_lines_clear(tc[j, 0:2], tc[j, 2:4], tc[j, 4], tc[i, 0:2], tc[i, 2:4],
                                                           tc[i, 4], tc[j, 5] + tc[i, 5] + scenarios.GETTING_CLOSE_RANGE)
                           for j in range(i))
    again, _ = _gen(lib, 1500, N, 5)
    other, _ = _gen(lib, 1500, N, 6)
    assert np.array_equal(again, rows) and not np.array_equal(other, rows)


def test_distribution_matches_the_reference_generator(lib):
    N, S = 10, 4000
    drows, dkind = _gen(lib, S, N, 11)
    hrows = scenarios.random_scenario_bank(S, N, seed=11)
    hkind = _host_kind(hrows)

    def summ(rows, kind):
        d = np.linalg.norm(rows[:, :, 2:4] - rows[:, :, 0:2], axis=-1)
        rnd = kind == 2
        return dict(swap=np.mean(kind == 0), circle=np.mean(kind == 1), radius=rows[..., 5].mean(),
                    speed=rows[..., 4].mean(), sg_random=d[rnd].mean(), sg_all=d.mean(),
                    extent=np.abs(rows[rnd][..., 0:4]).max(axis=(1, 2)).mean(),
                    ring=np.linalg.norm(rows[kind == 1][..., 0:2], axis=-1).mean())
    h, d = summ(hrows, hkind), summ(drows, dkind)
    tol = dict(swap=0.025, circle=0.025, radius=0.006, speed=0.012, sg_random=0.18, sg_all=0.18, extent=0.18, ring=0.1)
    bad = {k: (h[k], d[k]) for k in tol if abs(h[k] - d[k]) > tol[k]}
    assert not bad, bad
