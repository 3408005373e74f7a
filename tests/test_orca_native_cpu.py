"""The kernels' ORCA solver, compiled for the host, against the reference's own RVO2 library -- bit for bit.

gym_collision_avoidance_b200/csrc/cav_orca_core.cuh is host/device source: nvcc compiles it into k_policy_move_v2
(compile-time N <= 10 with minimal sorting networks, run-time N with insertion in kd-tree order); here g++
compiles the same file (tests/native/orca_host.cpp, -ffp-contract=off like the library's -fmad=false) and every solve is
compared with oracle/_ref/librvo2_ref.so, the unmodified reference C++ (RVO2 Agent.cpp computeNewVelocity + update).
Dense scenes make agents overlap (the 1/timeStep branch) and make linearProgram3 run."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cav_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gym_collision_avoidance_b200", "csrc")
REF_SO = os.path.join(os.path.dirname(os.path.abspath(cav_oracle.__file__)), "_ref", "librvo2_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/librvo2_ref.so not built")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "liborca_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-fPIC",
                           "-shared", "-x", "c++", os.path.join(HERE, "native", "orca_host.cpp"), "-I", CSRC, "-o", out])
    L = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    L.host_orca_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp] + [ctypes.c_float] * 7 + \
        [ctypes.c_int, fp, fp]
    L.host_orca_solve.restype = ctypes.c_int
    L.host_orca_solve_kd.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp] + [ctypes.c_float] * 7 + \
        [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, fp, fp]
    L.host_orca_solve_kd.restype = ctypes.c_int
    L.host_orca_solve_kd_ct.argtypes = L.host_orca_solve_kd.argtypes
    L.host_orca_solve_kd_ct.restype = ctypes.c_int
    return L


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("n,side,trials", [(2, 3.0, 500), (3, 2.5, 500), (4, 3.0, 500), (5, 3.0, 600), (6, 3.0, 600),
                                           (8, 3.5, 600), (10, 4.0, 800), (10, 8.0, 400), (12, 4.5, 500),
                                           (17, 5.0, 300), (24, 6.0, 200)])
def test_device_orca_source_bit_exact_vs_reference_rvo2(host, n, side, trials):
    import rvo2_ref
    R = rvo2_ref.lib()
    rng = np.random.RandomState(n * 977 + int(side * 10))
    dt, th = np.float32(0.1), np.float32(5.0)
    sim = R.rvo2ref_new(dt, np.float32(np.inf), 64, th, th, 0.0, 0.0, 0.0, 0.0)
    variants = [2] + ([3] if n <= 10 else [])
    overlaps = 0
    for trial in range(trials):
        pos = rng.uniform(-side / 2, side / 2, size=(n, 2)).astype(np.float32)
        vel = (rng.uniform(-1.5, 1.5, size=(n, 2)) * (rng.rand(n, 1) > 0.2)).astype(np.float32)
        rad = (rng.uniform(0.2, 0.8, size=n) * 1.05).astype(np.float32)
        msp = rng.uniform(0.5, 2.0, size=n).astype(np.float32)
        goal = rng.uniform(-side, side, size=(n, 2))
        pref = goal - pos
        pref = (msp[:, None] * pref / np.linalg.norm(pref, axis=1, keepdims=True)).astype(np.float32)
        ego = int(rng.randint(n))
        collab = np.float32(0.5)
        ref_pos = np.zeros(2, np.float32)
        ref_vel = np.zeros(2, np.float32)
        R.rvo2ref_solve_ego(sim, n, ego, _fp(pos), _fp(vel), _fp(pref), _fp(rad), _fp(msp), collab, _fp(ref_pos),
                            _fp(ref_vel))
        for v in variants:
            ovx, ovy = ctypes.c_float(), ctypes.c_float()
            rc = host.host_orca_solve(v, n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1], th,
                                      dt, collab, np.float32(np.inf), 64, ctypes.byref(ovx), ctypes.byref(ovy))
            assert rc == 0
            got = np.array([ovx.value, ovy.value], np.float32)
            assert got.tobytes() == ref_vel.tobytes(), (v, trial, got, ref_vel)
        d = np.linalg.norm(pos - pos[ego], axis=1)
        overlaps += int(((d < rad + rad[ego]) & (np.arange(n) != ego)).any())
    R.rvo2ref_delete(sim)
    if n >= 5 and side <= 6.0:
        assert overlaps > 0


def test_finite_neighbor_dist_and_max_neighbors_match_reference(host):
    """SENSING_HORIZON < inf -> neighborDist (RVOPolicy.py:12-16) and a neighbour cap below n-1."""
    import rvo2_ref
    R = rvo2_ref.lib()
    rng = np.random.RandomState(5)
    dt, th, nd = np.float32(0.1), np.float32(5.0), np.float32(3.0)
    n = 10
    for maxn in (64, 4):
        sim = R.rvo2ref_new(dt, nd, maxn, th, th, 0.0, 0.0, 0.0, 0.0)
        for trial in range(300):
            pos = rng.uniform(-3, 3, size=(n, 2)).astype(np.float32)
            vel = rng.uniform(-1.5, 1.5, size=(n, 2)).astype(np.float32)
            rad = (rng.uniform(0.2, 0.8, size=n) * 1.05).astype(np.float32)
            msp = rng.uniform(0.5, 2.0, size=n).astype(np.float32)
            pref = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
            ego = int(rng.randint(n))
            ref_pos, ref_vel = np.zeros(2, np.float32), np.zeros(2, np.float32)
            R.rvo2ref_solve_ego(sim, n, ego, _fp(pos), _fp(vel), _fp(pref), _fp(rad), _fp(msp), np.float32(0.5),
                                _fp(ref_pos), _fp(ref_vel))
            for v in (2, 3):
                ovx, ovy = ctypes.c_float(), ctypes.c_float()
                assert host.host_orca_solve(v, n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1],
                                            th, dt, np.float32(0.5), nd, maxn, ctypes.byref(ovx), ctypes.byref(ovy)) == 0
                got = np.array([ovx.value, ovy.value], np.float32)
                assert got.tobytes() == ref_vel.tobytes(), (maxn, v, trial, got, ref_vel)
        R.rvo2ref_delete(sim)


@pytest.mark.parametrize("n,kind", [(12, "circle"), (16, "circle"), (20, "circle"), (20, "lattice"), (33, "lattice"),
                                    (64, "lattice"), (11, "stack"), (24, "circle_finite"), (13, "lattice"), (16, "lattice"),
                                    (14, "circle_finite"), (15, "stack")])
def test_device_kdtree_source_tie_order_bit_exact_vs_reference_rvo2(host, n, kind):
    """cav_kdtree.cuh (the kernels' kd-tree: build over the persistent permutation + visiting order) and the generic
    solver's insertion in that order, on scenes made of exact float32 distance ties, stepped for 60 doSteps against ONE
    persistent reference simulator: bit-identical new velocity for every ego at every step (the same scenes as
    tests/test_oracle_rvo2.py::test_kdtree_tie_order_bit_exact_vs_reference_rvo2, which pins the oracle)."""
    import rvo2_ref
    R = rvo2_ref.lib()
    rng = np.random.RandomState(n)
    dt, th = np.float32(0.1), np.float32(5.0)
    nd = np.float32(3.0) if kind == "circle_finite" else np.float32(np.inf)
    sim = R.rvo2ref_new(dt, nd, 64, th, th, 0.0, 0.0, 0.0, 0.0)
    perm = np.arange(n, dtype=np.uint8)
    if kind.startswith("circle"):
        ang = 2 * np.pi * np.arange(n) / n
        pos0 = np.stack([4 * np.cos(ang), 4 * np.sin(ang)], 1)
        goal = -pos0
    elif kind == "lattice":
        side = int(np.ceil(np.sqrt(n)))
        pos0 = np.array([[i % side, i // side] for i in range(n)], np.float64) * 1.5
        pos0 = pos0[rng.permutation(n)]
        goal = pos0[::-1].copy()
    else:
        pos0 = np.zeros((n, 2))
        pos0[n // 2:] = [1.0, 0.0]
        goal = pos0 + [3.0, 1.0]
    pos = pos0.astype(np.float32)
    vel = np.zeros((n, 2), np.float32)
    rad = np.full(n, 0.5 * 1.05, np.float32)
    msp = np.ones(n, np.float32)
    for step in range(60):
        pref = goal - pos
        pref = (msp[:, None] * pref / np.maximum(np.linalg.norm(pref, axis=1, keepdims=True), 1e-9)).astype(np.float32)
        rp, rv = np.zeros(2, np.float32), np.zeros(2, np.float32)
        R.rvo2ref_solve_ego(sim, n, 0, _fp(pos), _fp(vel), _fp(pref), _fp(rad), _fp(msp), np.float32(0.5), _fp(rp), _fp(rv))
        all_ref = np.zeros((n, 2), np.float32)
        for a in range(n):
            R.rvo2ref_get_velocity(sim, a, _fp(all_ref[a]))
        for ego in range(n):
            ovx, ovy = ctypes.c_float(), ctypes.c_float()
            rc = host.host_orca_solve_kd(n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1], th, dt,
                                         np.float32(0.5), nd, 64, perm.ctypes.data, int(ego == 0), ctypes.byref(ovx),
                                         ctypes.byref(ovy))
            assert rc == 0
            got = np.array([ovx.value, ovy.value], np.float32)
            assert got.tobytes() == all_ref[ego].tobytes(), (kind, step, ego, got, all_ref[ego], perm.tolist())
            if 11 <= n <= 16:  # the compile-time solver (visiting rank as the tie-break field of a sorting-network key)
                cvx, cvy = ctypes.c_float(), ctypes.c_float()
                rc = host.host_orca_solve_kd_ct(n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1],
                                                th, dt, np.float32(0.5), nd, 64, perm.ctypes.data, 0, ctypes.byref(cvx),
                                                ctypes.byref(cvy))
                assert rc == 0
                got = np.array([cvx.value, cvy.value], np.float32)
                assert got.tobytes() == all_ref[ego].tobytes(), ("ct", kind, step, ego, got, all_ref[ego], perm.tolist())
        vel = all_ref
        pos = (pos + vel * dt).astype(np.float32)
    R.rvo2ref_delete(sim)
