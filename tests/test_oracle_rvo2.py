"""Pins the float32 ORCA restatement (oracle/cav_oracle.c: cavo_orca_solve) BIT-EXACTLY against the unmodified
reference RVO2 C++ (oracle/_ref/librvo2_ref.so, built from /root/reference/.../Python-RVO2/src by oracle/Makefile).
CPU only; skipped when the prebuilt reference library is absent."""
import ctypes
import os

import numpy as np
import pytest

import cav_oracle

REF_SO = os.path.join(os.path.dirname(os.path.abspath(cav_oracle.__file__)), "_ref", "librvo2_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/librvo2_ref.so not built")


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("n,side,trials", [(2, 3.0, 400), (5, 3.0, 600), (10, 4.0, 600), (10, 8.0, 400),
                                           (20, 6.0, 300), (40, 9.0, 100), (64, 12.0, 40)])
def test_orca_bit_exact_vs_reference_rvo2(n, side, trials):
    import rvo2_ref
    R = rvo2_ref.lib()
    L = cav_oracle.lib()
    rng = np.random.RandomState(n * 1000 + int(side))
    dt, th = np.float32(0.1), np.float32(5.0)
    sim = R.rvo2ref_new(dt, np.float32(np.inf), 64, th, th, 0.0, 0.0, 0.0, 0.0)
    n_lp3 = n_overlap = 0
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
        ovx, ovy = ctypes.c_float(), ctypes.c_float()
        nl = ctypes.c_int()
        lines = np.zeros((64, 4), np.float32)
        L.cavo_orca_solve(n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1], th, dt, collab,
                          np.float32(np.inf), 64, ctypes.byref(ovx), ctypes.byref(ovy), lines.ctypes.data,
                          ctypes.byref(nl))
        got = np.array([ovx.value, ovy.value], np.float32)
        assert got.tobytes() == ref_vel.tobytes(), (trial, got, ref_vel)
        # Agent::update (Agent.cpp:470-475): position_ += velocity_ * timeStep_
        new_pos = (pos[ego] + got * dt).astype(np.float32)
        assert new_pos.tobytes() == ref_pos.tobytes()
        assert nl.value == n - 1
        d = np.linalg.norm(pos - pos[ego], axis=1)
        n_overlap += int(((d < rad + rad[ego]) & (np.arange(n) != ego)).any())
    R.rvo2ref_delete(sim)
    if n >= 5 and side <= 6.0:
        assert n_overlap > 0  # the dense cases must exercise the already-colliding branch


def _solve_perm(L, n, ego, pos, vel, rad, msp, pref, th, dt, perm, rebuild, neighbor_dist=np.inf, max_nb=64):
    ovx, ovy = ctypes.c_float(), ctypes.c_float()
    L.cavo_orca_solve_perm(n, ego, _fp(pos), _fp(vel), _fp(rad), msp[ego], pref[ego, 0], pref[ego, 1], th, dt,
                           np.float32(0.5), np.float32(neighbor_dist), max_nb, perm.ctypes.data, int(rebuild),
                           ctypes.byref(ovx), ctypes.byref(ovy))
    return np.array([ovx.value, ovy.value], np.float32)


@pytest.mark.parametrize("n,kind", [(12, "circle"), (16, "circle"), (20, "circle"), (20, "lattice"), (33, "lattice"),
                                    (64, "lattice"), (11, "stack"), (24, "circle_finite")])
def test_kdtree_tie_order_bit_exact_vs_reference_rvo2(n, kind):
    """Exact float32 distance ties: the neighbour order then depends on the reference's kd-tree leaf order, i.e. on
    KdTree::agents_, which buildAgentTree re-partitions in place on every doStep and never resets (KdTree.cpp:47-111).
    One persistent reference simulator is stepped through a whole sequence of symmetric scenes; the restatement
    (oracle/cav_oracle.c: cavo_kd_build / kd_query_rec on one persistent permutation) must give bit-identical new
    velocities for EVERY ego at EVERY step."""
    import rvo2_ref
    R = rvo2_ref.lib()
    L = cav_oracle.lib()
    L.cavo_orca_solve_perm.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_float)] * 3 + \
        [ctypes.c_float] * 7 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
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
    else:  # several agents at the SAME position: left == begin special case (KdTree.cpp:99-102)
        pos0 = np.zeros((n, 2))
        pos0[n // 2:] = [1.0, 0.0]
        goal = pos0 + [3.0, 1.0]
    pos = pos0.astype(np.float32)
    vel = np.zeros((n, 2), np.float32)
    rad = np.full(n, 0.5 * 1.05, np.float32)
    msp = np.ones(n, np.float32)
    n_tie_steps = 0
    for step in range(60):
        pref = goal - pos
        nrm = np.maximum(np.linalg.norm(pref, axis=1, keepdims=True), 1e-9)
        pref = (msp[:, None] * pref / nrm).astype(np.float32)
        new_vel = np.zeros_like(vel)
        for ego in range(n):
            ref_pos = np.zeros(2, np.float32)
            ref_vel = np.zeros(2, np.float32)
            if ego == 0:  # one doStep per env step advances the persistent tree exactly once on both sides
                R.rvo2ref_solve_ego(sim, n, ego, _fp(pos), _fp(vel), _fp(pref), _fp(rad), _fp(msp), np.float32(0.5),
                                    _fp(ref_pos), _fp(ref_vel))
                all_ref = np.zeros((n, 2), np.float32)
                for a in range(n):  # doStep solved every agent with the same tree: read them all
                    R.rvo2ref_get_velocity(sim, a, _fp(all_ref[a]))
            got = _solve_perm(L, n, ego, pos, vel, rad, msp, pref, th, dt, perm, rebuild=(ego == 0), neighbor_dist=nd)
            assert got.tobytes() == all_ref[ego].tobytes(), (kind, step, ego, got, all_ref[ego], perm.tolist())
            new_vel[ego] = got
        d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1).astype(np.float32)
        n_tie_steps += int(any(len(np.unique(np.delete(d2[i], i))) < n - 1 for i in range(n)))
        vel = new_vel
        pos = (pos + vel * dt).astype(np.float32)
    R.rvo2ref_delete(sim)
    assert n_tie_steps > 0  # the scene really had exact distance ties
    assert n <= 10 or kind == "stack" or not np.array_equal(perm, np.arange(n)), "the tree never re-ordered agents_"
