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
