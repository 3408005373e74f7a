"""Pins the C restatement (oracle/cav_oracle.c, float64) against golden vectors minted from the unmodified
reference (tools/make_golden.py).  CPU only.

Tolerance: flags / done / game_over / step counters are compared EXACTLY.  Floats are compared at 1e-12
relative + 1e-12 absolute: the restatement follows the reference's expression order, but NumPy routes a few
operations through libm pow() (``x**2`` on NumPy scalars) and BLAS ddot (``np.dot`` / ``np.linalg.norm``), whose
last-bit rounding is platform-dependent, so a 1-ulp difference is legitimate.
"""
import numpy as np
import pytest

import common
from cav_oracle import Oracle
from gym_collision_avoidance_b200.state import HostState

RTOL = ATOL = 1e-12


@pytest.mark.parametrize("name", common.golden_names() + common.golden_names(cpu_only=True))
def test_teacher_forced_f64(name):
    """Every recorded step, from the reference's exact pre-step state: one oracle step == the reference's."""
    g = common.Golden(name)
    T = g.T
    c = g.cav_config(T, "f64", track_aux=True)
    st = HostState(c)
    for t in range(T):
        g.inject(st, t, t)
    ext = np.ascontiguousarray(g.z["ext_actions"][:T], dtype=np.float32)
    Oracle(c).step(st, ext)
    worst = common.compare_step(g, st, np.arange(T), np.arange(T), RTOL, ATOL)
    assert worst <= 1e-9


@pytest.mark.parametrize("name", [n for n in common.golden_names() + common.golden_names(cpu_only=True)
                                  if not n.startswith("mixed")])
def test_free_running_f64(name):
    """Reset from the scenario row and free-run: the oracle tracks the reference over whole episodes."""
    g = common.Golden(name)
    c = g.cav_config(1, "f64", track_aux=True, history_len=int(g.z["history"].shape[1]))
    st = HostState(c)
    st.static_map_bits = g.static_map_bits()
    orc = Oracle(c)
    orc.reset(st, g.bank())
    # reset state == golden step 0
    z = g.z
    n = g.n
    common.assert_close("reset.pv", st.pv[0, :n], z["st_pv"][0][:n], RTOL, ATOL)
    common.assert_close("reset.ht", st.ht[0, :n], z["st_ht"][0][:n], RTOL, ATOL)
    common.assert_close("reset.ego", st.ego[0, :n], z["st_ego"][0][:n], RTOL, ATOL)
    common.compare_obs(g, st.obs[0], z["obs"][0], RTOL, ATOL)
    for t in range(g.T):
        ext = np.ascontiguousarray(z["ext_actions"][t][None], dtype=np.float32)
        orc.step(st, ext)
        common.compare_step(g, st, np.array([0]), np.array([t]), 1e-7, 1e-7)
    # trajectories (Agent.global_state_history / ego_state_history)
    H = z["history"]
    for i in range(n):
        k = int(z["st_step_num"][g.T][i])
        common.assert_close("history", st.history[0, i, :k], H[i, :k], 1e-7, 1e-7)


def test_numpy_round_semantics():
    """round(np.float64, 2) == rint(x*100)/100 (what OtherAgentsStatesSensor.py:103 evaluates on NumPy floats)."""
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-20, 20, 200000), np.array([2.675, 0.285, 1.005, -0.125, 0.125, 2.5e-3])])
    want = np.array([round(np.float64(v), 2) for v in x])
    got = np.rint(x * 100.0) / 100.0
    assert np.array_equal(want, got)
