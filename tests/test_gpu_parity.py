"""-m gpu: parity of the sm_100a kernels (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star: 1e-5 relative, done / collision flags bit-exact):
  * CAV_F64 vs golden (reference, float64): flags / done / game_over / step counters EXACT; floats 1e-9 (the only
    legitimate difference is CUDA libm vs glibc sin/cos/atan2 at the last ulp).  This is the precision the
    north_star tolerance is claimed for.
  * CAV_F64 vs oracle f64 and CAV_F32 vs oracle f32 on the same injected state: flags EXACT; floats 1e-12 / 2e-6.
  * CAV_F32 (float4 columns) vs golden: flags EXACT on these fixtures; positions / distances / times 1e-4 relative;
    velocity-like values (velocities, commanded speed) 5e-4 absolute and angles 1e-3 rad: the reference derives the
    commanded speed / heading from (float32 new position - position) / dt, so a float32-rounded position carries
    ulp(pos)/dt ~ 1e-5..1e-4 m/s of noise through the ORCA solve, and the heading of a nearly
    stopped agent is atan2 of that noisy displacement; laser beams may differ where a
    float32-rounded position moves a sample across a 0.1 m cell border (<= 0.5 % of beams).
"""
import numpy as np
import pytest

import common
from cav_oracle import Oracle
from gym_collision_avoidance_b200.state import HostState

pytestmark = pytest.mark.gpu

GOLDEN = common.golden_names()


def _inject_all(g, c):
    st = HostState(c)
    for t in range(g.T):
        g.inject(st, t, t)
    return st


@pytest.fixture(params=["tpa", "group"])
def mapping(request, monkeypatch):
    """Run with both kernel families: one thread per agent (default: k_policy_move_v2 for N <= 32, pairwise+obs for
    N <= 24) and sub-warp groups."""
    monkeypatch.setenv("CAV_TPA_MAX_N", "32" if request.param == "tpa" else "0")
    return request.param


# The reference's exactly symmetric circle presets put agents ON branch cuts: headings of exactly +-pi, lateral velocity
# components of +-1e-10, turning_dir flipping between pi and -0.  One step from the reference's state is reproduced to
# 1e-9 with exact flags and turning_dir (teacher-forced tests below, float64) -- that is the parity claim -- but
#  * the float32 layout rounds those states onto the other side of a cut, and
#  * a free run amplifies the legal <= 2 ulp difference between CUDA's and glibc's sin / cos / atan2 across a cut
#    within a few hundred steps,
# so those two kinds of test skip the exactly symmetric fixtures (the perturbation-free CPU oracle tracks them over whole
# episodes, tests/test_oracle_golden.py).
SYMMETRIC = ("circle10_r5", "circle12_r6", "circle16_r8", "circle20_r10", "laser20_circle")


@pytest.mark.parametrize("name", GOLDEN)
def test_step_teacher_forced_f64_vs_reference_golden(name, mapping):
    from gpu_common import GpuRunner
    g = common.Golden(name)
    c = g.cav_config(g.T, "f64", track_aux=True)
    st = _inject_all(g, c)
    ext = np.ascontiguousarray(g.z["ext_actions"][:g.T], dtype=np.float32)
    out = GpuRunner(c).step(st, ext)
    common.compare_step(g, out, np.arange(g.T), np.arange(g.T), 1e-9, 1e-9)


@pytest.mark.parametrize("name", GOLDEN)
def test_step_teacher_forced_f32_vs_reference_golden(name, mapping):
    from gpu_common import GpuRunner
    if name in SYMMETRIC:
        pytest.skip("exactly symmetric preset: float32 state lands on the other side of +-pi / +-0 branch cuts")
    if name == "walls6":
        pytest.skip("float32-rounded positions move an agent across a 0.1 m map-cell border next to the wall, so the "
                    "wall-collision flag of the float64 reference is not reproducible in the float32 layout; the "
                    "float32 path is checked exactly against the float32 oracle in test_step_vs_oracle_same_state")
    g = common.Golden(name)
    c = g.cav_config(g.T, "f32", track_aux=True)
    st = _inject_all(g, c)
    ext = np.ascontiguousarray(g.z["ext_actions"][:g.T], dtype=np.float32)
    out = GpuRunner(c).step(st, ext)
    tol = common.Tol(1e-4, 1e-4, vel=(0.0, 5e-4), ang=(0.0, 1e-3), laser_frac=5e-3)
    common.compare_step(g, out, np.arange(g.T), np.arange(g.T), tol)


@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-6)])
@pytest.mark.parametrize("name", GOLDEN)
def test_step_vs_oracle_same_state(name, prec, tol, mapping):
    from gpu_common import GpuRunner
    if name in SYMMETRIC:
        if prec == "f32":
            pytest.skip("exactly symmetric preset: float32 state lands on the other side of +-pi / +-0 branch cuts")
        # whole 400-step episodes: the commanded heading change is stored as float32 (env.py:305-307), which turns a
        # 1-ulp float64 libm difference that straddles a float32 rounding boundary into one float32 ulp of the command
        tol = 1e-9
    g = common.Golden(name)
    c = g.cav_config(g.T, prec, track_aux=True)
    st = _inject_all(g, c)
    ext = np.ascontiguousarray(g.z["ext_actions"][:g.T], dtype=np.float32)
    want = st.copy()
    Oracle(c).step(want, ext)
    got = GpuRunner(c).step(st, ext)
    common.compare_states(got, want, g.n, tol, tol)


@pytest.mark.parametrize("name", ["rvo10_c0", "rvo4_c7", "rvo3_in8slots", "mixed5_train", "dense6_clip3_last", "tti8_clip4",
                                  "horizon8", "laser6", "walls6", "rvo20_rand", "circle12_r6", "single_agent_train"])
def test_f64_state_with_f32_observation_columns(name, mapping):
    """cfg.obs_f32: float64 state, `obs` and `reward` written as float32 by the kernels (the reference's DECLARED
    observation dtype, config.py:93-170).  State / flags as in the float64 tests; the float32 columns equal the float64
    values of the reference rounded to float32, up to one float32 ulp where CUDA's and glibc's float64 libm differ in the
    last bit (tolerance 2e-7 relative + 1e-7 absolute, far inside north_star's 1e-5); laser beams exact."""
    from gpu_common import GpuRunner
    g = common.Golden(name)
    c = g.cav_config(g.T, "f64", track_aux=True, obs_f32=True)
    st = _inject_all(g, c)
    assert st.obs.dtype == np.float32 and st.reward.dtype == np.float32 and st.pv.dtype == np.float64
    ext = np.ascontiguousarray(g.z["ext_actions"][:g.T], dtype=np.float32)
    want = st.copy()
    Oracle(c).step(want, ext)
    out = GpuRunner(c).step(st, ext)
    assert out.obs.dtype == np.float32
    tol = 1e-9 if name in SYMMETRIC else 1e-12
    common.compare_states(out, want, g.n, tol, tol, cols=[k for k in want.arrays if k not in (
        "obs", "reward", "stats", "episode", "kinds", "num_active")])
    f32 = common.Tol(2e-7, 1e-7)
    common.compare_obs(g, out.obs, want.obs, f32)
    common.assert_close("reward", out.reward[:, :g.n], want.reward[:, :g.n], 2e-7, 1e-7)
    # and against the reference's golden float64 observation
    common.compare_obs(g, out.obs, g.z["obs"][1:g.T + 1], f32)


@pytest.mark.parametrize("name", ["rvo10_c0", "mixed5_train", "dense6_clip3_last"])
def test_phase_entry_points_vs_oracle(name, mapping):
    """cav_k_policy / cav_k_dynamics / cav_k_pairwise_obs each equal the oracle's phase on the same state."""
    from gpu_common import GpuRunner, to_dev
    import torch
    g = common.Golden(name)
    c = g.cav_config(g.T, "f64", track_aux=True)
    st = _inject_all(g, c)
    ext = np.ascontiguousarray(g.z["ext_actions"][:g.T], dtype=np.float32)
    orc = Oracle(c)
    r = GpuRunner(c)
    r.upload(st)
    want = st.copy()
    orc.policy(want, ext)
    r.cav.k_policy(to_dev(ext))
    got = r.download()
    common.assert_close("actions", got.actions, want.actions, 1e-6, 1e-7)
    common.assert_close("goal(static)", got.goal, want.goal, 1e-12, 1e-12)
    orc.dynamics(want)
    r.upload(want_before := _with_actions(st, want.actions, want.goal))
    r.cav.k_dynamics()
    got = r.download()
    common.compare_states(got, want, g.n, 1e-12, 1e-12, cols=("pv", "ht", "ego", "flags", "step_num", "aux0",
                                                               "past_actions", "past_vel"))
    orc.pairwise_obs(want)
    r.upload(got)
    r.cav.k_pairwise_obs()
    got2 = r.download()
    common.compare_states(got2, want, g.n, 1e-12, 1e-12)


def _with_actions(st, actions, goal):
    s = st.copy()
    s.actions[...] = actions
    s.goal[...] = goal
    return s


@pytest.mark.parametrize("name", [n for n in GOLDEN if not n.startswith("mixed") and n != "circle20_r10"])
def test_free_running_f64_tracks_reference(name, mapping):
    """Reset from the scenario row and free-run whole episodes on the GPU (float64 state)."""
    from gpu_common import GpuRunner, to_dev
    g = common.Golden(name)
    c = g.cav_config(1, "f64", track_aux=True, history_len=int(g.z["history"].shape[1]))
    r = GpuRunner(c)
    r.set_static_map(g.static_map_bits())
    r.cav.reset(r.bank(g.bank()))
    out = r.download()
    n = g.n
    common.assert_close("reset.pv", out.pv[0, :n], g.z["st_pv"][0][:n], 1e-12, 1e-12)
    common.assert_close("reset.ht", out.ht[0, :n], g.z["st_ht"][0][:n], 1e-12, 1e-12)
    common.compare_obs(g, out.obs[0], g.z["obs"][0], 1e-9, 1e-9)
    for t in range(g.T):
        r.cav.step(to_dev(np.ascontiguousarray(g.z["ext_actions"][t][None], dtype=np.float32)))
        out = r.download()
        common.compare_step(g, out, np.array([0]), np.array([t]), 1e-6, 1e-6)
    H = g.z["history"]
    for i in range(n):
        k = int(g.z["st_step_num"][g.T][i])
        common.assert_close("history", out.history[0, i, :k], H[i, :k], 1e-6, 1e-6)
