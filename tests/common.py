"""Shared helpers for the parity tests: golden-fixture loading, config building, state injection, comparison."""
import glob
import json
import os

import numpy as np

from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config
from gym_collision_avoidance_b200.state import HostBank, HostState

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_COLS = ("pv", "goal", "ht", "ego", "flags", "step_num")
AUX_COLS = ("aux0", "past_actions", "past_vel")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden(object):
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(str(self.z["meta"]))
        self.N = self.meta["num_agents"]
        self.n = self.meta["num_active"]
        self.T = self.meta["steps"]

    def py_config(self):
        m = self.meta["config"]
        ov = dict(m)
        for k in ("SENSING_HORIZON", "WIGGLY_BEHAVIOR_THRESHOLD"):
            if ov[k] == "inf":
                ov[k] = np.inf
        evaluate = ov.pop("EVALUATE_MODE")
        use_map = ov.pop("USE_STATIC_MAP")
        ov.pop("MAX_NUM_AGENTS_IN_ENVIRONMENT")
        cfg = make_config(self.N, evaluate=False, use_static_map=use_map, states_in_obs=self.meta["states_in_obs"],
                          **ov)
        cfg.EVALUATE_MODE = evaluate
        return cfg

    def cav_config(self, num_envs, precision="f64", track_aux=True, history_len=0):
        return to_cav_config(self.py_config(), num_envs, self.N, precision=precision, track_aux=track_aux,
                             history_len=history_len, store_history=history_len > 0)

    def inject(self, st, e, t):
        """Write golden state at step t (after t steps) into env e of a HostState."""
        z = self.z
        for col in STATE_COLS + tuple(c for c in AUX_COLS if c in st.arrays):
            st.arrays[col][e] = z["st_" + col][t].astype(st.arrays[col].dtype)
        st.kinds[e] = z["kinds"]
        st.num_active[e] = self.n
        st.obs[e] = z["obs"][t].astype(st.obs.dtype)
        if t > 0:
            st.actions[e] = z["actions"][t - 1]
        else:
            st.actions[e] = 0
        st.done[e] = (z["st_flags"][t] & abi.CAV_FLAG_DONE) != 0
        if "laser_count" in st.arrays:
            st.laser_count[e, :self.n] = t + 1  # reset sensed once, then once per step

    def bank(self):
        rows = np.zeros((1, self.N, 8))
        rows[0, :, 0:2] = self.z["st_pv"][0][:, 0:2]
        rows[0, :, 2:4] = self.z["st_goal"][0][:, 0:2]
        rows[0, :, 4] = self.z["st_goal"][0][:, 3]
        rows[0, :, 5] = self.z["st_goal"][0][:, 2]
        rows[0, :, 6] = self.z["st_ht"][0][:, 0]
        rows[0, self.n:, 4] = 1.0  # keep absent slots finite
        return HostBank(rows, kinds=self.z["kinds"][None], num_active=np.array([self.n], np.int32))


def ang_diff(a, b):
    d = (a - b + np.pi) % (2 * np.pi) - np.pi
    return d


def assert_close(name, got, want, rtol, atol, angle=False):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    d = ang_diff(got, want) if angle else (got - want)
    d = np.where(both_inf, 0.0, d)
    err = np.abs(d)
    tol = atol + rtol * np.abs(np.where(both_inf, 0.0, want))
    bad = ~(err <= tol)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError("%s: %d/%d mismatches, max err %.3e; first at %s got %s want %s" % (
            name, bad.sum(), bad.size, np.nanmax(err), idx.tolist(), got[tuple(idx[0])], want[tuple(idx[0])]))
    return float(err.max()) if err.size else 0.0


def compare_step(g, st, envs, ts, rtol, atol, check_aux=True, obs_rtol=None, obs_atol=None):
    """Compare env e of `st` (after one step from golden state ts[e]) against golden state ts[e]+1."""
    z = g.z
    t1 = np.asarray(ts) + 1
    t0 = np.asarray(ts)
    n = g.n
    worst = 0.0
    want_flags = z["st_flags"][t1][:, :n] & 0x3f
    got_flags = st.flags[envs][:, :n] & 0x3f
    if not np.array_equal(got_flags, want_flags):
        bad = np.argwhere(got_flags != want_flags)[:5]
        raise AssertionError("flags mismatch at (step,agent) %s got %s want %s" % (
            [(int(t0[b[0]]), int(b[1])) for b in bad], got_flags[tuple(bad[0])], want_flags[tuple(bad[0])]))
    assert np.array_equal(st.done[envs][:, :n], z["dones"][t0][:, :n]), "done mismatch"
    assert np.array_equal(st.game_over[envs].astype(bool), z["game_over"][t0]), "game_over mismatch"
    assert np.array_equal(st.step_num[envs][:, :n], z["st_step_num"][t1][:, :n]), "step_num mismatch"
    for col in ("pv", "goal", "ht", "ego"):
        got, want = st.arrays[col][envs][:, :n], z["st_" + col][t1][:, :n]
        if col == "ht":
            worst = max(worst, assert_close("ht.heading", got[..., 0], want[..., 0], rtol, atol, angle=True))
            worst = max(worst, assert_close("ht.rest", got[..., 1:], want[..., 1:], rtol, atol))
        elif col == "ego":
            worst = max(worst, assert_close("ego.heading_ego", got[..., 1], want[..., 1], rtol, atol, angle=True))
            worst = max(worst, assert_close("ego.dist", got[..., 0], want[..., 0], rtol, atol))
            worst = max(worst, assert_close("ego.ref", got[..., 2:], want[..., 2:], rtol, atol))
        else:
            worst = max(worst, assert_close(col, got, want, rtol, atol))
    worst = max(worst, assert_close("reward", st.reward[envs][:, :n], z["rewards"][t0][:, :n], rtol, atol))
    live = (z["st_flags"][t0][:, :n] & abi.CAV_FLAG_DONE) == 0
    ga = st.actions[envs][:, :n].astype(np.float64)
    wa = z["actions"][t0][:, :n].astype(np.float64)
    worst = max(worst, assert_close("actions.speed", np.where(live, ga[..., 0], 0), np.where(live, wa[..., 0], 0),
                                    max(rtol, 2e-7), max(atol, 1e-7)))
    worst = max(worst, assert_close("actions.dheading", np.where(live, ga[..., 1], 0), np.where(live, wa[..., 1], 0),
                                    max(rtol, 2e-7), max(atol, 1e-7), angle=True))
    ortol = rtol if obs_rtol is None else obs_rtol
    oatol = atol if obs_atol is None else obs_atol
    worst = max(worst, compare_obs(g, st.obs[envs], z["obs"][t1], ortol, oatol))
    if check_aux and "aux0" in st.arrays:
        for col in AUX_COLS:
            worst = max(worst, assert_close(col, st.arrays[col][envs][:, :n], z["st_" + col][t1][:, :n], rtol, atol))
    return worst


def compare_obs(g, got, want, rtol, atol):
    """Packed observation compare; heading_ego_frame compared as an angle."""
    cfg = g.py_config()
    from gym_collision_avoidance_b200.envs.config import obs_layout
    off, total = obs_layout(cfg)
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    worst = 0.0
    mask = np.ones(total, bool)
    if "heading_ego_frame" in off:
        o = off["heading_ego_frame"]
        worst = max(worst, assert_close("obs.heading_ego", got[..., o], want[..., o], rtol, atol, angle=True))
        mask[o] = False
    worst = max(worst, assert_close("obs", got[..., mask], want[..., mask], rtol, atol))
    return worst
