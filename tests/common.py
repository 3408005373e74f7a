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


def golden_names(cpu_only=False):
    """Fixture names; cpu_only=True lists tests/golden/cpu_only/ (oracle-vs-reference pins at sizes whose GPU parity is
    covered against the oracle instead), as "cpu_only/<name>"."""
    if cpu_only:
        return sorted("cpu_only/" + os.path.splitext(os.path.basename(p))[0]
                      for p in glob.glob(os.path.join(GOLDEN_DIR, "cpu_only", "*.npz")))
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

    def cav_config(self, num_envs, precision="f64", track_aux=True, history_len=0, obs_f32=False):
        return to_cav_config(self.py_config(), num_envs, self.N, precision=precision, track_aux=track_aux,
                             history_len=history_len, store_history=history_len > 0, obs_f32=obs_f32)

    def static_map_bits(self):
        """uint32 bitmap of the fixture's static map (None when the fixture has none)."""
        if "static_map" not in self.z:
            return None
        from gym_collision_avoidance_b200.envs.Map import Map
        shape = tuple(self.meta["static_map_shape"])
        sm = np.unpackbits(self.z["static_map"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
        return Map(16, 16, 0.1, static_map=sm).bits()

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
        if st.static_map_bits is None and "static_map" in self.z:
            st.static_map_bits = self.static_map_bits()
        if "kd_order" in st.arrays:
            st.kd_order[e] = self.kd_order(t)

    def kd_order(self, t):
        """RVO2's persistent kd-tree agent permutation after t steps of this episode (hidden simulator state the golden
        cannot record: the reference library has no accessor for KdTree::agents_).  It is a pure function of the float32
        position history, replayed here with the oracle's cavo_kd_build (itself pinned bit for bit against the reference
        library on tie-heavy scenes, tests/test_oracle_rvo2.py).  The reference stops stepping an agent's simulator once
        that agent is done, so replay while any RVO agent is live -- afterwards nobody reads it."""
        if not hasattr(self, "_kd"):
            import ctypes
            import cav_oracle
            L = cav_oracle.lib()
            pos = np.ascontiguousarray(self.z["st_pv"][:, :self.n, 0:2].astype(np.float32))
            T = pos.shape[0]
            out = np.zeros((T + 1, self.N), np.uint8)
            out[:] = np.arange(self.N, dtype=np.uint8)
            perm = np.arange(self.n, dtype=np.uint8)
            tree = ctypes.create_string_buffer(1 << 16)
            rvo = (self.z["kinds"][:self.n] & 0xff) == abi.CAV_POLICY_RVO
            for k in range(T):
                live = (self.z["st_flags"][k][:self.n] & abi.CAV_FLAG_DONE) == 0
                if self.n > abi.CAV_KD_MAX_LEAF and (live & rvo).any():
                    L.cavo_kd_build(tree, pos[k].ctypes.data, perm.ctypes.data, self.n)
                out[k + 1, :self.n] = perm
            self._kd = out
        return self._kd[t]

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


class Tol(object):
    """Per-quantity tolerances.  `pos`: positions / distances / times; `ang`: angles (compared modulo 2*pi);
    `vel`: velocities, speeds and commanded actions; `laser_frac`: allowed fraction of differing laser beams."""

    def __init__(self, rtol, atol, vel=None, ang=None, laser_frac=0.0):
        self.pos = (rtol, atol)
        self.vel = vel if vel is not None else (rtol, atol)
        self.ang = ang if ang is not None else (rtol, atol)
        self.laser_frac = laser_frac


def _as_tol(rtol, atol):
    return rtol if isinstance(rtol, Tol) else Tol(rtol, atol)


def compare_step(g, st, envs, ts, rtol, atol=None, check_aux=True):
    """Compare env e of `st` (after one step from golden state ts[e]) against golden state ts[e]+1.
    Flags, done, game_over and step counters are compared exactly."""
    tol = _as_tol(rtol, atol)
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
    G = lambda col: st.arrays[col][envs][:, :n]  # noqa: E731
    W = lambda col: z["st_" + col][t1][:, :n]  # noqa: E731
    worst = max(worst, assert_close("pos", G("pv")[..., :2], W("pv")[..., :2], *tol.pos))
    worst = max(worst, assert_close("vel", G("pv")[..., 2:], W("pv")[..., 2:], *tol.vel))
    worst = max(worst, assert_close("goal", G("goal"), W("goal"), *tol.pos))
    worst = max(worst, assert_close("heading", G("ht")[..., 0], W("ht")[..., 0], *tol.ang, angle=True))
    worst = max(worst, assert_close("time", G("ht")[..., 1:3], W("ht")[..., 1:3], *tol.pos))
    worst = max(worst, assert_close("turning_dir", G("ht")[..., 3], W("ht")[..., 3], *tol.ang))
    worst = max(worst, assert_close("dist_to_goal", G("ego")[..., 0], W("ego")[..., 0], *tol.pos))
    worst = max(worst, assert_close("heading_ego", G("ego")[..., 1], W("ego")[..., 1], *tol.ang, angle=True))
    worst = max(worst, assert_close("ref_prll", G("ego")[..., 2:], W("ego")[..., 2:], *tol.ang))
    nr = 1 if g.meta["config"].get("TRAIN_SINGLE_AGENT") else n  # step() returns rewards[0] alone (env.py:454-455)
    worst = max(worst, assert_close("reward", st.reward[envs][:, :nr], z["rewards"][t0][:, :nr], *tol.pos))
    live = (z["st_flags"][t0][:, :n] & abi.CAV_FLAG_DONE) == 0
    ga = st.actions[envs][:, :n].astype(np.float64)
    wa = z["actions"][t0][:, :n].astype(np.float64)
    f32 = (2e-7, 1e-7)
    tv = (max(tol.vel[0], f32[0]), max(tol.vel[1], f32[1]))
    ta = (max(tol.ang[0], f32[0]), max(tol.ang[1], f32[1]))
    worst = max(worst, assert_close("actions.speed", np.where(live, ga[..., 0], 0), np.where(live, wa[..., 0], 0), *tv))
    worst = max(worst, assert_close("actions.dheading", np.where(live, ga[..., 1], 0), np.where(live, wa[..., 1], 0),
                                    *ta, angle=True))
    worst = max(worst, compare_obs(g, st.obs[envs], z["obs"][t1], tol))
    if check_aux and "aux0" in st.arrays:
        worst = max(worst, assert_close("aux.speed", G("aux0")[..., 0], W("aux0")[..., 0], *tv))
        worst = max(worst, assert_close("aux.dheading", G("aux0")[..., 1], W("aux0")[..., 1], *ta, angle=True))
        worst = max(worst, assert_close("aux.vel_ego", G("aux0")[..., 2:], W("aux0")[..., 2:], *tv))
        worst = max(worst, assert_close("past_actions.speed", G("past_actions")[..., 0::2], W("past_actions")[..., 0::2], *tv))
        worst = max(worst, assert_close("past_actions.dh", G("past_actions")[..., 1::2], W("past_actions")[..., 1::2], *ta))
        worst = max(worst, assert_close("past_vel", G("past_vel"), W("past_vel"), *tv))
    return worst


def compare_obs(g, got, want, rtol, atol=None):
    """Packed observation compare.  Row layout of other_agents_states: [p_prll, p_orth, v_prll, v_orth, r, r+r,
    dist] -> position / velocity tolerances per column; heading_ego_frame as an angle; laserscan beams must be
    equal (they are multiples of the range resolution) except for `laser_frac` of them."""
    tol = _as_tol(rtol, atol)
    cfg = g.py_config()
    from gym_collision_avoidance_b200.envs.config import obs_layout, state_size
    off, total = obs_layout(cfg)
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    worst = 0.0
    for s_, o in off.items():
        ln = state_size(cfg, s_)
        a, b = got[..., o:o + ln], want[..., o:o + ln]
        if s_ == "heading_ego_frame":
            worst = max(worst, assert_close("obs.heading_ego", a, b, *tol.ang, angle=True))
        elif s_ in ("other_agents_states", "other_agent_states"):
            a7, b7 = a.reshape(a.shape[:-1] + (-1, 7)), b.reshape(b.shape[:-1] + (-1, 7))
            worst = max(worst, assert_close("obs.rows.pos", a7[..., [0, 1, 4, 5, 6]], b7[..., [0, 1, 4, 5, 6]], *tol.pos))
            worst = max(worst, assert_close("obs.rows.vel", a7[..., [2, 3]], b7[..., [2, 3]], *tol.vel))
        elif s_ == "laserscan":
            bad = np.abs(a - b) > 1e-6
            frac = bad.mean() if bad.size else 0.0
            if frac > tol.laser_frac:
                idx = np.argwhere(bad)[0]
                raise AssertionError("obs.laserscan: %d/%d beams differ (%.4f > %.4f allowed); first at %s got %s "
                                     "want %s" % (bad.sum(), bad.size, frac, tol.laser_frac, idx.tolist(),
                                                  a[tuple(idx)], b[tuple(idx)]))
        else:
            worst = max(worst, assert_close("obs." + s_, a, b, *tol.pos))
    return worst


def compare_states(got, want, n, rtol, atol, cols=None):
    """Full HostState vs HostState compare (same config): flags / counters exact, floats within tolerance."""
    from gym_collision_avoidance_b200 import abi as _abi
    cols = cols or tuple(k for k in want.arrays if k not in ("stats", "episode", "kinds", "num_active"))
    worst = 0.0
    for k in cols:
        if k not in want.arrays:
            continue
        a, b = got.arrays[k], want.arrays[k]
        if a.ndim == 1:  # per-env columns (game_over, episode)
            assert np.array_equal(a, b), "%s mismatch" % k
            continue
        a, b = a[:, :n], b[:, :n]
        if k == "flags":
            if not np.array_equal(a & 0x1ff, b & 0x1ff):
                bad = np.argwhere((a & 0x1ff) != (b & 0x1ff))[:5]
                raise AssertionError("flags mismatch at %s: got %s want %s" % (bad.tolist(), a[tuple(bad[0])],
                                                                                b[tuple(bad[0])]))
        elif a.dtype.kind in "iu":
            assert np.array_equal(a, b), "%s mismatch" % k
        elif k == "actions":  # both sides keep the last live action for done agents
            worst = max(worst, assert_close(k, a, b, max(rtol, 2e-7), max(atol, 1e-7)))
        elif k == "ht":
            worst = max(worst, assert_close("ht.heading", a[..., 0], b[..., 0], rtol, atol, angle=True))
            worst = max(worst, assert_close("ht.rest", a[..., 1:], b[..., 1:], rtol, atol))
        elif k == "ego":
            worst = max(worst, assert_close("ego.heading", a[..., 1], b[..., 1], rtol, atol, angle=True))
            worst = max(worst, assert_close("ego.rest", a[..., [0, 2, 3]], b[..., [0, 2, 3]], rtol, atol))
        else:
            worst = max(worst, assert_close(k, a, b, rtol, atol))
    return worst
