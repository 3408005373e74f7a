#!/usr/bin/env python
"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

The reference's own tests pin nothing numerically on this path, so known-answer vectors are
produced by running the reference itself -- ``CollisionAvoidanceEnv.reset/step`` imported from /root/reference
through ``oracle/ref_harness.py`` (inert stubs for absent third-party imports; ``rvo2`` = the reference's own C++
compiled into ``oracle/_ref``) -- and dumping, per step, the full agent state before/after, rewards, done flags,
game_over and the packed observation.  Output: ``tests/golden/<name>.npz`` (float64, compressed).

    python tools/make_golden.py            # regenerate every fixture
    python tools/make_golden.py rvo10_c0   # just one

Agents are always built with ``np.float64`` headings: under NumPy >= 2 (NEP 50) a Python-float heading would
turn the reference's heading state into float32 after the first step, which is an artefact of the
NumPy version, not of the reference's intended float64 semantics.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

POLICY_IDS = {"RVO": 0, "NonCooperativePolicy": 1, "Static": 2, "External": 3, "learning": 4}

CFG_TEMPLATE = """
import numpy as np
from gym_collision_avoidance.envs.config import Config as _Base
class Cfg(_Base):
    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = {N}
        self.MAX_NUM_AGENTS_TO_SIM = {N}
        {pre}
        _Base.__init__(self)
        {post}
"""


def write_cfg(N, pre, post):
    body = CFG_TEMPLATE.format(N=N, pre="\n        ".join(pre) or "pass", post="\n        ".join(post) or "pass")
    scratch = os.path.join(ROOT, "gpurun_out")  # git-ignored scratch inside the repository
    os.makedirs(scratch, exist_ok=True)
    f = tempfile.NamedTemporaryFile("w", suffix=".py", delete=False, dir=scratch)
    f.write(body)
    f.close()
    return f.name


def agent_state(agents, N):
    n = len(agents)
    st = dict(pv=np.zeros((N, 4)), goal=np.zeros((N, 4)), ht=np.zeros((N, 4)), ego=np.zeros((N, 4)),
              flags=np.zeros(N, np.uint32), step_num=np.zeros(N, np.int32), aux0=np.zeros((N, 4)),
              past_actions=np.zeros((N, 4)), past_vel=np.zeros((N, 4)))
    for i, a in enumerate(agents):
        st["pv"][i] = [a.pos_global_frame[0], a.pos_global_frame[1], a.vel_global_frame[0], a.vel_global_frame[1]]
        st["goal"][i] = [a.goal_global_frame[0], a.goal_global_frame[1], a.radius, a.pref_speed]
        st["ht"][i] = [a.heading_global_frame, a.time_remaining_to_reach_goal, a.t, a.turning_dir]
        st["ego"][i] = [a.dist_to_goal, a.heading_ego_frame, a.ref_prll[0], a.ref_prll[1]]
        f = 0
        f |= 0x01 if a.is_at_goal else 0
        f |= 0x02 if a.was_at_goal_already else 0
        f |= 0x04 if a.in_collision else 0
        f |= 0x08 if a.was_in_collision_already else 0
        f |= 0x10 if a.ran_out_of_time else 0
        f |= 0x20 if a.is_done else 0
        st["flags"][i] = f
        st["step_num"][i] = a.step_num
        st["aux0"][i] = [a.speed_global_frame, a.delta_heading_global_frame, a.vel_ego_frame[0], a.vel_ego_frame[1]]
        st["past_actions"][i] = a.past_actions.reshape(-1)
        st["past_vel"][i] = a.past_global_velocities.reshape(-1)
    return st


def pack_obs(obs, Config, N):
    rows = []
    for i in range(N):
        parts = [np.asarray(obs[i][s], dtype=np.float64).reshape(-1) for s in Config.STATES_IN_OBS]
        rows.append(np.concatenate(parts))
    return np.stack(rows)


def run_scenario(name, N, agents_fn, pre=(), post=(), max_steps=400, ext_action_fn=None, seed=0, stop_after=None,
                 static_map=None, subdir=""):
    cfg_path = write_cfg(N, list(pre), list(post))
    ref_harness.load_reference(cfg_path, "Cfg")
    from gym_collision_avoidance.envs import Config
    from gym_collision_avoidance.envs.collision_avoidance_env import CollisionAvoidanceEnv
    from gym_collision_avoidance.envs import test_cases as tc
    np.random.seed(seed)
    env = CollisionAvoidanceEnv()
    agents = agents_fn(tc, Config)
    n = len(agents)
    for a in agents:
        assert isinstance(a.heading_global_frame, np.float64), type(a.heading_global_frame)
    env.set_agents(agents)
    obs, _ = env.reset()
    if static_map is not None:
        # the reference loads static maps from image files through imageio/scipy.misc (absent here, Map.py:17-21);
        # writing the boolean layer into the Map object reset() just created is equivalent for the step path
        assert env.map is not None and env.map.static_map.shape == static_map.shape
        env.map.static_map[...] = static_map
    dyn_ids = {"UnicycleDynamics": 0, "UnicycleDynamicsMaxTurnRate": 1, "ExternalDynamics": 2}
    kinds = np.array([POLICY_IDS[a.policy.str] | (dyn_ids[type(a.dynamics_model).__name__] << 8) for a in agents]
                     + [0] * (N - n), np.uint32)
    states = [agent_state(env.agents, N)]
    obs_seq = [pack_obs(obs, Config, N)]
    rewards, dones, gos, exts, acts = [], [], [], [], []
    rng = np.random.RandomState(seed + 1)
    t = 0
    while True:
        ext = np.zeros((N, 2), np.float32)
        actions = {}
        for i, a in enumerate(env.agents):
            if a.policy.is_external:
                e = ext_action_fn(rng, i, t) if ext_action_fn else np.zeros(2, np.float32)
                ext[i] = e
                actions[i] = np.asarray(ext[i], dtype=np.float64)  # float32-representable float64 values
        was_live = [not a.is_done for a in env.agents]
        obs, rew, game_over, _, info = env.step(actions)
        st = agent_state(env.agents, N)
        act = np.zeros((N, 2), np.float32)
        for i, a in enumerate(env.agents):
            act[i] = a.past_actions[0, :]  # last live action (float32-representable)
        states.append(st)
        obs_seq.append(pack_obs(obs, Config, N))
        r = np.zeros(N)
        if np.ndim(rew):
            r[:n] = np.asarray(rew, dtype=np.float64).reshape(-1)[:n]
        else:  # Config.TRAIN_SINGLE_AGENT: step() returns rewards[0] only (env.py:454-455)
            r[0] = rew
        rewards.append(r)
        d = np.zeros(N, np.uint8)
        for i, a in enumerate(env.agents):
            d[i] = info["which_agents_done"][a.id]
        dones.append(d)
        gos.append(bool(game_over))
        exts.append(ext)
        acts.append(act)
        t += 1
        if game_over or t >= max_steps or (stop_after is not None and t >= stop_after):
            break
    T = t
    hist_len = max(int(a.step_num) for a in env.agents)
    hist = np.zeros((N, max(hist_len, 1), 14))
    for i, a in enumerate(env.agents):
        k = a.step_num
        hist[i, :k, :11] = a.global_state_history[:k]
        hist[i, :k, 11:] = a.ego_state_history[:k]
    meta = dict(
        name=name, num_agents=N, num_active=n, steps=T, obs_len=int(obs_seq[0].shape[1]),
        states_in_obs=list(Config.STATES_IN_OBS),
        config=dict(
            DT=Config.DT, NEAR_GOAL_THRESHOLD=Config.NEAR_GOAL_THRESHOLD, MAX_TIME_RATIO=Config.MAX_TIME_RATIO,
            EVALUATE_MODE=bool(Config.EVALUATE_MODE), TRAIN_SINGLE_AGENT=bool(Config.TRAIN_SINGLE_AGENT),
            SENSING_HORIZON=float(Config.SENSING_HORIZON) if np.isfinite(Config.SENSING_HORIZON) else "inf",
            MAX_NUM_OTHER_AGENTS_OBSERVED=int(Config.MAX_NUM_OTHER_AGENTS_OBSERVED),
            MAX_NUM_AGENTS_IN_ENVIRONMENT=int(Config.MAX_NUM_AGENTS_IN_ENVIRONMENT),
            AGENT_SORTING_METHOD=Config.AGENT_SORTING_METHOD, USE_STATIC_MAP=bool(Config.USE_STATIC_MAP),
            LASERSCAN_LENGTH=int(Config.LASERSCAN_LENGTH), LASERSCAN_NUM_PAST=int(Config.LASERSCAN_NUM_PAST),
            RVO_TIME_HORIZON=Config.RVO_TIME_HORIZON, RVO_COLLAB_COEFF=Config.RVO_COLLAB_COEFF,
            RVO_ANTI_COLLAB_T=Config.RVO_ANTI_COLLAB_T,
            REWARD_AT_GOAL=Config.REWARD_AT_GOAL, REWARD_COLLISION_WITH_AGENT=Config.REWARD_COLLISION_WITH_AGENT,
            REWARD_COLLISION_WITH_WALL=Config.REWARD_COLLISION_WITH_WALL, REWARD_TIME_STEP=Config.REWARD_TIME_STEP,
            REWARD_WIGGLY_BEHAVIOR=Config.REWARD_WIGGLY_BEHAVIOR,
            WIGGLY_BEHAVIOR_THRESHOLD=(float(Config.WIGGLY_BEHAVIOR_THRESHOLD)
                                       if np.isfinite(Config.WIGGLY_BEHAVIOR_THRESHOLD) else "inf"),
            GETTING_CLOSE_RANGE=Config.GETTING_CLOSE_RANGE, STORE_HISTORY=bool(Config.STORE_HISTORY),
        ),
    )
    extra = {}
    if static_map is not None:
        extra["static_map"] = np.packbits(static_map.astype(np.uint8), axis=None)
        meta["static_map_shape"] = list(static_map.shape)
        meta_json = json.dumps(meta)
    out = dict(meta=np.array(json.dumps(meta)), kinds=kinds, **extra,
               rewards=np.stack(rewards), dones=np.stack(dones), game_over=np.array(gos),
               ext_actions=np.stack(exts), actions=np.stack(acts), obs=np.stack(obs_seq), history=hist)
    for k in states[0]:
        out["st_" + k] = np.stack([s[k] for s in states])
    out_dir = os.path.join(OUT, subdir)
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    os.unlink(cfg_path)
    print("%-22s N=%d n=%d steps=%d game_over=%s  %.0f KB" % (
        name, N, n, T, gos[-1], os.path.getsize(os.path.join(out_dir, name + ".npz")) / 1024.0))


EVAL_POST = ["self.EVALUATE_MODE = True", "self.TRAIN_MODE = False", "self.DT = 0.1", "self.MAX_TIME_RATIO = 8."]


def mk_agents(tc, rows, policies, sensors=("other_agents_states",), headings=None, dynamics=None):
    from gym_collision_avoidance.envs.agent import Agent
    agents = []
    for i, row in enumerate(rows):
        px, py, gx, gy, pref_speed, radius = [np.float64(v) for v in row[:6]]
        if headings is None:
            heading = np.arctan2(gy - py, gx - px)
        else:
            heading = np.float64(headings[i])
        pol = policies if isinstance(policies, str) else policies[i]
        agents.append(Agent(px, py, gx, gy, radius, pref_speed, heading, tc.policy_dict[pol],
                            dynamics or tc.dynamics_dict["unicycle"], [tc.sensor_dict[s] for s in sensors], i))
    return agents


def fixture_rows(tc, n_agents, idx):
    return tc.preset_testCases(n_agents, full_test_suite=True)[idx]


SCENARIOS = {}


def scenario(fn):
    SCENARIOS[fn.__name__] = fn
    return fn


@scenario
def swap2_rvo():  # BASELINE config[0]: 2-agent swap, RVOPolicy (reference get_testcase_two_agents geometry)
    rows = [[-3, -3, 3, 3, 1.0, 0.5], [3, 3, -3, -3, 1.0, 0.5]]
    run_scenario("swap2_rvo", 2, lambda tc, C: mk_agents(tc, rows, "RVO", headings=[0.0, np.pi]), post=EVAL_POST,
                 stop_after=80)


for _i in range(3):
    def _mk(i):
        def f():
            run_scenario("rvo10_c%d" % i, 10, lambda tc, C: mk_agents(tc, fixture_rows(tc, 10, i), "RVO"),
                         post=EVAL_POST)
        f.__name__ = "rvo10_c%d" % i
        return f
    scenario(_mk(_i))

for _i in range(3):
    def _mk2(i):
        def f():
            run_scenario("noncoop10_c%d" % i, 10,
                         lambda tc, C: mk_agents(tc, fixture_rows(tc, 10, i), "noncoop"), post=EVAL_POST)
        f.__name__ = "noncoop10_c%d" % i
        return f
    scenario(_mk2(_i))


@scenario
def rvo4_c7():
    run_scenario("rvo4_c7", 4, lambda tc, C: mk_agents(tc, fixture_rows(tc, 4, 7), "RVO"), post=EVAL_POST)


@scenario
def rvo3_in8slots():  # fewer agents than slots: obs rows for absent slots stay zero, K = 7
    run_scenario("rvo3_in8slots", 8, lambda tc, C: mk_agents(tc, fixture_rows(tc, 3, 5), "RVO"), post=EVAL_POST)


@scenario
def mixed5_train():  # training-mode config: learning + RVO + noncoop + static + external, random headings
    def agents(tc, C):
        rows = fixture_rows(tc, 5, 11)
        hd = np.random.uniform(-np.pi, np.pi, size=5)
        return mk_agents(tc, rows, ["learning", "RVO", "noncoop", "static", "external"], headings=hd)

    def ext(rng, i, t):
        if i == 0:
            return np.array([rng.uniform(0.3, 1.0), rng.uniform(0.2, 0.8)], np.float32)
        return np.array([rng.uniform(0.2, 1.2), rng.uniform(-0.5, 0.5)], np.float32)

    run_scenario("mixed5_train", 5, agents, ext_action_fn=ext, max_steps=120)


@scenario
def dense6_clip3_last():  # K=3 < N-1 with closest_last ordering, dense start (collisions), wiggle reward on
    def agents(tc, C):
        rng = np.random.RandomState(5)
        rows = []
        for i in range(6):
            th = 2 * np.pi * i / 6 + rng.uniform(-0.2, 0.2)
            r = rng.uniform(1.2, 2.0)
            rows.append([r * np.cos(th), r * np.sin(th), -r * np.cos(th), -r * np.sin(th), rng.uniform(0.6, 1.4),
                         rng.uniform(0.3, 0.6)])
        return mk_agents(tc, rows, ["noncoop", "RVO", "noncoop", "RVO", "noncoop", "RVO"])

    run_scenario("dense6_clip3_last", 6, agents, pre=["self.MAX_NUM_OTHER_AGENTS_OBSERVED = 3"],
                 post=EVAL_POST + ["self.AGENT_SORTING_METHOD = 'closest_last'",
                                   "self.WIGGLY_BEHAVIOR_THRESHOLD = 0.3", "self.REWARD_WIGGLY_BEHAVIOR = -0.05",
                                   "self.REWARD_TIME_STEP = -0.01"])


@scenario
def tti8_clip4():  # AGENT_SORTING_METHOD = 'time_to_impact' (util.compute_time_to_impact), K = 4 of 7 neighbours
    run_scenario("tti8_clip4", 8, lambda tc, C: mk_agents(tc, fixture_rows(tc, 8, 6), "RVO"),
                 pre=["self.MAX_NUM_OTHER_AGENTS_OBSERVED = 4"],
                 post=EVAL_POST + ["self.AGENT_SORTING_METHOD = 'time_to_impact'"], stop_after=120)


@scenario
def tti6_mixed():  # time_to_impact with non-cooperative agents (collisions: time_to_impact 0) and full K
    run_scenario("tti6_mixed", 6, lambda tc, C: mk_agents(tc, fixture_rows(tc, 6, 2),
                                                           ["noncoop", "RVO", "noncoop", "RVO", "RVO", "noncoop"]),
                 post=EVAL_POST + ["self.AGENT_SORTING_METHOD = 'time_to_impact'"], stop_after=100)


@scenario
def maxturn6_ext():  # UnicycleDynamicsMaxTurnRate (clip(dtheta/dt, +-3) dt) under RVO / external commands, and one
    # ExternalDynamics agent (state never moves; the others still avoid it)
    def agents(tc, C):
        from gym_collision_avoidance.envs.dynamics.UnicycleDynamicsMaxTurnRate import UnicycleDynamicsMaxTurnRate
        from gym_collision_avoidance.envs.dynamics.ExternalDynamics import ExternalDynamics
        from gym_collision_avoidance.envs.dynamics.UnicycleDynamics import UnicycleDynamics
        rows = fixture_rows(tc, 6, 4)
        pol = ["RVO", "RVO", "external", "noncoop", "RVO", "external"]
        dyn = [UnicycleDynamicsMaxTurnRate, UnicycleDynamics, UnicycleDynamicsMaxTurnRate, UnicycleDynamicsMaxTurnRate,
               UnicycleDynamicsMaxTurnRate, ExternalDynamics]
        out = []
        for i in range(6):
            out += mk_agents(tc, [rows[i]], pol[i], dynamics=dyn[i])
            out[-1].id = i
        return out

    def ext(rng, i, t):
        return np.array([rng.uniform(0.2, 1.2), rng.uniform(-0.6, 0.6)], np.float32)

    run_scenario("maxturn6_ext", 6, agents, post=EVAL_POST, ext_action_fn=ext, stop_after=100)


# ---- tests/golden/cpu_only/: pin the ORACLE on the rest of BASELINE's N range (the GPU kernels are compared with the
# oracle at these sizes on the same states, tests/test_gpu_scale.py); not part of the GPU parity parametrisation ----
def _rand_rows(n, seed, side):
    np.random.seed(seed)
    from gym_collision_avoidance.envs.policies.CADRL.scripts.multi import gen_rand_testcases as g
    return g.generate_rand_test_case_multi(n, side, [0.5, 2.0], [0.2, 0.8])


@scenario
def rvo32_rand():  # N = 32 (sub-warp-group kernels' range): the reference's kd-tree with several leaf splits
    run_scenario("rvo32_rand", 32, lambda tc, C: mk_agents(tc, _rand_rows(32, 321, 7.0), "RVO"), post=EVAL_POST,
                 stop_after=30, subdir="cpu_only")


@scenario
def rvo64_rand():  # N = 64 = CAV_MAX_AGENTS, K = 63
    run_scenario("rvo64_rand", 64, lambda tc, C: mk_agents(tc, _rand_rows(64, 641, 7.0), "RVO"), post=EVAL_POST,
                 stop_after=12, subdir="cpu_only")


@scenario
def noncoop16_tti():  # 16 agents, time_to_impact ordering clipped to K = 6, collisions
    run_scenario("noncoop16_tti", 16, lambda tc, C: mk_agents(tc, _rand_rows(16, 161, 5.0), "noncoop"),
                 pre=["self.MAX_NUM_OTHER_AGENTS_OBSERVED = 6"],
                 post=EVAL_POST + ["self.AGENT_SORTING_METHOD = 'time_to_impact'"], stop_after=60, subdir="cpu_only")


@scenario
def horizon8():  # finite SENSING_HORIZON: obs skips far agents, RVO neighborDist finite
    run_scenario("horizon8", 8, lambda tc, C: mk_agents(tc, fixture_rows(tc, 8, 3), "RVO"),
                 post=EVAL_POST + ["self.SENSING_HORIZON = 3.0"], stop_after=90)


@scenario
def rvo20_rand():  # N > 10: the reference's kd-tree splits (KdTree.cpp:76); random geometry -> no distance ties
    def agents(tc, C):
        np.random.seed(123)
        from gym_collision_avoidance.envs.policies.CADRL.scripts.multi import gen_rand_testcases as g
        rows = g.generate_rand_test_case_multi(20, 11.0, [0.5, 2.0], [0.2, 0.8])
        return mk_agents(tc, rows, "RVO")

    run_scenario("rvo20_rand", 20, agents, post=EVAL_POST, stop_after=60)


@scenario
def laser6():  # laserscan (512 beams x 3 past) + neighbour rows, RVO, agents near the map centre
    def agents(tc, C):
        rows = fixture_rows(tc, 6, 2)
        return mk_agents(tc, rows, "RVO", sensors=("other_agents_states", "laserscan"))

    run_scenario("laser6", 6, agents,
                 pre=["self.USE_STATIC_MAP = True",
                      "self.STATES_IN_OBS = ['is_learning','num_other_agents','dist_to_goal','heading_ego_frame',"
                      "'pref_speed','radius','other_agents_states','laserscan']"],
                 post=EVAL_POST, stop_after=12)


@scenario
def laser20_circle():  # BASELINE config[3] shape: 20 agents RVO + 512-beam laserscan, the reference's own symmetric
    # gen_circle_test_case (exact float32 distance ties -> kd-tree leaf order decides the ORCA neighbour order)
    def agents(tc, C):
        return mk_agents(tc, tc.gen_circle_test_case(20, 7.0), "RVO", sensors=("other_agents_states", "laserscan"))

    run_scenario("laser20_circle", 20, agents,
                 pre=["self.USE_STATIC_MAP = True",
                      "self.STATES_IN_OBS = ['is_learning','num_other_agents','dist_to_goal','heading_ego_frame',"
                      "'pref_speed','radius','other_agents_states','laserscan']"],
                 post=EVAL_POST, stop_after=8)


# ---- the reference's own symmetric presets, UN-perturbed, whole episodes: exact float32 distance ties at every step,
# so the ORCA neighbour order is decided by the kd-tree's persistent agent permutation (KdTree.cpp:47-111,266-298) ----
def _circle(name, n, radius, **kw):
    def f():
        run_scenario(name, n, lambda tc, C: mk_agents(tc, tc.gen_circle_test_case(n, radius), "RVO"), post=EVAL_POST,
                     **kw)
    f.__name__ = name
    return scenario(f)


_circle("circle20_r10", 20, 10.0)   # preset_testCases(20) (test_cases.py:884-889)
_circle("circle16_r8", 16, 8.0)
_circle("circle12_r6", 12, 6.0)
_circle("circle10_r5", 10, 5.0, stop_after=150)     # single leaf (n <= MAX_LEAF_SIZE): index order


# ---- anti-collaborative RVO (RVOPolicy.py:77-88, R/Agent.cpp:368-415) with RVO_COLLAB_COEFF = -1: the policy re-draw is
# then deterministic (np.random.choice([True, False], p=[0, 1]) is always False), so the reference pins the arithmetic of
# the "steer into the closest neighbour" branch bit for bit; other coefficients are random in the reference ----
@scenario
def anticollab6():
    run_scenario("anticollab6", 6, lambda tc, C: mk_agents(tc, fixture_rows(tc, 6, 1), "RVO"),
                 post=EVAL_POST + ["self.RVO_COLLAB_COEFF = -1.0"], stop_after=120)


@scenario
def anticollab12_mixed():  # N > 10 (kd-tree order), anti-collaborative RVO agents among non-cooperative ones
    run_scenario("anticollab12_mixed", 12,
                 lambda tc, C: mk_agents(tc, _rand_rows(12, 77, 6.0), ["RVO", "noncoop"] * 6),
                 post=EVAL_POST + ["self.RVO_COLLAB_COEFF = -1.0"], stop_after=80)


@scenario
def single_agent_train():  # Config.TRAIN_SINGLE_AGENT: game_over = done[0] (env.py:541-542), scalar reward rewards[0]
    # (env.py:454-455); agent 0 learning (external actions), the others RVO / noncoop
    def agents(tc, C):
        rows = np.array(fixture_rows(tc, 4, 3), dtype=np.float64)
        rows[0, 2:4] = rows[0, 0:2] + [1.6, 0.9]  # agent 0 arrives while the others are still under way
        hd = np.random.uniform(-np.pi, np.pi, size=4)
        hd[0] = np.arctan2(0.9, 1.6)
        return mk_agents(tc, rows, ["learning", "RVO", "noncoop", "RVO"], headings=hd)

    def ext(rng, i, t):
        return np.array([rng.uniform(0.8, 1.0), rng.uniform(0.45, 0.55)], np.float32)

    run_scenario("single_agent_train", 4, agents, post=["self.TRAIN_SINGLE_AGENT = True"], ext_action_fn=ext,
                 max_steps=200)


@scenario
def walls6():  # static-map wall collisions (env.py:494-506): noncoop agents drive through wall blocks
    sm = np.zeros((160, 160), bool)
    sm[70:74, 60:100] = True     # horizontal wall just above the x axis (rows = y, Map.py:28)
    sm[90:110, 120:123] = True   # vertical wall at x ~ +4
    sm[20:30, 20:30] = True

    def agents(tc, C):
        rows = [[-4.0, 2.0, 4.0, -2.0, 1.0, 0.4], [3.0, -3.5, 5.5, 2.0, 1.2, 0.3], [-5.0, -4.0, -5.5, 5.0, 0.8, 0.5],
                [0.0, 3.0, 0.0, -3.0, 1.0, 0.35], [6.0, 6.0, -6.0, 5.5, 1.5, 0.6], [-2.0, -6.0, 2.0, -6.5, 0.9, 0.25]]
        return mk_agents(tc, rows, "noncoop")

    run_scenario("walls6", 6, agents, pre=["self.USE_STATIC_MAP = True"], post=EVAL_POST, static_map=sm)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SCENARIOS)
    for nme in names:
        SCENARIOS[nme]()
