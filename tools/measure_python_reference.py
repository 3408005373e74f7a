#!/usr/bin/env python
"""agent-steps/s of the UNMODIFIED Python reference (``CollisionAvoidanceEnv.step`` imported from /root/reference through
oracle/ref_harness.py; rvo2 = the reference's own C++ in oracle/_ref) in THIS container, for the bench workloads: N agents,
RVO or NonCooperative policies, random-start / random-goal cases from the reference's own generator, EvaluateConfig
(DT 0.1, MAX_TIME_RATIO 8), serial env loop with auto-reset on game_over (the DummyVecEnv pattern, vec_env.py:120-128), one
process on one core.  The reference cannot travel to the GPU box, so bench.py quotes these numbers
(``cpu_baseline.python_reference``) next to the C port it times there.

    python tools/measure_python_reference.py            # writes profiles/python_reference_baseline.json
"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_harness  # noqa: E402
from make_golden import write_cfg, mk_agents, EVAL_POST  # noqa: E402


def measure(N, policy, seconds=20.0, laser=False):
    pre = []
    sensors = ("other_agents_states",)
    if laser:
        pre = ["self.USE_STATIC_MAP = True",
               "self.STATES_IN_OBS = ['is_learning','num_other_agents','dist_to_goal','heading_ego_frame',"
               "'pref_speed','radius','other_agents_states','laserscan']"]
        sensors = ("other_agents_states", "laserscan")
    cfg_path = write_cfg(N, pre, list(EVAL_POST))
    ref_harness.load_reference(cfg_path, "Cfg")
    from gym_collision_avoidance.envs.collision_avoidance_env import CollisionAvoidanceEnv
    from gym_collision_avoidance.envs import test_cases as tc
    from gym_collision_avoidance.envs.policies.CADRL.scripts.multi import gen_rand_testcases as g
    np.random.seed(0)
    env = CollisionAvoidanceEnv()
    side = 7.0 if N >= 5 else 4.5

    def new_episode():
        rows = g.generate_rand_test_case_multi(N, side if N <= 10 else 2.5 * np.sqrt(N), [0.5, 2.0], [0.2, 0.8])
        env.set_agents(mk_agents(tc, rows, policy, sensors=sensors))
        env.reset()
    new_episode()
    for _ in range(5):
        env.step({})
    steps = 0
    t_step = 0.0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        t0 = time.perf_counter()
        obs, rew, over, _, info = env.step({})
        t_step += time.perf_counter() - t0
        steps += 1
        if over:
            new_episode()
    os.unlink(cfg_path)
    return N * steps / t_step, steps


if __name__ == "__main__":
    out = {"how": "unmodified Python reference, one process / one core, step() only (reset excluded), build container",
           "cpu": platform.processor() or open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"),
           "numpy": np.__version__}
    for N, pol, laser in ((10, "RVO", False), (10, "noncoop", False), (20, "RVO", True), (2, "RVO", False)):
        rate, steps = measure(N, pol, laser=laser)
        key = "%dagents_%s%s" % (N, pol, "_laser" if laser else "")
        out[key] = rate
        print(key, "%.1f agent-steps/s" % rate, "(%d steps)" % steps, flush=True)
    with open(os.path.join(ROOT, "profiles", "python_reference_baseline.json"), "w") as f:
        json.dump(out, f, indent=1)
