#!/usr/bin/env python
"""Agent-count x env-count sweep (BASELINE.json configs[4]) on one GPU: step time and agent-steps/s per (N, E),
optionally for both kernel mappings.  Prints one JSON line per point (CUDA events, graph replay, auto-reset)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_collision_avoidance_b200.envs.config import make_config  # noqa: E402
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank  # noqa: E402
from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv  # noqa: E402


def point(E, N, prec, policy, steps, mapping):
    if mapping == "group":
        os.environ["CAV_TPA_MAX_N"] = "0"
    elif mapping == "tpa":
        os.environ["CAV_TPA_MAX_N"] = "25"
    else:
        os.environ.pop("CAV_TPA_MAX_N", None)
    env = VecCollisionAvoidanceEnv(E, make_config(N, evaluate=True), N, precision=prec,
                                   scenario_bank=random_scenario_bank(256, N, seed=N), policies=policy)
    env.reset()
    for _ in range(10):
        env.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        env.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    env.close()
    return {"E": E, "N": N, "precision": prec, "policy": policy, "mapping": mapping, "ms_per_step": ms,
            "agent_steps_per_s": E * N / (ms * 1e-3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", default="2,4,8,10,16,20,24,32,64")
    ap.add_argument("--envs", default="16384")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--policy", default="RVO")
    ap.add_argument("--mappings", default="auto")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    for N in map(int, a.agents.split(",")):
        for E in map(int, a.envs.split(",")):
            for m in a.mappings.split(","):
                if m == "tpa" and N > 25:
                    continue
                print(json.dumps(point(E, N, a.precision, a.policy, a.steps, m)), flush=True)
