#!/usr/bin/env python
"""Device-timed step with GA3C-CADRL agents (the network runs inside cav_step): ms per step, agent-steps/s, and the
k_ga3c kernel alone with its float32 FLOP rate.   python tools/bench_ga3c.py [--envs 16384] [--agents 10] [--mix]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_collision_avoidance_b200.envs.config import make_config  # noqa: E402
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank  # noqa: E402
from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--agents", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mix", action="store_true", help="every other agent RVO instead of GA3C-CADRL")
    a = ap.parse_args()
    E, N = a.envs, a.agents
    blob = np.load(os.path.join(ROOT, "tests", "golden", "ga3c", "ga3c_iros18.npz"))["blob"]
    pol = ["GA3C_CADRL" if (not a.mix or i % 2 == 0) else "RVO" for i in range(N)]
    env = VecCollisionAvoidanceEnv(E, make_config(N, evaluate=True), N, policies=pol, ga3c=blob, obs_f32=True,
                                   scenario_bank=random_scenario_bank(256, N, seed=N))
    env.reset()
    for _ in range(5):
        env.step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for s in range(a.steps):
        ev[s][0].record()
        env.step()
        ev[s][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([x.elapsed_time(y) for x, y in ev]))
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s in range(5):
        kev[s][0].record()
        env.cav.k_ga3c()
        kev[s][1].record()
    torch.cuda.synchronize()
    kms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    nga = E * sum(p == "GA3C_CADRL" for p in pol)
    others = min(N - 1, 19)
    flop = nga * 2.0 * (others * 71 * 256 + 68 * 256 + 2 * 256 * 256 + 256 * 11)
    print(json.dumps({"E": E, "N": N, "policies": "mixed GA3C/RVO" if a.mix else "GA3C_CADRL", "ms_per_step": ms,
                      "agent_steps_per_s": E * N / (ms * 1e-3), "k_ga3c_ms": kms, "ga3c_agents": nga,
                      "k_ga3c_TFLOPs_f32": flop / (kms * 1e-3) / 1e12}))
