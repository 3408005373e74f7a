#!/usr/bin/env python
"""Ad-hoc kernel timing (CUDA events) used while developing; bench.py is the contract benchmark."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_collision_avoidance_b200.device import Cavoid, DeviceBank  # noqa: E402
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config  # noqa: E402
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank  # noqa: E402


def run(E, N, prec, policy, steps, laser=False):
    states = None
    if laser:
        states = ['is_learning', 'num_other_agents', 'dist_to_goal', 'heading_ego_frame', 'pref_speed', 'radius',
                  'other_agents_states', 'laserscan']
    cfg = make_config(N, evaluate=True, use_static_map=laser, states_in_obs=states)
    c = to_cav_config(cfg, E, N, precision=prec)
    rows = random_scenario_bank(512, N, seed=1)
    kinds = np.full((512, N), policy, np.uint32)
    cav = Cavoid(c, "cuda:0")
    bank = DeviceBank(rows, kinds=kinds, device="cuda:0")
    cav.reset(bank)
    torch.cuda.synchronize()
    for _ in range(5):
        cav.step()
        cav.reset(bank, env_mask=cav.state.game_over, accumulate_stats=True)
    torch.cuda.synchronize()
    names = ["policy_move", "pairwise_obs"] + (["laser"] if laser else [])
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(steps)]
    t0 = time.time()
    for s in range(steps):
        ev[s][0].record()
        cav.k_policy()
        cav.k_dynamics()
        ev[s][1].record()
        cav.k_pairwise_obs()
        ev[s][2].record()
        if laser:
            cav.k_laserscan()
            ev[s][3].record()
        cav.reset(bank, env_mask=cav.state.game_over, accumulate_stats=True)
    torch.cuda.synchronize()
    wall = time.time() - t0
    per = [np.median([ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(steps)]) for i in range(len(names))]
    # fused step timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        cav.step()
        cav.reset(bank, env_mask=cav.state.game_over, accumulate_stats=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("E=%d N=%d %s policy=%d laser=%s: " % (E, N, prec, policy, laser) +
          " ".join("%s=%.1fus" % (n, p * 1e3) for n, p in zip(names, per)) +
          " | step+reset=%.1fus -> %.3e agent-steps/s | wall/step %.1fus | stats %s" % (
              ms * 1e3, E * N / (ms * 1e-3), wall / steps * 1e6, cav.stats()))
    cav.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    for prec in ("f32", "f64"):
        run(16384, 10, prec, 0, a.steps)
        run(32768, 10, prec, 1, a.steps)
    run(4096, 10, "f32", 0, a.steps)
    run(1024, 20, "f64", 0, 10, laser=True)
    run(1024, 20, "f32", 0, 10, laser=True)
    run(2048, 64, "f32", 0, 10)
    run(65536, 2, "f32", 0, 10)
