#!/usr/bin/env python
"""How much of k_policy_move's time is lock-step waste between UNLIKE environments in a warp, and how much is the tail of
long warps?  Same S scenarios, same number of environments each, three placements over the E environments:
  interleaved   scenario of env e = e % S            (a warp's environments all differ -- what a random bank gives)
  blocked       scenario of env e = e * S // E       (a warp's environments are alike -- what sorting by cost would give)
  blocked_lpt   blocks in order of decreasing single-scenario cost (longest first)
plus every scenario alone (E copies).  Kernel times over episode steps 5..24 (all agents live), CUDA events."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_collision_avoidance_b200.device import Cavoid, DeviceBank  # noqa: E402
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config  # noqa: E402
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank  # noqa: E402


def measure(cav, bank, sidx, steps=20, skip=5):
    cav.reset(bank, scenario_idx=sidx)
    for _ in range(skip):
        cav.step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for s in range(steps):
        ev[s][0].record()
        cav.k_policy()
        cav.k_dynamics()
        ev[s][1].record()
        cav.k_pairwise_obs()
        ev[s][2].record()
    torch.cuda.synchronize()
    pm = np.array([e[0].elapsed_time(e[1]) for e in ev]) * 1e3
    po = np.array([e[1].elapsed_time(e[2]) for e in ev]) * 1e3
    return float(pm.mean()), float(po.mean())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--agents", type=int, default=10)
    ap.add_argument("--scenarios", type=int, default=16)
    ap.add_argument("--seed", type=int, default=10)
    a = ap.parse_args()
    E, N, S = a.envs, a.agents, a.scenarios
    dev = "cuda:0"
    c = to_cav_config(make_config(N, evaluate=True), E, N, precision="f64", obs_f32=True)
    cav = Cavoid(c, dev)
    rows = random_scenario_bank(S, N, seed=a.seed)
    bank = DeviceBank(rows, kinds=np.full((S, N), 0, np.uint32), device=dev)  # kind 0 = RVO
    e = torch.arange(E, dtype=torch.int64)
    place = {"interleaved": (e % S), "blocked": (e * S // E)}
    out = {}
    single = []
    for s in range(S):
        single.append(measure(cav, bank, torch.full((E,), s, dtype=torch.int32, device=dev))[0])
    order = np.argsort(-np.array(single))
    rank = torch.from_numpy(order.copy())
    place["blocked_lpt"] = rank[(e * S // E)]
    place["blocked_spt"] = rank.flip(0)[(e * S // E)]
    for k, v in place.items():
        out[k] = measure(cav, bank, v.to(torch.int32).to(dev))
    print("E=%d N=%d S=%d  single-scenario policy us: mean %.1f min %.1f max %.1f" % (E, N, S, np.mean(single), np.min(single), np.max(single)))
    for k, (pm, po) in out.items():
        print("  %-12s policy+move %.1f us   pairwise+obs %.1f us" % (k, pm, po))
