#!/usr/bin/env python
"""Agent-count x env-count sweep (BASELINE.json configs[4]: N in {2..64} x E in {1k..131k}) -- the roofline curve.

One JSON line per (N, E) point: device-timed step (CUDA events, CUDA-graph replay, auto-reset, float64 state + float32
observation unless --precision f32), the fused pairwise+obs kernel timed alone with its achieved fraction of the measured
HBM peak (algorithmic bytes with the actual element sizes, as bench.py), the policy+move kernel, and agent-steps/s.
Under torchrun every rank runs the same point on its own GPU (environments are independent: weak scaling, no collective on
the step path); rank 0 prints the job total (sum of the per-rank rates computed from the max-over-ranks step time).

    python tools/sweep.py                                   # full grid on one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/sweep.py --agents 2,8,32 --envs 4096,32768
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (measured_peak, pairwise_obs_bytes)
from gym_collision_avoidance_b200.envs.config import make_config  # noqa: E402
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank  # noqa: E402
from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv  # noqa: E402


def point(E, N, prec, policy, steps, dev, flush):
    side = None
    env = VecCollisionAvoidanceEnv(E, make_config(N, evaluate=True), N, device=dev, precision=prec,
                                   scenario_bank=random_scenario_bank(256, N, seed=N), policies=policy,
                                   obs_f32=(prec == "f64"))
    env.reset()
    for _ in range(8):
        env.step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s in range(steps):
        flush()
        ev[s][0].record()
        env.step()
        ev[s][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    cav = env.cav
    reps = max(3, min(steps, 10))
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
    for s in range(reps):
        flush()
        kev[s][0].record()
        cav.k_policy()
        cav.k_dynamics()
        kev[s][1].record()
        flush()
        kev[s][2].record()
        cav.k_pairwise_obs()
        kev[s][3].record()
        cav.reset(env.bank, env_mask=env.state.game_over, accumulate_stats=True)
    torch.cuda.synchronize()
    pm = float(np.mean([k[0].elapsed_time(k[1]) for k in kev]))
    po = float(np.mean([k[2].elapsed_time(k[3]) for k in kev]))
    bpa = bench.pairwise_obs_bytes(env.c)
    peak, _ = bench.measured_peak()
    out = {"E": E, "N": N, "state_precision": prec, "obs_precision": "f32", "policy": policy, "ms_per_step": ms,
           "agent_steps_per_s": E * N / (ms * 1e-3), "policy_move_kernel_ms": pm, "pairwise_obs_kernel_ms": po,
           "pairwise_obs_bytes_per_agent_step": bpa,
           "pairwise_obs_roofline_frac": bpa * E * N / (po * 1e-3) / 1e9 / peak,
           "step_algorithmic_GBps": (bpa + 3 * 4 * (8 if prec == "f64" else 4) * 2) * E * N / (ms * 1e-3) / 1e9}
    env.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", default="2,4,8,16,32,64")
    ap.add_argument("--envs", default="1024,2048,4096,8192,16384,32768,65536,131072")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--policy", default="RVO")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--max-agents-total", type=int, default=4 << 20, help="skip points with E*N above this")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fr = torch.empty(64 << 20, dtype=torch.float32, device=dev)

    def flush():
        fl.zero_()
        fr.sum()
    for N in map(int, a.agents.split(",")):
        for E in map(int, a.envs.split(",")):
            if E * N > a.max_agents_total:
                continue
            r = point(E, N, a.precision, a.policy, a.steps, dev, flush)
            if world > 1:
                t = torch.tensor([r["ms_per_step"]], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r["ms_per_step"] = float(t.item())
                r["agent_steps_per_s"] = world * E * N / (r["ms_per_step"] * 1e-3)
            r["n_gpus"] = world
            if rank == 0:
                print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()
