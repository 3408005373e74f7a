#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the vectorized collision-avoidance step on B200 (contract benchmark).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--envs E] [--agents A] [--policy RVO|noncoop] [--precision f64|f32] [--laser] [--history T]

A "step" is one pass of the hot path over one batch: every agent of every environment selects an action (RVO/ORCA
LP by default), moves, the pairwise collision / reward / sorted-neighbour observation / done pass runs, finished
environments are reset from the device-resident scenario bank and observed (auto-reset).  Default workload:
16384 envs x 10 agents, RVOPolicy, random-start / random-goal scenarios -- BASELINE.json configs[1]'s shape
(10 agents, random goals, RVOPolicy) at the env count north_star's single-GPU target is quoted on (16384; pass
--envs 4096 for configs[1] verbatim).  --gpus N (under torchrun) shards environments: every rank steps its own
E environments (weak scaling), no data-path collective; one NCCL all-gather of the episode statistics at the end.

State is float64 (the reference's NumPy arithmetic); the observation / reward columns are float32 by default -- the
dtype the reference declares for its observation space (config.py:93-170), written directly by the kernels
(cfg.obs_f32; --obs-f64 keeps float64 columns).

Episode phase: the timed region starts W (+ --preroll) steps after a reset of every environment, i.e. with (nearly) all
agents still live -- the most expensive phase of an episode (every live RVO agent runs an ORCA solve); `live_agent_steps`
/ `live_fraction` report how many agent slots were live.  --preroll 150 measures the stationary mix instead.

Timing: W untimed warm-up steps; K timed steps, each bracketed by CUDA events on the launching stream, with a
256 MiB memset + a 256 MiB read pass between steps to flush the 126 MB L2 (outside the event brackets); value =
agent-steps / sum of event times, max over ranks; a barrier + synchronize brackets the whole timed loop.
`e2e` is the same workload through VecCollisionAvoidanceEnv.step_host: pinned-host actions in, pinned-host
observations / rewards / done flags out, every step (wall clock, includes the copies).
`roofline` times the dominant fused pairwise+obs kernel alone (CUDA events around its launches, L2 flushed).
`cpu_baseline` / `--impl reference`: the C restatement of the reference's step (oracle/, kind "port"; the Python
reference cannot travel to the GPU box; its own rate measured in the build container is quoted from
profiles/python_reference_baseline.json) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=16384, help="environments PER GPU")
    ap.add_argument("--agents", type=int, default=10)
    ap.add_argument("--policy", default="RVO", choices=["RVO", "noncoop"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--laser", action="store_true", help="add LaserScanSensor 512 beams x 3 (BASELINE configs[3])")
    ap.add_argument("--history", type=int, default=256, help="trajectory rows kept per agent (0 = off)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="cpu_baseline sample steps (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--obs-f64", action="store_true", help="float64 observation / reward columns (default float32)")
    ap.add_argument("--preroll", type=int, default=0, help="untimed steps before the warm-up (episode phase)")
    ap.add_argument("--no-lookahead", action="store_true", help="e2e: wait for every step's download before the next launch")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    return ap.parse_args()


def workload_config(a):
    from gym_collision_avoidance_b200.envs.config import make_config
    states = None
    if a.laser:
        states = ["is_learning", "num_other_agents", "dist_to_goal", "heading_ego_frame", "pref_speed", "radius",
                  "other_agents_states", "laserscan"]
    return make_config(a.agents, evaluate=True, use_static_map=a.laser, states_in_obs=states)


def workload_name(a):
    return "%d envs/GPU x %d agents, %sPolicy, UnicycleDynamics, random start/goal bank%s, sorted-neighbour obs, " \
           "auto-reset, DT=0.1" % (a.envs, a.agents, "RVO" if a.policy == "RVO" else "NonCooperative",
                                   " + LaserScanSensor 512x3" if a.laser else "")


def obs_f32(a):
    return a.precision == "f64" and not a.obs_f64


def config_dict(a, world):
    """The `config` object of the JSON line -- identical for both arms (ours / --impl reference)."""
    return {"workload": workload_name(a), "envs_per_gpu": a.envs, "agents": a.agents, "policy": a.policy,
            "state_precision": a.precision, "obs_precision": "f32" if (a.precision == "f32" or obs_f32(a)) else "f64",
            "orca_precision": "f32", "history_rows": a.history,
            "episode_phase": "timed steps start %d steps after a full reset (all agents live at reset)" % (
                max(a.warmup, 3) + a.preroll),
            "parallelism": "envs sharded x%d, no per-step collective" % world,
            "l2": "flushed between timed steps (256 MiB memset + 256 MiB read pass, outside the CUDA-event brackets)",
            "cuda_graph": not a.no_graph}


def bind_numa(local_rank):
    """Bind this process (and so its pinned allocations, first touch) to the CPUs of the GPU's NUMA node: on the 8-GPU
    boxes GPUs 0-3 hang off socket 0 and 4-7 off socket 1, and eight ranks downloading 46 MB per step across the
    socket interconnect is what held the round-1 8-GPU e2e at 0.79 efficiency.  Returns a description (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        node = int(open("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


def scenario_bank(a, seed=1234):
    from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank
    return random_scenario_bank(1024, a.agents, seed=seed)


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline_run(a, steps, envs=None, threads=0, budget_s=15.0):
    """Oracle (C port of the reference step) on the host cores; returns (agent_steps_per_s, cores, sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cav_oracle
    from gym_collision_avoidance_b200.envs.config import to_cav_config
    from gym_collision_avoidance_b200.envs.vec_env import kinds_array
    from gym_collision_avoidance_b200.state import HostBank, HostState
    cores = len(os.sched_getaffinity(0))
    E = envs or max(1024, 256 * cores)
    if a.laser:
        E = max(16, 2 * cores)
    cfg = workload_config(a)
    rows = scenario_bank(a)
    c = to_cav_config(cfg, E, a.agents, precision=a.precision, history_len=a.history)  # float64 columns: no rounding pass
    hs = HostState(c)
    bank = HostBank(rows, kinds=kinds_array(rows.shape[0], a.agents, a.policy))
    orc = cav_oracle.Oracle(c, threads=cores)
    orc.reset(hs, bank)

    def run(k):
        for _ in range(k):
            orc.step(hs)
            orc.reset(hs, bank, env_mask=hs.game_over.copy(), accumulate_stats=True)

    run(3)  # warm-up (thread pool start, first touch)
    if threads == 0:
        # "all the host threads it can use": the affinity mask of a shared box can be larger than its CPU quota, and
        # oversubscribed OpenMP teams run far slower, so pick the fastest team size (torchrun also exports
        # OMP_NUM_THREADS=1, hence the explicit count)
        best = None
        t = cores
        while t >= 1:
            orc.threads = t
            run(1)
            t0 = time.perf_counter()
            run(2)
            dt_ = (time.perf_counter() - t0) / 2
            if best is None or dt_ < best[0]:
                best = (dt_, t)
            t //= 2
        threads = best[1]
    orc.threads = threads
    run(1)
    t0 = time.perf_counter()
    run(3)
    per = (time.perf_counter() - t0) / 3
    if budget_s:  # bound the sample to about budget_s seconds of wall time
        steps = int(max(3, min(steps, budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(hs)
        orc.reset(hs, bank, env_mask=hs.game_over.copy(), accumulate_stats=True)
    dt = time.perf_counter() - t0
    used = threads
    return E * a.agents * steps / dt, used, "%d envs x %d agents x %d steps (C port of the reference step, OpenMP over " \
        "envs, 1 ORCA solve per agent; the Python reference itself measured ~3e3 agent-steps/s/core, BASELINE.md)" % (
            E, a.agents, steps), dt / steps


def auto_cpu_steps(a):
    if a.cpu_steps:
        return a.cpu_steps
    return 20 if a.laser else 2000  # upper bound; cpu_baseline_run caps the sample at ~15 s of wall time


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): `nvidia-smi -lms 100`
    as the recipe says, plus -- the timed region of the default run lasts only ~0.1 s -- an NVML poll every 5 ms from a
    thread (pynvml, read-only queries).  Both feed the same summary."""

    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = []  # (sm_mhz, sm_max_mhz, reasons bitmask)
        self._stop = threading.Event()
        self._nvml_thread = None

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                self.nvml.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), float(smax),
                                  int(reasons_fn(h))))
                time.sleep(0.005)
        except Exception:
            return

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active," \
            "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        try:
            self._nvml_thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._nvml_thread.start()
        except Exception:
            self._nvml_thread = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        self._stop.set()
        if self._nvml_thread is not None:
            self._nvml_thread.join(timeout=1.0)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        n_smi = len(sm)
        for clk, cmax, mask in list(self.nvml):
            sm.append(clk)
            smax.append(cmax)
            for name, bit in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                              ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP)):
                if mask & bit:
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock samples (nvidia-smi / NVML unavailable)"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax), "samples": len(sm),
                "samples_nvidia_smi": n_smi, "samples_nvml": len(sm) - n_smi, "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def measured_traffic(a):
    """(DRAM bytes per launch of the dominant kernel, source) from the committed `ncu --set full` capture of this
    kernel build (profiles/roofline_traffic.json names the capture), or (None, None): ncu cannot run inside the bench."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        key = "%s%s_%dx%d_%s" % (a.precision, "_obsf32" if obs_f32(a) else "", a.envs, a.agents, a.policy)
        return d.get(key), d.get("source")
    except Exception:
        return None, None


def pairwise_obs_bytes(c):
    """Algorithmic bytes per agent-step of the fused pairwise+obs kernel (state read 3 x real4 + flags,
    write obs + reward + done + flags) with the ACTUAL element sizes: state `real` = 4 or 8 bytes, observation and
    reward 4 bytes when the columns are float32 (precision f32, or obs_f32 next to float64 state)."""
    rb = 8 if c.precision == 1 else 4
    ob = 4 if (c.precision != 1 or c.obs_f32) else 8
    K = c.max_num_other_agents_observed
    read = 3 * 4 * rb + 4
    write = ob * (6 + 7 * K) + ob + 1 + 4
    return read + write


def python_reference_rate(a):
    """agent-steps/s of the UNMODIFIED Python reference measured in the build container (it cannot travel to the GPU box;
    tools/measure_python_reference.py writes profiles/python_reference_baseline.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "python_reference_baseline.json")))
        return d.get("%dagents_%s%s" % (a.agents, a.policy, "_laser" if a.laser else ""))
    except Exception:
        return None


def run_ours(a):
    import torch
    import torch.distributed as dist
    from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = None if a.no_numa else bind_numa(local)  # before any pinned allocation
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = workload_config(a)
    env = VecCollisionAvoidanceEnv(a.envs, cfg, a.agents, device=dev, precision=a.precision,
                                   scenario_bank=scenario_bank(a, seed=1234 + rank), policies=a.policy,
                                   history_len=a.history, use_cuda_graph=not a.no_graph, obs_f32=obs_f32(a),
                                   host_lookahead=False if a.no_lookahead else None)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.empty(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB

    def l2_flush():
        """Write a 256 MiB buffer (> 126 MB L2), then stream another 256 MiB through L2 with reads so the timed
        kernels start from an L2 that holds neither their data nor a backlog of dirty flush lines."""
        flush.zero_()
        flush_rd.sum()
    for _ in range(a.preroll + max(a.warmup, 3)):
        env.step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    live = torch.zeros((), dtype=torch.int64, device=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for s in range(a.steps):
        live += (env.state.flags.view(-1) & 0x20).eq(0).sum()  # agent slots that are live going into this step
        l2_flush()  # outside the timed bracket
        ev[s][0].record()
        env.step()
        ev[s][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms_rank = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    per_rank_ms = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank_ms[rank] = dev_ms_rank
    live_t = live.to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(per_rank_ms, op=dist.ReduceOp.SUM)
        dist.all_reduce(live_t, op=dist.ReduceOp.SUM)
    per_rank_ms = [float(x) for x in per_rank_ms.tolist()]
    dev_ms = max(per_rank_ms)
    launches = env.launches_per_step * a.steps

    # ---- roofline of the dominant kernel (fused pairwise+obs), timed alone ----
    cav = env.cav
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    pm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for s in range(50):
        l2_flush()
        pm[s][0].record()
        cav.k_policy()
        cav.k_dynamics()
        pm[s][1].record()
        l2_flush()
        kev[s][0].record()
        cav.k_pairwise_obs()
        kev[s][1].record()
        cav.reset(env.bank, env_mask=env.state.game_over, accumulate_stats=True)
    torch.cuda.synchronize()
    k_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in kev]))
    pm_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in pm]))
    peak, peak_src = measured_peak()
    bpa = pairwise_obs_bytes(env.c)
    achieved = bpa * a.envs * a.agents / (k_ms * 1e-3) / 1e9

    # ---- e2e through the public host-buffer API ----
    h_act = np.zeros((a.envs, a.agents, 2), np.float32)
    env.reset()
    for _ in range(3):
        env.step_host(h_act)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(10, min(a.steps, 50))
    # Wall-clock timed (host work is part of the path), so one descheduled host thread in a 17 ms window moves the
    # number by 15 % (seen once in ~20 runs): three back-to-back windows of e2e_steps calls, the MEDIAN window is reported
    # and all three are kept in the JSON line.
    win = torch.zeros(3, dtype=torch.float64, device=dev)
    for w in range(3):
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            env.step_host(h_act)
        torch.cuda.synchronize()
        win[w] = time.perf_counter() - t0
    win_local = win.clone()
    if world > 1:
        dist.all_reduce(win, op=dist.ReduceOp.MAX)  # per window: the slowest rank
    e2e_windows = [float(x) for x in win.tolist()]
    mid = sorted(range(3), key=lambda i: e2e_windows[i])[1]
    e2e_t = win[mid:mid + 1].clone()
    e2e_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    e2e_rank[rank] = win_local[mid]
    numa_all = [numa]
    if world > 1:
        dist.all_reduce(e2e_rank, op=dist.ReduceOp.SUM)
        try:
            numa_all = [None] * world
            dist.all_gather_object(numa_all, numa)
        except Exception:  # diagnostics only
            numa_all = [numa]
    e2e_val = world * a.envs * a.agents * e2e_steps / float(e2e_t.item())

    totals, per_rank = env.gather_episode_stats()  # the only collective (NCCL all-gather), not on the step path
    out = None
    if rank == 0:
        value = world * a.envs * a.agents * a.steps / (dev_ms * 1e-3)
        traffic, traffic_src = measured_traffic(a)
        live_steps = float(live_t.item())
        out = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": config_dict(a, world),
            "live_agent_steps": live_steps, "live_fraction": live_steps / (world * a.envs * a.agents * a.steps),
            "live_agent_steps_per_s": live_steps / (dev_ms * 1e-3),
            "per_rank_ms_per_step": [x / a.steps for x in per_rank_ms],
            "wall_ms_per_step_incl_flush": t_wall / a.steps * 1e3,
            "clocks": clocks, "numa_binding": numa, "numa_binding_per_rank": numa_all,
            "e2e": {"value": e2e_val, "unit": "agent-steps/s", "h2d_bytes_per_step": world * env.h2d_bytes_per_step,
                    "d2h_bytes_per_step": world * env.d2h_bytes_per_step, "bytes_scope": "whole job (all ranks)",
                    "steps": e2e_steps, "ms_per_step": float(e2e_t.item()) / e2e_steps * 1e3,
                    "windows_ms_per_step": [x / e2e_steps * 1e3 for x in e2e_windows], "window": "median of 3",
                    "per_rank_ms_per_step": [float(x) / e2e_steps * 1e3 for x in e2e_rank.tolist()],
                    "lookahead": bool(env.host_lookahead),
                    "note": "every call uploads the [E,N,2] action array from pinned memory and downloads the float32 obs / "
                            "reward / done arrays; with lookahead (no external-policy agent reads the actions) the next "
                            "step's kernels run under this step's download"},
            "gpu_launches": world * launches,
            "roofline": {"kernel": "k_pairwise_obs (fused pairwise collision+reward+sorted-neighbour obs+done)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_agent_step": bpa, "kernel_ms": k_ms, "policy_move_kernel_ms": pm_ms},
            "episode_stats": totals,
        }
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()  # the CPU leg below runs with no rank spinning in a collective
    if out is not None:
        if not a.no_cpu_baseline:
            v, cores, sample, _ = cpu_baseline_run(a, auto_cpu_steps(a))
            out["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                   "sample": sample, "python_reference": python_reference_rate(a)}
        print(json.dumps(out))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # K timed "steps", each a bounded sample of the workload (max(1024, 256 x cores) envs instead of --envs); the
    # warm-up is inside cpu_baseline_run (3 steps + the team-size probe); capped at ~30 s of wall time
    v, cores, sample, sec_per_step = cpu_baseline_run(a, max(3, a.steps), budget_s=30.0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference", "metric": "agent-steps/sec", "value": v, "unit": "agent-steps/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic",
        "config": config_dict(a, world),
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "python_reference": python_reference_rate(a)},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
