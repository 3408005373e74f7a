"""Vectorized multi-agent collision-avoidance environment on one B200.

``VecCollisionAvoidanceEnv`` is the batched counterpart of the reference's ``DummyVecEnv`` over
``CollisionAvoidanceEnv`` (reference ``envs/vec_env.py:86-159``: a serial Python loop over env instances with
auto-reset on done, ``step_wait`` :120-128) with the observation already packed like
``MultiagentDictToMultiagentArrayWrapper`` (``envs/wrappers.py:149-173``): ``[num_envs, num_agents, obs_len]`` in
``Config.STATES_IN_OBS`` order.  One ``step`` is one CUDA-graph launch of the library's kernels
(policy+move, pairwise+obs, [laserscan], masked reset of finished envs, their first observation); nothing on the
step path runs in Python or on the CPU.
"""
import numpy as np
import torch

from .. import abi
from ..device import Cavoid, DeviceBank
from . import config as _config
from .scenarios import random_scenario_bank

POLICY_IDS = {"RVO": abi.CAV_POLICY_RVO, "noncoop": abi.CAV_POLICY_NONCOOP, "static": abi.CAV_POLICY_STATIC,
              "external": abi.CAV_POLICY_EXTERNAL, "learning": abi.CAV_POLICY_LEARNING,
              "GA3C_CADRL": abi.CAV_POLICY_GA3C}
DYNAMICS_IDS = {"unicycle": abi.CAV_DYN_UNICYCLE, "unicycle_max_turn_rate": abi.CAV_DYN_UNICYCLE_MAX_TURN_RATE,
                "external": abi.CAV_DYN_EXTERNAL}


def kinds_array(num_scenarios, num_agents, policies="RVO", dynamics="unicycle"):
    """uint32 [S, N] kinds column from policy / dynamics names (one name, or one per agent slot)."""
    pol = [policies] * num_agents if isinstance(policies, str) else list(policies)
    dyn = [dynamics] * num_agents if isinstance(dynamics, str) else list(dynamics)
    if len(pol) != num_agents or len(dyn) != num_agents:
        raise ValueError("need one policy / dynamics name per agent slot")
    for p in pol:
        if p not in POLICY_IDS:
            raise abi.CavoidError("policy %r has no device implementation (no CPU fallback)" % (p,))
    row = np.array([abi.make_kind(POLICY_IDS[p], DYNAMICS_IDS[d]) for p, d in zip(pol, dyn)], np.uint32)
    return np.tile(row[None], (num_scenarios, 1))


class VecCollisionAvoidanceEnv(object):
    """``num_envs`` independent environments of ``num_agents`` agent slots on one GPU.

    Args mirror the reference where it has the notion: ``config`` is a reference-style ``Config`` object (ours or
    the reference's own); ``policies`` / ``dynamics`` use the keys of the reference's ``policy_dict`` /
    ``dynamics_dict`` (``test_cases.py:68-96``); ``scenario_bank`` holds "cadrl test case" rows
    ``[px,py,gx,gy,pref_speed,radius(,heading)]`` (``test_cases.py:508``), env ``e`` playing scenario
    ``(e + k*num_envs) mod S`` in its ``k``-th episode.
    """

    def __init__(self, num_envs, config=None, num_agents=None, device="cuda:0", precision="f64", scenario_bank=None,
                 bank_kinds=None, bank_num_active=None, policies="RVO", dynamics="unicycle", auto_reset=True,
                 use_cuda_graph=True, history_len=0, track_aux=False, seed=0, num_scenarios=None,
                 scenario_source="host", obs_f32=False, host_lookahead=None, ga3c=None):
        if config is None:
            config = _config.make_config(num_agents or 10, evaluate=True)
        self.config = config
        self.num_envs = int(num_envs)
        self.num_agents = int(num_agents if num_agents is not None else config.MAX_NUM_AGENTS_IN_ENVIRONMENT)
        self.device = torch.device(device)
        self.auto_reset = bool(auto_reset)
        # obs_f32: float64 state with float32 `obs` / `reward` columns written by the kernels themselves (the dtype the
        # reference declares for its observation space, config.py:93-170): half the HBM and PCIe bytes of the
        # observation, no cast pass in the host-buffer API
        self.c = _config.to_cav_config(config, self.num_envs, self.num_agents, precision=precision,
                                       history_len=history_len, track_aux=track_aux, obs_f32=obs_f32, seed=seed)
        self.obs_offsets, self.obs_len = _config.obs_layout(config)
        if scenario_source not in ("host", "device"):
            raise ValueError("scenario_source must be 'host' or 'device'")
        self.scenario_source = scenario_source
        torch.cuda.set_device(self.device)
        self.cav = Cavoid(self.c, self.device)
        if scenario_source == "device" and scenario_bank is None:
            # bank generated in HBM by cav_generate_scenarios (reference distribution, Philox stream); call
            # regenerate_scenarios(seed) for fresh ones
            S = num_scenarios or max(256, min(16384, self.num_envs))
            self.bank = DeviceBank.empty(S, self.num_agents, device=self.device)
            if bank_kinds is None:
                bank_kinds = kinds_array(S, self.num_agents, policies, dynamics)
            self.bank.kinds = torch.from_numpy(np.ascontiguousarray(bank_kinds).view(np.int32).copy()).to(self.device)
            if bank_num_active is not None:
                self.bank.num_active = torch.from_numpy(np.asarray(bank_num_active, np.int32)).to(self.device)
            self.cav.generate_scenarios(self.bank, seed, random_heading=not config.EVALUATE_MODE)
        else:
            if scenario_bank is None:
                S = num_scenarios or max(256, min(4096, self.num_envs))
                # training mode (EVALUATE_MODE False): random initial headings like cadrl_test_case_to_agents
                scenario_bank = random_scenario_bank(S, self.num_agents, seed=seed,
                                                     random_heading=not config.EVALUATE_MODE)
            scenario_bank = np.asarray(scenario_bank, dtype=np.float64)
            S = scenario_bank.shape[0]
            if bank_kinds is None:
                bank_kinds = kinds_array(S, self.num_agents, policies, dynamics)
            self.bank = DeviceBank(scenario_bank, kinds=bank_kinds, num_active=bank_num_active, device=self.device)
        st = self.cav.state
        self.state = st
        self.ext_actions = torch.zeros((self.num_envs, self.num_agents, 2), dtype=torch.float32, device=self.device)
        self.agent_dones = torch.zeros_like(st.done)  # which_agents_done of the step (kept across auto-reset)
        kinds_np = np.asarray(bank_kinds, dtype=np.uint32) & 0xff
        # GA3C-CADRL agents ("GA3C_CADRL"): `ga3c` is the weight blob (float32 array / tensor) or the prefix of a TF1
        # checkpoint of the reference (".../IROS18/network_01900000"), read without TensorFlow
        if ga3c is not None:
            if isinstance(ga3c, str):
                from .policies.GA3C_CADRL import network as _ga3c_network
                ga3c = _ga3c_network.load_blob(ga3c)
            self.cav.bind_ga3c(ga3c)
        elif (kinds_np == abi.CAV_POLICY_GA3C).any():
            raise abi.CavoidError("GA3C_CADRL agents need the network: pass ga3c=<checkpoint prefix or weight blob>")
        self.has_external_agents = bool(np.isin(kinds_np, [abi.CAV_POLICY_EXTERNAL, abi.CAV_POLICY_LEARNING]).any())
        # ---- host-buffer API (step_host): pinned mirrors, a copy stream, optional one-step lookahead ----
        # The host API hands observations / rewards out as float32.  With float32 columns (precision f32 or obs_f32) the
        # download reads the state columns directly; with float64 columns a cast pass into float32 mirrors runs first.
        self._obs_is_f32 = st.obs.dtype == torch.float32
        # lookahead: with no external / learning agent the caller's actions are never read, so the next step can be
        # launched BEFORE the current results are waited for; its kernels then run under the current step's download
        # (46 MB at 16384 x 10: ~0.87 ms over PCIe 5 x16 against ~0.1 ms of kernels).  Results are snapshotted on the
        # device into two alternating staging sets, so the running step never overwrites what is being downloaded.
        self.host_lookahead = (not self.has_external_agents) if host_lookahead is None else bool(host_lookahead)
        if self.host_lookahead and self.has_external_agents:
            raise abi.CavoidError("host_lookahead needs an environment without external / learning agents")
        nslot = 2 if self.host_lookahead else 1
        self._h_actions = torch.zeros((self.num_envs, self.num_agents, 2), dtype=torch.float32).pin_memory()
        self._h_actions_np = self._h_actions.numpy()
        f32 = torch.float32
        self._d_stage = []
        self._h_out = []
        for k in range(nslot):
            need_stage = self.host_lookahead or not self._obs_is_f32
            self._d_stage.append(None if not need_stage else (
                torch.zeros(tuple(st.obs.shape), dtype=f32, device=self.device),
                torch.zeros(tuple(st.reward.shape), dtype=f32, device=self.device),
                torch.zeros_like(st.game_over), torch.zeros_like(st.done)))
            self._h_out.append((torch.zeros(tuple(st.obs.shape), dtype=f32).pin_memory(),
                                torch.zeros(tuple(st.reward.shape), dtype=f32).pin_memory(),
                                torch.zeros(tuple(st.game_over.shape), dtype=torch.uint8).pin_memory(),
                                torch.zeros(tuple(st.done.shape), dtype=torch.uint8).pin_memory()))
        self._h_obs, self._h_reward, self._h_game_over, self._h_agent_dones = self._h_out[0]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._ev_ready = [torch.cuda.Event() for _ in range(nslot)]   # results of a slot are on the device
        self._ev_home = [torch.cuda.Event() for _ in range(nslot)]    # ... and in pinned host memory
        self._ahead = None  # slot of a step that has been launched but not yet handed to the caller
        self._graph = None
        self._warm = False
        self._use_graph = bool(use_cuda_graph)
        self._stream = torch.cuda.Stream(device=self.device)
        self.launches_per_step = 0
        self.steps = 0

    # ------------------------------------------------------------------------------------------------------
    @property
    def observation(self):
        return self.state.obs

    def regenerate_scenarios(self, seed):
        """Overwrite the bank in place with fresh device-generated scenarios (same buffers, so a captured step graph
        stays valid); they take effect at the following resets."""
        if not hasattr(self.bank, "case_kind"):
            raise abi.CavoidError("regenerate_scenarios needs scenario_source='device'")
        self.cav.generate_scenarios(self.bank, seed, random_heading=not self.config.EVALUATE_MODE)

    def reset(self):
        """Reset every environment (reference ``reset`` -> ``_get_obs``); returns ``(obs, {})``."""
        st = self.state
        st.episode.zero_()
        st.stats.zero_()
        st.game_over.zero_()
        self._ahead = None
        self.cav.reset(self.bank)
        self.agent_dones.zero_()
        return st.obs, {}

    def _enqueue_step(self):
        """The per-step kernel sequence (captured once into a CUDA graph)."""
        before = self.cav.launches
        self.cav.step(self.ext_actions)
        self.agent_dones.copy_(self.state.done)
        if self.auto_reset:
            # DummyVecEnv.step_wait (vec_env.py:124-125): finished envs are reset and return their first obs
            self.cav.reset(self.bank, env_mask=self.state.game_over, accumulate_stats=True)
        self.launches_per_step = self.cav.launches - before

    def _launch(self):
        if not self._use_graph:
            self._enqueue_step()
            return
        if self._graph is None:
            if not self._warm:
                # the first step runs eagerly (lazy CUDA module loading must not happen during capture) ...
                self._enqueue_step()
                self._warm = True
                return
            # ... the second one is captured (capture records, it does not execute) and replayed from then on
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._enqueue_step()
            self._graph = g
        self._graph.replay()

    def step(self, actions=None):
        """One step for every environment.  ``actions``: ``[E, N, 2]`` float32 (device tensor, host tensor or
        NumPy) holding ``[speed, delta_heading]`` (external policies) or the ``[0,1]^2`` learning action
        (``LearningPolicy.py:30-32``) for the agents whose policy is external; ignored for the others.

        Returns ``(obs, rewards, game_over, info)``, all device tensors owned by the env (overwritten by the next
        step): obs ``[E,N,obs_len]``, rewards ``[E,N]``, game_over ``[E]`` uint8 (envs that finished this step and
        whose obs is already their reset obs), info ``{"which_agents_done": [E,N] uint8}``.
        """
        if actions is not None:
            if isinstance(actions, np.ndarray):
                actions = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32))
            self.ext_actions.copy_(actions.reshape(self.ext_actions.shape), non_blocking=True)
        if self._ahead is not None:
            self._ahead = None  # step_host already ran this step (lookahead): its results are in the state columns
        else:
            self._launch()
            self.steps += 1
        st = self.state
        return st.obs, st.reward, st.game_over, {"which_agents_done": self.agent_dones}

    def _launch_to_host(self, slot):
        """Enqueue one step and the download of its results into pinned slot `slot` (nothing waits here)."""
        self._launch()
        self.steps += 1
        st = self.state
        main = torch.cuda.current_stream(self.device)
        stage = self._d_stage[slot]
        if stage is not None:  # device-side snapshot (and float64 -> float32 cast when the columns are float64)
            stage[0].copy_(st.obs)
            stage[1].copy_(st.reward)
            stage[2].copy_(st.game_over)
            stage[3].copy_(self.agent_dones)
            src = stage
        else:
            src = (st.obs, st.reward, st.game_over, self.agent_dones)
        self._ev_ready[slot].record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_ready[slot])
            for h, d in zip(self._h_out[slot], src):
                h.copy_(d, non_blocking=True)
            self._ev_home[slot].record(self._copy_stream)
        if stage is None:
            main.wait_event(self._ev_home[slot])  # the next step must not overwrite columns that are still being read

    def step_host(self, actions=None):
        """Host-buffer API: pinned-host actions in, pinned-host ``(obs, rewards, game_over, which_agents_done)`` out
        (float32 obs / rewards), one call per step.  The returned tensors are owned by the env and are overwritten two
        calls later (lookahead) / by the next call.

        With ``host_lookahead`` (default when no agent has an external or learning policy -- ``actions`` are then never
        read) the NEXT step is launched before this one's results are waited for, so its kernels run under this
        step's device->host copy; the device state is then one step ahead of what has been returned (``step()``
        returns that step without running another one; ``reset()`` discards it)."""
        if actions is not None:
            # plain single-threaded memcpy into the pinned staging buffer: torch's CPU copy_ fans a 1.3 MB copy out
            # over the intra-op thread pool, whose spinning workers then starve the CUDA graph launch (measured on
            # the B200 host: 2.9 ms per step instead of 1.07 ms)
            if isinstance(actions, torch.Tensor):
                actions = actions.detach().cpu().numpy()
            np.copyto(self._h_actions_np, np.asarray(actions, dtype=np.float32).reshape(self._h_actions_np.shape))
            self.ext_actions.copy_(self._h_actions, non_blocking=True)
        if not self.host_lookahead:
            self._launch_to_host(0)
            self._ev_home[0].synchronize()
            return self._h_out[0]
        if self._ahead is None:  # first call after reset() / step(): run the step this call is about
            self._launch_to_host(0)
            self._ahead = 0
        cur = self._ahead
        self._launch_to_host(cur ^ 1)  # the step the NEXT call will return
        self._ahead = cur ^ 1
        self._ev_home[cur].synchronize()
        self._h_obs, self._h_reward, self._h_game_over, self._h_agent_dones = self._h_out[cur]
        return self._h_out[cur]

    @property
    def h2d_bytes_per_step(self):
        return self._h_actions.numel() * 4

    @property
    def d2h_bytes_per_step(self):
        h = self._h_out[0]
        return sum(t.numel() * t.element_size() for t in h)

    # ------------------------------------------------------------------------------------------------------
    def episode_stats(self):
        """Per-rank end-of-episode aggregates accumulated on device (metric definitions: reference
        ``experiments/src/env_utils.py:54-87``)."""
        return self.cav.stats()

    def gather_episode_stats(self):
        """All-gather the per-rank statistics vectors (NCCL over NVLink when torch.distributed is initialised
        with the nccl backend; gloo on CPU tensors in tests) and return (total dict, per-rank array).  This is the
        ONLY collective of the framework and it is not on the per-step path."""
        return gather_stats(self.cav.stats(), self.device)

    def trajectories(self, env_ids=None):
        """Episode trajectories of the CURRENT episode of the given environments, in the reference's per-agent array
        format (needs ``history_len > 0``): a list (per env) of lists (per agent) of dicts with
        ``global_state_history`` ``[T, 11] = [t, px, py, gx, gy, radius, pref_speed, vx, vy, speed, heading]`` and
        ``ego_state_history`` ``[T, 3] = [t, dist_to_goal, heading_ego_frame]`` (reference ``agent.py:257-289``, rows
        ``[:step_num]``), which is what ``visualize.plot_episode`` and
        ``experiments/src/run_trajectory_dataset_creator.py:43-104`` consume.  With ``auto_reset`` a finished
        episode's rows are overwritten by the next one, so collectors run with ``auto_reset=False`` and read at
        ``game_over``.  One device->host copy of the selected envs' ring buffers."""
        st = self.state
        if st.t.get("history") is None:
            raise abi.CavoidError("trajectories() needs history_len > 0")
        ids = list(range(self.num_envs)) if env_ids is None else [int(e) for e in env_ids]
        idx = torch.as_tensor(ids, dtype=torch.long, device=self.device)
        hist = st.history.index_select(0, idx).cpu().numpy().astype(np.float64)
        steps = st.step_num.index_select(0, idx).cpu().numpy()
        nact = st.num_active.index_select(0, idx).cpu().numpy()
        out = []
        for k in range(len(ids)):
            n = int(nact[k])
            agents = []
            for i in range(n):
                T = min(int(steps[k, i]), hist.shape[2])
                agents.append({"global_state_history": hist[k, i, :T, :11].copy(),
                               "ego_state_history": hist[k, i, :T, 11:14].copy()})
            out.append(agents)
        return out

    def obs_slice(self, state_name):
        o = self.obs_offsets[state_name]
        return slice(o, o + _config.state_size(self.config, state_name))

    def close(self):
        self.cav.close()


def shard_envs(total_envs, rank, world_size):
    """Contiguous block of environments owned by `rank`: returns (first_env, num_envs).  Whole
    environments only -- an environment's agents never straddle GPUs."""
    base, rem = divmod(int(total_envs), int(world_size))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def stats_vector(stats):
    return np.array([float(stats[n]) for n in abi.STAT_NAMES], dtype=np.float64)


def gather_stats(stats, device=None):
    """All-gather of the (tiny) per-rank episode-statistics vector; sums are the job totals."""
    import torch.distributed as dist
    v = stats_vector(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(zip(abi.STAT_NAMES, v.tolist())), v[None]
    backend = dist.get_backend()
    dev = torch.device(device) if (backend == "nccl" and device is not None) else torch.device("cpu")
    t = torch.from_numpy(v).to(dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    per_rank = torch.stack(out).cpu().numpy()
    tot = per_rank.sum(axis=0)
    return dict(zip(abi.STAT_NAMES, tot.tolist())), per_rank


class VecEnv(object):
    """Reference ``envs/vec_env.py:6-84`` interface: ``reset`` / ``step_async`` / ``step_wait`` / ``step`` / ``close``."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    @property
    def unwrapped(self):
        return self


class DummyVecEnv(VecEnv):
    """Drop-in for the reference's ``DummyVecEnv`` (``envs/vec_env.py:86-159``): ``DummyVecEnv([make_env] * E)`` steps
    E environments and auto-resets the finished ones (``step_wait`` :120-128) -- here as ONE batched
    ``VecCollisionAvoidanceEnv`` on the GPU instead of a Python loop over env objects.  ``env_fns[0]()`` is called once
    to read the environment's Config (its wrappers, if any, only fix the observation view: the packed
    ``(max_num_agents, obs_len)`` array of ``MultiagentDictToMultiagentArrayWrapper``, or agent 0's row for
    ``FlattenDictWrapper``); episodes come from the Config's ``TEST_CASE_ARGS`` distribution (device scenario generator)
    or an explicit ``scenario_bank``.  Returns host NumPy arrays like the reference: ``obs [E, ...]`` float32, ``rews``
    (``[E]`` = agent 0's reward under ``TRAIN_SINGLE_AGENT``, else ``[E, N]``), ``dones [E]`` bool, ``infos`` list of
    dicts with ``which_agents_done``."""

    def __init__(self, env_fns, device="cuda:0", scenario_bank=None, policies=None, precision="f64", seed=0):
        from .wrappers import FlattenDictWrapper
        proto = env_fns[0]()
        base = getattr(proto, "unwrapped", proto)
        C = getattr(base, "Config", None)
        if C is None:
            from . import Config as C
        self.single_agent_view = isinstance(proto, FlattenDictWrapper) or bool(C.TRAIN_SINGLE_AGENT)
        VecEnv.__init__(self, len(env_fns), getattr(proto, "observation_space", None), getattr(proto, "action_space", None))
        N = C.MAX_NUM_AGENTS_IN_ENVIRONMENT
        if policies is None:
            pol = C.TEST_CASE_ARGS.get("policies", "RVO")
            policies = pol if isinstance(pol, str) else (["learning"] + ["RVO"] * (N - 1))
        if hasattr(base, "close"):
            base.close()
        self.venv = VecCollisionAvoidanceEnv(self.num_envs, C, N, device=device, precision=precision,
                                             scenario_bank=scenario_bank, policies=policies, seed=seed,
                                             scenario_source="host" if scenario_bank is not None else "device")
        self.Config = C
        self.actions = None

    def _view(self, obs):
        return obs[:, 0] if self.single_agent_view else obs

    def reset(self):
        self.venv.reset()
        torch.cuda.current_stream(self.venv.device).synchronize()
        return self._view(self.venv.state.obs.float().cpu().numpy())

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        a = None
        if self.actions is not None:
            a = np.asarray(self.actions, dtype=np.float32)
            if a.ndim == 2:  # one action per environment = agent 0's (single-agent training)
                full = np.zeros((self.num_envs, self.venv.num_agents, 2), np.float32)
                full[:, 0] = a
                a = full
        obs, rew, over, dones = self.venv.step_host(a)
        rew = rew.numpy()
        infos = [{"which_agents_done": dict(enumerate(d.astype(bool)))} for d in dones.numpy()]
        return (np.copy(self._view(obs.numpy())), np.copy(rew[:, 0] if self.Config.TRAIN_SINGLE_AGENT else rew),
                over.numpy().astype(bool), infos)

    def close(self):
        self.venv.close()
