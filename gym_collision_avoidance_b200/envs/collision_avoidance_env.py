"""Single-environment drop-in for the reference's ``CollisionAvoidanceEnv``
(reference ``envs/collision_avoidance_env.py:33-646``): same constructor, ``reset`` / ``step`` / ``set_agents`` /
``set_testcase`` / ``set_static_map`` / ``set_plot_save_dir`` methods, same return values (observation
dict-of-dicts, reward array, ``game_over``, ``False``, info dict), same ``env.agents`` list of ``Agent`` objects.

The step itself is one ``cav_step`` on a 1-environment handle of ``libcavoid_sm100a.so`` (float64 state);
afterwards the device columns are copied back into the ``Agent`` objects.  It exists for API parity and for tests
that read like the reference's; throughput comes from :class:`vec_env.VecCollisionAvoidanceEnv`.
"""
import inspect

import numpy as np
import torch

from .. import abi
from ..device import Cavoid
from ..state import HostState
from . import config as _config
from . import test_cases as tc
from .Map import Map
from .wrappers import packed_to_dict


class _Space(object):
    """Minimal stand-in for gym.spaces.Box / Dict (gym itself is not a dependency of the hot path)."""

    def __init__(self, low=None, high=None, shape=None, dtype=None, spaces=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype
        self.spaces = spaces if spaces is not None else {}

    def __getitem__(self, k):
        return self.spaces[k]

    def __setitem__(self, k, v):
        self.spaces[k] = v


_HOST_HOOKS = ("find_next_action", "external_action_to_action", "step", "update_ego_frame")


def _kernel_id(obj, what):
    """Device kernel id of a built-in plugin object.  A user subclass that overrides a method the reference would call
    on the step path (``find_next_action`` ...) cannot be honoured -- the device runs the built-in arithmetic -- so it
    is rejected instead of being silently replaced (there is no CPU fallback)."""
    cls = type(obj)
    kid = getattr(cls, "kernel_id", None)
    if kid is None:
        raise abi.CavoidError("%s %s has no device implementation (kernel_id); this framework has no CPU fallback "
                              "for the step path" % (what, cls.__name__))
    owner = next(k for k in cls.__mro__ if "kernel_id" in vars(k))
    for k in cls.__mro__:
        if k is owner:
            break
        over = [h for h in _HOST_HOOKS if h in vars(k)]
        if over:
            raise abi.CavoidError("%s %s overrides %s of the built-in %s: the step path runs on the device and cannot "
                                  "call Python code (no CPU fallback)" % (what, cls.__name__, over, owner.__name__))
    return kid


def _action_of(actions, i):
    """actions[i] for a dict keyed by agent index or any sequence / array indexed by it (the reference indexes
    ``actions[agent_index]``, env.py:316)."""
    try:
        if isinstance(actions, dict):
            return actions[i]
        if actions is None or i >= len(actions):
            raise KeyError(i)
        return actions[i]
    except (KeyError, IndexError, TypeError):
        raise KeyError("agent %d has an external policy but no action was supplied" % i)


class CollisionAvoidanceEnv(object):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 30}

    def __init__(self, config=None, device="cuda:0"):
        from . import Config as _Singleton
        self.Config = config if config is not None else _Singleton
        C = self.Config
        self.id = 0
        self.device = device
        self._initialize_rewards()
        self.num_agents = C.MAX_NUM_AGENTS_IN_ENVIRONMENT
        self.dt_nominal = C.DT
        self.collision_dist = C.COLLISION_DIST
        self.getting_close_range = C.GETTING_CLOSE_RANGE
        self.evaluate = C.EVALUATE_MODE
        self.plot_episodes = C.SHOW_EPISODE_PLOTS or C.SAVE_EPISODE_PLOTS
        self.plt_limits, self.plt_fig_size = C.PLT_LIMITS, C.PLT_FIG_SIZE
        self.test_case_index = 0
        self.set_testcase(C.TEST_CASE_FN, dict(C.TEST_CASE_ARGS))
        self.animation_period_steps = C.ANIMATION_PERIOD_STEPS
        self.max_heading_change = np.pi / 3
        self.min_heading_change = -self.max_heading_change
        self.min_speed, self.max_speed = 0.0, 1.0
        self.action_space_type = C.ACTION_SPACE_TYPE
        self.low_action = np.array([self.min_speed, self.min_heading_change])
        self.high_action = np.array([self.max_speed, self.max_heading_change])
        self.action_space = _Space(self.low_action, self.high_action, dtype=np.float32)
        self.observation = {}
        self.observation_space = _Space()
        for a in range(C.MAX_NUM_AGENTS_IN_ENVIRONMENT):
            self.observation[a] = {}
            self.observation_space.spaces[a] = _Space()
            for state in C.STATES_IN_OBS:
                info = C.STATE_INFO_DICT[state]
                self.observation[a][state] = np.zeros(info["size"], dtype=info["dtype"])
                self.observation_space.spaces[a][state] = _Space(info["bounds"][0] * np.ones(info["size"]),
                                                                 info["bounds"][1] * np.ones(info["size"]),
                                                                 dtype=info["dtype"])
        self.agents = None
        self.default_agents = None
        self.prev_episode_agents = None
        self.static_map_filename = None
        self.static_map_array = None
        self.map = None
        self.episode_step_number = None
        self.episode_number = 0
        self.plot_save_dir = None
        self.plot_policy_name = None
        self.perturbed_obs = None
        self._cav = None
        self._hist_len = 0

    # ---- reference setters ----------------------------------------------------------------------------------
    def set_agents(self, agents):
        self.default_agents = agents

    def set_static_map(self, map_filename):
        """Path of a ``.npy`` boolean map, or a boolean array (reference ``env.py:369-376`` takes an image path)."""
        if isinstance(map_filename, np.ndarray):
            self.static_map_array = map_filename
        else:
            self.static_map_filename = map_filename

    def set_plot_save_dir(self, plot_save_dir):
        import os
        os.makedirs(plot_save_dir, exist_ok=True)
        self.plot_save_dir = plot_save_dir

    def set_perturbed_info(self, perturbed_obs):
        self.perturbed_obs = perturbed_obs

    def set_testcase(self, test_case_fn_str, test_case_args):
        test_case_fn = getattr(tc, test_case_fn_str, None)
        assert callable(test_case_fn), "unknown test case function %r" % (test_case_fn_str,)
        params = inspect.signature(test_case_fn).parameters
        for key in list(test_case_args.keys()):
            if key not in params:
                del test_case_args[key]
        self.test_case_fn = test_case_fn
        self.test_case_args = test_case_args

    def _initialize_rewards(self):
        C = self.Config
        self.reward_at_goal = C.REWARD_AT_GOAL
        self.reward_collision_with_agent = C.REWARD_COLLISION_WITH_AGENT
        self.reward_collision_with_wall = C.REWARD_COLLISION_WITH_WALL
        self.reward_getting_close = C.REWARD_GETTING_CLOSE
        self.reward_entered_norm_zone = C.REWARD_ENTERED_NORM_ZONE
        self.reward_time_step = C.REWARD_TIME_STEP
        self.reward_wiggly_behavior = C.REWARD_WIGGLY_BEHAVIOR
        self.wiggly_behavior_threshold = C.WIGGLY_BEHAVIOR_THRESHOLD
        self.possible_reward_values = np.array([self.reward_at_goal, self.reward_collision_with_agent,
                                                self.reward_time_step, self.reward_collision_with_wall,
                                                self.reward_wiggly_behavior])
        self.min_possible_reward = np.min(self.possible_reward_values)
        self.max_possible_reward = np.max(self.possible_reward_values)

    # ---- device plumbing -------------------------------------------------------------------------------------
    def _ensure_handle(self, hist_len):
        C = self.Config
        if self._cav is not None and self._hist_len >= hist_len:
            return
        if self._cav is not None:
            self._cav.close()
        self._hist_len = max(hist_len, 64)
        c = _config.to_cav_config(C, 1, C.MAX_NUM_AGENTS_IN_ENVIRONMENT, precision="f64",
                                  history_len=self._hist_len, store_history=bool(C.STORE_HISTORY), track_aux=True,
                                  seed=int(np.random.randint(0, 2 ** 31 - 1)) if C.RVO_COLLAB_COEFF < 0 else 0)
        self._c = c
        self._cav = Cavoid(c, self.device)
        self._ga3c_bound = None
        if C.USE_STATIC_MAP:
            self.map = Map(16, 16, 0.1, self.static_map_filename, static_map=self.static_map_array)
            self._cav.state.set_static_map(self.map.bits())
            self._cav.bind()

    # columns that hold Agent attributes; everything else (observation incl. the laserscan history rows, trajectory ring,
    # laser_count, min_dist, done / game_over, episode bookkeeping, kd_order) is device-side history that an upload of
    # agent STATES must leave alone
    _AGENT_COLS = ("pv", "goal", "ht", "ego", "flags", "step_num", "kinds", "num_active", "aux0", "past_actions",
                   "past_vel", "actions")

    def _upload_agents(self, full=True):
        """Agent objects -> device columns (Agent attributes as they are: the reference's reset does not
        re-initialise ``default_agents`` either, env.py:362-363).  ``full=False`` (Agent.set_state between steps, the
        ExternalDynamics workflow) writes only the agent-state columns."""
        hs = HostState(self._c)
        n = len(self.agents)
        hs.num_active[0] = n
        ga3c = None
        for i, a in enumerate(self.agents):
            hs.pv[0, i] = [a.pos_global_frame[0], a.pos_global_frame[1], a.vel_global_frame[0], a.vel_global_frame[1]]
            hs.goal[0, i] = [a.goal_global_frame[0], a.goal_global_frame[1], a.radius, a.pref_speed]
            hs.ht[0, i] = [a.heading_global_frame, a.time_remaining_to_reach_goal, a.t, a.turning_dir]
            hs.ego[0, i] = [a.dist_to_goal, a.heading_ego_frame, a.ref_prll[0], a.ref_prll[1]]
            f = 0
            f |= abi.CAV_FLAG_AT_GOAL if a.is_at_goal else 0
            f |= abi.CAV_FLAG_WAS_AT_GOAL_ALREADY if a.was_at_goal_already else 0
            f |= abi.CAV_FLAG_IN_COLLISION if a.in_collision else 0
            f |= abi.CAV_FLAG_WAS_IN_COLLISION_ALREADY if a.was_in_collision_already else 0
            f |= abi.CAV_FLAG_RAN_OUT_OF_TIME if a.ran_out_of_time else 0
            f |= abi.CAV_FLAG_DONE if a.is_done else 0
            hs.flags[0, i] = f
            hs.step_num[0, i] = a.step_num
            hs.kinds[0, i] = abi.make_kind(_kernel_id(a.policy, "policy"), _kernel_id(a.dynamics_model, "dynamics"))
            if hs.kinds[0, i] & 0xff == abi.CAV_POLICY_GA3C:
                ga3c = a.policy
            hs.aux0[0, i] = [a.speed_global_frame, a.delta_heading_global_frame, a.vel_ego_frame[0], a.vel_ego_frame[1]]
            hs.past_actions[0, i] = a.past_actions.reshape(-1)
            hs.past_vel[0, i] = a.past_global_velocities.reshape(-1)
            hs.actions[0, i] = a.past_actions[0]
        if ga3c is not None and full:
            # one network per environment (the reference loads one TF session per GA3CCADRLPolicy object; agents that were
            # initialised from different checkpoints cannot share a device network)
            if ga3c.weight_blob is None:
                raise abi.CavoidError("GA3CCADRLPolicy: call agent.policy.initialize_network() before env.set_agents / reset")
            if getattr(self, "_ga3c_bound", None) is not ga3c.weight_blob:  # upload the 683 KB once, not at every reset
                self._cav.bind_ga3c(ga3c.weight_blob)
                self._ga3c_bound = ga3c.weight_blob
        self._cav.state.load_host(hs, cols=None if full else self._AGENT_COLS)

    def _download(self):
        torch.cuda.synchronize()
        hs = self._cav.state.to_host()
        C = self.Config
        K = C.MAX_NUM_OTHER_AGENTS_OBSERVED
        off, _ = _config.obs_layout(C)
        for i, a in enumerate(self.agents):
            pv, gl, ht, ego, f = hs.pv[0, i], hs.goal[0, i], hs.ht[0, i], hs.ego[0, i], int(hs.flags[0, i])
            a.pos_global_frame = np.array(pv[0:2])
            a.vel_global_frame = np.array(pv[2:4])
            a.goal_global_frame = np.array(gl[0:2])
            a.heading_global_frame = np.float64(ht[0])
            a.time_remaining_to_reach_goal, a.t, a.turning_dir = float(ht[1]), float(ht[2]), float(ht[3])
            a.dist_to_goal, a.heading_ego_frame = float(ego[0]), float(ego[1])
            a.ref_prll = np.array(ego[2:4])
            a.ref_orth = np.array([-ego[3], ego[2]])
            a.is_at_goal = bool(f & abi.CAV_FLAG_AT_GOAL)
            a.was_at_goal_already = bool(f & abi.CAV_FLAG_WAS_AT_GOAL_ALREADY)
            a.in_collision = bool(f & abi.CAV_FLAG_IN_COLLISION)
            a.was_in_collision_already = bool(f & abi.CAV_FLAG_WAS_IN_COLLISION_ALREADY)
            a.ran_out_of_time = bool(f & abi.CAV_FLAG_RAN_OUT_OF_TIME)
            a.is_done = bool(f & abi.CAV_FLAG_DONE)
            a.step_num = int(hs.step_num[0, i])
            a.speed_global_frame, a.delta_heading_global_frame = float(hs.aux0[0, i, 0]), float(hs.aux0[0, i, 1])
            a.vel_ego_frame = np.array(hs.aux0[0, i, 2:4])
            a.past_actions = np.array(hs.past_actions[0, i]).reshape(2, 2)
            a.past_global_velocities = np.array(hs.past_vel[0, i]).reshape(2, 2)
            a.min_dist_to_other_agents = float(hs.min_dist[0, i])
            a.sensor_data = {}
            if "other_agents_states" in off:
                o = off["other_agents_states"]
                rows = np.array(hs.obs[0, i, o:o + 7 * K]).reshape(K, 7)
                a.sensor_data["other_agents_states"] = rows
                cnt = (int(hs.obs[0, i, off["num_other_agents"]]) if "num_other_agents" in off
                       else int((rows[:, 5] != 0).sum()))
                a.num_other_agents_observed = cnt
                if cnt > 0:
                    a.other_agent_states = rows[0].copy()
            if "laserscan" in off:
                o = off["laserscan"]
                shape = (C.LASERSCAN_NUM_PAST, C.LASERSCAN_LENGTH)
                a.sensor_data["laserscan"] = np.array(hs.obs[0, i, o:o + shape[0] * shape[1]]).reshape(shape)
            if C.STORE_HISTORY and "history" in hs.arrays:
                k = min(a.step_num, self._hist_len, a.global_state_history.shape[0])
                a.global_state_history[:k] = hs.history[0, i, :k, :11]
                a.ego_state_history[:k] = hs.history[0, i, :k, 11:]
        return hs

    def _obs_dict(self, hs):
        C = self.Config
        n = len(self.agents)
        d = packed_to_dict(C, hs.obs[0], num_slots=C.MAX_NUM_AGENTS_IN_ENVIRONMENT)
        for i in range(n, C.MAX_NUM_AGENTS_IN_ENVIRONMENT):
            for state in C.STATES_IN_OBS:
                d[i][state] = np.zeros(C.STATE_INFO_DICT[state]["size"], dtype=C.STATE_INFO_DICT[state]["dtype"])
        self.observation = d
        return d

    # ---- reference API ---------------------------------------------------------------------------------------
    def reset(self):
        """Reference ``reset`` (``env.py:236-282``): (re)build the agents, zero the observation, sense."""
        C = self.Config
        if self.episode_step_number is not None and self.episode_step_number > 0:
            self.episode_number += 1
        self.begin_episode = True
        self.episode_step_number = 0
        if self.evaluate and self.agents is not None:
            import copy
            self.prev_episode_agents = copy.deepcopy(self.agents)
        if self.default_agents is None:
            self.agents = self.test_case_fn(**self.test_case_args)
        else:
            self.agents = self.default_agents
        if len(self.agents) > C.MAX_NUM_AGENTS_IN_ENVIRONMENT:
            raise abi.CavoidError("more agents than Config.MAX_NUM_AGENTS_IN_ENVIRONMENT")
        for agent in self.agents:
            agent.max_heading_change = self.max_heading_change
            agent.max_speed = self.max_speed
            for s in agent.sensors:
                if s.name not in ("other_agents_states", "laserscan"):
                    raise abi.CavoidError("sensor %r has no device implementation" % (s.name,))
        hist = max([a.num_states_in_history for a in self.agents] + [1]) if C.STORE_HISTORY else 0
        self._ensure_handle(hist)
        st = self._cav.state
        st.obs.zero_()
        if "laser_count" in st.t:
            st.laser_count.zero_()
        self._upload_agents()
        self._cav.k_observe()
        hs = self._download()
        return self._obs_dict(hs), {}

    def step(self, actions, dt=None):
        """Reference ``step`` (``env.py:156-234``)."""
        C = self.Config
        if dt is None:
            dt = self.dt_nominal
        self.episode_step_number += 1
        N = C.MAX_NUM_AGENTS_IN_ENVIRONMENT
        ext = np.zeros((1, N, 2), np.float32)
        dirty = False
        for i, a in enumerate(self.agents):
            if a.policy.is_external and not a.is_done:
                ext[0, i] = np.asarray(_action_of(actions, i), dtype=np.float64).reshape(-1)[:2]
            dirty = dirty or getattr(a, "_state_dirty", False)
        if dirty:  # Agent.set_state was used (ExternalDynamics agents): push the new states first
            self._upload_agents(full=False)
            for a in self.agents:
                a._state_dirty = False
        self._cav.step(torch.from_numpy(ext).to(self.device), float(dt))
        hs = self._download()
        n = len(self.agents)
        rewards = np.array(hs.reward[0, :n], dtype=np.float64)
        if C.TRAIN_SINGLE_AGENT:
            rewards = rewards[0]
        game_over = bool(hs.game_over[0])
        which_agents_done = {a.id: bool(hs.done[0, i]) for i, a in enumerate(self.agents)}
        which_agents_learning = {a.id: a.policy.is_still_learning for a in self.agents}
        return (self._obs_dict(hs), rewards, game_over, False,
                {"which_agents_done": which_agents_done, "which_agents_learning": which_agents_learning})

    def close(self):
        if self._cav is not None:
            self._cav.close()
            self._cav = None
