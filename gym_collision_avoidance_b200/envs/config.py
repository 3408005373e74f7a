"""Configuration objects with the reference's attribute names (reference ``envs/config.py:3-304``).

``Config`` keeps the reference's public attribute surface (``DT``, ``MAX_TIME_RATIO``, ``STATES_IN_OBS``,
``STATE_INFO_DICT`` ...) so code written against the reference keeps working; subclasses may pre-set
``MAX_NUM_AGENTS_IN_ENVIRONMENT`` / ``MAX_NUM_OTHER_AGENTS_OBSERVED`` / ``USE_STATIC_MAP`` / ``STATES_IN_OBS``
before calling ``Config.__init__`` exactly like the reference's ``hasattr`` guards (``config.py:21,64-70,178``).
:func:`to_cav_config` flattens any such object (ours or the reference's own) into the POD ``cav_config_t``.
"""
import ctypes
import math

import numpy as np

from .. import abi

_OBS_ROW = 7  # [p_prll, p_orth, v_prll, v_orth, radius_other, combined_radius, dist_2_other]

_SCALAR_DEFAULTS = dict(
    # display (host-side only; the device path never plots)
    ANIMATE_EPISODES=False, SHOW_EPISODE_PLOTS=False, SAVE_EPISODE_PLOTS=False, ANIMATION_PERIOD_STEPS=5,
    PLT_LIMITS=None, PLT_FIG_SIZE=(10, 8),
    # modes
    TRAIN_MODE=True, PLAY_MODE=False, EVALUATE_MODE=False, TRAIN_SINGLE_AGENT=False,
    # rewards (reference config.py:30-39)
    REWARD_AT_GOAL=1.0, REWARD_COLLISION_WITH_AGENT=-0.25, REWARD_COLLISION_WITH_WALL=-0.25,
    REWARD_GETTING_CLOSE=-0.1, REWARD_ENTERED_NORM_ZONE=-0.05, REWARD_TIME_STEP=0.0, REWARD_WIGGLY_BEHAVIOR=0.0,
    WIGGLY_BEHAVIOR_THRESHOLD=np.inf, COLLISION_DIST=0.0, GETTING_CLOSE_RANGE=0.2, SOCIAL_NORMS="none",
    # simulation (config.py:43-47)
    DT=0.2, NEAR_GOAL_THRESHOLD=0.2, MAX_TIME_RATIO=2.0,
    # sensors (config.py:72-79)
    SENSING_HORIZON=np.inf, LASERSCAN_LENGTH=512, LASERSCAN_NUM_PAST=3, NUM_STEPS_IN_OBS_HISTORY=1,
    NUM_PAST_ACTIONS_IN_STATE=0,
    # RVO (config.py:83-86)
    RVO_TIME_HORIZON=5.0, RVO_COLLAB_COEFF=0.5, RVO_ANTI_COLLAB_T=1.0,
    STORE_HISTORY=True, PLOT_EVERY_N_EPISODES=100, COLLISION_AVOIDANCE=True,
    AGENT_SORTING_METHOD="closest_first",
)


class Config(object):
    def __init__(self):
        self.continuous, self.discrete = range(2)
        self.ACTION_SPACE_TYPE = self.continuous
        for k, v in _SCALAR_DEFAULTS.items():
            setattr(self, k, v)
        for name, default in (("PLOT_CIRCLES_ALONG_TRAJ", True), ("USE_STATIC_MAP", False),
                              ("MAX_NUM_AGENTS_IN_ENVIRONMENT", 4), ("MAX_NUM_AGENTS_TO_SIM", 4)):
            if not hasattr(self, name):
                setattr(self, name, default)
        self.MAX_NUM_OTHER_AGENTS_IN_ENVIRONMENT = self.MAX_NUM_AGENTS_IN_ENVIRONMENT - 1
        if not hasattr(self, "MAX_NUM_OTHER_AGENTS_OBSERVED"):
            self.MAX_NUM_OTHER_AGENTS_OBSERVED = self.MAX_NUM_AGENTS_IN_ENVIRONMENT - 1
        self.TEST_CASE_FN = "get_testcase_random"
        self.TEST_CASE_ARGS = {
            "policy_to_ensure": "learning",
            "policies": ["noncoop", "learning", "static"],
            "policy_distr": [0.05, 0.9, 0.05],
            "speed_bnds": [0.5, 2.0],
            "radius_bnds": [0.2, 0.8],
            "side_length": [{"num_agents": [0, 5], "side_length": [4, 5]},
                            {"num_agents": [5, np.inf], "side_length": [6, 8]}],
        }
        K = self.MAX_NUM_OTHER_AGENTS_OBSERVED
        row_std = np.array([5.0, 5.0, 1.0, 1.0, 1.0, 5.0, 1.0], dtype=np.float32)
        row_mean = np.array([0.0, 0.0, 0.0, 0.0, 0.5, 0.0, 1.0], dtype=np.float32)

        def scalar(attr, bounds, std=None, mean=None):
            d = {"dtype": np.float32, "size": 1, "bounds": bounds, "attr": attr}
            if std is not None:
                d["std"] = np.array([std], dtype=np.float32)
                d["mean"] = np.array([mean], dtype=np.float32)
            return d

        scan_shape = (self.LASERSCAN_NUM_PAST, self.LASERSCAN_LENGTH)
        self.STATE_INFO_DICT = {
            "dist_to_goal": scalar('get_agent_data("dist_to_goal")', [-np.inf, np.inf], 5.0, 0.0),
            "radius": scalar('get_agent_data("radius")', [0, np.inf], 1.0, 0.5),
            "heading_ego_frame": scalar('get_agent_data("heading_ego_frame")', [-np.pi, np.pi], 3.14, 0.0),
            "pref_speed": scalar('get_agent_data("pref_speed")', [0, np.inf], 1.0, 1.0),
            "num_other_agents": scalar('get_agent_data("num_other_agents_observed")', [0, np.inf], 1.0, 1.0),
            "other_agent_states": {"dtype": np.float32, "size": _OBS_ROW, "bounds": [-np.inf, np.inf],
                                   "attr": 'get_agent_data("other_agent_states")', "std": row_std, "mean": row_mean},
            "other_agents_states": {"dtype": np.float32, "size": (K, _OBS_ROW), "bounds": [-np.inf, np.inf],
                                    "attr": 'get_sensor_data("other_agents_states")',
                                    "std": np.tile(row_std, (K, 1)), "mean": np.tile(row_mean, (K, 1))},
            "laserscan": {"dtype": np.float32, "size": scan_shape, "bounds": [0.0, 6.0],
                          "attr": 'get_sensor_data("laserscan")',
                          "std": 5.0 * np.ones(scan_shape, dtype=np.float32),
                          "mean": 5.0 * np.ones(scan_shape, dtype=np.float32)},
            "is_learning": scalar('get_agent_data_equiv("policy.str", "learning")', [0.0, 1.0]),
        }
        self.setup_obs()

    def setup_obs(self):
        if not hasattr(self, "STATES_IN_OBS"):
            self.STATES_IN_OBS = ["is_learning", "num_other_agents", "dist_to_goal", "heading_ego_frame",
                                  "pref_speed", "radius", "other_agents_states"]
        if not hasattr(self, "STATES_NOT_USED_IN_POLICY"):
            self.STATES_NOT_USED_IN_POLICY = ["is_learning"]
        self.MEAN_OBS, self.STD_OBS = {}, {}
        for state in self.STATES_IN_OBS:
            info = self.STATE_INFO_DICT[state]
            if "mean" in info:
                self.MEAN_OBS[state] = info["mean"]
            if "std" in info:
                self.STD_OBS[state] = info["std"]


class EvaluateConfig(Config):
    """Reference ``config.py:193-200``: DT=0.1, MAX_TIME_RATIO=8, EVALUATE_MODE, up to 19 agents."""

    def __init__(self):
        if not hasattr(self, "MAX_NUM_AGENTS_IN_ENVIRONMENT"):
            self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 19
        Config.__init__(self)
        self.EVALUATE_MODE = True
        self.TRAIN_MODE = False
        self.DT = 0.1
        self.MAX_TIME_RATIO = 8.0


class Example(EvaluateConfig):
    """Reference ``config.py:202-210`` (the plotting switches are host-side only)."""

    def __init__(self):
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS = True
        self.PLOT_CIRCLES_ALONG_TRAJ = True
        self.ANIMATE_EPISODES = True


class Formations(EvaluateConfig):
    """Reference ``config.py:212-224``."""

    def __init__(self):
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS, self.SHOW_EPISODE_PLOTS, self.ANIMATE_EPISODES = True, False, True
        self.NEAR_GOAL_THRESHOLD = 0.2
        self.PLT_LIMITS, self.PLT_FIG_SIZE = [[-5, 6], [-2, 7]], (10, 10)
        self.PLOT_CIRCLES_ALONG_TRAJ = False
        self.NUM_AGENTS_TO_TEST, self.POLICIES_TO_TEST, self.NUM_TEST_CASES = [6], ["GA3C-CADRL-10"], 2
        self.LETTERS = ["C", "A", "D", "R", "L"]


class SmallTestSuite(EvaluateConfig):
    """Reference ``config.py:226-233``."""

    def __init__(self):
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS, self.SHOW_EPISODE_PLOTS, self.ANIMATE_EPISODES = True, False, False
        self.PLOT_CIRCLES_ALONG_TRAJ = True
        self.NUM_TEST_CASES = 4


class LargeNumAgents(EvaluateConfig):
    """Reference ``config.py:235-251``: observe 39 others (the reference still caps the environment at 19 agents)."""

    def __init__(self):
        self.MAX_NUM_OTHER_AGENTS_OBSERVED = 39
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS, self.SHOW_EPISODE_PLOTS, self.ANIMATE_EPISODES = True, False, False
        self.PLOT_CIRCLES_ALONG_TRAJ = False
        self.PLT_LIMITS = [[-20, 20], [-15, 15]]
        self.NUM_TEST_CASES, self.NUM_AGENTS_TO_TEST = 10, [40]
        self.RECORD_PICKLE_FILES = False
        self.POLICIES_TO_TEST = ["GA3C-CADRL-10"]
        self.FIXED_RADIUS_AND_VPREF = False
        self.NEAR_GOAL_THRESHOLD = 0.2


class FullTestSuite(EvaluateConfig):
    """Reference ``config.py:253-287``."""

    def __init__(self):
        self.MAX_NUM_OTHER_AGENTS_OBSERVED = 19
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS, self.SHOW_EPISODE_PLOTS, self.ANIMATE_EPISODES = True, False, False
        self.PLOT_CIRCLES_ALONG_TRAJ = True
        self.NUM_TEST_CASES, self.NUM_AGENTS_TO_TEST = 4, [2, 3, 4]
        self.RECORD_PICKLE_FILES = False
        self.POLICIES_TO_TEST = ["CADRL", "RVO", "GA3C-CADRL-10"]
        self.FIXED_RADIUS_AND_VPREF = False
        self.NEAR_GOAL_THRESHOLD = 0.2


class CollectRegressionDataset(EvaluateConfig):
    """Reference ``config.py:290-304``."""

    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 4
        self.MAX_NUM_AGENTS_TO_SIM = 4
        self.DATASET_NAME = ""
        EvaluateConfig.__init__(self)
        self.TEST_CASE_ARGS["policies"] = "CADRL"
        self.AGENT_SORTING_METHOD = "closest_first"


def make_config(num_agents, evaluate=True, use_static_map=False, states_in_obs=None, **overrides):
    """Convenience: a Config for exactly ``num_agents`` slots (what the oracle harness does with
    ``GYM_CONFIG_PATH``)."""
    base = EvaluateConfig if evaluate else Config

    class _Cfg(base):
        def __init__(self):
            self.MAX_NUM_AGENTS_IN_ENVIRONMENT = int(num_agents)
            self.MAX_NUM_AGENTS_TO_SIM = int(num_agents)
            if use_static_map:
                self.USE_STATIC_MAP = True
            if states_in_obs is not None:
                self.STATES_IN_OBS = list(states_in_obs)
            if "MAX_NUM_OTHER_AGENTS_OBSERVED" in overrides:
                self.MAX_NUM_OTHER_AGENTS_OBSERVED = overrides["MAX_NUM_OTHER_AGENTS_OBSERVED"]
            base.__init__(self)
            for k, v in overrides.items():
                setattr(self, k, v)

    return _Cfg()


def state_size(cfg, state):
    size = cfg.STATE_INFO_DICT[state]["size"]
    return int(np.prod(size))


def obs_layout(cfg):
    """(offsets dict, total length) of the packed per-agent observation in STATES_IN_OBS order
    (reference ``wrappers.py:149-173``)."""
    off, pos = {}, 0
    for s in cfg.STATES_IN_OBS:
        if s not in abi.OBS_STATE_IDS:
            raise abi.CavoidError("observation state %r has no device implementation" % (s,))
        off[s] = pos
        pos += state_size(cfg, s)
    return off, pos


_SORT = {"closest_first": abi.CAV_SORT_CLOSEST_FIRST, "closest_last": abi.CAV_SORT_CLOSEST_LAST,
         "time_to_impact": abi.CAV_SORT_TIME_TO_IMPACT}


def parse_precision(p):
    if isinstance(p, str):
        if p in ("f64", "float64", "double"):
            return abi.CAV_F64
        if p in ("f32", "float32", "float"):
            return abi.CAV_F32
        raise ValueError("unknown precision %r" % (p,))
    if p in (np.float64,):
        return abi.CAV_F64
    if p in (np.float32,):
        return abi.CAV_F32
    return abi.CAV_F64 if int(p) == abi.CAV_F64 else abi.CAV_F32


def to_cav_config(cfg, num_envs, num_agents=None, precision="f64", store_history=None, history_len=0,
                  track_aux=False, obs_f32=False, seed=0):
    """Flatten a reference-style Config object into ``cav_config_t``.  ``obs_f32``: float64 state with float32
    observation / reward columns (the reference's declared observation dtype)."""
    c = abi.CavConfig()
    c.struct_size = ctypes.sizeof(abi.CavConfig)
    c.precision = parse_precision(precision)
    c.num_envs = int(num_envs)
    c.num_agents = int(num_agents if num_agents is not None else cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT)
    c.dt = float(cfg.DT)
    c.near_goal_threshold = float(cfg.NEAR_GOAL_THRESHOLD)
    c.max_time_ratio = float(cfg.MAX_TIME_RATIO)
    c.max_heading_change = math.pi / 3  # env.max_heading_change, reference env.py:88
    c.max_speed = 1.0
    c.reward_at_goal = float(cfg.REWARD_AT_GOAL)
    c.reward_collision_with_agent = float(cfg.REWARD_COLLISION_WITH_AGENT)
    c.reward_collision_with_wall = float(cfg.REWARD_COLLISION_WITH_WALL)
    c.reward_time_step = float(cfg.REWARD_TIME_STEP)
    c.reward_wiggly_behavior = float(cfg.REWARD_WIGGLY_BEHAVIOR)
    c.wiggly_behavior_threshold = float(cfg.WIGGLY_BEHAVIOR_THRESHOLD)
    c.getting_close_range = float(cfg.GETTING_CLOSE_RANGE)
    c.sensing_horizon = float(cfg.SENSING_HORIZON)
    c.max_num_other_agents_observed = int(cfg.MAX_NUM_OTHER_AGENTS_OBSERVED)
    c.agent_sorting_method = _SORT[cfg.AGENT_SORTING_METHOD]
    if cfg.EVALUATE_MODE:
        c.game_over_mode = abi.CAV_GAME_OVER_ALL_DONE
    elif cfg.TRAIN_SINGLE_AGENT:
        c.game_over_mode = abi.CAV_GAME_OVER_AGENT0
    else:
        c.game_over_mode = abi.CAV_GAME_OVER_LEARNING_AGENTS
    c.train_single_agent = int(bool(cfg.TRAIN_SINGLE_AGENT))
    c.rvo_time_horizon = float(cfg.RVO_TIME_HORIZON)
    c.rvo_collab_coeff = float(cfg.RVO_COLLAB_COEFF)
    c.rvo_neighbor_dist = float(cfg.SENSING_HORIZON)
    c.rvo_max_delta_heading = math.pi / 6
    c.rvo_radius_scale = 1 + 5e-2
    c.rvo_max_neighbors = int(cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT)
    c.use_static_map = int(bool(cfg.USE_STATIC_MAP))
    c.laserscan_length = int(cfg.LASERSCAN_LENGTH)
    c.laserscan_num_past = int(cfg.LASERSCAN_NUM_PAST)
    c.laser_min_angle, c.laser_max_angle = -math.pi / 2, math.pi / 2
    c.laser_min_range, c.laser_max_range, c.laser_range_resolution = 0.0, 6.0, 0.1
    c.map_x_width, c.map_y_width, c.map_grid_cell_size = 16.0, 16.0, 0.1
    off, total = obs_layout(cfg)
    c.obs_len = total
    for i in range(abi.CAV_OBS_NUM_STATES):
        c.obs_offset[i] = -1
    for s, o in off.items():
        c.obs_offset[abi.OBS_STATE_IDS[s]] = o
    sh = bool(cfg.STORE_HISTORY) if store_history is None else bool(store_history)
    c.store_history = int(sh and history_len > 0)
    c.history_len = int(history_len)
    c.track_aux = int(bool(track_aux))
    c.obs_f32 = int(bool(obs_f32) and c.precision == abi.CAV_F64)
    c.rvo_anti_collab_t = float(getattr(cfg, "RVO_ANTI_COLLAB_T", 1.0))
    c.rvo_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return c
