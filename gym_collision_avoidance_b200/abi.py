"""ctypes mirror of ``include/cavoid.h`` (the C ABI of ``libcavoid_sm100a.so``) and the library loader.

The product path has NO CPU fallback: :func:`load_library` raises if the CUDA library has not been built
(``python __graft_entry__.py build`` / ``python -m gym_collision_avoidance_b200.build``).
"""
import ctypes
import os

CAV_ABI_VERSION = 2
CAV_GA3C_BLOB_FLOATS = 170784

# cav_status_t
CAV_OK = 0
CAV_ERR_BAD_CONFIG = -1
STATUS_NAMES = {0: "CAV_OK", -1: "CAV_ERR_BAD_CONFIG", -2: "CAV_ERR_NULL_POINTER", -3: "CAV_ERR_MISALIGNED",
                -4: "CAV_ERR_NOT_BOUND", -5: "CAV_ERR_CUDA", -6: "CAV_ERR_BAD_HANDLE", -7: "CAV_ERR_UNSUPPORTED"}

CAV_F32, CAV_F64 = 0, 1

CAV_POLICY_RVO, CAV_POLICY_NONCOOP, CAV_POLICY_STATIC, CAV_POLICY_EXTERNAL, CAV_POLICY_LEARNING, CAV_POLICY_GA3C = range(6)
CAV_DYN_UNICYCLE, CAV_DYN_UNICYCLE_MAX_TURN_RATE, CAV_DYN_EXTERNAL = range(3)


def make_kind(policy, dynamics):
    return int(policy) | (int(dynamics) << 8)


CAV_FLAG_AT_GOAL = 0x01
CAV_FLAG_WAS_AT_GOAL_ALREADY = 0x02
CAV_FLAG_IN_COLLISION = 0x04
CAV_FLAG_WAS_IN_COLLISION_ALREADY = 0x08
CAV_FLAG_RAN_OUT_OF_TIME = 0x10
CAV_FLAG_DONE = 0x20
CAV_FLAG_COLLIDED_AGENT = 0x40
CAV_FLAG_COLLIDED_WALL = 0x80
CAV_FLAG_RVO_NON_COOP = 0x100

CAV_SORT_CLOSEST_FIRST, CAV_SORT_CLOSEST_LAST, CAV_SORT_TIME_TO_IMPACT = range(3)
CAV_GAME_OVER_ALL_DONE, CAV_GAME_OVER_AGENT0, CAV_GAME_OVER_LEARNING_AGENTS = range(3)

(CAV_OBS_IS_LEARNING, CAV_OBS_NUM_OTHER_AGENTS, CAV_OBS_DIST_TO_GOAL, CAV_OBS_HEADING_EGO_FRAME, CAV_OBS_PREF_SPEED,
 CAV_OBS_RADIUS, CAV_OBS_OTHER_AGENTS_STATES, CAV_OBS_OTHER_AGENT_STATES, CAV_OBS_LASERSCAN) = range(9)
CAV_OBS_NUM_STATES = 9
OBS_STATE_IDS = {
    "is_learning": CAV_OBS_IS_LEARNING, "num_other_agents": CAV_OBS_NUM_OTHER_AGENTS,
    "dist_to_goal": CAV_OBS_DIST_TO_GOAL, "heading_ego_frame": CAV_OBS_HEADING_EGO_FRAME,
    "pref_speed": CAV_OBS_PREF_SPEED, "radius": CAV_OBS_RADIUS, "other_agents_states": CAV_OBS_OTHER_AGENTS_STATES,
    "other_agent_states": CAV_OBS_OTHER_AGENT_STATES, "laserscan": CAV_OBS_LASERSCAN,
}

(CAV_STAT_EPISODES, CAV_STAT_EPISODES_COLLISION, CAV_STAT_EPISODES_ALL_AT_GOAL, CAV_STAT_EPISODES_ANY_STUCK,
 CAV_STAT_AGENTS, CAV_STAT_AGENTS_AT_GOAL, CAV_STAT_AGENTS_COLLISION, CAV_STAT_AGENTS_STUCK, CAV_STAT_AGENT_STEPS,
 CAV_STAT_SUM_TIME_TO_GOAL, CAV_STAT_SUM_EXTRA_TIME) = range(11)
CAV_STAT_NUM = 16
STAT_NAMES = ["episodes", "episodes_collision", "episodes_all_at_goal", "episodes_any_stuck", "agents",
              "agents_at_goal", "agents_collision", "agents_stuck", "agent_steps", "sum_time_to_goal",
              "sum_extra_time_to_goal"]
STAT_IS_DOUBLE = (CAV_STAT_SUM_TIME_TO_GOAL, CAV_STAT_SUM_EXTRA_TIME)

CAV_HISTORY_ROW = 14
CAV_BANK_ROW = 8
CAV_MAX_AGENTS = 64
CAV_KD_MAX_LEAF = 10  # RVO2 KdTree.h MAX_LEAF_SIZE: more agents than this split the agent kd-tree

_d, _i = ctypes.c_double, ctypes.c_int32
_vp = ctypes.c_void_p


class CavConfig(ctypes.Structure):
    _fields_ = [
        ("struct_size", _i), ("precision", _i), ("num_envs", _i), ("num_agents", _i),
        ("dt", _d), ("near_goal_threshold", _d), ("max_time_ratio", _d), ("max_heading_change", _d),
        ("max_speed", _d),
        ("reward_at_goal", _d), ("reward_collision_with_agent", _d), ("reward_collision_with_wall", _d),
        ("reward_time_step", _d), ("reward_wiggly_behavior", _d), ("wiggly_behavior_threshold", _d),
        ("getting_close_range", _d),
        ("sensing_horizon", _d), ("max_num_other_agents_observed", _i), ("agent_sorting_method", _i),
        ("game_over_mode", _i), ("train_single_agent", _i),
        ("rvo_time_horizon", _d), ("rvo_collab_coeff", _d), ("rvo_neighbor_dist", _d),
        ("rvo_max_delta_heading", _d), ("rvo_radius_scale", _d), ("rvo_max_neighbors", _i),
        ("use_static_map", _i), ("laserscan_length", _i), ("laserscan_num_past", _i),
        ("laser_min_angle", _d), ("laser_max_angle", _d),
        ("laser_min_range", _d), ("laser_max_range", _d), ("laser_range_resolution", _d),
        ("map_x_width", _d), ("map_y_width", _d), ("map_grid_cell_size", _d),
        ("obs_len", _i), ("obs_offset", _i * CAV_OBS_NUM_STATES),
        ("store_history", _i), ("history_len", _i), ("track_aux", _i), ("obs_f32", _i),
        ("rvo_anti_collab_t", _d), ("rvo_seed", ctypes.c_uint64),
    ]


class CavBuffers(ctypes.Structure):
    _fields_ = [(n, _vp) for n in (
        "pv", "goal", "ht", "ego", "flags", "kinds", "step_num", "num_active", "actions", "reward", "min_dist",
        "done", "game_over", "obs", "laser_count", "static_map_bits", "history", "aux0", "past_actions", "past_vel",
        "episode", "stats", "kd_order")]


BUFFER_FIELDS = [n for n, _ in CavBuffers._fields_]


class CavScenarioBank(ctypes.Structure):
    _fields_ = [("rows", _vp), ("kinds", _vp), ("num_active", _vp), ("num_scenarios", _i), ("reserved", _i)]


class CavScenarioGen(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("side_length", ctypes.c_double * 2), ("speed_bnds", ctypes.c_double * 2),
                ("radius_bnds", ctypes.c_double * 2), ("random_heading", ctypes.c_int32), ("reserved", ctypes.c_int32)]


# every symbol include/cavoid.h declares
EXPORTED_SYMBOLS = [
    "cav_abi_version", "cav_create", "cav_real_bytes", "cav_workspace_bytes", "cav_bind", "cav_reset", "cav_step",
    "cav_k_policy", "cav_k_dynamics", "cav_k_pairwise_obs", "cav_k_laserscan", "cav_k_observe",
    "cav_generate_scenarios", "cav_last_launch_count", "cav_env_schedule", "cav_bind_ga3c", "cav_k_ga3c",
    "cav_ga3c_forward",
    "cav_last_error", "cav_destroy",
]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libcavoid_sm100a.so")
_LIB = None


class CavoidError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen ``libcavoid_sm100a.so`` and declare prototypes.  Raises if it is not built (no CPU fallback)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise CavoidError(
            "CUDA library %s not found: build it with `python __graft_entry__.py build`. "
            "There is no CPU fallback for the step path." % p)
    L = ctypes.CDLL(p)
    h = _vp
    L.cav_abi_version.restype = ctypes.c_int
    L.cav_abi_version.argtypes = []
    L.cav_create.restype = ctypes.c_int
    L.cav_create.argtypes = [ctypes.POINTER(CavConfig), ctypes.POINTER(h)]
    L.cav_real_bytes.restype = ctypes.c_size_t
    L.cav_real_bytes.argtypes = [h]
    L.cav_workspace_bytes.restype = ctypes.c_size_t
    L.cav_workspace_bytes.argtypes = [h]
    L.cav_bind.restype = ctypes.c_int
    L.cav_bind.argtypes = [h, ctypes.POINTER(CavBuffers)]
    L.cav_reset.restype = ctypes.c_int
    L.cav_reset.argtypes = [h, _vp, ctypes.POINTER(CavScenarioBank), _vp, ctypes.c_int, _vp]
    L.cav_step.restype = ctypes.c_int
    L.cav_step.argtypes = [h, _vp, ctypes.c_double, _vp]
    L.cav_k_policy.restype = ctypes.c_int
    L.cav_k_policy.argtypes = [h, _vp, _vp]
    L.cav_k_dynamics.restype = ctypes.c_int
    L.cav_k_dynamics.argtypes = [h, ctypes.c_double, _vp]
    L.cav_k_pairwise_obs.restype = ctypes.c_int
    L.cav_k_pairwise_obs.argtypes = [h, _vp]
    L.cav_k_laserscan.restype = ctypes.c_int
    L.cav_k_laserscan.argtypes = [h, _vp]
    L.cav_k_observe.restype = ctypes.c_int
    L.cav_k_observe.argtypes = [h, _vp, _vp]
    L.cav_last_launch_count.restype = ctypes.c_int
    L.cav_generate_scenarios.restype = ctypes.c_int
    L.cav_generate_scenarios.argtypes = [h, ctypes.POINTER(CavScenarioGen), _vp, _vp, ctypes.c_int32, _vp]
    L.cav_last_launch_count.argtypes = [h]
    L.cav_bind_ga3c.restype = ctypes.c_int
    L.cav_bind_ga3c.argtypes = [h, ctypes.c_void_p]
    L.cav_k_ga3c.restype = ctypes.c_int
    L.cav_k_ga3c.argtypes = [h, ctypes.c_void_p, ctypes.c_void_p]
    L.cav_ga3c_forward.restype = ctypes.c_int
    L.cav_ga3c_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]
    L.cav_env_schedule.restype = ctypes.c_int
    L.cav_env_schedule.argtypes = [h, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.cav_last_error.restype = ctypes.c_char_p
    L.cav_last_error.argtypes = [h]
    L.cav_destroy.restype = ctypes.c_int
    L.cav_destroy.argtypes = [h]
    if L.cav_abi_version() != CAV_ABI_VERSION:
        raise CavoidError("ABI version mismatch: library %d, python %d" % (L.cav_abi_version(), CAV_ABI_VERSION))
    if path is None:
        _LIB = L
    return L
