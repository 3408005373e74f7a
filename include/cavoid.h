/*
 * cavoid.h -- C ABI of libcavoid_sm100a.so, the B200-native batched step path for
 * multi-agent collision-avoidance environments.
 *
 * This is the drop-in boundary for ONE hot path of mit-acl/gym-collision-avoidance: the per-step body
 * of CollisionAvoidanceEnv.step (reference gym_collision_avoidance/envs/collision_avoidance_env.py:156-234)
 * and reset (:236-282), batched over [num_envs x num_agents].  Every entry point below names the reference
 * interface it replaces.  The only native boundary in the reference on this path is the Cython binding
 * envs/policies/Python-RVO2/src/rvo2.pyx:80-230 (PyRVOSimulator: addAgent / setAgent* / doStep /
 * getAgentPosition); cav_k_policy subsumes it.  INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs.  No torch / C++ types cross this boundary.
 *   - every function returns 0 (CAV_OK) or a negative cav_status_t; nothing throws across the ABI;
 *     cav_last_error(h) returns a human-readable message for the last failure on that handle.
 *   - all device pointers are BORROWED (e.g. torch tensor.data_ptr()); the library never frees them.
 *     Vector columns (pv / goal / ht / ego / aux*) and obs must be 16-byte aligned.  The library itself allocates
 *     two small scratch arrays with cudaMalloc (cav_workspace_bytes): the environment schedule of the policy + move
 *     kernel at cav_create and the live-agent list of the GA3C-CADRL network at cav_bind_ga3c; cav_destroy frees them.
 *   - all work is ordered on the caller's CUDA stream; no implicit synchronisation, no host allocation and
 *     no host<->device copy on the step path, so a step can be captured into a CUDA graph.  (cav_step may run ranges of
 *     the environments and the environment-ordering kernel on library-owned helper streams that are forked from the
 *     caller's stream with an event and joined back into it before cav_step returns to the stream's next operation:
 *     invisible to the caller, graph capture records it as forks / joins; CAV_STEP_SPLIT=0 in the environment disables
 *     the ranges.)
 *   - agents whose policy id is CAV_POLICY_GA3C need a network (cav_bind_ga3c) before the first cav_step; without one
 *     they keep their previous action (the host classes refuse to build such an environment).
 *   - a handle is bound to the CUDA device that was current at cav_create; use it from one thread at a time.
 *     Multi-GPU = one handle per device / rank (environments are independent; no collective per step).
 *   - `real` below means float when cfg.precision == CAV_F32 and double when CAV_F64.  CAV_F64 reproduces the
 *     reference's float64 NumPy state arithmetic; CAV_F32 is the bandwidth-lean layout (float4 columns).
 *     The ORCA/RVO2 solve is float32 in both (the reference's C++ solver is float32, Vector2.h:241-242).
 *
 * SoA layout: every per-agent column is indexed [e * num_agents + i] (env-major), so one environment's agent
 * block is contiguous in every column and a tile of consecutive environments is one contiguous byte range.
 */
#ifndef CAVOID_H_
#define CAVOID_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAV_ABI_VERSION 2

typedef enum {
    CAV_OK = 0,
    CAV_ERR_BAD_CONFIG = -1,   /* inconsistent / unsupported cav_config_t */
    CAV_ERR_NULL_POINTER = -2, /* required pointer is NULL */
    CAV_ERR_MISALIGNED = -3,   /* vector column not 16-byte aligned */
    CAV_ERR_NOT_BOUND = -4,    /* cav_bind has not been called */
    CAV_ERR_CUDA = -5,         /* CUDA runtime / launch failure (see cav_last_error) */
    CAV_ERR_BAD_HANDLE = -6,
    CAV_ERR_UNSUPPORTED = -7   /* plugin kind with no device implementation (no CPU fallback exists) */
} cav_status_t;

typedef enum { CAV_F32 = 0, CAV_F64 = 1 } cav_precision_t;

/* Policy plugin ids (reference envs/test_cases.py:68-85 policy_dict; envs/policies/*.py). */
typedef enum {
    CAV_POLICY_RVO = 0,      /* RVOPolicy.py:50-122 + RVO2 Agent.cpp:287-630 */
    CAV_POLICY_NONCOOP = 1,  /* NonCooperativePolicy.py:21 */
    CAV_POLICY_STATIC = 2,   /* StaticPolicy.py:21-22 (goal := pos) */
    CAV_POLICY_EXTERNAL = 3, /* ExternalPolicy.py:14-16 (passthrough of the caller's action) */
    CAV_POLICY_LEARNING = 4, /* LearningPolicy.py:30-32 (scaled external action, is_still_learning) */
    CAV_POLICY_GA3C = 5      /* GA3CCADRLPolicy.py:49-84: LSTM network on the agent's own observation (cav_bind_ga3c) */
} cav_policy_t;

/* Dynamics plugin ids (reference envs/test_cases.py:93-96 dynamics_dict; envs/dynamics/*.py). */
typedef enum {
    CAV_DYN_UNICYCLE = 0,               /* UnicycleDynamics.py:27-47 */
    CAV_DYN_UNICYCLE_MAX_TURN_RATE = 1, /* UnicycleDynamicsMaxTurnRate.py:30-43 */
    CAV_DYN_EXTERNAL = 2                /* ExternalDynamics.py:10-13 (no-op) */
} cav_dynamics_t;

/* kinds column: policy | dynamics << 8 */
#define CAV_KIND_POLICY(k) ((k) & 0xffu)
#define CAV_KIND_DYNAMICS(k) (((k) >> 8) & 0xffu)
#define CAV_MAKE_KIND(policy, dynamics) (((uint32_t)(policy)) | (((uint32_t)(dynamics)) << 8))

/* flags column bits (reference Agent attributes, envs/agent.py:109-113,138). */
#define CAV_FLAG_AT_GOAL 0x01u                  /* is_at_goal */
#define CAV_FLAG_WAS_AT_GOAL_ALREADY 0x02u      /* was_at_goal_already */
#define CAV_FLAG_IN_COLLISION 0x04u             /* in_collision */
#define CAV_FLAG_WAS_IN_COLLISION_ALREADY 0x08u /* was_in_collision_already */
#define CAV_FLAG_RAN_OUT_OF_TIME 0x10u          /* ran_out_of_time */
#define CAV_FLAG_DONE 0x20u                     /* is_done (set by the done check, env.py:514-553) */
#define CAV_FLAG_COLLIDED_AGENT 0x40u           /* this step's collision_with_agent[i] (env.py:490-493) */
#define CAV_FLAG_COLLIDED_WALL 0x80u            /* this step's collision_with_wall[i] (env.py:494-506) */
#define CAV_FLAG_RVO_NON_COOP 0x100u            /* RVOPolicy.use_non_coop_policy (RVOPolicy.py:32,77-88; anti-collaborative RVO) */

/* AGENT_SORTING_METHOD (OtherAgentsStatesSensor.py:33-52). */
typedef enum { CAV_SORT_CLOSEST_FIRST = 0, CAV_SORT_CLOSEST_LAST = 1, CAV_SORT_TIME_TO_IMPACT = 2 } cav_sorting_t;

/* game_over rule (env.py:538-551). */
typedef enum {
    CAV_GAME_OVER_ALL_DONE = 0,       /* Config.EVALUATE_MODE */
    CAV_GAME_OVER_AGENT0 = 1,         /* Config.TRAIN_SINGLE_AGENT */
    CAV_GAME_OVER_LEARNING_AGENTS = 2 /* all agents whose policy.is_still_learning */
} cav_game_over_t;

/* Observation states (Config.STATE_INFO_DICT keys, envs/config.py:93-170).  cfg.obs_offset[s] is the offset of
 * state s inside one agent's packed observation vector (Config.STATES_IN_OBS order, exactly the packing of
 * MultiagentDictToMultiagentArrayWrapper, envs/wrappers.py:149-173), or -1 when s is not observed. */
typedef enum {
    CAV_OBS_IS_LEARNING = 0,
    CAV_OBS_NUM_OTHER_AGENTS = 1,
    CAV_OBS_DIST_TO_GOAL = 2,
    CAV_OBS_HEADING_EGO_FRAME = 3,
    CAV_OBS_PREF_SPEED = 4,
    CAV_OBS_RADIUS = 5,
    CAV_OBS_OTHER_AGENTS_STATES = 6, /* (K,7) rows */
    CAV_OBS_OTHER_AGENT_STATES = 7,  /* (7,) = closest row, Agent.other_agent_states */
    CAV_OBS_LASERSCAN = 8,           /* (num_past, length) */
    CAV_OBS_NUM_STATES = 9
} cav_obs_state_t;

/* Per-episode statistics accumulated on device when an environment is reset after game_over
 * (metric definitions: reference experiments/src/env_utils.py:54-87).  uint64/double pairs in `stats`. */
typedef enum {
    CAV_STAT_EPISODES = 0,          /* episodes finished */
    CAV_STAT_EPISODES_COLLISION = 1, /* episodes where any agent ended in collision */
    CAV_STAT_EPISODES_ALL_AT_GOAL = 2,
    CAV_STAT_EPISODES_ANY_STUCK = 3, /* any agent ran out of time */
    CAV_STAT_AGENTS = 4,            /* agent-episodes finished */
    CAV_STAT_AGENTS_AT_GOAL = 5,
    CAV_STAT_AGENTS_COLLISION = 6,
    CAV_STAT_AGENTS_STUCK = 7,
    CAV_STAT_AGENT_STEPS = 8,       /* sum of Agent.step_num */
    CAV_STAT_SUM_TIME_TO_GOAL = 9,  /* double: sum of Agent.t over agents that reached their goal */
    CAV_STAT_SUM_EXTRA_TIME = 10,   /* double: sum of (t - straight_line_time_to_reach_goal), same agents */
    CAV_STAT_NUM = 16
} cav_stat_t;

/* History row layouts (Agent.to_vector, envs/agent.py:275-289): 11 global + 3 ego values. */
#define CAV_HISTORY_ROW 14

typedef struct cav_config {
    int32_t struct_size; /* sizeof(cav_config_t), checked by cav_create */
    int32_t precision;   /* cav_precision_t */
    int32_t num_envs;    /* E */
    int32_t num_agents;  /* N: agent slots per env, 1..64 (Config.MAX_NUM_AGENTS_IN_ENVIRONMENT) */

    /* simulation (envs/config.py:43-47) */
    double dt;                  /* Config.DT */
    double near_goal_threshold; /* Config.NEAR_GOAL_THRESHOLD */
    double max_time_ratio;      /* Config.MAX_TIME_RATIO (reset only) */
    double max_heading_change;  /* env.max_heading_change = pi/3 (env.py:88), LearningPolicy scaling */
    double max_speed;           /* env.max_speed = 1.0 (env.py:91); stored, unused on the path like the reference */

    /* rewards (envs/config.py:30-39; env.py:577-599) */
    double reward_at_goal;
    double reward_collision_with_agent;
    double reward_collision_with_wall;
    double reward_time_step;
    double reward_wiggly_behavior;
    double wiggly_behavior_threshold;
    double getting_close_range;

    /* sensing (envs/config.py:64-79,172-174) */
    double sensing_horizon;               /* Config.SENSING_HORIZON (inf by default) */
    int32_t max_num_other_agents_observed; /* K = Config.MAX_NUM_OTHER_AGENTS_OBSERVED */
    int32_t agent_sorting_method;          /* cav_sorting_t */
    int32_t game_over_mode;                /* cav_game_over_t */
    int32_t train_single_agent;            /* reserved (reward vector is always per agent) */

    /* RVO (RVOPolicy.py:13-28; envs/config.py:83-86) */
    double rvo_time_horizon;      /* Config.RVO_TIME_HORIZON */
    double rvo_collab_coeff;      /* Config.RVO_COLLAB_COEFF; < 0 = anti-collaborative RVO (RVOPolicy.py:77-88, Agent.cpp:368-415) */
    double rvo_neighbor_dist;     /* neighborDist = Config.SENSING_HORIZON */
    double rvo_max_delta_heading; /* pi/6 */
    double rvo_radius_scale;      /* 1 + 5e-2 */
    int32_t rvo_max_neighbors;    /* maxNeighbors = Config.MAX_NUM_AGENTS_IN_ENVIRONMENT */

    /* laserscan + map (LaserScanSensor.py:24-47; env.py:389-392; Map.py:6-24) */
    int32_t use_static_map;     /* Config.USE_STATIC_MAP */
    int32_t laserscan_length;   /* Config.LASERSCAN_LENGTH (beams) */
    int32_t laserscan_num_past; /* Config.LASERSCAN_NUM_PAST */
    double laser_min_angle, laser_max_angle;               /* -pi/2, +pi/2 */
    double laser_min_range, laser_max_range, laser_range_resolution; /* 0, 6, 0.1 */
    double map_x_width, map_y_width, map_grid_cell_size;   /* 16, 16, 0.1 */

    /* observation packing */
    int32_t obs_len;                        /* floats per agent */
    int32_t obs_offset[CAV_OBS_NUM_STATES]; /* -1 = state not in Config.STATES_IN_OBS */

    /* storage */
    int32_t store_history; /* Config.STORE_HISTORY: write one CAV_HISTORY_ROW row per live agent-step */
    int32_t history_len;   /* T: rows per agent in `history` (rows >= T are dropped) */
    int32_t track_aux;     /* keep speed / delta_heading / vel_ego / past_actions / past_global_velocities */
    int32_t obs_f32;       /* precision == CAV_F64 only: `obs` and `reward` are float32 columns -- the dtype the reference
                            * DECLARES for its observation space (np.float32, envs/config.py:93-170) -- written directly by
                            * the kernels from the float64 state (half the bytes to HBM and to the host); 0 = `real` */
    /* anti-collaborative RVO (rvo_collab_coeff < 0): every rvo_anti_collab_t seconds of agent time the agent re-draws
     * whether it ignores the others (probability 1 - |coeff|) or steers INTO its closest neighbour (RVOPolicy.py:77-88).
     * The reference draws from NumPy's global stream; here the draw is Philox4x32-10 keyed by (rvo_seed; env, agent,
     * episode, step_num): same distribution, reproducible, independent of launch geometry. */
    double rvo_anti_collab_t; /* Config.RVO_ANTI_COLLAB_T */
    uint64_t rvo_seed;
} cav_config_t;

/* Device buffers (all borrowed).  "real4" = 4 consecutive reals, 16-byte aligned. */
typedef struct cav_buffers {
    void *pv;   /* real4 [E*N]: pos_global_frame.xy, vel_global_frame.xy                    (agent.py:72-75) */
    void *goal; /* real4 [E*N]: goal_global_frame.xy, radius, pref_speed                    (agent.py:74,92-95) */
    void *ht;   /* real4 [E*N]: heading_global_frame, time_remaining_to_reach_goal, t, turning_dir */
    void *ego;  /* real4 [E*N]: dist_to_goal, heading_ego_frame, ref_prll.xy                (Dynamics.py:31-41) */
    uint32_t *flags;     /* [E*N] CAV_FLAG_* */
    uint32_t *kinds;     /* [E*N] CAV_MAKE_KIND(policy, dynamics) */
    int32_t *step_num;   /* [E*N] Agent.step_num */
    int32_t *num_active; /* [E] number of agents present in env e (<= N), or NULL = N everywhere */
    float *actions;      /* [E*N*2] float32 actions chosen this step (env.py:305-307 all_actions) */
    void *reward;        /* real [E*N] (env.py:394-456); float32 when cfg.obs_f32 */
    void *min_dist;      /* real [E*N] dist_btwn_nearest_agent (env.py:482-487); may be NULL */
    uint8_t *done;       /* [E*N] which_agents_done (env.py:529-536) */
    uint8_t *game_over;  /* [E] (env.py:538-551) */
    void *obs;           /* real [E*N*obs_len] packed observation, STATES_IN_OBS order; 16-byte aligned; float32 when
                          * cfg.obs_f32 */
    uint32_t *laser_count;          /* [E*N] LaserScanSensor.num_measurements_made; NULL when no laserscan */
    const uint32_t *static_map_bits; /* [rows * ceil(cols/32)] row-major bitmap of Map.static_map, or NULL (empty) */
    void *history;       /* real [E*N*T*CAV_HISTORY_ROW] or NULL (agent.py:257-260,275-289) */
    void *aux0;          /* real4 [E*N]: speed_global_frame, delta_heading_global_frame, vel_ego_frame.xy; or NULL */
    void *past_actions;  /* real4 [E*N]: past_actions[0,:], past_actions[1,:]; or NULL */
    void *past_vel;      /* real4 [E*N]: past_global_velocities[0,:], [1,:]; or NULL */
    int32_t *episode;    /* [E] episodes started per env (reset bookkeeping) */
    uint64_t *stats;     /* [CAV_STAT_NUM] per-handle episode statistics (doubles bit-cast for the SUM_* slots) */
    uint8_t *kd_order;   /* [E*N] RVO2 KdTree::agents_ (KdTree.cpp:47-111): the simulators' persistent agent permutation,
                          * re-partitioned in place by every step's buildAgentTree and reset to identity by cav_reset;
                          * it fixes the neighbour order at exactly equal float distances (Agent.cpp:424-448).  Required
                          * when num_agents > 10 (= MAX_LEAF_SIZE: below that the tree is one leaf and the order is the
                          * agent index); may be NULL otherwise. */
} cav_buffers_t;

/* Scenario bank for reset: S scenarios x N agent rows.  Row = 8 doubles
 * [start_x, start_y, goal_x, goal_y, pref_speed, radius, initial_heading, unused]: the reference's "cadrl test
 * case" row (envs/test_cases.py:508, [px,py,gx,gy,pref_speed,radius]) plus the heading computed by
 * cadrl_test_case_to_agents (test_cases.py:553-559). */
#define CAV_BANK_ROW 8
typedef struct cav_scenario_bank {
    const double *rows;        /* [S*N*CAV_BANK_ROW] device */
    const uint32_t *kinds;     /* [S*N] device, or NULL = keep the env's current kinds */
    const int32_t *num_active; /* [S] device, or NULL = N */
    int32_t num_scenarios;     /* S */
    int32_t reserved;
} cav_scenario_bank_t;

typedef struct cav_handle_s *cav_handle_t;
typedef void *cav_stream_t; /* cudaStream_t */

/* Library / ABI version.  (No reference equivalent.) */
int cav_abi_version(void);

/* Replaces CollisionAvoidanceEnv.__init__ (env.py:52-154) + RVOPolicy.__init__ (RVOPolicy.py:10-32,
 * i.e. rvo2.PyRVOSimulator.__cinit__, rvo2.pyx:83-90): validates the config, creates a handle on the
 * current CUDA device. */
int cav_create(const cav_config_t *cfg, cav_handle_t *out);

/* Bytes of `real` for this handle (4 or 8). */
size_t cav_real_bytes(cav_handle_t h);

/* Device scratch the library holds beyond cav_buffers_t (SURVEY.md §8b lists the entry point): the kernels work out of
 * registers and their own shared memory; the one exception is the environment schedule of the policy + move kernel --
 * 6 bytes per environment (an int32 order and a uint16 cost), allocated by cav_create with cudaMalloc on the current
 * device when num_envs >= 2048 and freed by cav_destroy (CAV_ENV_ORDER=0 in the environment disables it) -- and, once
 * cav_bind_ga3c has been called, the list of live GA3C-CADRL agents (4 bytes per agent slot).  Returns those bytes, 0 when
 * there are none. */
size_t cav_workspace_bytes(cav_handle_t h);

/* Diagnostics: copies the environment schedule into caller-owned DEVICE buffers on `stream` (either may be NULL):
 * order[q] = environment processed at position q of the policy + move kernel (int32 [num_envs]), cost[e] = scheduling key
 * environment e earned in its last step (uint16 [num_envs]).  Returns CAV_ERR_BAD_CONFIG when the schedule is disabled.
 * The schedule only decides WHICH warp computes an environment -- results do not depend on it. */
int cav_env_schedule(cav_handle_t h, int32_t *order, uint16_t *cost, cav_stream_t stream);

/* Attaches caller-owned device columns (replaces the per-Agent Python attribute storage, agent.py:59-138). */
int cav_bind(cav_handle_t h, const cav_buffers_t *bufs);

/* Replaces CollisionAvoidanceEnv.reset (env.py:236-282) -> _init_agents (:345-367) -> Agent.reset
 * (agent.py:59-138) -> _get_obs (:555-575) for every env e with env_mask[e] != 0 (NULL mask = all envs).
 * scenario index for env e: scenario_idx[e] if given, else (e + episode[e] * E) mod S.  When `accumulate_stats`
 * is non-zero the finished episode's statistics are added to `stats` before the state is overwritten
 * (DummyVecEnv.step_wait auto-reset, envs/vec_env.py:120-128 + env_utils.run_episode, env_utils.py:54-87). */
int cav_reset(cav_handle_t h, const uint8_t *env_mask, const cav_scenario_bank_t *bank, const int32_t *scenario_idx,
              int accumulate_stats, cav_stream_t stream);

/* Replaces CollisionAvoidanceEnv.step (env.py:156-234): policy -> move -> collisions/rewards -> observation
 * -> done.  ext_actions: [E*N*2] float32 device (read only for EXTERNAL / LEARNING agents), may be NULL when
 * no agent has an external policy.  dt <= 0 means cfg.dt. */
int cav_step(cav_handle_t h, const float *ext_actions, double dt, cav_stream_t stream);

/* Per-phase entry points (unit parity against the oracle); cav_step == policy+move fused, then pairwise_obs,
 * then laserscan when observed.  SURVEY.md §8b also sketches a `cav_k_done`: there is none, by design --
 * _check_which_agents_done (env.py:514-553) needs exactly the flags the collision / reward pass has just computed, so it
 * is the tail of cav_k_pairwise_obs (done + game_over written there) instead of one more launch and one more read of flags.
 *   cav_k_policy      : env.py:304-323 action selection (+ rvo2 doStep)        -> bufs.actions
 *   cav_k_dynamics    : env.py:326-327 Agent.take_action (agent.py:192-241) reading bufs.actions
 *   cav_k_pairwise_obs: env.py:394-512 collisions+rewards, OtherAgentsStatesSensor.sense
 *                       (OtherAgentsStatesSensor.py:58-144), get_observation_dict (agent.py:323-327),
 *                       _check_which_agents_done (env.py:514-553)
 *   cav_k_laserscan   : Map.add_agents_to_map (Map.py:46-50) + LaserScanSensor.sense (LaserScanSensor.py:49-101)
 *   cav_k_observe     : env.py:555-575 _get_obs alone (sensors + packing, no rewards / flags), for the envs selected
 *                       by env_mask (NULL = all): what reset() returns for externally injected agent states */
int cav_k_policy(cav_handle_t h, const float *ext_actions, cav_stream_t stream);
int cav_k_dynamics(cav_handle_t h, double dt, cav_stream_t stream);
int cav_k_pairwise_obs(cav_handle_t h, cav_stream_t stream);
int cav_k_laserscan(cav_handle_t h, cav_stream_t stream);
int cav_k_observe(cav_handle_t h, const uint8_t *env_mask, cav_stream_t stream);

/* Device-side scenario generation: fills `rows` ([num_scenarios * N * CAV_BANK_ROW] doubles, device) with what the
 * reference's get_testcase_random (envs/test_cases.py:212-253) -> generate_rand_test_case_multi
 * (policies/CADRL/scripts/multi/gen_rand_testcases.py:111-142: 15 % swap :322-377, 15 % circle :380-423, 70 %
 * rejection-sampled random cases :144-233 incl. the straight-line test :425-444) would build, one scenario per
 * thread, same draw sequence per scenario but from a Philox4x32-10 stream keyed by (seed, scenario): the rows follow
 * the reference's DISTRIBUTION (statistical parity with the bit-exact host generator, tests/test_gpu_scenarios.py),
 * not NumPy's MT19937 stream.  case_kind ([num_scenarios] device, may be NULL) receives 0 = swap, 1 = circle,
 * 2 = random.  The result is a valid cav_scenario_bank_t.rows for cav_reset. */
typedef struct cav_scenario_gen {
    uint64_t seed;
    double side_length[2]; /* side ~ U(lo, hi): Config.TEST_CASE_ARGS['side_length'] (config.py:56-59) */
    double speed_bnds[2];  /* config.py:54 */
    double radius_bnds[2]; /* config.py:55 */
    int32_t random_heading; /* 0: initial heading toward the goal (Config.EVALUATE_MODE, test_cases.py:553-557);
                             * 1: heading ~ U(-pi, pi) per agent (training mode, :558-559) */
    int32_t reserved;
} cav_scenario_gen_t;
int cav_generate_scenarios(cav_handle_t h, const cav_scenario_gen_t *gen, double *rows, int32_t *case_kind,
                           int32_t num_scenarios, cav_stream_t stream);

/* ---- GA3C-CADRL policy network (reference policies/GA3CCADRLPolicy.py:49-84, policies/GA3C_CADRL/network.py:37-74) ------
 * The reference restores a TF1 meta-graph + checkpoint and calls sess.run('Softmax:0', {X: vec_obs}).  The graph (read
 * from network_01900000.meta; identical in the three shipped checkpoints) is:
 *   X [rows, 138] = [num_other_agents, dist_to_goal, heading_ego_frame, pref_speed, radius, 19 x 7 other-agent states]
 *   Xn = (X - avg) / std                                       (Const, Const_1: sub, div)
 *   LSTMCell(64, forget_bias 1) over Xn[:, 5:] as 19 steps of 7, sequence_length = int32(X[:, 0]), zero initial state:
 *       [i j f o] = [x_t, h] @ Wl + bl;  c = sigmoid(f + 1) * c + sigmoid(i) * tanh(j);  h = sigmoid(o) * tanh(c)
 *   a1 = relu([Xn[:, 1:5], h] @ W1 + b1); a2 = relu(a1 @ W2 + b2); a3 = relu(a2 @ W3 + b3); p = softmax(a3 @ Wo + bo)
 * and the policy takes actions[argmax p] of the 11-entry table (network.py:7-16), speed scaled by the observed pref_speed.
 * All float32.  The weights travel as ONE float32 blob in device memory (host memory for the oracle), sections in this
 * order, each in TF's row-major [in][out] layout except the LSTM kernel, whose 256 columns are regrouped per hidden unit:
 *   avg[138] std[138] Wl[71][64][4 = i,j,f,o] bl[64][4] W1[68][256] b1[256] W2[256][256] b2[256] W3[256][256] b3[256]
 *   Wo[256][11] bo[11] pad[1]                                     = CAV_GA3C_BLOB_FLOATS floats
 * (gym_collision_avoidance_b200/envs/policies/GA3C_CADRL/network.py reads the reference's checkpoint files into it). */
#define CAV_GA3C_X_LEN 138
#define CAV_GA3C_MAX_OTHERS 19
#define CAV_GA3C_HIDDEN 64
#define CAV_GA3C_FC 256
#define CAV_GA3C_ACTIONS 11
#define CAV_GA3C_BLOB_FLOATS 170784

/* Attaches the network to a handle: agents whose policy id is CAV_POLICY_GA3C then get their action from it inside
 * cav_step / cav_k_policy.  X is gathered from the agent's packed observation of the previous step (what the reference
 * passes to find_next_action, env.py:313-323): the states of Config.STATES_IN_OBS in order without is_learning,
 * cropped / zero-padded to 138 (network.py:24-35 crop_x).  `blob` is borrowed device memory (float32, 16-byte aligned). */
int cav_bind_ga3c(cav_handle_t h, const float *blob);

/* Phase entry point: the network for every live CAV_POLICY_GA3C agent -> bufs.actions ([pref_speed * a0, a1] as float32,
 * GA3CCADRLPolicy.py:80-84); cav_step and cav_k_policy run it themselves when a network is bound.  probs (nullable):
 * float32 [E*N*11] device, receives the softmax of the agents it ran for (other rows untouched). */
int cav_k_ga3c(cav_handle_t h, float *probs, cav_stream_t stream);

/* The network alone on a batch (replaces NetworkVPCore.predict_p, network.py:37-41): X float32 [rows][138] device ->
 * probs float32 [rows][11] (nullable) and action_index int32 [rows] (nullable, argmax: first maximum). */
int cav_ga3c_forward(const float *blob, const float *X, int32_t rows, float *probs, int32_t *action_index,
                     cav_stream_t stream);

/* Number of kernels the last cav_step / cav_reset call on this handle enqueued (for launch accounting). */
int cav_last_launch_count(cav_handle_t h);

const char *cav_last_error(cav_handle_t h);
int cav_destroy(cav_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* CAVOID_H_ */
