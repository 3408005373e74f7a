// cav_abi.cu -- the C ABI of libcavoid_sm100a.so (include/cavoid.h): handle management, config validation and
// flattening, and the entry points that enqueue the sm_100a kernels on the caller's stream.  Nothing here
// allocates device memory, copies host<->device or synchronises: every buffer is borrowed from the caller.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "cav_kernels.h"
#include "cav_kdtree.cuh"

using namespace cav;

struct cav_handle_s {
    uint32_t magic;
    cav_config_t cfg;
    cav_buffers_t bufs;
    bool bound;
    int device;
    int sm_count;
    KParams<float> pf;
    KParams<double> pd;
    ObsPlan plan;
    LaserParams lp;
    ResetParams rp;
    bool has_laser;
    bool use_tpa;  // one-thread-per-agent kernels (small N) instead of the sub-warp-group kernels
    bool use_v2_policy;  // k_policy_move_v2 (cav_policy.cu, thread per agent, N <= 32) instead of the sub-warp-group kernel
    bool use_tpa_n;  // ... and, where instantiated, the compile-time-N pairwise+obs kernel (CAV_TPA_N=0 disables)
    // cav_step in 2..4 env ranges (see step_split): helper streams forked off / joined back into the caller's stream
    cudaStream_t aux[3];
    cudaEvent_t ev_fork, ev_join[3];
    int split_parts;   // number of env ranges (0 = no split)
    int split_at[5];   // range k = [split_at[k], split_at[k + 1])
    // environment scheduling of k_policy_move_v2 (cav_policy.cu, EnvSchedule): the only device memory the library owns
    const float *ga3c_blob;  // GA3C-CADRL weights (cav_bind_ga3c), borrowed
    int32_t *ga3c_list;      // scratch of k_ga3c_compact: E*N + 1 ints, owned (allocated by cav_bind_ga3c)
    EnvSchedule sched;
    cudaStream_t sort_stream;
    cudaEvent_t ev_policy[4], ev_sorted;
    int launches;
    char err[512];
};

#define CAV_MAGIC 0xCA70D001u

static thread_local char g_err[512] = "";

static int fail(cav_handle_t h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) {
        strncpy(h->err, buf, sizeof(h->err) - 1);
        h->err[sizeof(h->err) - 1] = 0;
    }
    strncpy(g_err, buf, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
    return code;
}

static bool valid(cav_handle_t h) { return h && h->magic == CAV_MAGIC; }

template <typename T>
static void fill_params(const cav_config_t &c, KParams<T> &p) {
    p.E = c.num_envs;
    p.N = c.num_agents;
    p.dt = (T)c.dt;
    p.near_goal_thr = (T)c.near_goal_threshold;
    p.max_heading_change = (T)c.max_heading_change;
    p.reward_at_goal = (T)c.reward_at_goal;
    p.reward_coll_agent = (T)c.reward_collision_with_agent;
    p.reward_coll_wall = (T)c.reward_collision_with_wall;
    p.reward_time_step = (T)c.reward_time_step;
    p.reward_wiggly = (T)c.reward_wiggly_behavior;
    p.wiggly_thr = (T)c.wiggly_behavior_threshold;
    p.getting_close = (T)c.getting_close_range;
    // env.py:589-599: clip bounds = min / max of the possible reward values
    const T vals[5] = {p.reward_at_goal, p.reward_coll_agent, p.reward_time_step, p.reward_coll_wall, p.reward_wiggly};
    p.reward_min = vals[0];
    p.reward_max = vals[0];
    for (int k = 1; k < 5; ++k) {
        if (vals[k] < p.reward_min) p.reward_min = vals[k];
        if (vals[k] > p.reward_max) p.reward_max = vals[k];
    }
    p.sensing_horizon = (T)c.sensing_horizon;
    p.K = c.max_num_other_agents_observed;
    p.sorting = c.agent_sorting_method;
    p.game_over_mode = c.game_over_mode;
    p.rvo_inv_time_horizon_num = (float)c.rvo_time_horizon;
    p.rvo_time_step = (float)c.dt;  // PyRVOSimulator(timeStep=Config.DT), RVOPolicy.py:13,25
    p.rvo_collab = (float)c.rvo_collab_coeff;
    p.rvo_neighbor_dist = (float)c.rvo_neighbor_dist;
    p.rvo_max_delta_heading = (T)c.rvo_max_delta_heading;
    p.rvo_radius_scale = (T)c.rvo_radius_scale;
    p.rvo_inv_dt = (T)1 / (T)c.dt;  // 1/self.dt, RVOPolicy.py:105
    p.rvo_max_neighbors = c.rvo_max_neighbors;
    p.rvo_anti_collab_t = (T)c.rvo_anti_collab_t;
    p.rvo_anti_p = (float)fabs(c.rvo_collab_coeff);
    p.rvo_seed = c.rvo_seed;
    p.store_history = c.store_history;
    p.history_len = c.history_len;
    p.obs_len = c.obs_len;
}

static int state_len(const cav_config_t &c, int s) {
    switch (s) {
        case CAV_OBS_OTHER_AGENTS_STATES: return 7 * c.max_num_other_agents_observed;
        case CAV_OBS_OTHER_AGENT_STATES: return 7;
        case CAV_OBS_LASERSCAN: return c.laserscan_num_past * c.laserscan_length;
        default: return 1;
    }
}

static int build_plan(const cav_config_t &c, ObsPlan &pl) {
    memset(&pl, 0, sizeof(pl));
    for (int s = 0; s < CAV_OBS_NUM_STATES; ++s) pl.c_state[s] = -1;
    pl.oas7_goff = c.obs_offset[CAV_OBS_OTHER_AGENT_STATES];
    // states staged through shared memory by the pairwise+obs kernel, in ascending global offset
    int order[CAV_OBS_NUM_STATES], no = 0;
    for (int s = 0; s < CAV_OBS_NUM_STATES; ++s) {
        if (c.obs_offset[s] < 0) continue;
        if (c.obs_offset[s] + state_len(c, s) > c.obs_len) return -1;
        if (s == CAV_OBS_OTHER_AGENT_STATES || s == CAV_OBS_LASERSCAN) continue;
        order[no++] = s;
    }
    for (int i = 1; i < no; ++i)
        for (int j = i; j > 0 && c.obs_offset[order[j]] < c.obs_offset[order[j - 1]]; --j) {
            const int t = order[j];
            order[j] = order[j - 1];
            order[j - 1] = t;
        }
    int coff = 0;
    for (int i = 0; i < no; ++i) {
        const int s = order[i], go = c.obs_offset[s], ln = state_len(c, s);
        if (i > 0 && go < c.obs_offset[order[i - 1]] + state_len(c, order[i - 1])) return -1;  // overlap
        pl.c_state[s] = coff;
        if (pl.nseg > 0 && pl.goff[pl.nseg - 1] + pl.len[pl.nseg - 1] == go) {
            pl.len[pl.nseg - 1] += ln;
        } else {
            if (pl.nseg >= CAV_MAX_OBS_SEG) return -1;
            pl.goff[pl.nseg] = go;
            pl.coff[pl.nseg] = coff;
            pl.len[pl.nseg] = ln;
            pl.nseg++;
        }
        coff += ln;
    }
    pl.compact_len = coff > 0 ? coff : 1;
    return 0;
}

extern "C" {

int cav_abi_version(void) { return CAV_ABI_VERSION; }

// the environment ranges the schedule is sorted in: cav_step's launch ranges (one range without the split)
static void sched_ranges(cav_handle_t h, int *at, int *n) {
    if (h->split_parts > 1) {
        *n = h->split_parts;
        for (int k = 0; k <= h->split_parts; ++k) at[k] = h->split_at[k];
    } else {
        *n = 1;
        at[0] = 0;
        at[1] = h->cfg.num_envs;
    }
}

int cav_create(const cav_config_t *cfg, cav_handle_t *out) {
    if (!cfg || !out) return fail(nullptr, CAV_ERR_NULL_POINTER, "cav_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(cav_config_t))
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: struct_size %d != %zu", cfg->struct_size,
                    sizeof(cav_config_t));
    if (cfg->precision != CAV_F32 && cfg->precision != CAV_F64)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: bad precision %d", cfg->precision);
    if (cfg->num_envs < 1) return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: num_envs %d < 1", cfg->num_envs);
    if (cfg->num_agents < 1 || cfg->num_agents > CAV_MAX_AGENTS)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: num_agents %d outside [1, %d]", cfg->num_agents,
                    CAV_MAX_AGENTS);
    if (!(cfg->dt > 0.0)) return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: dt must be > 0");
    if (cfg->max_num_other_agents_observed < 0 || cfg->obs_len < 1)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: bad observation sizes");
    if (cfg->agent_sorting_method != CAV_SORT_CLOSEST_FIRST && cfg->agent_sorting_method != CAV_SORT_CLOSEST_LAST &&
        cfg->agent_sorting_method != CAV_SORT_TIME_TO_IMPACT)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: unknown agent_sorting_method %d", cfg->agent_sorting_method);
    if (cfg->game_over_mode < 0 || cfg->game_over_mode > 2)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: bad game_over_mode");
    if (cfg->rvo_collab_coeff < 0.0 && !(cfg->rvo_anti_collab_t > 0.0))
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: anti-collaborative RVO needs rvo_anti_collab_t > 0");
    if (cfg->store_history && cfg->history_len < 1)
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: store_history needs history_len >= 1");
    cav_handle_s *h = new (std::nothrow) cav_handle_s;
    if (!h) return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->magic = CAV_MAGIC;
    h->cfg = *cfg;
    if (build_plan(*cfg, h->plan) != 0) {
        delete h;
        return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: observation offsets overlap or exceed obs_len");
    }
    h->has_laser = cfg->obs_offset[CAV_OBS_LASERSCAN] >= 0;
    if (h->has_laser) {
        if (!cfg->use_static_map) {
            delete h;
            return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: laserscan needs use_static_map "
                                                      "(LaserScanSensor.py:25-27)");
        }
        LaserParams &lp = h->lp;
        lp.E = cfg->num_envs;
        lp.N = cfg->num_agents;
        lp.B = cfg->laserscan_length;
        lp.P = cfg->laserscan_num_past;
        lp.obs_len = cfg->obs_len;
        lp.obs_off = cfg->obs_offset[CAV_OBS_LASERSCAN];
        lp.cell = cfg->map_grid_cell_size;
        lp.inv_cell = 1.0 / lp.cell;
        lp.rows = (int)(cfg->map_x_width / lp.cell);  // Map.py:13
        lp.cols = (int)(cfg->map_y_width / lp.cell);
        lp.wpr = (lp.cols + 31) / 32;
        lp.origin0 = (cfg->map_x_width / 2.) / lp.cell;  // Map.py:23
        lp.origin1 = (cfg->map_y_width / 2.) / lp.cell;
        lp.min_angle = cfg->laser_min_angle;
        lp.max_angle = cfg->laser_max_angle;
        lp.ang_step = lp.B > 1 ? (lp.max_angle - lp.min_angle) / (double)(lp.B - 1) : 0.0;
        lp.min_range = cfg->laser_min_range;
        lp.max_range = cfg->laser_max_range;
        lp.range_res = cfg->laser_range_resolution;
        lp.nr = (int)ceil((lp.max_range - lp.min_range) / lp.range_res);  // len(np.arange(min, max, res))
        if (lp.B < 1 || lp.P < 1 || lp.rows < 1 || lp.cols < 1 || lp.nr < 1 || (size_t)lp.rows * lp.wpr * 4 > 200 * 1024) {
            delete h;
            return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_create: bad laserscan / map geometry");
        }
    }
    h->rp.E = cfg->num_envs;
    h->rp.N = cfg->num_agents;
    h->rp.obs_len = cfg->obs_len;
    h->rp.dt = cfg->dt;
    h->rp.near_goal_threshold = cfg->near_goal_threshold;
    h->rp.max_time_ratio = cfg->max_time_ratio;
    fill_params<float>(*cfg, h->pf);
    fill_params<double>(*cfg, h->pd);
    {
        // mapping: thread-per-agent up to CAV_TPA_MAX_N agents (default 24, measured crossover),
        // sub-warp groups with warp-shuffle bitonic sorts above.  CAV_TPA_MAX_N=0 forces the group kernels.
        int tpa_max = 24;
        if (const char *ev = getenv("CAV_TPA_MAX_N")) tpa_max = atoi(ev);
        if (tpa_max > 25) tpa_max = 25;  // the sorted-neighbour list is packed 5 bits x 24 in two registers
        h->use_tpa = cfg->num_agents <= tpa_max &&
                     tpa_smem_bytes(cfg->num_agents, cfg->precision == CAV_F64, h->plan,
                                    cfg->max_num_other_agents_observed) <= 200 * 1024;
        // compile-time-N kernels wherever they apply (16384 x 10: float32 24 us vs 38 us generic, float64 39 us vs 50 us)
        h->use_tpa_n = h->use_tpa;
        if (const char *ev = getenv("CAV_TPA_N")) h->use_tpa_n = h->use_tpa && atoi(ev) != 0;
        // policy + move: thread per agent up to 32 agents (k_policy_move_v2), sub-warp groups above (and under
        // CAV_TPA_MAX_N=0, which the tests use to run the group kernels at every N)
        h->use_v2_policy = cfg->num_agents >= 2 && cfg->num_agents <= 32 && tpa_max > 0;
        if (const char *ev = getenv("CAV_POLICY_GROUP")) h->use_v2_policy = h->use_v2_policy && atoi(ev) == 0;
    }
    cudaError_t ce = cudaGetDevice(&h->device);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (ce != cudaSuccess) {
        const int rc = fail(nullptr, CAV_ERR_CUDA, "cav_create: %s", cudaGetErrorString(ce));
        delete h;
        return rc;
    }
    // Two halves: policy + move is 1.5 waves of CTAs and pairwise + obs 1.8 at 16384 x 10, so each kernel ends with the
    // GPU half empty for a quarter of its run time (profiles/r02e_ncu_f64_16384x10.txt: waves_per_multiprocessor).
    // Environments are independent, so cav_step enqueues policy(A) -> pairwise(A) on the caller's stream and
    // policy(B) -> pairwise(B) on a helper stream forked from it with an event and joined back with another: pairwise(A)
    // fills the SMs that policy(B)'s tail leaves idle.  Caller-visible semantics are unchanged (everything is ordered
    // after the caller's earlier work and before its later work; graph capture records the fork / join as edges).
    // CAV_STEP_SPLIT=0 disables it, 2..4 picks the number of ranges (default 2).
    h->ga3c_blob = nullptr;
    h->ga3c_list = nullptr;
    h->split_parts = 0;
    for (int k = 0; k < 3; ++k) h->aux[k] = nullptr, h->ev_join[k] = nullptr;
    h->ev_fork = nullptr;
    {
        bool want = cfg->num_envs >= 4096 && h->use_v2_policy && h->use_tpa_n && !h->has_laser &&
                    cfg->agent_sorting_method != CAV_SORT_TIME_TO_IMPACT;
        int parts = 2;
        if (const char *ev = getenv("CAV_STEP_SPLIT")) parts = atoi(ev);
        if (parts > 4) parts = 4;
        want = want && parts >= 2;
        if (want) {
            int pw_tile = 0;
            const WallMap nomap{};
            const cudaError_t q = (cfg->precision == CAV_F64)
                ? (cfg->obs_f32 ? run_pairwise_obs_tpa_n<double, float>(h->bufs, h->pd, h->plan, nomap, nullptr, 0, h->sm_count, nullptr, 0, -1, &pw_tile)
                                : run_pairwise_obs_tpa_n<double, double>(h->bufs, h->pd, h->plan, nomap, nullptr, 0, h->sm_count, nullptr, 0, -1, &pw_tile))
                : run_pairwise_obs_tpa_n<float, float>(h->bufs, h->pf, h->plan, nomap, nullptr, 0, h->sm_count, nullptr, 0, -1, &pw_tile);
            const int pm_tile = policy_move_v2_tile_envs(cfg->num_agents);
            if (q == cudaSuccess && pw_tile > 0 && pm_tile > 0) {
                int a = pw_tile, b = pm_tile;
                while (b) {
                    const int t = a % b;
                    a = b;
                    b = t;
                }
                const int unit = pw_tile / a * pm_tile;  // lcm: a boundary both kernels' tiles respect
                bool ok = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
                h->split_at[0] = 0;
                for (int k = 1; k < parts && ok; ++k) {
                    h->split_at[k] = (int)((long long)cfg->num_envs * k / parts) / unit * unit;
                    ok = h->split_at[k] > h->split_at[k - 1] &&
                         cudaStreamCreateWithFlags(&h->aux[k - 1], cudaStreamNonBlocking) == cudaSuccess &&
                         cudaEventCreateWithFlags(&h->ev_join[k - 1], cudaEventDisableTiming) == cudaSuccess;
                }
                h->split_at[parts] = cfg->num_envs;
                if (ok) h->split_parts = parts;
            }
        }
    }
    // Environment schedule (see cav_policy.cu): 6 bytes per environment of library-owned scratch, identity order to start
    // with.  CAV_ENV_ORDER=0 disables it (the kernels then process environment q at position q).
    h->sort_stream = nullptr;
    h->ev_sorted = nullptr;
    for (int k = 0; k < 4; ++k) h->ev_policy[k] = nullptr;
    {
        bool want = h->use_v2_policy && cfg->num_envs >= 2048;
        if (const char *ev = getenv("CAV_ENV_ORDER")) want = want && atoi(ev) != 0;
        if (want) {
            int *order = nullptr;
            unsigned short *cost = nullptr;
            const size_t E = (size_t)cfg->num_envs;
            bool ok = cudaMalloc(&order, E * sizeof(int)) == cudaSuccess && cudaMalloc(&cost, E * sizeof(unsigned short)) == cudaSuccess &&
                      cudaMemset(cost, 0, E * sizeof(unsigned short)) == cudaSuccess;
            if (ok) {
                h->sched.order = order;
                h->sched.cost = cost;
                // the same ranges cav_step launches: order[range] must always be a permutation OF THAT RANGE, or one
                // range's policy kernel would move environments the other range's pairwise kernel is reading
                int at[5], nr;
                sched_ranges(h, at, &nr);
                ok = run_env_order(h->sched, at, nr, nullptr) == cudaSuccess && cudaStreamSynchronize(nullptr) == cudaSuccess;
            }
            if (ok && h->split_parts > 1) {
                ok = cudaStreamCreateWithFlags(&h->sort_stream, cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&h->ev_sorted, cudaEventDisableTiming) == cudaSuccess;
                for (int k = 0; k < h->split_parts && ok; ++k)
                    ok = cudaEventCreateWithFlags(&h->ev_policy[k], cudaEventDisableTiming) == cudaSuccess;
            }
            if (!ok) {
                cudaGetLastError();
                if (order) cudaFree(order);
                if (cost) cudaFree(cost);
                h->sched = EnvSchedule{};
            }
        }
    }
    *out = h;
    return CAV_OK;
}

size_t cav_real_bytes(cav_handle_t h) { return valid(h) ? (h->cfg.precision == CAV_F64 ? 8 : 4) : 0; }

size_t cav_workspace_bytes(cav_handle_t h) {
    if (!valid(h)) return 0;
    size_t b = h->sched.order ? (size_t)h->cfg.num_envs * (sizeof(int) + sizeof(unsigned short)) : 0;
    if (h->ga3c_list) b += ((size_t)h->cfg.num_envs * h->cfg.num_agents + 1) * sizeof(int32_t);
    return b;
}

int cav_env_schedule(cav_handle_t h, int32_t *order, uint16_t *cost, cav_stream_t stream) {
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, "cav_env_schedule: bad handle");
    if (!h->sched.order) return fail(h, CAV_ERR_BAD_CONFIG, "cav_env_schedule: no environment schedule on this handle");
    const size_t E = (size_t)h->cfg.num_envs;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaSuccess;
    if (order) ce = cudaMemcpyAsync(order, h->sched.order, E * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess && cost) ce = cudaMemcpyAsync(cost, h->sched.cost, E * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) return fail(h, CAV_ERR_CUDA, "cav_env_schedule: %s", cudaGetErrorString(ce));
    return CAV_OK;
}

int cav_bind(cav_handle_t h, const cav_buffers_t *bufs) {
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, "cav_bind: bad handle");
    if (!bufs) return fail(h, CAV_ERR_NULL_POINTER, "cav_bind: null buffers");
    const struct {
        const void *p;
        const char *name;
        bool vec;
    } req[] = {{bufs->pv, "pv", true},         {bufs->goal, "goal", true},       {bufs->ht, "ht", true},
               {bufs->ego, "ego", true},       {bufs->flags, "flags", false},    {bufs->kinds, "kinds", false},
               {bufs->step_num, "step_num", false}, {bufs->actions, "actions", false},
               {bufs->reward, "reward", false}, {bufs->done, "done", false},     {bufs->game_over, "game_over", false},
               {bufs->obs, "obs", false}};
    for (const auto &r : req) {
        if (!r.p) return fail(h, CAV_ERR_NULL_POINTER, "cav_bind: %s is NULL", r.name);
        if (r.vec && ((uintptr_t)r.p & 15u)) return fail(h, CAV_ERR_MISALIGNED, "cav_bind: %s not 16-byte aligned", r.name);
    }
    if (((uintptr_t)bufs->actions & 7u)) return fail(h, CAV_ERR_MISALIGNED, "cav_bind: actions not 8-byte aligned");
    // the compile-time-N kernels stream the observation out with 16-byte stores phased against the column's base
    if (((uintptr_t)bufs->obs & 15u)) return fail(h, CAV_ERR_MISALIGNED, "cav_bind: obs not 16-byte aligned");
    const void *optvec[] = {bufs->aux0, bufs->past_actions, bufs->past_vel};
    for (const void *p : optvec)
        if (p && ((uintptr_t)p & 15u)) return fail(h, CAV_ERR_MISALIGNED, "cav_bind: aux column not 16-byte aligned");
    if (h->has_laser && !bufs->laser_count) return fail(h, CAV_ERR_NULL_POINTER, "cav_bind: laser_count is NULL");
    if (h->cfg.num_agents > CAV_KD_MAX_LEAF && !bufs->kd_order)
        return fail(h, CAV_ERR_NULL_POINTER, "cav_bind: kd_order is NULL (required when num_agents > %d)", CAV_KD_MAX_LEAF);
    if (h->cfg.store_history && !bufs->history) return fail(h, CAV_ERR_NULL_POINTER, "cav_bind: history is NULL");
    h->bufs = *bufs;
    h->bound = true;
    return CAV_OK;
}

#define CAV_CHECK_READY(h, name)                                                    \
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, name ": bad handle");   \
    if (!(h)->bound) return fail(h, CAV_ERR_NOT_BOUND, name ": cav_bind has not been called")

// CAV_DEBUG_SYNC=1 (debugging only; breaks graph capture): name every launch on stderr and wait for it
static bool debug_sync() {
    static const bool on = [] {
        const char *e = getenv("CAV_DEBUG_SYNC");
        return e && atoi(e) != 0;
    }();
    return on;
}
static cudaError_t debug_after(const char *name, cudaError_t ce) {
    if (!debug_sync() || ce != cudaSuccess) return ce;
    fprintf(stderr, "[cavoid] %s launched ... ", name);
    fflush(stderr);
    ce = cudaDeviceSynchronize();
    fprintf(stderr, "%s\n", ce == cudaSuccess ? "done" : cudaGetErrorString(ce));
    return ce;
}

#define CAV_CUDA(h, name, expr)                                                                \
    do {                                                                                       \
        cudaError_t ce_ = debug_after(name, (expr));                                           \
        if (ce_ != cudaSuccess) return fail(h, CAV_ERR_CUDA, name ": %s", cudaGetErrorString(ce_)); \
        (h)->launches++;                                                                       \
    } while (0)

// the network for the CAV_POLICY_GA3C agents -> bufs.actions, ahead of the kernels that consume the actions
static int do_ga3c(cav_handle_t h, float *probs, cudaStream_t st) {
    CAV_CUDA(h, "ga3c_policy", run_ga3c_policy(h->bufs, h->cfg, h->ga3c_blob, probs, h->ga3c_list, st));
    return CAV_OK;
}

static int do_policy_move(cav_handle_t h, const float *ext, int phases, double dt, cudaStream_t st) {
    if (h->ga3c_blob && (phases & 1)) {
        const int rc = do_ga3c(h, nullptr, st);
        if (rc) return rc;
    }
    if (h->use_v2_policy) {
        const cudaError_t ce = debug_after(
            "policy_move_v2", (h->cfg.precision == CAV_F64)
                                  ? run_policy_move_v2<double>(h->bufs, h->pd, ext, phases, dt, h->sm_count, st, 0, -1, h->sched)
                                  : run_policy_move_v2<float>(h->bufs, h->pf, ext, phases, dt, h->sm_count, st, 0, -1, h->sched));
        if (ce == cudaSuccess) {
            h->launches++;
            return CAV_OK;
        }
        if (ce != cudaErrorNotSupported) return fail(h, CAV_ERR_CUDA, "policy_move_v2: %s", cudaGetErrorString(ce));
        h->use_v2_policy = false;
    }
    if (h->cfg.precision == CAV_F64)
        CAV_CUDA(h, "policy_move", run_policy_move<double>(h->bufs, h->pd, ext, phases, dt, h->sm_count, st));
    else
        CAV_CUDA(h, "policy_move", run_policy_move<float>(h->bufs, h->pf, ext, phases, dt, h->sm_count, st));
    return CAV_OK;
}
static WallMap wall_map(cav_handle_t h) {
    WallMap wm;
    memset(&wm, 0, sizeof(wm));
    const cav_config_t &c = h->cfg;
    if (c.use_static_map && h->bufs.static_map_bits && c.map_grid_cell_size > 0.0) {
        wm.bits = h->bufs.static_map_bits;
        wm.cell = c.map_grid_cell_size;
        wm.inv_cell = 1.0 / wm.cell;
        wm.rows = (int)(c.map_x_width / wm.cell);
        wm.cols = (int)(c.map_y_width / wm.cell);
        wm.wpr = (wm.cols + 31) / 32;
        wm.origin0 = (c.map_x_width / 2.) / wm.cell;
        wm.origin1 = (c.map_y_width / 2.) / wm.cell;
    }
    return wm;
}

// (state precision, observation precision) of a handle: float/float, double/double, or double/float (cfg.obs_f32)
#define CAV_BY_PRECISION(h, CALL_F, CALL_D, CALL_DF) \
    ((h)->cfg.precision != CAV_F64 ? (CALL_F) : ((h)->cfg.obs_f32 ? (CALL_DF) : (CALL_D)))

static int do_pairwise_obs(cav_handle_t h, const uint8_t *mask, int obs_only, cudaStream_t st) {
    const WallMap wm = wall_map(h);
    // 'time_to_impact' ordering (three-level real key) lives in the sub-warp-group kernel only, for every N
    if (h->use_tpa && h->cfg.agent_sorting_method != CAV_SORT_TIME_TO_IMPACT) {
        if (h->use_tpa_n) {  // compile-time-N specialisation when the configuration allows it
            const cudaError_t ce = debug_after("pairwise_obs_tpa_n", CAV_BY_PRECISION(
                h, (run_pairwise_obs_tpa_n<float, float>(h->bufs, h->pf, h->plan, wm, mask, obs_only, h->sm_count, st)),
                (run_pairwise_obs_tpa_n<double, double>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st)),
                (run_pairwise_obs_tpa_n<double, float>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st))));
            if (ce == cudaSuccess) {
                h->launches++;
                return CAV_OK;
            }
            if (ce != cudaErrorNotSupported) return fail(h, CAV_ERR_CUDA, "pairwise_obs_tpa_n: %s", cudaGetErrorString(ce));
            h->use_tpa_n = false;  // not applicable to this handle: stop asking
        }
        CAV_CUDA(h, "pairwise_obs_tpa",
                 CAV_BY_PRECISION(
                     h, (run_pairwise_obs_tpa<float, float>(h->bufs, h->pf, h->plan, wm, mask, obs_only, h->sm_count, st)),
                     (run_pairwise_obs_tpa<double, double>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st)),
                     (run_pairwise_obs_tpa<double, float>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st))));
        return CAV_OK;
    }
    CAV_CUDA(h, "pairwise_obs",
             CAV_BY_PRECISION(h, (run_pairwise_obs<float, float>(h->bufs, h->pf, h->plan, wm, mask, obs_only, h->sm_count, st)),
                              (run_pairwise_obs<double, double>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st)),
                              (run_pairwise_obs<double, float>(h->bufs, h->pd, h->plan, wm, mask, obs_only, h->sm_count, st))));
    return CAV_OK;
}
static int do_laser(cav_handle_t h, const uint8_t *mask, cudaStream_t st) {
    CAV_CUDA(h, "laserscan",
             CAV_BY_PRECISION(h, (run_laserscan<float, float>(h->bufs, h->lp, mask, h->sm_count, st)),
                              (run_laserscan<double, double>(h->bufs, h->lp, mask, h->sm_count, st)),
                              (run_laserscan<double, float>(h->bufs, h->lp, mask, h->sm_count, st))));
    return CAV_OK;
}

// cav_step as two independent halves of the environments on two streams (see cav_create)
static int step_split(cav_handle_t h, const float *ext, double dt, cudaStream_t st) {
    const int P = h->split_parts;
    const WallMap wm = wall_map(h);
    auto policy = [&](cudaStream_t s, int e0, int e1) {
        return (h->cfg.precision == CAV_F64) ? run_policy_move_v2<double>(h->bufs, h->pd, ext, 3, dt, h->sm_count, s, e0, e1, h->sched)
                                             : run_policy_move_v2<float>(h->bufs, h->pf, ext, 3, dt, h->sm_count, s, e0, e1, h->sched);
    };
    auto pairwise = [&](cudaStream_t s, int e0, int e1) {
        return CAV_BY_PRECISION(
            h, (run_pairwise_obs_tpa_n<float, float>(h->bufs, h->pf, h->plan, wm, nullptr, 0, h->sm_count, s, e0, e1, nullptr)),
            (run_pairwise_obs_tpa_n<double, double>(h->bufs, h->pd, h->plan, wm, nullptr, 0, h->sm_count, s, e0, e1, nullptr)),
            (run_pairwise_obs_tpa_n<double, float>(h->bufs, h->pd, h->plan, wm, nullptr, 0, h->sm_count, s, e0, e1, nullptr)));
    };
    // range 0 runs on the caller's stream, range k > 0 on helper stream k - 1
    auto stream_of = [&](int k) { return k == 0 ? st : h->aux[k - 1]; };
    if (h->ga3c_blob) {
        const int rc = do_ga3c(h, nullptr, st);
        if (rc) return rc;
    }
    CAV_CUDA(h, "step fork", cudaEventRecord(h->ev_fork, st));
    h->launches--;
    for (int k = 1; k < P; ++k) {
        CAV_CUDA(h, "step fork", cudaStreamWaitEvent(h->aux[k - 1], h->ev_fork, 0));
        h->launches--;
    }
    for (int k = P - 1; k >= 0; --k) CAV_CUDA(h, "policy_move_v2 (env range)", policy(stream_of(k), h->split_at[k], h->split_at[k + 1]));
    if (h->sched.order) {  // next step's schedule, beside the pairwise kernels
        for (int k = 0; k < P; ++k) {
            CAV_CUDA(h, "env order fork", cudaEventRecord(h->ev_policy[k], stream_of(k)));
            h->launches--;
            CAV_CUDA(h, "env order fork", cudaStreamWaitEvent(h->sort_stream, h->ev_policy[k], 0));
            h->launches--;
        }
        CAV_CUDA(h, "env_order", run_env_order(h->sched, h->split_at, P, h->sort_stream));
        CAV_CUDA(h, "env order join", cudaEventRecord(h->ev_sorted, h->sort_stream));
        h->launches--;
    }
    for (int k = P - 1; k >= 0; --k) CAV_CUDA(h, "pairwise_obs_tpa_n (env range)", pairwise(stream_of(k), h->split_at[k], h->split_at[k + 1]));
    if (h->sched.order) {
        CAV_CUDA(h, "env order join", cudaStreamWaitEvent(st, h->ev_sorted, 0));
        h->launches--;
    }
    for (int k = 1; k < P; ++k) {
        CAV_CUDA(h, "step join", cudaEventRecord(h->ev_join[k - 1], h->aux[k - 1]));
        h->launches--;
        CAV_CUDA(h, "step join", cudaStreamWaitEvent(st, h->ev_join[k - 1], 0));
        h->launches--;
    }
    return CAV_OK;
}

int cav_reset(cav_handle_t h, const uint8_t *env_mask, const cav_scenario_bank_t *bank, const int32_t *scenario_idx,
              int accumulate_stats, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_reset");
    if (!bank || !bank->rows) return fail(h, CAV_ERR_NULL_POINTER, "cav_reset: scenario bank is NULL");
    if (bank->num_scenarios < 1) return fail(h, CAV_ERR_BAD_CONFIG, "cav_reset: empty scenario bank");
    cudaStream_t st = (cudaStream_t)stream;
    h->launches = 0;
    CAV_CUDA(h, "reset",
             CAV_BY_PRECISION(h, (run_reset<float, float>(h->bufs, h->rp, env_mask, *bank, scenario_idx, accumulate_stats, st)),
                              (run_reset<double, double>(h->bufs, h->rp, env_mask, *bank, scenario_idx, accumulate_stats, st)),
                              (run_reset<double, float>(h->bufs, h->rp, env_mask, *bank, scenario_idx, accumulate_stats, st))));
    int rc = do_pairwise_obs(h, env_mask, 1, st);  // reset returns self._get_obs() (env.py:282)
    if (rc) return rc;
    if (h->has_laser) rc = do_laser(h, env_mask, st);
    return rc;
}

int cav_generate_scenarios(cav_handle_t h, const cav_scenario_gen_t *gen, double *rows, int32_t *case_kind,
                           int32_t num_scenarios, cav_stream_t stream) {
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, "cav_generate_scenarios: bad handle");
    if (!gen || !rows) return fail(h, CAV_ERR_NULL_POINTER, "cav_generate_scenarios: gen / rows is NULL");
    if (num_scenarios < 1) return fail(h, CAV_ERR_BAD_CONFIG, "cav_generate_scenarios: num_scenarios < 1");
    if (!(gen->side_length[0] > 0.0) || gen->side_length[1] < gen->side_length[0] ||
        !(gen->speed_bnds[0] > 0.0) || gen->speed_bnds[1] < gen->speed_bnds[0] || !(gen->radius_bnds[0] > 0.0) ||
        gen->radius_bnds[1] < gen->radius_bnds[0])
        return fail(h, CAV_ERR_BAD_CONFIG, "cav_generate_scenarios: bounds must be positive and ordered");
    if ((reinterpret_cast<uintptr_t>(rows) & 7u) != 0) return fail(h, CAV_ERR_MISALIGNED, "cav_generate_scenarios: rows misaligned");
    h->launches = 0;
    CAV_CUDA(h, "generate_scenarios",
             run_generate_scenarios(rows, case_kind, num_scenarios, h->cfg.num_agents, gen->seed, gen->side_length[0],
                                    gen->side_length[1], gen->speed_bnds, gen->radius_bnds, 0.2 /* GETTING_CLOSE_RANGE,
                                    CADRL/scripts/multi/global_var.py:8 */, gen->random_heading, (cudaStream_t)stream));
    return CAV_OK;
}

int cav_step(cav_handle_t h, const float *ext_actions, double dt, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_step");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(dt > 0.0)) dt = h->cfg.dt;
    h->launches = 0;
    if (h->split_parts > 1 && h->use_v2_policy && h->use_tpa_n && !debug_sync()) return step_split(h, ext_actions, dt, st);
    int rc = do_policy_move(h, ext_actions, 3, dt, st);
    if (rc) return rc;
    if (h->sched.order && h->use_v2_policy) {
        int at[5], nr;
        sched_ranges(h, at, &nr);
        CAV_CUDA(h, "env_order", run_env_order(h->sched, at, nr, st));
    }
    rc = do_pairwise_obs(h, nullptr, 0, st);
    if (rc) return rc;
    if (h->has_laser) rc = do_laser(h, nullptr, st);
    return rc;
}

int cav_k_policy(cav_handle_t h, const float *ext_actions, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_policy");
    h->launches = 0;
    return do_policy_move(h, ext_actions, 1, h->cfg.dt, (cudaStream_t)stream);
}
int cav_k_dynamics(cav_handle_t h, double dt, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_dynamics");
    if (!(dt > 0.0)) dt = h->cfg.dt;
    h->launches = 0;
    return do_policy_move(h, nullptr, 2, dt, (cudaStream_t)stream);
}
int cav_k_pairwise_obs(cav_handle_t h, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_pairwise_obs");
    h->launches = 0;
    return do_pairwise_obs(h, nullptr, 0, (cudaStream_t)stream);
}
int cav_k_laserscan(cav_handle_t h, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_laserscan");
    if (!h->has_laser) return fail(h, CAV_ERR_BAD_CONFIG, "cav_k_laserscan: laserscan is not in the observation");
    h->launches = 0;
    return do_laser(h, nullptr, (cudaStream_t)stream);
}

int cav_k_observe(cav_handle_t h, const uint8_t *env_mask, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_observe");
    h->launches = 0;
    int rc = do_pairwise_obs(h, env_mask, 1, (cudaStream_t)stream);
    if (rc) return rc;
    if (h->has_laser) rc = do_laser(h, env_mask, (cudaStream_t)stream);
    return rc;
}

int cav_bind_ga3c(cav_handle_t h, const float *blob) {
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, "cav_bind_ga3c: bad handle");
    if (!blob) {
        h->ga3c_blob = nullptr;  // detach
        return CAV_OK;
    }
    if ((reinterpret_cast<uintptr_t>(blob) & 15u) != 0) return fail(h, CAV_ERR_MISALIGNED, "cav_bind_ga3c: blob must be 16-byte aligned");
    const cav_config_t &c = h->cfg;
    // GA3CCADRLPolicy.find_next_action reads obs['pref_speed'] and feeds the flattened observation: the network's first five
    // inputs must be the scalars it was trained on
    if (c.obs_offset[CAV_OBS_PREF_SPEED] < 0 || c.obs_offset[CAV_OBS_NUM_OTHER_AGENTS] < 0)
        return fail(h, CAV_ERR_BAD_CONFIG, "cav_bind_ga3c: STATES_IN_OBS must hold num_other_agents and pref_speed");
    if (!h->ga3c_list) {
        const size_t n = (size_t)c.num_envs * c.num_agents + 1;
        if (cudaMalloc(&h->ga3c_list, n * sizeof(int32_t)) != cudaSuccess) {
            h->ga3c_list = nullptr;
            return fail(h, CAV_ERR_CUDA, "cav_bind_ga3c: cudaMalloc of the agent list failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    h->ga3c_blob = blob;
    return CAV_OK;
}

int cav_k_ga3c(cav_handle_t h, float *probs, cav_stream_t stream) {
    CAV_CHECK_READY(h, "cav_k_ga3c");
    if (!h->ga3c_blob) return fail(h, CAV_ERR_NOT_BOUND, "cav_k_ga3c: no network bound (cav_bind_ga3c)");
    h->launches = 0;
    return do_ga3c(h, probs, (cudaStream_t)stream);
}

int cav_ga3c_forward(const float *blob, const float *X, int32_t rows, float *probs, int32_t *action_index,
                     cav_stream_t stream) {
    if (!blob || (!X && rows > 0)) return fail(nullptr, CAV_ERR_NULL_POINTER, "cav_ga3c_forward: blob / X is NULL");
    if (rows < 0) return fail(nullptr, CAV_ERR_BAD_CONFIG, "cav_ga3c_forward: rows < 0");
    if ((reinterpret_cast<uintptr_t>(blob) & 15u) != 0) return fail(nullptr, CAV_ERR_MISALIGNED, "cav_ga3c_forward: blob must be 16-byte aligned");
    const cudaError_t ce = run_ga3c_forward(blob, X, rows, probs, action_index, (cudaStream_t)stream);
    if (ce != cudaSuccess) return fail(nullptr, CAV_ERR_CUDA, "cav_ga3c_forward: %s", cudaGetErrorString(ce));
    return CAV_OK;
}

int cav_last_launch_count(cav_handle_t h) { return valid(h) ? h->launches : 0; }

const char *cav_last_error(cav_handle_t h) { return valid(h) ? h->err : g_err; }

int cav_destroy(cav_handle_t h) {
    if (!valid(h)) return fail(nullptr, CAV_ERR_BAD_HANDLE, "cav_destroy: bad handle");
    h->magic = 0;
    for (int k = 0; k < 3; ++k) {
        if (h->aux[k]) cudaStreamDestroy(h->aux[k]);
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->sort_stream) cudaStreamDestroy(h->sort_stream);
    if (h->ev_sorted) cudaEventDestroy(h->ev_sorted);
    for (int k = 0; k < 4; ++k)
        if (h->ev_policy[k]) cudaEventDestroy(h->ev_policy[k]);
    if (h->ga3c_list) cudaFree(h->ga3c_list);
    if (h->sched.order) cudaFree(const_cast<int *>(h->sched.order));
    if (h->sched.cost) cudaFree(h->sched.cost);
    delete h;
    return CAV_OK;
}

}  // extern "C"
