// cav_reset.cu -- masked environment reset from a device-resident scenario bank, with per-episode statistics.
//
// Replaces CollisionAvoidanceEnv.reset -> _init_agents -> Agent.reset (reference collision_avoidance_env.py:236-282,
// :345-367; agent.py:59-138; test_cases.py:495-590 scenario row -> agents) for the environments selected by a
// device mask (the DummyVecEnv auto-reset pattern, vec_env.py:120-128), and accumulates the finished episode's
// statistics (experiments/src/env_utils.py:54-87) before the state is overwritten.  One warp per environment.
#include "cav_kernels.h"

namespace cav {

struct ResetPtrs {
    void *pv, *goal, *ht, *ego;
    uint32_t *flags, *kinds;
    int32_t *step_num, *num_active;
    float *actions;
    void *reward, *min_dist;
    uint8_t *done;
    void *obs;
    uint32_t *laser_count;
    void *aux0, *past_actions, *past_vel;
    int32_t *episode;
    unsigned long long *stats;
    uint8_t *kd_order;
};

template <typename V>
__device__ __forceinline__ V warp_sum(V v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <typename T, typename OT>
__global__ void __launch_bounds__(256) k_reset(ResetPtrs b, ResetParams rp, const uint8_t *__restrict__ env_mask,
                                                 cav_scenario_bank_t bank, const int32_t *__restrict__ scenario_idx,
                                                 int accumulate_stats) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int N = rp.N;
    R4<T> *g_pv = reinterpret_cast<R4<T> *>(b.pv);
    R4<T> *g_goal = reinterpret_cast<R4<T> *>(b.goal);
    R4<T> *g_ht = reinterpret_cast<R4<T> *>(b.ht);
    R4<T> *g_ego = reinterpret_cast<R4<T> *>(b.ego);
    for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < rp.E; e += gridDim.x * warps_per_block) {
        if (env_mask && !env_mask[e]) continue;  // warp-uniform
        const size_t base = (size_t)e * N;
        // ---- statistics of the episode that just finished (env_utils.py:61-84) ----
        if (accumulate_stats && b.stats) {
            const int n_old = b.num_active ? b.num_active[e] : N;
            unsigned agents = 0, at_goal = 0, coll = 0, stuck = 0;
            unsigned long long steps = 0;
            double sum_t = 0.0, sum_extra = 0.0;
            for (int a = lane; a < n_old; a += 32) {
                const uint32_t f = b.flags[base + a];
                const R4<T> ht = ld4(g_ht + base + a);
                agents += 1;
                steps += (unsigned long long)b.step_num[base + a];
                if (f & CAV_FLAG_IN_COLLISION) coll += 1;
                if (f & CAV_FLAG_RAN_OUT_OF_TIME) stuck += 1;
                if (f & CAV_FLAG_AT_GOAL) {
                    at_goal += 1;
                    sum_t += (double)ht.z;
                    // straight_line_time_to_reach_goal = (time_remaining + t) / MAX_TIME_RATIO (agent.py:98-103)
                    sum_extra += (double)ht.z - ((double)ht.y + (double)ht.z) / rp.max_time_ratio;
                }
            }
            agents = warp_sum(agents);
            at_goal = warp_sum(at_goal);
            coll = warp_sum(coll);
            stuck = warp_sum(stuck);
            steps = warp_sum(steps);
            sum_t = warp_sum(sum_t);
            sum_extra = warp_sum(sum_extra);
            if (lane == 0) {
                atomicAdd(&b.stats[CAV_STAT_EPISODES], 1ull);
                if (coll) atomicAdd(&b.stats[CAV_STAT_EPISODES_COLLISION], 1ull);
                if (at_goal == agents) atomicAdd(&b.stats[CAV_STAT_EPISODES_ALL_AT_GOAL], 1ull);
                if (stuck) atomicAdd(&b.stats[CAV_STAT_EPISODES_ANY_STUCK], 1ull);
                atomicAdd(&b.stats[CAV_STAT_AGENTS], (unsigned long long)agents);
                atomicAdd(&b.stats[CAV_STAT_AGENTS_AT_GOAL], (unsigned long long)at_goal);
                atomicAdd(&b.stats[CAV_STAT_AGENTS_COLLISION], (unsigned long long)coll);
                atomicAdd(&b.stats[CAV_STAT_AGENTS_STUCK], (unsigned long long)stuck);
                atomicAdd(&b.stats[CAV_STAT_AGENT_STEPS], steps);
                atomicAdd(reinterpret_cast<double *>(&b.stats[CAV_STAT_SUM_TIME_TO_GOAL]), sum_t);
                atomicAdd(reinterpret_cast<double *>(&b.stats[CAV_STAT_SUM_EXTRA_TIME]), sum_extra);
            }
        }
        // ---- pick the scenario ----
        int sidx = 0;
        if (lane == 0) {
            const int ep = b.episode ? b.episode[e] : 0;
            sidx = scenario_idx ? scenario_idx[e]
                                : (int)(((long long)e + (long long)ep * rp.E) % (long long)bank.num_scenarios);
            if (b.episode) b.episode[e] = ep + 1;
        }
        sidx = __shfl_sync(0xffffffffu, sidx, 0);
        const int n = bank.num_active ? bank.num_active[sidx] : N;
        if (lane == 0 && b.num_active) b.num_active[e] = n;
        // ---- env.reset zeroes the whole observation (env.py:276-281): coalesced over the env's block ----
        {
            OT *o = reinterpret_cast<OT *>(b.obs) + base * (size_t)rp.obs_len;  // OT: observation precision in HBM
            const int total = N * rp.obs_len;
            for (int k = lane; k < total; k += 32) o[k] = (OT)0;
        }
        // ---- Agent.reset (agent.py:59-138) ----
        for (int a = lane; a < N; a += 32) {
            const size_t gi = base + a;
            const double *row = bank.rows + ((size_t)sidx * N + a) * CAV_BANK_ROW;
            if (bank.kinds) b.kinds[gi] = bank.kinds[(size_t)sidx * N + a];
            R4<T> pv{(T)0, (T)0, (T)0, (T)0}, gl{(T)0, (T)0, (T)0, (T)0}, ht{(T)0, (T)0, (T)0, (T)0},
                ego{(T)0, (T)0, (T)0, (T)0};
            if (a < n) {
                pv.x = (T)row[0];
                pv.y = (T)row[1];
                gl.x = (T)row[2];
                gl.y = (T)row[3];
                gl.w = (T)row[4];  // pref_speed
                gl.z = (T)row[5];  // radius
                ht.x = (T)row[6];  // initial heading
                const T dx = pv.x - gl.x, dy = pv.y - gl.y;
                const T slt = (M<T>::sqrt_(dx * dx + dy * dy) - (T)rp.near_goal_threshold) / gl.w;  // agent.py:98
                T tr = (T)rp.max_time_ratio * slt;
                tr = tr > (T)rp.dt ? tr : (T)rp.dt;  // agent.py:103
                ht.y = tr;
                T vex, vey;
                update_ego<T>(pv, gl, ht.x, ego, vex, vey);  // agent.py:126
            }
            st4(g_pv + gi, pv);
            st4(g_goal + gi, gl);
            st4(g_ht + gi, ht);
            st4(g_ego + gi, ego);
            b.flags[gi] = (a < n) ? CAV_FLAG_RVO_NON_COOP : 0u;  // RVOPolicy.__init__: use_non_coop_policy = True (RVOPolicy.py:32)
            b.step_num[gi] = 0;
            b.done[gi] = 0;
            reinterpret_cast<float2 *>(b.actions)[gi] = make_float2(0.f, 0.f);
            const R4<T> z{(T)0, (T)0, (T)0, (T)0};
            if (b.aux0) st4(reinterpret_cast<R4<T> *>(b.aux0) + gi, z);
            if (b.past_actions) st4(reinterpret_cast<R4<T> *>(b.past_actions) + gi, z);
            if (b.past_vel) st4(reinterpret_cast<R4<T> *>(b.past_vel) + gi, z);
            if (b.laser_count) b.laser_count[gi] = 0u;
            if (b.kd_order) b.kd_order[gi] = (uint8_t)a;  // new episode = new simulators: KdTree::agents_ is the identity
            if (b.min_dist) reinterpret_cast<T *>(b.min_dist)[gi] = (a < n) ? M<T>::inf() : (T)0;
        }
    }
}

template <typename T, typename OT>
cudaError_t run_reset(const cav_buffers_t &bufs, const ResetParams &rp, const uint8_t *env_mask,
                      const cav_scenario_bank_t &bank, const int32_t *scenario_idx, int accumulate_stats,
                      cudaStream_t st) {
    ResetPtrs p;
    p.pv = bufs.pv;
    p.goal = bufs.goal;
    p.ht = bufs.ht;
    p.ego = bufs.ego;
    p.flags = bufs.flags;
    p.kinds = bufs.kinds;
    p.step_num = bufs.step_num;
    p.num_active = bufs.num_active;
    p.actions = bufs.actions;
    p.reward = bufs.reward;
    p.min_dist = bufs.min_dist;
    p.done = bufs.done;
    p.obs = bufs.obs;
    p.laser_count = bufs.laser_count;
    p.aux0 = bufs.aux0;
    p.past_actions = bufs.past_actions;
    p.past_vel = bufs.past_vel;
    p.episode = bufs.episode;
    p.stats = reinterpret_cast<unsigned long long *>(bufs.stats);
    p.kd_order = bufs.kd_order;
    const int wpb = 8;
    int grid = (rp.E + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    k_reset<T, OT><<<grid, wpb * 32, 0, st>>>(p, rp, env_mask, bank, scenario_idx, accumulate_stats);
    return cudaGetLastError();
}

template cudaError_t run_reset<float, float>(const cav_buffers_t &, const ResetParams &, const uint8_t *,
                                             const cav_scenario_bank_t &, const int32_t *, int, cudaStream_t);
template cudaError_t run_reset<double, double>(const cav_buffers_t &, const ResetParams &, const uint8_t *,
                                               const cav_scenario_bank_t &, const int32_t *, int, cudaStream_t);
template cudaError_t run_reset<double, float>(const cav_buffers_t &, const ResetParams &, const uint8_t *,
                                              const cav_scenario_bank_t &, const int32_t *, int, cudaStream_t);

}  // namespace cav
