// cav_kernels.h -- internal launcher interface between the C-ABI layer (cav_abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "cav_common.cuh"

namespace cav {

// G lanes per agent, SL neighbours per lane, GPE groups per env, EPB envs per CTA, CTA size
void launch_geometry(int N, int &G, int &SL, int &GPE, int &EPB, int &threads);

// phases: 1 = policy (action selection), 2 = move, 3 = both (fused)
template <typename T>
cudaError_t run_policy_move(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                            int sm_count, cudaStream_t st);
// OT = precision of the packed observation / reward columns in HBM: T, or float next to float64 state (cfg.obs_f32)
template <typename T, typename OT>
cudaError_t run_pairwise_obs(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan, const WallMap &wm,
                             const uint8_t *env_mask, int obs_only, int sm_count, cudaStream_t st);

// one-thread-per-agent pairwise+obs kernel (cav_tpa.cu), any observation layout, N <= 24
template <typename T, typename OT>
cudaError_t run_pairwise_obs_tpa(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan,
                                 const WallMap &wm, const uint8_t *env_mask, int obs_only, int sm_count,
                                 cudaStream_t st);
size_t tpa_smem_bytes(int N, int precision_f64, const ObsPlan &plan, int K);
// compile-time-N, register-resident specialisation (cav_tpa_n.cu); cudaErrorNotSupported = use the generic kernel
template <typename T, typename OT>
cudaError_t run_pairwise_obs_tpa_n(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan, const WallMap &wm,
                                   const uint8_t *env_mask, int obs_only, int sm_count, cudaStream_t st, int env0 = 0,
                                   int env1 = -1, int *tile_envs = nullptr);


// Scheduling order of the environments inside k_policy_move_v2 (library-owned scratch, see cav_policy.cu): position q of
// a launch's range processes environment order[q]; cost[e] is what environment e's linear programs cost last step.
struct EnvSchedule {
    const int *order = nullptr;
    unsigned short *cost = nullptr;
};

// thread-per-agent policy + move kernel (cav_policy.cu): every N in [2, 32], kd-tree order for N > 10;
// cudaErrorNotSupported otherwise (N = 1 and N > 32: the sub-warp-group kernel above)
template <typename T>
cudaError_t run_policy_move_v2(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                               int sm_count, cudaStream_t st, int env0 = 0, int env1 = -1, EnvSchedule sched = EnvSchedule{});
int policy_move_v2_tile_envs(int N);  // environments per CTA tile (a launch's env0 must be a multiple of it)
// order[lo, hi) := the environments lo .. hi - 1 by decreasing cost, for each of the nranges ranges [at[k], at[k + 1])
cudaError_t run_env_order(const EnvSchedule &sched, const int *at, int nranges, cudaStream_t st);

// GA3C-CADRL policy network (cav_ga3c.cu): on a batch X [rows][138], and for the CAV_POLICY_GA3C agents of bound buffers
cudaError_t run_ga3c_forward(const float *blob, const float *X, int rows, float *probs, int32_t *action_index, cudaStream_t st);
cudaError_t run_ga3c_policy(const cav_buffers_t &bufs, const cav_config_t &cfg, const float *blob, float *probs,
                            int32_t *list /* scratch: num_envs * num_agents + 1 ints */, cudaStream_t st);

struct LaserParams {
    int E, N, B, P, obs_len, obs_off;
    int rows, cols, wpr, nr;
    double origin0, origin1, cell, inv_cell;
    double min_angle, max_angle, ang_step, min_range, max_range, range_res;
};
template <typename T, typename OT>
cudaError_t run_laserscan(const cav_buffers_t &bufs, const LaserParams &lp, const uint8_t *env_mask, int sm_count,
                          cudaStream_t st);

struct ResetParams {
    int E, N, obs_len;
    double dt, near_goal_threshold, max_time_ratio;
};
template <typename T, typename OT>
cudaError_t run_reset(const cav_buffers_t &bufs, const ResetParams &rp, const uint8_t *env_mask,
                      const cav_scenario_bank_t &bank, const int32_t *scenario_idx, int accumulate_stats,
                      cudaStream_t st);

// device-side scenario banks (cav_scenario.cu)
cudaError_t run_generate_scenarios(double *rows, int32_t *kind_out, int S, int N, uint64_t seed, double side_lo,
                                   double side_hi, const double speed_bnds[2], const double radius_bnds[2],
                                   double close_range, int random_heading, cudaStream_t st);

}  // namespace cav
