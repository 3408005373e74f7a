// cav_policy.cu -- k_policy_move_v2: action selection + move for every agent count up to 32 (reference
// collision_avoidance_env.py:304-327 `_take_action`: RVOPolicy.py:50-122 + RVO2 Agent.cpp:287-630 / KdTree.cpp:47-111,
// 266-298, NonCooperativePolicy.py:21, StaticPolicy.py:21-22, ExternalPolicy.py:14-16, LearningPolicy.py:30-32, then
// Agent.take_action agent.py:192-241).
//
// Mapping (as k_pairwise_obs_tpa_n): environments never straddle a warp -- a warp owns EPW = 32 / N consecutive
// environments, one lane per agent, neighbour states go through the warp's own shared-memory slice behind __syncwarp,
// there is no block barrier.  What changed against the first compile-time-N kernel, and why (ncu,
// profiles/r01g_ncu_f64_16384x10.txt: 92 registers -> 20 warps / SM, ~90 KB of unrolled code -> `no_instruction` 2.2
// warps per issue, 42 % of the issue slots used):
//   * the ORCA half-planes live in shared memory [line][lane] instead of 36 registers, the neighbour sort keys die
//     before the linear programs start, and the float64 state (pv / goal / ht: 24 registers) is NOT carried across
//     the solve -- Agent.take_action re-reads it (L1 hits);
//   * every loop of the solve is rolled: the whole kernel is a few KB of code;
//   * one sqrt and one reciprocal per half-plane whatever branch of Agent.cpp:299-365 a lane is in
//     (orca_line_shared), a branch-free linearProgram1 loop (lp1_flat);
//   * measured and dropped (git history, commit "Warp-cooperative linearProgram2"): rounds in which every lane finds ITS
//     next violated line and the owners' (i, j) line pairs are spread over lane groups with shuffle folds -- bit-identical
//     on a CPU emulation of the warp, but 5 % SLOWER than the lane-private loops once those had their float64 spills
//     removed (78.8 vs 74.7 us at 16384 x 10, all agents live): ~30 shuffles + 4 __syncwarp per round outweigh the
//     ~3-of-30-lanes inefficiency of a loop whose body is software-pipelined four deep;
//   * N > 10: the reference's kd-tree decides the neighbour order at exactly equal distances -- the environment's
//     persistent agent permutation (`kd_order`) is re-partitioned by ONE lane per environment (cav_kdtree.cuh,
//     operation for operation KdTree.cpp:47-111) while the other lanes wait at a __syncwarp, then every lane inserts
//     its neighbours in the tree's visiting order.
//   * environment scheduling (EnvSchedule): a warp's cost is that of its costliest lane at every line index, so a warp of
//     UNLIKE environments pays for all of them, and the kernel ends when its longest warp does.  Measured with 16
//     scenarios x 1024 copies at 16384 x 10 (profiles/r02i_grouping.txt): scenarios interleaved over the environments
//     64.9 us, like next to like 52.0 us, like next to like AND costliest first 48.5 us.  So every step each lane counts
//     (4 + line index) per linearProgram1 it runs (+64 for a linearProgram3), the environment's sum goes to cost[e],
//     k_env_order sorts each launch range by decreasing cost (one CTA per range: 1024-bucket counting sort in shared
//     memory), and position q of the next step's launch processes environment order[q].  Environments are independent:
//     the order changes which warp computes an environment, never what it computes.
// NT in 2..10: compile-time agent count, sorting network in registers, no kd-tree needed (single leaf).
// NT in 11..16: compile-time agent count, the same network with each neighbour's kd-tree VISITING RANK as the tie-break
//   field of its key (the lane walks the tree once and notes the ranks, one byte per agent) -- CAV_PM_CTK.
// NT == 0: run-time agent count 17..32 (insertion sort through shared memory in kd-tree order).
#include "cav_orca_core.cuh"
#include "cav_tpa_common.cuh"

namespace cav {

#ifndef CAV_PM_WARPS
#define CAV_PM_WARPS 2  // warps per CTA
#endif
#ifndef CAV_PM_MINB
// resident CTAs per SM the register allocation must allow.  Sweep at 16384 x 10, float64 (profiles/r02c_libsweep.txt):
// (warps, CTAs) = (2, 12) -> 85 registers 56.8 us; (4, 6) 57.4; (1, 24) 60.0; (4, 8) 62.0; (2, 16) -> 64 registers 62.6;
// (1, 32) 68.8: avoiding the float64 spills is worth more than the last third of the occupancy
#define CAV_PM_MINB 12
#endif

#ifndef CAV_PM_CTK
#define CAV_PM_CTK 1  // compile-time instantiations for 11 .. 16 agents (kd-tree rank in a register sorting network); 0: run-time-n path
#endif

struct PmLayout {  // per-warp shared-memory slice (bytes)
    int N, EPW, LW, NO;
    int off_rad, off_lines, off_keys, off_box, off_meta, off_perm, warp_bytes;
};

__host__ __device__ constexpr int pm_r16(int x) { return (x + 15) & ~15; }

__host__ __device__ constexpr PmLayout pm_layout(int N, int mode) {
    PmLayout L{};
    L.N = N;
    L.EPW = 32 / N;
    L.LW = L.EPW * N;
    L.NO = N - 1;
    int o = 16 * 32;  // float4 state [32]
    L.off_rad = o;
    o += 4 * 32;
    L.off_lines = o;
    o += pm_r16(16 * (L.NO > 2 ? L.NO : 2) * L.LW);  // >= 2 slots: the anti-collaborative branch builds two half-planes
    L.off_keys = o;
    L.off_box = o;
    L.off_meta = o;
    L.off_perm = o;
    if (mode != 0) {
        // mode 1: 64-bit insertion-sort keys [NO][LW]; mode 2: one byte of kd-tree visiting rank per (agent, lane) [N][32]
        o += mode == 1 ? pm_r16(8 * L.NO * L.LW) : pm_r16(32 * N);
        L.off_box = o;
        o += 16 * L.EPW * (2 * N - 1);
        L.off_meta = o;
        o += pm_r16(4 * L.EPW * (2 * N - 1));
        L.off_perm = o;
        o += 32;
    }
    L.warp_bytes = o;
    return L;
}

// a second read of a column the thread has read before: `volatile` keeps the compiler from carrying the first copy in
// registers across the solve
template <typename T>
__device__ __forceinline__ R4<T> ld4_again(const R4<T> *p);
template <>
__device__ __forceinline__ R4<float> ld4_again<float>(const R4<float> *p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return R4<float>{v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ R4<double> ld4_again<double>(const R4<double> *p) {
    double a, b, c, d;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2+16];" : "=d"(c), "=d"(d) : "l"(p));
    return R4<double>{a, b, c, d};
}

template <typename T, int NT>
__global__ void __launch_bounds__(32 * CAV_PM_WARPS, CAV_PM_MINB)
    k_policy_move_v2(TpaBuffers b, KParams<T> p, const float *__restrict__ ext, uint8_t *__restrict__ kd_order, int phases,
                     T dt, int tile0, int ntiles, PmLayout rl, EnvSchedule sched) {
    constexpr bool CN = NT > 0;                 // compile-time agent count
    constexpr bool CT = NT > 0 && NT <= CAV_KD_MAX_LEAF;  // ... in one kd-tree leaf: the neighbour order needs no tree
    constexpr bool CTK = NT > CAV_KD_MAX_LEAF;  // ... 11 .. 16: kd-tree visiting ranks break distance ties
    constexpr PmLayout cl = pm_layout(CN ? NT : 32, CTK ? 2 : 0);
    const int N = CN ? NT : rl.N;
    const int EPW = CN ? cl.EPW : rl.EPW;
    const int LW = CN ? cl.LW : rl.LW;
    const int NO = CN ? cl.NO : rl.NO;
    const int off_rad = CN ? cl.off_rad : rl.off_rad, off_lines = CN ? cl.off_lines : rl.off_lines;
    const int warp_bytes = CN ? cl.warp_bytes : rl.warp_bytes;
    const int off_keys = CN ? cl.off_keys : rl.off_keys, off_box = CN ? cl.off_box : rl.off_box;
    const int off_meta = CN ? cl.off_meta : rl.off_meta, off_perm = CN ? cl.off_perm : rl.off_perm;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned char *sp = smem_raw + (size_t)warp * warp_bytes;
    float4 *s_f4 = reinterpret_cast<float4 *>(sp);
    float *s_frad = reinterpret_cast<float *>(sp + off_rad);
    float4 *s_lines = reinterpret_cast<float4 *>(sp + off_lines);  // [NO][LW]

    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    R4<T> *g_goal = reinterpret_cast<R4<T> *>(b.goal);
    const R4<T> *g_ht = reinterpret_cast<const R4<T> *>(b.ht);
    const R4<T> *g_ego = reinterpret_cast<const R4<T> *>(b.ego);
    const int el = lane / N, a = lane - el * N;  // environment within the warp, agent within the environment
    const int eb = el * N;

    for (int tile = tile0 + blockIdx.x; tile < ntiles; tile += gridDim.x) {  // tiles [tile0, ntiles) of this launch
        const int q = (tile * CAV_PM_WARPS + warp) * EPW + el;  // position in the schedule
        const bool have = lane < LW && q < p.E;
        const int e = (have && sched.order) ? sched.order[q] : q;
        const size_t gi = (size_t)e * N + a;
        int work = 0;
        uint32_t f = 0, kind = 0;
        int n = 0;
        float prefx = 0.f, prefy = 0.f, maxSpeed = 0.f;
        T heading = (T)0;
        if (have) {
            const R4<T> pv = ld4(g_pv + gi);
            const R4<T> gl = ld4(g_goal + gi);
            f = b.flags[gi];
            kind = b.kinds[gi];
            n = b.num_active ? b.num_active[e] : N;
            s_f4[lane] = make_float4((float)pv.x, (float)pv.y, (float)pv.z, (float)pv.w);  // RVOPolicy.py:57-74
            s_frad[lane] = (float)(p.rvo_radius_scale * gl.z);
            if ((phases & 1) && a < n && !(f & CAV_FLAG_DONE) && CAV_KIND_POLICY(kind) == CAV_POLICY_RVO) {
                const T gx = gl.x - pv.x, gy = gl.y - pv.y;  // RVOPolicy.py:66
                const T nrm = M<T>::sqrt_(gx * gx + gy * gy);
                const T sc = gl.w / nrm;  // :67
                prefx = (float)(sc * gx);
                prefy = (float)(sc * gy);
                maxSpeed = (float)gl.w;
                heading = ld4(g_ht + gi).x;
            }
        }
        const bool live = have && a < n && (phases & 1) && !(f & CAV_FLAG_DONE);
        const bool rvo = live && CAV_KIND_POLICY(kind) == CAV_POLICY_RVO;
        __syncwarp();

        KdTree kd{nullptr, nullptr, nullptr};
        bool use_kd = false;
        if (!CT) {
            // doStep -> buildAgentTree once per environment step, if any of its RVO agents is still stepping its simulator
            const unsigned rvo_mask = __ballot_sync(0xffffffffu, rvo);
            const unsigned env_bits = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << eb;
            use_kd = have && n > CAV_KD_MAX_LEAF && (rvo_mask & env_bits) != 0u && kd_order != nullptr;
            const int slots = 2 * N - 1;
            kd.box = reinterpret_cast<float4 *>(sp + off_box) + el * slots;
            kd.meta = reinterpret_cast<KdMeta *>(sp + off_meta) + el * slots;
            kd.perm = reinterpret_cast<uint8_t *>(sp + off_perm) + eb;
            if (use_kd && a < n) kd.perm[a] = kd_order[gi];
            __syncwarp();
            if (use_kd && a == 0) kd_build(s_f4 + eb, n, kd);
            __syncwarp();
            if (use_kd && a < n) kd_order[gi] = kd.perm[a];
        }

        float2 act = make_float2(0.f, 0.f);
        // ---- RVOPolicy.find_next_action: neighbour order + half-planes per lane, then the linear programs ----
        int nl = 0;
        if (rvo) {
            // RVOPolicy.py:76-90: the ego's collaboration coefficient; negative = anti-collaborative, re-drawn every
            // rvo_anti_collab_t seconds of agent time (within dt of a multiple of it) between "ignore the others"
            // (probability 1 - |coeff|: coefficient 0) and "steer into the closest neighbour" (the negative coefficient)
            float ego_collab = p.rvo_collab;
            if (p.rvo_collab < 0.0f) {
                const T tm = pymod<T>(ld4(g_ht + gi).z, p.rvo_anti_collab_t);
                const T r1 = M<T>::rint_(tm * (T)1000) / (T)1000, r2 = M<T>::rint_((p.rvo_anti_collab_t - tm) * (T)1000) / (T)1000;
                if (r1 < p.dt || r2 < p.dt) {
                    const double u = philox_u01(p.rvo_seed, (uint32_t)e, (uint32_t)a, (uint32_t)(b.episode ? b.episode[e] : 0),
                                                (uint32_t)b.step_num[gi]);
                    f = (u < 1.0 - (double)p.rvo_anti_p) ? (f | CAV_FLAG_RVO_NON_COOP) : (f & ~(uint32_t)CAV_FLAG_RVO_NON_COOP);
                    if (!(phases & 2)) b.flags[gi] = f;  // (the move phase stores the flags otherwise)
                }
                if (f & CAV_FLAG_RVO_NON_COOP) ego_collab = 0.0f;
            }
            if (CT) {
                nl = orca_lines_n<(CT ? NT - 1 : 1)>(n, a, s_f4 + eb, s_frad + eb, prefx, prefy, p.rvo_inv_time_horizon_num,
                                                     p.rvo_time_step, ego_collab, p.rvo_neighbor_dist, p.rvo_max_neighbors,
                                                     SmemLines{s_lines + lane, LW});
            } else if (CTK) {
                // this ego's kd-tree visiting ranks, one byte per agent in the lane's column of [N][32]
                uint8_t *s_rank = reinterpret_cast<uint8_t *>(sp + off_keys) + lane;
                if (use_kd) kd_visit_ranks(kd, s_f4[lane].x, s_f4[lane].y, s_rank, 32);
                nl = orca_lines_nk<(CTK ? NT - 1 : 1)>(n, a, s_f4 + eb, s_frad + eb, prefx, prefy, p.rvo_inv_time_horizon_num,
                                                       p.rvo_time_step, ego_collab, p.rvo_neighbor_dist, p.rvo_max_neighbors,
                                                       use_kd ? s_rank : nullptr, 32, SmemLines{s_lines + lane, LW});
            } else {
                unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(sp + off_keys);
                nl = orca_lines_rt(n, a, s_f4 + eb, s_frad + eb, prefx, prefy, p.rvo_inv_time_horizon_num, p.rvo_time_step,
                                   ego_collab, p.rvo_neighbor_dist, p.rvo_max_neighbors, s_keys + lane, LW,
                                   use_kd ? &kd : nullptr, SmemLines{s_lines + lane, LW});
            }
        }
        float nvx = 0.f, nvy = 0.f;
        if (rvo) {
            const int lp_fail = lp2_flat(SmemLines{s_lines + lane, LW}, nl, maxSpeed, prefx, prefy, false, nvx, nvy, &work);
            if (lp_fail < nl) {  // infeasible: linearProgram3 (Agent.cpp:419-421)
#if CAV_SCHED_KEY == 1
                work |= 1 << 30;
#else
                work += 64;
#endif
                Line projbuf[CN ? (NT > 2 ? NT - 1 : 2) : 31];
                lp3_flat(SmemLines{s_lines + lane, LW}, nl, lp_fail, maxSpeed, LocalLines{projbuf}, nvx, nvy);
            }
            // Agent::update (Agent.cpp:470-475) in float, then RVOPolicy.py:96-111 in `real`
            const float4 me = s_f4[lane];
            const R4<T> pv = ld4_again<T>(g_pv + gi);
            const float npx = me.x + nvx * p.rvo_time_step, npy = me.y + nvy * p.rvo_time_step;
            const T dpx = (T)npx - pv.x, dpy = (T)npy - pv.y;
            const T ang1 = M<T>::atan2_(dpy, dpx);
            const T new_heading = pymod<T>(ang1 - (T)0, two_pi<T>());
            T dh = wrap<T>(new_heading - heading);
            T spd = p.rvo_inv_dt * M<T>::sqrt_(dpx * dpx + dpy * dpy);
            if (M<T>::abs_(dh) > p.rvo_max_delta_heading) {
                dh = sign<T>(dh) * p.rvo_max_delta_heading;
                spd = (T)0;
            }
            act = make_float2((float)spd, (float)dh);  // float32 store, env.py:305-307
        } else if (live) {
            const uint32_t pol = CAV_KIND_POLICY(kind);
            T a0 = (T)0, a1 = (T)0;
            if (pol == CAV_POLICY_NONCOOP) {  // NonCooperativePolicy.py:21
                a0 = ld4(g_goal + gi).w;
                a1 = -ld4(g_ego + gi).y;
            } else if (pol == CAV_POLICY_STATIC) {  // StaticPolicy.py:21-22: goal := pos
                R4<T> gl = ld4(g_goal + gi);
                const R4<T> pv = ld4(g_pv + gi);
                gl.x = pv.x;
                gl.y = pv.y;
                st4(g_goal + gi, gl);
            } else if (pol == CAV_POLICY_EXTERNAL) {  // ExternalPolicy.py:14-16
                const float2 x = ext ? reinterpret_cast<const float2 *>(ext)[gi] : make_float2(0.f, 0.f);
                a0 = (T)x.x;
                a1 = (T)x.y;
            } else if (pol == CAV_POLICY_GA3C) {  // GA3CCADRLPolicy.py:80-84: k_ga3c (cav_ga3c.cu) wrote it this step
                const float2 x = reinterpret_cast<const float2 *>(b.actions)[gi];
                a0 = (T)x.x;
                a1 = (T)x.y;
            } else {  // CAV_POLICY_LEARNING, LearningPolicy.py:30-32
                const float2 x = ext ? reinterpret_cast<const float2 *>(ext)[gi] : make_float2(0.f, 0.f);
                a1 = p.max_heading_change * ((T)2. * (T)x.y - (T)1.);
                a0 = ld4(g_goal + gi).w * (T)x.x;
            }
            act = make_float2((float)a0, (float)a1);
        }
        if (live) reinterpret_cast<float2 *>(b.actions)[gi] = act;
        if ((phases & 2) && have && a < n) {  // Agent.take_action on freshly read state
            const R4<T> pv = ld4_again<T>(g_pv + gi);
            const R4<T> gl = ld4_again<T>(g_goal + gi);
            const R4<T> ht = ld4_again<T>(g_ht + gi);
            tpa_move<T>(b, p, gi, pv, gl, ht, f, kind, act, live, dt);
        }
        __syncwarp();
        if (sched.cost && (phases & 1)) {  // the environment's scheduling key for the next step
            const unsigned env_bits = lane < LW ? ((N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << eb) : (1u << lane);
#if CAV_SCHED_KEY == 1
            // union over the environment's agents of the line indices that ran linearProgram1 (what a warp of such
            // environments executes), the highest nine indices as the key, linearProgram3 on top
            const unsigned uni = __reduce_or_sync(env_bits, (unsigned)work);
            const int sh = NO > 9 ? NO - 9 : 0;
            const int total = (int)((uni & 0x3fffffffu) >> sh & 511u) | ((uni >> 30) ? 512 : 0);
#else
            const int total = __reduce_add_sync(env_bits, work);
#endif
            if (have && a == 0) sched.cost[e] = (unsigned short)(total < 1023 ? total : 1023);
        }
    }
}

// order[lo, hi) := environments lo .. hi - 1 by decreasing cost (ties in any order); one CTA per range
#define CAV_ORDER_BUCKETS 1024
__global__ void __launch_bounds__(1024) k_env_order(const unsigned short *__restrict__ cost, int *__restrict__ order, int4 at0,
                                                    int at4) {
    __shared__ int s_cnt[CAV_ORDER_BUCKETS];
    __shared__ int s_warp[32];
    const int at[5] = {at0.x, at0.y, at0.z, at0.w, at4};
    const int lo = at[blockIdx.x], hi = at[blockIdx.x + 1];
    const int tid = threadIdx.x;
    s_cnt[tid] = 0;
    __syncthreads();
    for (int e = lo + tid; e < hi; e += 1024) {
        const int c = cost[e];
        atomicAdd(&s_cnt[CAV_ORDER_BUCKETS - 1 - (c < CAV_ORDER_BUCKETS - 1 ? c : CAV_ORDER_BUCKETS - 1)], 1);
    }
    __syncthreads();
    // exclusive prefix sum over the buckets (bucket 0 = highest cost)
    const int v = s_cnt[tid];
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((tid & 31) >= d) inc += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        int w = s_warp[tid];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, d);
            if (tid >= d) w += t;
        }
        s_warp[tid] = w;
    }
    __syncthreads();
    s_cnt[tid] = inc - v + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
    __syncthreads();
    for (int e = lo + tid; e < hi; e += 1024) {
        const int c = cost[e];
        const int slot = atomicAdd(&s_cnt[CAV_ORDER_BUCKETS - 1 - (c < CAV_ORDER_BUCKETS - 1 ? c : CAV_ORDER_BUCKETS - 1)], 1);
        order[lo + slot] = e;
    }
}

cudaError_t run_env_order(const EnvSchedule &sched, const int *at, int nranges, cudaStream_t st) {
    if (!sched.order || !sched.cost || nranges < 1 || nranges > 4) return cudaErrorInvalidValue;
    int a[5];
    for (int k = 0; k < 5; ++k) a[k] = at[k < nranges ? k : nranges];
    k_env_order<<<nranges, 1024, 0, st>>>(sched.cost, const_cast<int *>(sched.order), make_int4(a[0], a[1], a[2], a[3]), a[4]);
    return cudaGetLastError();
}

// Environments [env0, env1) (env0 a multiple of policy_move_v2_tile_envs); env1 < 0 = all.
template <typename T, int NT>
static cudaError_t launch_v2(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                             int sm_count, cudaStream_t st, int env0, int env1, const EnvSchedule &sched) {
    const PmLayout L = pm_layout(p.N, NT == 0 ? 1 : (NT > CAV_KD_MAX_LEAF ? 2 : 0));
    const int epb = CAV_PM_WARPS * L.EPW;
    if (env1 < 0 || env1 > p.E) env1 = p.E;
    const int tile0 = env0 / epb;
    const int ntiles = (env1 + epb - 1) / epb;  // one past the last tile of this launch
    if (ntiles <= tile0) return cudaSuccess;
    const size_t smem = (size_t)CAV_PM_WARPS * L.warp_bytes;
    auto kern = k_policy_move_v2<T, NT>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int cap = sm_count * 32;
    const int grid = (ntiles - tile0) < cap ? (ntiles - tile0) : cap;
    kern<<<grid, 32 * CAV_PM_WARPS, smem, st>>>(make_tpa(bufs), p, ext, bufs.kd_order, phases, (T)dt, tile0, ntiles, L, sched);
    return cudaGetLastError();
}

// every N in [2, 32]; cudaErrorNotSupported otherwise (N = 1 and N > 32 take the sub-warp-group kernel)
int policy_move_v2_tile_envs(int N) { return CAV_PM_WARPS * pm_layout(N, N > 10 ? 1 : 0).EPW; }

template <typename T>
cudaError_t run_policy_move_v2(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                               int sm_count, cudaStream_t st, int env0, int env1, EnvSchedule sched) {
    switch (p.N) {
#define CAV_CASE(NN) \
    case NN: return launch_v2<T, NN>(bufs, p, ext, phases, dt, sm_count, st, env0, env1, sched);
        CAV_CASE(2)
        CAV_CASE(3)
        CAV_CASE(4)
        CAV_CASE(5)
        CAV_CASE(6)
        CAV_CASE(7)
        CAV_CASE(8)
        CAV_CASE(9)
        CAV_CASE(10)
#if CAV_PM_CTK
        CAV_CASE(11)
        CAV_CASE(12)
        CAV_CASE(13)
        CAV_CASE(14)
        CAV_CASE(15)
        CAV_CASE(16)
#endif
#undef CAV_CASE
        default: break;
    }
    if (p.N > 10 && p.N <= 32) return launch_v2<T, 0>(bufs, p, ext, phases, dt, sm_count, st, env0, env1, sched);
    return cudaErrorNotSupported;
}

template cudaError_t run_policy_move_v2<float>(const cav_buffers_t &, const KParams<float> &, const float *, int, double,
                                               int, cudaStream_t, int, int, EnvSchedule);
template cudaError_t run_policy_move_v2<double>(const cav_buffers_t &, const KParams<double> &, const float *, int, double,
                                                int, cudaStream_t, int, int, EnvSchedule);

}  // namespace cav
