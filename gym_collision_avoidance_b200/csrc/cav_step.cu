// cav_step.cu -- the two per-step kernels of libcavoid_sm100a.so (sm_100a):
//
//   k_policy_move   reference collision_avoidance_env.py:304-327 (_take_action): action selection for every agent
//                   from the pre-move state of its whole environment (RVOPolicy.py:50-122 + RVO2 Agent.cpp:287-630,
//                   NonCooperativePolicy.py:21, StaticPolicy.py:21-22, ExternalPolicy.py:14-16,
//                   LearningPolicy.py:30-32), block barrier, then Agent.take_action (agent.py:192-241:
//                   UnicycleDynamics.py:27-47, Dynamics.py:31-41, goal / timeout checks, history row).
//   k_pairwise_obs  env.py:394-512 (_compute_rewards + _check_for_collisions), OtherAgentsStatesSensor.sense
//                   (OtherAgentsStatesSensor.py:58-144) + get_observation_dict (agent.py:323-327), and
//                   _check_which_agents_done (env.py:514-553): the fused "pairwise + obs" kernel.
//
// Layout / mapping: a CTA owns a tile of EPB consecutive environments.  Their agent blocks are contiguous in every
// SoA column, so staging a tile is a coalesced 16-byte-per-lane copy into shared memory.  All O(N^2) work then
// runs out of shared memory with one sub-warp group of G lanes per (env, agent): lanes own the N-1 neighbours
// (SL per lane when N-1 > 32), distances / ORCA lines / sort keys live in registers, reductions and the bitonic
// neighbour sort are warp shuffles.  The purely per-agent move phase runs one thread per agent.
#include "cav_kernels.h"
#include "cav_tti.cuh"
#include "cav_orca.cuh"

namespace cav {

enum { PH_POLICY = 1, PH_MOVE = 2 };

struct DevBuffers {
    void *pv, *goal, *ht, *ego;
    uint32_t *flags;
    uint32_t *kinds;
    int32_t *step_num;
    const int32_t *num_active;
    float *actions;
    void *reward, *min_dist;
    uint8_t *done, *game_over;
    void *obs;
    void *history, *aux0, *past_actions, *past_vel;
    const int32_t *episode;
};

static inline DevBuffers make_dev(const cav_buffers_t &b) {
    DevBuffers d;
    d.pv = b.pv;
    d.goal = b.goal;
    d.ht = b.ht;
    d.ego = b.ego;
    d.flags = b.flags;
    d.kinds = b.kinds;
    d.step_num = b.step_num;
    d.num_active = b.num_active;
    d.actions = b.actions;
    d.reward = b.reward;
    d.min_dist = b.min_dist;
    d.done = b.done;
    d.game_over = b.game_over;
    d.obs = b.obs;
    d.history = b.history;
    d.aux0 = b.aux0;
    d.past_actions = b.past_actions;
    d.past_vel = b.past_vel;
    d.episode = b.episode;
    return d;
}

__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// =============================================================================================================
// kernel A: policy + move
// =============================================================================================================
template <typename T, int G, int SL>
__global__ void __launch_bounds__(512) k_policy_move(DevBuffers b, KParams<T> p, const float *__restrict__ ext,
                                                       uint8_t *__restrict__ kd_order, int phases, T dt, int EPB, int GPE,
                                                       int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N;
    const int cap = EPB * N;
    unsigned char *sp = smem_raw;
    R4<T> *s_pv = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    R4<T> *s_goal = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    R4<T> *s_ht = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    R4<T> *s_ego = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    float4 *s_f4 = reinterpret_cast<float4 *>(sp);
    sp += align16(sizeof(float4) * cap);
    float2 *s_act = reinterpret_cast<float2 *>(sp);
    sp += align16(sizeof(float2) * cap);
    float *s_frad = reinterpret_cast<float *>(sp);
    sp += align16(sizeof(float) * cap);
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(sp);
    sp += align16(sizeof(uint32_t) * cap);
    uint32_t *s_kinds = reinterpret_cast<uint32_t *>(sp);
    sp += align16(sizeof(uint32_t) * cap);
    int *s_n = reinterpret_cast<int *>(sp);
    sp += align16(sizeof(int) * EPB);
    // agent kd-tree per environment (N > MAX_LEAF_SIZE only; cav_kdtree.cuh): boxes, node records, permutation, in-use flag
    const bool kd_cfg = N > CAV_KD_MAX_LEAF && kd_order != nullptr;
    const int kd_slots_n = 2 * N - 1;
    float4 *s_kdbox = reinterpret_cast<float4 *>(sp);
    sp += kd_cfg ? align16(sizeof(float4) * (size_t)EPB * kd_slots_n) : 0;
    KdMeta *s_kdmeta = reinterpret_cast<KdMeta *>(sp);
    sp += kd_cfg ? align16(sizeof(KdMeta) * (size_t)EPB * kd_slots_n) : 0;
    uint8_t *s_kdperm = reinterpret_cast<uint8_t *>(sp);
    sp += kd_cfg ? align16((size_t)cap) : 0;
    uint8_t *s_kduse = reinterpret_cast<uint8_t *>(sp);

    const int tid = threadIdx.x;
    const unsigned gm = group_mask<G>();
    const int glane = tid & (G - 1);
    const int grp = tid / G;
    const int g_el = grp / GPE;        // env slot of this group inside the tile
    const int g_first = grp - g_el * GPE;  // first agent handled by this group

    R4<T> *g_pv = reinterpret_cast<R4<T> *>(b.pv);
    R4<T> *g_goal = reinterpret_cast<R4<T> *>(b.goal);
    R4<T> *g_ht = reinterpret_cast<R4<T> *>(b.ht);
    R4<T> *g_ego = reinterpret_cast<R4<T> *>(b.ego);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int e0 = tile * EPB;
        const int ne = min(EPB, p.E - e0);
        const int count = ne * N;
        const size_t base = (size_t)e0 * N;
        // ---- stage the tile's agent blocks (coalesced: the tile is one contiguous range in every column) ----
        for (int idx = tid; idx < count; idx += blockDim.x) {
            const R4<T> pv = ld4(g_pv + base + idx);
            const R4<T> gl = ld4(g_goal + base + idx);
            s_pv[idx] = pv;
            s_goal[idx] = gl;
            s_ht[idx] = ld4(g_ht + base + idx);
            s_ego[idx] = ld4(g_ego + base + idx);
            s_flags[idx] = b.flags[base + idx];
            s_kinds[idx] = b.kinds[base + idx];
            // what RVOPolicy.py:57-74 pushes into the float32 simulator
            s_f4[idx] = make_float4((float)pv.x, (float)pv.y, (float)pv.z, (float)pv.w);
            s_frad[idx] = (float)(p.rvo_radius_scale * gl.z);
            if (!(phases & PH_POLICY)) s_act[idx] = reinterpret_cast<const float2 *>(b.actions)[base + idx];
        }
        if (tid < ne) s_n[tid] = b.num_active ? b.num_active[e0 + tid] : N;
        __syncthreads();

        // ---- doStep -> buildAgentTree (KdTree.cpp:47-111): one thread per environment advances the persistent agent
        // permutation, if any RVO agent of the environment is still stepping its simulator --------------------------
        if (kd_cfg && (phases & PH_POLICY)) {
            if (tid < ne) {
                const int n = s_n[tid], ebk = tid * N;
                bool have_rvo = false;
                for (int i = 0; i < n; ++i)
                    have_rvo = have_rvo || (!(s_flags[ebk + i] & CAV_FLAG_DONE) &&
                                            CAV_KIND_POLICY(s_kinds[ebk + i]) == CAV_POLICY_RVO);
                const bool use = have_rvo && n > CAV_KD_MAX_LEAF;
                s_kduse[tid] = use ? 1 : 0;
                if (use) {
                    KdTree t{s_kdbox + tid * kd_slots_n, s_kdmeta + tid * kd_slots_n, s_kdperm + ebk};
                    for (int i = 0; i < n; ++i) t.perm[i] = kd_order[base + ebk + i];
                    kd_build(s_f4 + ebk, n, t);
                    for (int i = 0; i < n; ++i) kd_order[base + ebk + i] = t.perm[i];
                }
            }
            __syncthreads();
        }

        // ---- policy phase: one group per (env, agent) -------------------------------------------------------
        if ((phases & PH_POLICY) && g_el < ne && grp < EPB * GPE) {
            const int n = s_n[g_el];
            const int eb = g_el * N;
            const KdTree kdt{s_kdbox + g_el * kd_slots_n, s_kdmeta + g_el * kd_slots_n, s_kdperm + eb};
            const KdTree *kd = (kd_cfg && s_kduse[g_el]) ? &kdt : nullptr;
            for (int a = g_first; a < n; a += GPE) {
                const int li = eb + a;
                const uint32_t f = s_flags[li];
                if (f & CAV_FLAG_DONE) continue;  // env.py:311-312
                const uint32_t pol = CAV_KIND_POLICY(s_kinds[li]);
                T a0 = (T)0, a1 = (T)0;
                if (pol == CAV_POLICY_RVO) {
                    const R4<T> pv = s_pv[li];
                    const R4<T> gl = s_goal[li];
                    const T gx = gl.x - pv.x, gy = gl.y - pv.y;  // RVOPolicy.py:66
                    const T nrm = M<T>::sqrt_(gx * gx + gy * gy);
                    const T sc = gl.w / nrm;  // :67
                    const float prefx = (float)(sc * gx), prefy = (float)(sc * gy);
                    float nvx, nvy;
                    // RVOPolicy.py:76-90: anti-collaborative agents re-draw their mode every rvo_anti_collab_t seconds
                    float ego_collab = p.rvo_collab;
                    if (p.rvo_collab < 0.0f) {
                        uint32_t fl = f;
                        const T tm = pymod<T>(s_ht[li].z, p.rvo_anti_collab_t);
                        const T r1 = M<T>::rint_(tm * (T)1000) / (T)1000,
                                r2 = M<T>::rint_((p.rvo_anti_collab_t - tm) * (T)1000) / (T)1000;
                        if (r1 < p.dt || r2 < p.dt) {
                            const double u = philox_u01(p.rvo_seed, (uint32_t)(e0 + g_el), (uint32_t)a,
                                                        (uint32_t)(b.episode ? b.episode[e0 + g_el] : 0),
                                                        (uint32_t)b.step_num[base + li]);
                            fl = (u < 1.0 - (double)p.rvo_anti_p) ? (fl | CAV_FLAG_RVO_NON_COOP)
                                                                   : (fl & ~(uint32_t)CAV_FLAG_RVO_NON_COOP);
                            __syncwarp(gm);  // every lane of the group has read s_flags[li]
                            if (glane == 0) {
                                s_flags[li] = fl;
                                b.flags[base + li] = fl;
                            }
                        }
                        if (fl & CAV_FLAG_RVO_NON_COOP) ego_collab = 0.0f;
                    }
                    orca_group<G, SL>(gm, glane, n, a, s_f4 + eb, s_frad + eb, prefx, prefy, (float)gl.w,
                                      p.rvo_inv_time_horizon_num, p.rvo_time_step, ego_collab, p.rvo_neighbor_dist,
                                      p.rvo_max_neighbors, kd, nvx, nvy);
                    // Agent::update (Agent.cpp:470-475) in float, then RVOPolicy.py:96-111 in `real`
                    const float4 me = s_f4[li];
                    const float npx = me.x + nvx * p.rvo_time_step, npy = me.y + nvy * p.rvo_time_step;
                    const T dpx = (T)npx - pv.x, dpy = (T)npy - pv.y;
                    const T ang1 = M<T>::atan2_(dpy, dpx);
                    const T new_heading = pymod<T>(ang1 - (T)0, two_pi<T>());
                    T dh = wrap<T>(new_heading - s_ht[li].x);
                    T spd = p.rvo_inv_dt * M<T>::sqrt_(dpx * dpx + dpy * dpy);
                    if (M<T>::abs_(dh) > p.rvo_max_delta_heading) {
                        dh = sign<T>(dh) * p.rvo_max_delta_heading;
                        spd = (T)0;
                    }
                    a0 = spd;
                    a1 = dh;
                } else if (pol == CAV_POLICY_NONCOOP) {  // NonCooperativePolicy.py:21
                    a0 = s_goal[li].w;
                    a1 = -s_ego[li].y;
                } else if (pol == CAV_POLICY_STATIC) {  // StaticPolicy.py:21-22: goal := pos
                    if (glane == 0) {
                        R4<T> gl = s_goal[li];
                        gl.x = s_pv[li].x;
                        gl.y = s_pv[li].y;
                        s_goal[li] = gl;
                        st4(g_goal + base + li, gl);
                    }
                } else if (pol == CAV_POLICY_EXTERNAL) {  // ExternalPolicy.py:14-16
                    const float2 x = ext ? reinterpret_cast<const float2 *>(ext)[base + li] : make_float2(0.f, 0.f);
                    a0 = (T)x.x;
                    a1 = (T)x.y;
                } else if (pol == CAV_POLICY_GA3C) {  // GA3CCADRLPolicy.py:80-84: k_ga3c (cav_ga3c.cu) wrote it this step
                    const float2 x = reinterpret_cast<const float2 *>(b.actions)[base + li];
                    a0 = (T)x.x;
                    a1 = (T)x.y;
                } else {  // CAV_POLICY_LEARNING, LearningPolicy.py:30-32
                    const float2 x = ext ? reinterpret_cast<const float2 *>(ext)[base + li] : make_float2(0.f, 0.f);
                    a1 = p.max_heading_change * ((T)2. * (T)x.y - (T)1.);
                    a0 = s_goal[li].w * (T)x.x;
                }
                if (glane == 0) {
                    const float2 act = make_float2((float)a0, (float)a1);  // float32 store, env.py:305-307
                    s_act[li] = act;
                    reinterpret_cast<float2 *>(b.actions)[base + li] = act;
                }
            }
        }
        __syncthreads();

        // ---- move phase: one thread per agent (Agent.take_action, agent.py:192-241) -------------------------
        if ((phases & PH_MOVE) && tid < count) {
            const int li = tid;
            const int el = li / N, a = li - el * N;
            if (a < s_n[el]) {
                const size_t gi = base + li;
                uint32_t f = s_flags[li];
                R4<T> pv = s_pv[li];
                if (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) {  // agent.py:202-209
                    if (f & CAV_FLAG_AT_GOAL) f |= CAV_FLAG_WAS_AT_GOAL_ALREADY;
                    if (f & CAV_FLAG_IN_COLLISION) f |= CAV_FLAG_WAS_IN_COLLISION_ALREADY;
                    pv.z = (T)0;
                    pv.w = (T)0;
                    st4(g_pv + gi, pv);
                    b.flags[gi] = f;
                    if (b.past_vel) {
                        R4<T> *pvel = reinterpret_cast<R4<T> *>(b.past_vel) + gi;
                        R4<T> q = ld4(pvel);
                        q.z = q.x;
                        q.w = q.y;
                        q.x = (T)0;
                        q.y = (T)0;
                        st4(pvel, q);
                    }
                } else {
                    const R4<T> gl = s_goal[li];
                    R4<T> ht = s_ht[li];
                    const float2 act = s_act[li];
                    const T a0 = (T)act.x;
                    T a1 = (T)act.y;
                    const uint32_t dyn = CAV_KIND_DYNAMICS(s_kinds[li]);
                    if (b.past_actions) {  // agent.py:212-213
                        R4<T> *pa = reinterpret_cast<R4<T> *>(b.past_actions) + gi;
                        R4<T> q = ld4(pa);
                        q.z = q.x;
                        q.w = q.y;
                        q.x = a0;
                        q.y = a1;
                        st4(pa, q);
                    }
                    T speed_out = (T)0, dhead_out = (T)0;
                    if (b.aux0) {
                        const R4<T> q = ld4(reinterpret_cast<R4<T> *>(b.aux0) + gi);
                        speed_out = q.x;
                        dhead_out = q.y;
                    }
                    if (dyn != CAV_DYN_EXTERNAL) {
                        if (dyn == CAV_DYN_UNICYCLE_MAX_TURN_RATE) {  // UnicycleDynamicsMaxTurnRate.py:30-32
                            // `action` is a float32 array (env.py:305-307) and dt a Python float, so under NumPy >= 2 (NEP 50) the
                            // turning rate, its clip and `turning_rate*dt` are float32 operations; the sum with the heading is float64
                            float tr = (float)a1 / (float)dt;
                            tr = tr < -3.0f ? -3.0f : (tr > 3.0f ? 3.0f : tr);
                            a1 = (T)(tr * (float)dt);
                        }
                        const T heading = ht.x;
                        const T sel = wrap<T>(a1 + heading);  // UnicycleDynamics.py:28
                        const T c = M<T>::cos_(sel), s = M<T>::sin_(sel);
                        const T dx = a0 * c * dt, dy = a0 * s * dt;  // :30-31
                        pv.x += dx;
                        pv.y += dy;
                        pv.z = a0 * c;  // :34-35
                        pv.w = a0 * s;
                        speed_out = a0;
                        dhead_out = wrap<T>(sel - heading);  // :37-38
                        ht.x = sel;
                        if (dyn == CAV_DYN_UNICYCLE) {  // turning_dir, :42-47
                            T td = ht.w;
                            if (M<T>::abs_(td) < (T)1e-5) {
                                td = (T)0.11 * sign<T>(sel);
                            } else if (td * sel < (T)0) {
                                T v = -td + sel;
                                v = v < pi<T>() ? v : pi<T>();
                                v = v > -pi<T>() ? v : -pi<T>();
                                td = v;
                            } else {
                                T m = M<T>::abs_(td) - (T)0.1;
                                m = m > (T)0 ? m : (T)0.0;
                                td = sign<T>(td) * m;
                            }
                            ht.w = td;
                        }
                    }
                    R4<T> ego;
                    T vex, vey;
                    update_ego<T>(pv, gl, ht.x, ego, vex, vey, b.aux0 != nullptr);  // agent.py:227
                    if (b.aux0) st4(reinterpret_cast<R4<T> *>(b.aux0) + gi, R4<T>{speed_out, dhead_out, vex, vey});
                    const int sn = b.step_num[gi];
                    if (p.store_history && b.history && sn < p.history_len) {  // agent.py:257-260,275-289
                        T *row = reinterpret_cast<T *>(b.history) + ((size_t)gi * p.history_len + sn) * CAV_HISTORY_ROW;
                        const T v[CAV_HISTORY_ROW] = {ht.z, pv.x, pv.y, gl.x, gl.y, gl.z, gl.w, pv.z, pv.w,
                                                      (dyn != CAV_DYN_EXTERNAL) ? a0 : speed_out, ht.x, ht.z, ego.x, ego.y};
                        store_history_row(row, v);
                    }
                    // _check_if_at_goal, agent.py:150-153
                    const T ex = pv.x - gl.x, ey = pv.y - gl.y;
                    if (ex * ex + ey * ey <= p.near_goal_thr * p.near_goal_thr) f |= CAV_FLAG_AT_GOAL;
                    else f &= ~CAV_FLAG_AT_GOAL;
                    if (b.past_vel) {
                        R4<T> *pvel = reinterpret_cast<R4<T> *>(b.past_vel) + gi;
                        R4<T> q = ld4(pvel);
                        q.z = q.x;
                        q.w = q.y;
                        q.x = pv.z;
                        q.y = pv.w;
                        st4(pvel, q);
                    }
                    ht.y -= dt;  // agent.py:235-239
                    ht.z += dt;
                    b.step_num[gi] = sn + 1;
                    if (ht.y <= (T)0) f |= CAV_FLAG_RAN_OUT_OF_TIME;
                    st4(g_pv + gi, pv);
                    st4(g_ht + gi, ht);
                    st4(g_ego + gi, ego);
                    b.flags[gi] = f;
                }
            }
        }
        __syncthreads();  // shared memory is re-staged by the next tile
    }
}

// =============================================================================================================
// kernel B: pairwise collisions + rewards + sorted-neighbour observation + done
// =============================================================================================================
template <typename T>
struct ObsKey {
    T tt;  // -time_to_impact for AGENT_SORTING_METHOD 'time_to_impact' (OtherAgentsStatesSensor.py:36-38), else 0
    T kr;  // round(dist_2_other, 2) (:103); negated for 'time_to_impact' (key -x[1])
    T po;  // p_orthog_ego_frame
    int idx;
};

// time_to_impact<T> (util.compute_time_to_impact) lives in cav_tti.cuh: host/device source, also compiled by the CPU tests
template <typename T>
struct ObsKeyOps {
    static __device__ __forceinline__ bool less(const ObsKey<T> &a, const ObsKey<T> &b) {
        if (a.tt < b.tt) return true;
        if (a.tt > b.tt) return false;
        if (a.kr < b.kr) return true;
        if (a.kr > b.kr) return false;
        if (a.po < b.po) return true;
        if (a.po > b.po) return false;
        return a.idx < b.idx;  // Python's sorted() is stable: ties keep agent-index order
    }
    static __device__ __forceinline__ ObsKey<T> shfl_xor(unsigned gm, const ObsKey<T> &v, int j, int w) {
        ObsKey<T> o;
        o.tt = __shfl_xor_sync(gm, v.tt, j, w);
        o.kr = __shfl_xor_sync(gm, v.kr, j, w);
        o.po = __shfl_xor_sync(gm, v.po, j, w);
        o.idx = __shfl_xor_sync(gm, v.idx, j, w);
        return o;
    }
};

template <typename T, int G, int SL, typename OT>
__global__ void __launch_bounds__(512) k_pairwise_obs(DevBuffers b, KParams<T> p, ObsPlan plan, WallMap wm,
                                                        const uint8_t *__restrict__ env_mask, int obs_only, int EPB,
                                                        int GPE, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N;
    const int cap = EPB * N;
    unsigned char *sp = smem_raw;
    R4<T> *s_pv = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    R4<T> *s_goal = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    R4<T> *s_ego = reinterpret_cast<R4<T> *>(sp);
    sp += align16(sizeof(R4<T>) * cap);
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(sp);
    sp += align16(sizeof(uint32_t) * cap);
    uint32_t *s_kinds = reinterpret_cast<uint32_t *>(sp);
    sp += align16(sizeof(uint32_t) * cap);
    float *s_adh = reinterpret_cast<float *>(sp);
    sp += align16(sizeof(float) * cap);
    int *s_n = reinterpret_cast<int *>(sp);
    sp += align16(sizeof(int) * EPB);
    T *s_obs = reinterpret_cast<T *>(sp);  // [EPB*GPE groups][compact_len]

    const int tid = threadIdx.x;
    const unsigned gm = group_mask<G>();
    const int glane = tid & (G - 1);
    const int grp = tid / G;
    const int g_el = grp / GPE;
    const int g_first = grp - g_el * GPE;
    const int CL = plan.compact_len;
    T *my_obs = s_obs + (size_t)grp * CL;

    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    const R4<T> *g_goal = reinterpret_cast<const R4<T> *>(b.goal);
    const R4<T> *g_ego = reinterpret_cast<const R4<T> *>(b.ego);
    OT *g_obs = reinterpret_cast<OT *>(b.obs);  // OT: precision of the packed observation / reward in HBM
    OT *g_reward = reinterpret_cast<OT *>(b.reward);
    T *g_mind = reinterpret_cast<T *>(b.min_dist);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int e0 = tile * EPB;
        const int ne = min(EPB, p.E - e0);
        const int count = ne * N;
        const size_t base = (size_t)e0 * N;
        for (int idx = tid; idx < count; idx += blockDim.x) {
            s_pv[idx] = ld4(g_pv + base + idx);
            s_goal[idx] = ld4(g_goal + base + idx);
            s_ego[idx] = ld4(g_ego + base + idx);
            s_flags[idx] = b.flags[base + idx];
            s_kinds[idx] = b.kinds[base + idx];
            s_adh[idx] = b.actions[2 * (base + idx) + 1];
        }
        if (tid < ne) {
            int n = b.num_active ? b.num_active[e0 + tid] : N;
            if (env_mask && !env_mask[e0 + tid]) n = -1;  // env not selected: nothing to do
            s_n[tid] = n;
        }
        __syncthreads();

        if (g_el < ne && grp < EPB * GPE) {
            const int n = s_n[g_el];
            const int eb = g_el * N;
            for (int h = g_first; h < n; h += GPE) {
                const int li = eb + h;
                const R4<T> pvh = s_pv[li];
                const R4<T> glh = s_goal[li];
                const R4<T> egh = s_ego[li];
                const T prx = egh.z, pry = egh.w;
                const T orx = -pry, ory = prx;  // ref_orth, agent.py:347
                ObsKey<T> key[SL];
                T md = M<T>::inf();
                bool coll = false;
                int nc = 0;
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int e = s * G + glane;
                    const int o = (e < h) ? e : e + 1;
                    key[s].tt = M<T>::inf();
                    key[s].kr = M<T>::inf();
                    key[s].po = M<T>::inf();
                    key[s].idx = CAV_MAX_AGENTS + e;
                    bool seen = false;
                    if (e < n - 1) {
                        const R4<T> pvo = s_pv[eb + o];
                        const T ro = s_goal[eb + o].z;
                        const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                        const T dc = M<T>::sqrt_(rx * rx + ry * ry);  // util.l2norm == util.vec2_l2_norm
                        const T cr = glh.z + ro;  // env.py:481
                        const T gap = dc - cr;
                        md = min_keep(md, gap);
                        coll = coll || (dc <= cr);  // env.py:490
                        if (!(dc > p.sensing_horizon)) {  // OtherAgentsStatesSensor.py:90
                            const T d2o = dc - glh.z - ro;  // :87
                            key[s].tt = (T)0;
                            key[s].kr = M<T>::rint_(d2o * (T)100) / (T)100;
                            if (p.sorting == CAV_SORT_TIME_TO_IMPACT) {  // key (-tti, -round(dist, 2), p_orth), :36-38
                                key[s].tt = -time_to_impact<T>(pvh.x, pvh.y, pvo.x, pvo.y, pvh.z, pvh.w, pvo.z, pvo.w, cr);
                                key[s].kr = -key[s].kr;
                            }
                            key[s].po = rx * orx + ry * ory;
                            key[s].idx = o;
                            seen = true;
                        }
                    }
                    nc += __popc(gballot<G>(gm, seen));
                }
                // ---- collisions + reward (env.py:394-456) ----
                if (!obs_only) {
#pragma unroll
                    for (int off = G >> 1; off > 0; off >>= 1) md = min_keep(md, gshfl_xor<G>(gm, md, off));
                    md = gshfl<G>(gm, md, 0);
                    const bool anycoll = gballot<G>(gm, coll) != 0u;
                    uint32_t f = s_flags[li] & ~(uint32_t)(CAV_FLAG_COLLIDED_AGENT | CAV_FLAG_COLLIDED_WALL);
                    if (anycoll) f |= CAV_FLAG_COLLIDED_AGENT;
                    // wall collision (env.py:494-506): the group's lanes split the cells under the disc
                    bool wall = false;
                    if (wm.bits) wall = gballot<G>(gm, wall_hit(wm, (double)pvh.x, (double)pvh.y, (double)glh.z, glane, G)) != 0u;
                    if (wall) f |= CAV_FLAG_COLLIDED_WALL;
                    T r = p.reward_time_step;
                    if (f & CAV_FLAG_AT_GOAL) {
                        if (!(f & CAV_FLAG_WAS_AT_GOAL_ALREADY)) r = p.reward_at_goal;
                    } else if (!(f & CAV_FLAG_WAS_IN_COLLISION_ALREADY)) {
                        if (anycoll) {
                            r = p.reward_coll_agent;
                            f |= CAV_FLAG_IN_COLLISION;
                        } else if (wall) {
                            r = p.reward_coll_wall;
                            f |= CAV_FLAG_IN_COLLISION;
                        } else {
                            if (md <= p.getting_close) r = (T)-0.1 - md / (T)2.0;
                            if (M<T>::abs_((T)s_adh[li]) > p.wiggly_thr) r += p.reward_wiggly;
                        }
                    }
                    r = r < p.reward_min ? p.reward_min : (r > p.reward_max ? p.reward_max : r);  // np.clip
                    __syncwarp(gm);
                    if (glane == 0) {
                        s_flags[li] = f;
                        g_reward[base + li] = (OT)r;
                        if (g_mind) g_mind[base + li] = md;
                    }
                }
                // ---- neighbour ordering: warp-shuffle bitonic sort (OtherAgentsStatesSensor.py:33-52) ----
                bitonic_sort<G, SL, ObsKeyOps<T>>(gm, glane, key);
                if (nc > p.K) nc = p.K;  // :39
                if (p.sorting == CAV_SORT_CLOSEST_LAST) {
#pragma unroll
                    for (int s = 0; s < SL; ++s) {
                        const int rank = s * G + glane;
                        if (rank < nc) {
                            key[s].kr = -key[s].kr;
                        } else {
                            key[s].tt = M<T>::inf();
                            key[s].kr = M<T>::inf();
                            key[s].po = M<T>::inf();
                            key[s].idx = CAV_MAX_AGENTS + rank;
                        }
                    }
                    bitonic_sort<G, SL, ObsKeyOps<T>>(gm, glane, key);
                }
                // ---- assemble the packed observation in shared memory, then copy out coalesced ----
                const int c_oas = plan.c_state[CAV_OBS_OTHER_AGENTS_STATES];
                if (c_oas >= 0)
                    for (int k = glane; k < 7 * p.K; k += G) my_obs[c_oas + k] = (T)0;  // :108 zero rows
                __syncwarp(gm);
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int rank = s * G + glane;
                    if (rank < nc) {
                        const int o = key[s].idx;
                        const R4<T> pvo = s_pv[eb + o];
                        const T ro = s_goal[eb + o].z;
                        const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                        T row[7];
                        row[0] = rx * prx + ry * pry;
                        row[1] = rx * orx + ry * ory;
                        row[2] = pvo.z * prx + pvo.w * pry;
                        row[3] = pvo.z * orx + pvo.w * ory;
                        row[4] = ro;
                        row[5] = glh.z + ro;
                        row[6] = M<T>::sqrt_(rx * rx + ry * ry) - glh.z - ro;
                        if (c_oas >= 0) {
#pragma unroll
                            for (int k = 0; k < 7; ++k) my_obs[c_oas + 7 * rank + k] = row[k];
                        }
                        if (rank == 0 && plan.oas7_goff >= 0) {
                            // Agent.other_agent_states (persists when nobody is observed): direct global write
                            OT *dst = g_obs + (base + li) * p.obs_len + plan.oas7_goff;
#pragma unroll
                            for (int k = 0; k < 7; ++k) dst[k] = (OT)row[k];
                        }
                    }
                }
                if (glane == 0) {
                    if (plan.c_state[CAV_OBS_IS_LEARNING] >= 0)
                        my_obs[plan.c_state[CAV_OBS_IS_LEARNING]] =
                            (CAV_KIND_POLICY(s_kinds[li]) == CAV_POLICY_LEARNING) ? (T)1 : (T)0;
                    if (plan.c_state[CAV_OBS_NUM_OTHER_AGENTS] >= 0) my_obs[plan.c_state[CAV_OBS_NUM_OTHER_AGENTS]] = (T)nc;
                    if (plan.c_state[CAV_OBS_DIST_TO_GOAL] >= 0) my_obs[plan.c_state[CAV_OBS_DIST_TO_GOAL]] = egh.x;
                    if (plan.c_state[CAV_OBS_HEADING_EGO_FRAME] >= 0)
                        my_obs[plan.c_state[CAV_OBS_HEADING_EGO_FRAME]] = egh.y;
                    if (plan.c_state[CAV_OBS_PREF_SPEED] >= 0) my_obs[plan.c_state[CAV_OBS_PREF_SPEED]] = glh.w;
                    if (plan.c_state[CAV_OBS_RADIUS] >= 0) my_obs[plan.c_state[CAV_OBS_RADIUS]] = glh.z;
                }
                __syncwarp(gm);
                OT *dst = g_obs + (base + li) * p.obs_len;
                for (int sg = 0; sg < plan.nseg; ++sg) {
                    const int go = plan.goff[sg], co = plan.coff[sg], ln = plan.len[sg];
                    for (int k = glane; k < ln; k += G) dst[go + k] = (OT)my_obs[co + k];
                }
                __syncwarp(gm);  // my_obs is reused by this group's next agent
            }
        }
        __syncthreads();

        // ---- done flags + game_over (env.py:514-553) ----
        if (!obs_only) {
            if (tid < count) {
                const int el = tid / N, a = tid - el * N;
                if (a < s_n[el]) {
                    uint32_t f = s_flags[tid];
                    const bool d = (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) != 0;
                    f = d ? (f | CAV_FLAG_DONE) : (f & ~(uint32_t)CAV_FLAG_DONE);
                    s_flags[tid] = f;
                    b.flags[base + tid] = f;
                    b.done[base + tid] = d ? 1 : 0;
                }
            }
            __syncthreads();
            if (tid < ne && s_n[tid] >= 0) {
                const int n = s_n[tid];
                bool all_done = true, all_learning_done = true;
                for (int a = 0; a < n; ++a) {
                    const bool d = (s_flags[tid * N + a] & CAV_FLAG_DONE) != 0;
                    if (!d) {
                        all_done = false;
                        if (CAV_KIND_POLICY(s_kinds[tid * N + a]) == CAV_POLICY_LEARNING) all_learning_done = false;
                    }
                }
                bool go;
                if (p.game_over_mode == CAV_GAME_OVER_ALL_DONE) go = all_done;
                else if (p.game_over_mode == CAV_GAME_OVER_AGENT0) go = (s_flags[tid * N] & CAV_FLAG_DONE) != 0;
                else go = all_learning_done;
                b.game_over[e0 + tid] = go ? 1 : 0;
            }
        }
        __syncthreads();
    }
}

// =============================================================================================================
// host-side launchers
// =============================================================================================================
static void pick_shape(int N, int &G, int &SL) {
    const int others = N - 1 > 2 ? N - 1 : 2;  // >= 2 line slots per group: the anti-collaborative branch builds two
    SL = others > 32 ? 2 : 1;
    G = 1;
    while (G * SL < others && G < 32) G <<= 1;
}

void launch_geometry(int N, int &G, int &SL, int &GPE, int &EPB, int &threads) {
    pick_shape(N, G, SL);
    const int passes = (N * G + 511) / 512;
    GPE = (N + passes - 1) / passes;
    const int tpe = GPE * G;
    EPB = (GPE == N) ? (256 / tpe > 0 ? 256 / tpe : 1) : 1;
    threads = ((EPB * tpe + 31) / 32) * 32;
    if (threads < 32) threads = 32;
}

template <typename T>
size_t smem_policy_move(int cap, int EPB, int N) {
    auto a16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t kd = 0;
    if (N > CAV_KD_MAX_LEAF)
        kd = a16(sizeof(float4) * (size_t)EPB * (2 * N - 1)) + a16(sizeof(KdMeta) * (size_t)EPB * (2 * N - 1)) + a16(cap) +
             a16(EPB);
    return 4 * a16(sizeof(R4<T>) * cap) + a16(sizeof(float4) * cap) + a16(sizeof(float2) * cap) +
           3 * a16(sizeof(float) * cap) + a16(sizeof(int) * EPB) + kd + 64;
}
template <typename T>
size_t smem_pairwise_obs(int cap, int EPB, int groups, int CL) {
    auto a16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    return 3 * a16(sizeof(R4<T>) * cap) + 3 * a16(sizeof(float) * cap) + a16(sizeof(int) * EPB) +
           a16(sizeof(T) * (size_t)groups * CL) + 64;
}

template <typename T, int G, int SL>
static cudaError_t launch_pm(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                             int EPB, int GPE, int threads, int sm_count, cudaStream_t st) {
    const int ntiles = (p.E + EPB - 1) / EPB;
    const size_t smem = smem_policy_move<T>(EPB * p.N, EPB, p.N);
    auto kern = k_policy_move<T, G, SL>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = ntiles < sm_count * 16 ? ntiles : sm_count * 16;
    kern<<<grid, threads, smem, st>>>(make_dev(bufs), p, ext, bufs.kd_order, phases, (T)dt, EPB, GPE, ntiles);
    return cudaGetLastError();
}

template <typename T, typename OT, int G, int SL>
static cudaError_t launch_po(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan, const WallMap &wm,
                             const uint8_t *env_mask, int obs_only, int EPB, int GPE, int threads, int sm_count,
                             cudaStream_t st) {
    const int ntiles = (p.E + EPB - 1) / EPB;
    const size_t smem = smem_pairwise_obs<T>(EPB * p.N, EPB, threads / G, plan.compact_len);
    auto kern = k_pairwise_obs<T, G, SL, OT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = ntiles < sm_count * 16 ? ntiles : sm_count * 16;
    kern<<<grid, threads, smem, st>>>(make_dev(bufs), p, plan, wm, env_mask, obs_only, EPB, GPE, ntiles);
    return cudaGetLastError();
}

#define CAV_DISPATCH_SHAPE(FN, TT, ...)                             \
    do {                                                            \
        if (SL == 2) return FN<TT, 32, 2>(__VA_ARGS__);             \
        switch (G) {                                                \
            case 1: return FN<TT, 1, 1>(__VA_ARGS__);               \
            case 2: return FN<TT, 2, 1>(__VA_ARGS__);               \
            case 4: return FN<TT, 4, 1>(__VA_ARGS__);               \
            case 8: return FN<TT, 8, 1>(__VA_ARGS__);               \
            case 16: return FN<TT, 16, 1>(__VA_ARGS__);             \
            default: return FN<TT, 32, 1>(__VA_ARGS__);             \
        }                                                           \
    } while (0)
#define CAV_COMMA ,

template <typename T>
cudaError_t run_policy_move(const cav_buffers_t &bufs, const KParams<T> &p, const float *ext, int phases, double dt,
                            int sm_count, cudaStream_t st) {
    int G, SL, GPE, EPB, threads;
    launch_geometry(p.N, G, SL, GPE, EPB, threads);
    CAV_DISPATCH_SHAPE(launch_pm, T, bufs, p, ext, phases, dt, EPB, GPE, threads, sm_count, st);
}
template <typename T, typename OT>
cudaError_t run_pairwise_obs(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan, const WallMap &wm,
                             const uint8_t *env_mask, int obs_only, int sm_count, cudaStream_t st) {
    int G, SL, GPE, EPB, threads;
    launch_geometry(p.N, G, SL, GPE, EPB, threads);
    CAV_DISPATCH_SHAPE(launch_po, T CAV_COMMA OT, bufs, p, plan, wm, env_mask, obs_only, EPB, GPE, threads, sm_count, st);
}

template cudaError_t run_policy_move<float>(const cav_buffers_t &, const KParams<float> &, const float *, int, double,
                                            int, cudaStream_t);
template cudaError_t run_policy_move<double>(const cav_buffers_t &, const KParams<double> &, const float *, int,
                                             double, int, cudaStream_t);
template cudaError_t run_pairwise_obs<float, float>(const cav_buffers_t &, const KParams<float> &, const ObsPlan &,
                                                    const WallMap &, const uint8_t *, int, int, cudaStream_t);
template cudaError_t run_pairwise_obs<double, double>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                      const WallMap &, const uint8_t *, int, int, cudaStream_t);
template cudaError_t run_pairwise_obs<double, float>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                     const WallMap &, const uint8_t *, int, int, cudaStream_t);

}  // namespace cav
