// cav_tpa.cu -- one-thread-per-agent ("TPA") variants of the two per-step kernels for small agent counts.
//
//   k_policy_move_tpa   reference collision_avoidance_env.py:304-327 (_take_action): RVOPolicy.py:50-122 + RVO2
//                       Agent.cpp:287-630, NonCooperativePolicy.py:21, StaticPolicy.py:21-22, ExternalPolicy.py:14-16,
//                       LearningPolicy.py:30-32, then Agent.take_action (agent.py:192-241)
//   k_pairwise_obs_tpa  env.py:394-512 (collisions + rewards), OtherAgentsStatesSensor.sense
//                       (OtherAgentsStatesSensor.py:58-144) + get_observation_dict (agent.py:323-327),
//                       _check_which_agents_done (env.py:514-553)
//
// Why a second mapping: with N = 10 the sub-warp-group kernels (cav_step.cu) keep 9 of 16 lanes busy, replicate
// every per-agent scalar 16x and pay ~10 shuffle stages per sort; ncu showed them issue-bound at ~1000 warp
// instructions per agent pair (profiles/r01a_*).  Here a thread owns an agent and walks its N-1 neighbours out of
// the shared-memory copy of its environment's agent block; per-thread scratch (sort keys, ORCA lines, centre
// distances) lives in shared memory laid out [slot][thread] (conflict-free).  The packed observation is
// transposed through a small per-warp shared-memory arena in chunks of three neighbour rows, so that global
// stores are runs of >= 21 consecutive reals per agent instead of one real per 276-byte stride.
// Arithmetic (expression order, -fmad=false, IEEE sqrt/div) is identical to the group kernels and the oracle.
#include <cstdlib>

#include "cav_tpa_common.cuh"

namespace cav {


// =============================================================================================================
// kernel B (TPA): pairwise collisions + rewards + sorted-neighbour observation + done
// =============================================================================================================
// Neighbour ordering key, one 64-bit word per neighbour:  [ kint | p_orth (order-preserving bits) | agent index ].
//   kint = rint(dist_2_other * 100): ordering by round(dist_2_other, 2) == ordering by kint (NumPy rounds as
//   rint(x*100)/100, OtherAgentsStatesSensor.py:103), biased and clamped into KB bits.
//   float32 state: 23 + 32 + 6 bits -- exact.   float64 state: 24 + top 34 bits of p_orth + 6 bits; keys whose kint
//   fields are equal are re-compared exactly (kint as a real, full p_orth, index) by a bounded bubble pass, so the
//   final order is exactly the reference's stable two-key sort also for clamped / truncated fields.
// OT: precision of the packed observation / reward in HBM (the staging arena stays in T; see k_pairwise_obs_tpa_n)
template <typename T, typename OT>
__global__ void __launch_bounds__(128) k_pairwise_obs_tpa(TpaBuffers b, KParams<T> p, ObsPlan plan, WallMap wm,
                                                            const uint8_t *__restrict__ env_mask, int obs_only, int EPB,
                                                            int ntiles, int region_bytes, ChunkPlan chunks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N;
    const int cap = EPB * N;
    const int nt = blockDim.x;
    const int nwarps = (nt + 31) >> 5;
    unsigned char *sp = smem_raw;
    R4<T> *s_pv = reinterpret_cast<R4<T> *>(sp);
    sp += a16(sizeof(R4<T>) * cap);
    T *s_rad = reinterpret_cast<T *>(sp);
    sp += a16(sizeof(T) * cap);
    T *s_dc = reinterpret_cast<T *>(sp);  // [N][nt] centre distance to agent o
    sp += a16(sizeof(T) * (size_t)N * nt);
    unsigned char *s_region = sp;  // [nwarps][region_bytes]: sort keys [slot][lane], later the output arena
    sp += a16((size_t)nwarps * region_bytes);
    uint8_t *s_done = reinterpret_cast<uint8_t *>(sp);  // [cap]
    sp += a16((size_t)cap);
    int *s_n = reinterpret_cast<int *>(sp);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wlanes = min(32, nt - (warp << 5));  // lanes that exist in this warp (the last warp may be partial)
    const unsigned all_lanes = (wlanes == 32) ? 0xffffffffu : ((1u << wlanes) - 1u);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_region + (size_t)warp * region_bytes) + lane;
    T *arena = reinterpret_cast<T *>(s_region + (size_t)warp * region_bytes);
    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    const R4<T> *g_goal = reinterpret_cast<const R4<T> *>(b.goal);
    const R4<T> *g_ego = reinterpret_cast<const R4<T> *>(b.ego);
    OT *g_obs = reinterpret_cast<OT *>(b.obs);
    OT *g_reward = reinterpret_cast<OT *>(b.reward);
    T *g_mind = reinterpret_cast<T *>(b.min_dist);
    const int K = p.K;
    const int CL = plan.compact_len;
    const int c_oas = plan.c_state[CAV_OBS_OTHER_AGENTS_STATES];
    const bool closest_last = (p.sorting == CAV_SORT_CLOSEST_LAST);
    constexpr int KSHIFT = KeyCodec<T>::PB + 6;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int e0 = tile * EPB;
        const int ne = min(EPB, p.E - e0);
        const int count = ne * N;
        const size_t base = (size_t)e0 * N;
        const bool have = tid < count;
        const size_t gi = base + tid;
        R4<T> pvh{}, glh{}, egh{};
        uint32_t f = 0, kind = 0;
        float adh = 0.f;
        if (have) {
            pvh = ld4(g_pv + gi);
            glh = ld4(g_goal + gi);
            egh = ld4(g_ego + gi);
            f = b.flags[gi];
            kind = b.kinds[gi];
            adh = b.actions[2 * gi + 1];
            s_pv[tid] = pvh;
            s_rad[tid] = glh.z;
        }
        if (tid < ne) {
            int n = b.num_active ? b.num_active[e0 + tid] : N;
            if (env_mask && !env_mask[e0 + tid]) n = -1;  // env not selected
            s_n[tid] = n;
        }
        __syncthreads();

        const int el = have ? tid / N : 0, h = tid - el * N;
        const int n = have ? s_n[el] : -1;
        const bool active = have && h < n;
        const int eb = el * N;
        int nc = 0;
        unsigned long long order0 = 0ull, order1 = 0ull;  // sorted neighbour indices, 5 bits each (12 per word)
        const T prx = egh.z, pry = egh.w;
        const T orx = -pry, ory = prx;  // ref_orth, agent.py:347
        if (active) {
            T md = M<T>::inf();
            bool coll = false;
#pragma unroll 1
            for (int o = 0; o < n; ++o) {
                if (o == h) continue;
                const R4<T> pvo = s_pv[eb + o];
                const T ro = s_rad[eb + o];
                const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                const T dc = M<T>::sqrt_(rx * rx + ry * ry);  // util.l2norm == util.vec2_l2_norm
                const T cr = glh.z + ro;                       // env.py:481
                md = min_keep(md, dc - cr);                    // :482-487
                coll = coll || (dc <= cr);                     // :490
                s_dc[o * nt + tid] = dc;
                if (dc > p.sensing_horizon) continue;  // OtherAgentsStatesSensor.py:90
                const T d2o = dc - glh.z - ro;         // :87
                const T po = (rx * orx + ry * ory) + (T)0;  // +0: -0.0 and +0.0 compare equal in the reference
                const unsigned long long key = make_key<T>(M<T>::rint_(d2o * (T)100), po, o, false);
                int i = nc;
#pragma unroll 1
                while (i > 0) {  // insertion, ascending
                    const unsigned long long prev = keys[(i - 1) * 32];
                    if (prev < key) break;
                    keys[i * 32] = prev;
                    --i;
                }
                keys[i * 32] = key;
                ++nc;
            }
            // exact re-comparison inside runs of equal kint (see KeyCodec): bounded bubble passes
#pragma unroll 1
            for (int pass = 0; pass < nc; ++pass) {
                bool swapped = false;
#pragma unroll 1
                for (int r = 0; r + 1 < nc; ++r) {
                    const unsigned long long ka = keys[r * 32], kb = keys[(r + 1) * 32];
                    if ((ka >> KSHIFT) != (kb >> KSHIFT)) continue;
                    const int oa = (int)(ka & 63ull), ob = (int)(kb & 63ull);
                    const R4<T> pa = s_pv[eb + oa], pb = s_pv[eb + ob];
                    const T kia = M<T>::rint_((s_dc[oa * nt + tid] - glh.z - s_rad[eb + oa]) * (T)100);
                    const T kib = M<T>::rint_((s_dc[ob * nt + tid] - glh.z - s_rad[eb + ob]) * (T)100);
                    const T poa = (pa.x - pvh.x) * orx + (pa.y - pvh.y) * ory;
                    const T pob = (pb.x - pvh.x) * orx + (pb.y - pvh.y) * ory;
                    const bool b_first = (kib < kia) || (kib == kia && (pob < poa || (pob == poa && ob < oa)));
                    if (b_first) {
                        keys[r * 32] = kb;
                        keys[(r + 1) * 32] = ka;
                        swapped = true;
                    }
                }
                if (!swapped) break;
            }
            if (nc > K) nc = K;  // :39
            if (closest_last) {
                // re-sort the clipped list by (-kr, po), stable (:42-44): flip the kint field, insertion again
#pragma unroll 1
                for (int j = 0; j < nc; ++j) {
                    unsigned long long key = keys[j * 32];
                    const unsigned long long kb = ((1ull << KeyCodec<T>::KB) - 1ull) - (key >> KSHIFT);
                    key = (kb << KSHIFT) | (key & ((1ull << KSHIFT) - 1ull));
                    int i = j;
#pragma unroll 1
                    while (i > 0) {
                        const unsigned long long prev = keys[(i - 1) * 32];
                        if (prev < key) break;
                        keys[i * 32] = prev;
                        --i;
                    }
                    keys[i * 32] = key;
                }
#pragma unroll 1
                for (int pass = 0; pass < nc; ++pass) {
                    bool swapped = false;
#pragma unroll 1
                    for (int r = 0; r + 1 < nc; ++r) {
                        const unsigned long long ka = keys[r * 32], kb = keys[(r + 1) * 32];
                        if ((ka >> KSHIFT) != (kb >> KSHIFT)) continue;
                        const int oa = (int)(ka & 63ull), ob = (int)(kb & 63ull);
                        const R4<T> pa = s_pv[eb + oa], pb = s_pv[eb + ob];
                        const T kia = M<T>::rint_((s_dc[oa * nt + tid] - glh.z - s_rad[eb + oa]) * (T)100);
                        const T kib = M<T>::rint_((s_dc[ob * nt + tid] - glh.z - s_rad[eb + ob]) * (T)100);
                        const T poa = (pa.x - pvh.x) * orx + (pa.y - pvh.y) * ory;
                        const T pob = (pb.x - pvh.x) * orx + (pb.y - pvh.y) * ory;
                        // descending kr; ties keep the order of the first (closest-first) sort, i.e. (po, index)
                        const bool b_first = (kib > kia) || (kib == kia && (pob < poa || (pob == poa && ob < oa)));
                        if (b_first) {
                            keys[r * 32] = kb;
                            keys[(r + 1) * 32] = ka;
                            swapped = true;
                        }
                    }
                    if (!swapped) break;
                }
            }
#pragma unroll 1
            for (int r = 0; r < nc; ++r) {
                const unsigned long long o = keys[r * 32] & 31ull;
                if (r < 12) order0 |= o << (5 * r);
                else order1 |= o << (5 * (r - 12));
            }
            // ---- collisions + reward (env.py:394-456) ----
            if (!obs_only) {
                f &= ~(uint32_t)(CAV_FLAG_COLLIDED_AGENT | CAV_FLAG_COLLIDED_WALL);
                if (coll) f |= CAV_FLAG_COLLIDED_AGENT;
                const bool wall = wm.bits && wall_hit(wm, (double)pvh.x, (double)pvh.y, (double)glh.z, 0, 1);  // :494-506
                if (wall) f |= CAV_FLAG_COLLIDED_WALL;
                T r = p.reward_time_step;
                if (f & CAV_FLAG_AT_GOAL) {
                    if (!(f & CAV_FLAG_WAS_AT_GOAL_ALREADY)) r = p.reward_at_goal;
                } else if (!(f & CAV_FLAG_WAS_IN_COLLISION_ALREADY)) {
                    if (coll) {
                        r = p.reward_coll_agent;
                        f |= CAV_FLAG_IN_COLLISION;
                    } else if (wall) {
                        r = p.reward_coll_wall;
                        f |= CAV_FLAG_IN_COLLISION;
                    } else {
                        if (md <= p.getting_close) r = (T)-0.1 - md / (T)2.0;
                        if (M<T>::abs_((T)adh) > p.wiggly_thr) r += p.reward_wiggly;
                    }
                }
                r = r < p.reward_min ? p.reward_min : (r > p.reward_max ? p.reward_max : r);  // np.clip
                g_reward[gi] = (OT)r;
                if (g_mind) g_mind[gi] = md;
                // done (env.py:514-536)
                const bool d = (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) != 0;
                f = d ? (f | CAV_FLAG_DONE) : (f & ~(uint32_t)CAV_FLAG_DONE);
                b.flags[gi] = f;
                b.done[gi] = d ? 1 : 0;
                s_done[tid] = d ? 1 : 0;
            }
        }
        // ---- packed observation, transposed through the per-warp arena (aliases the sort keys) ----
        // Chunk c covers compact elements [c0, c1) (whole neighbour rows, at most CW reals): every lane deposits its
        // own slice in arena row `lane` (stride CW+1: conflict-free both ways), then the warp stores agent-major:
        // CW consecutive lanes write consecutive reals of one agent (one <= 128/256-byte run per agent and chunk).
        const unsigned vmask = __ballot_sync(0xffffffffu, active);
        if (vmask != 0u) {  // warp-uniform
            __syncwarp();   // every lane is done with its keys before the arena overwrites them
            constexpr int CW = TpaChunk<T>::CW;
            constexpr int AST = CW + 1;
            constexpr int APL = 32 / CW;  // agents per store instruction
            OT *gwarp = g_obs + (base + (size_t)(warp << 5)) * (size_t)p.obs_len;
            const int sub = lane / CW, kk = lane - sub * CW;
            const bool all_active = (vmask == 0xffffffffu);
#pragma unroll 1
            for (int c = 0; c < chunks.n; ++c) {
                const int c0 = chunks.c0[c], c1 = chunks.c1[c];
                if (active) {
                    T *my = arena + lane * AST - c0;  // my[e] for compact element e in [c0, c1)
                    // scalar states that fall into this chunk (agent.py:323-327)
                    int cs;
                    cs = plan.c_state[CAV_OBS_IS_LEARNING];
                    if (cs >= c0 && cs < c1) my[cs] = (CAV_KIND_POLICY(kind) == CAV_POLICY_LEARNING) ? (T)1 : (T)0;
                    cs = plan.c_state[CAV_OBS_NUM_OTHER_AGENTS];
                    if (cs >= c0 && cs < c1) my[cs] = (T)nc;
                    cs = plan.c_state[CAV_OBS_DIST_TO_GOAL];
                    if (cs >= c0 && cs < c1) my[cs] = egh.x;
                    cs = plan.c_state[CAV_OBS_HEADING_EGO_FRAME];
                    if (cs >= c0 && cs < c1) my[cs] = egh.y;
                    cs = plan.c_state[CAV_OBS_PREF_SPEED];
                    if (cs >= c0 && cs < c1) my[cs] = glh.w;
                    cs = plan.c_state[CAV_OBS_RADIUS];
                    if (cs >= c0 && cs < c1) my[cs] = glh.z;
                    if (c_oas >= 0 && K > 0 && c1 > c_oas && c0 < c_oas + 7 * K) {
                        const int r_lo = (c0 > c_oas) ? (c0 - c_oas + 6) / 7 : 0;     // first row starting in the chunk
                        const int r_hi = min(K, (c1 - c_oas) / 7);                    // one past the last whole row
#pragma unroll 1
                        for (int r = r_lo; r < r_hi; ++r) {
                            T *row = my + (c_oas + 7 * r);
                            if (r < nc) {
                                const int o = (int)(((r < 12) ? (order0 >> (5 * r)) : (order1 >> (5 * (r - 12)))) & 31ull);
                                const R4<T> pvo = s_pv[eb + o];
                                const T ro = s_rad[eb + o];
                                const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                                const T r0 = rx * prx + ry * pry, r1 = rx * orx + ry * ory;
                                const T r2 = pvo.z * prx + pvo.w * pry, r3 = pvo.z * orx + pvo.w * ory;
                                const T r5 = glh.z + ro;
                                const T r6 = s_dc[o * nt + tid] - glh.z - ro;  // np.linalg.norm(..) - r - r (:125)
                                row[0] = r0;
                                row[1] = r1;
                                row[2] = r2;
                                row[3] = r3;
                                row[4] = ro;
                                row[5] = r5;
                                row[6] = r6;
                                if (r == 0 && plan.oas7_goff >= 0) {
                                    // Agent.other_agent_states persists when nobody is observed: direct write
                                    OT *dst = g_obs + gi * (size_t)p.obs_len + plan.oas7_goff;
                                    dst[0] = (OT)r0;
                                    dst[1] = (OT)r1;
                                    dst[2] = (OT)r2;
                                    dst[3] = (OT)r3;
                                    dst[4] = (OT)ro;
                                    dst[5] = (OT)r5;
                                    dst[6] = (OT)r6;
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 7; ++k) row[k] = (T)0;  // zero rows (:108)
                            }
                        }
                    }
                }
                __syncwarp();
                if (c0 + kk < c1) {
                    const int gk = obs_gmap(plan, c0 + kk);
                    const T *src = arena + sub * AST + kk;
                    OT *dst = gwarp + sub * p.obs_len + gk;
                    const int dstep = APL * p.obs_len;
                    if (all_active) {
#pragma unroll 4
                        for (int j = sub; j < 32; j += APL) {
                            *dst = (OT)*src;
                            src += APL * AST;
                            dst += dstep;
                        }
                    } else {
#pragma unroll 1
                        for (int j = sub; j < 32; j += APL) {
                            if ((vmask >> j) & 1u) *dst = (OT)*src;
                            src += APL * AST;
                            dst += dstep;
                        }
                    }
                }
                __syncwarp();
            }
        }
        // ---- game_over (env.py:538-551) ----
        if (!obs_only) {
            __syncthreads();
            if (tid < ne && s_n[tid] >= 0) {
                const int nn = s_n[tid];
                bool all_done = true, all_learning_done = true;
                for (int a = 0; a < nn; ++a) {
                    const bool d = s_done[tid * N + a] != 0;
                    if (!d) {
                        all_done = false;
                        if (CAV_KIND_POLICY(b.kinds[base + (size_t)tid * N + a]) == CAV_POLICY_LEARNING)
                            all_learning_done = false;
                    }
                }
                bool go;
                if (p.game_over_mode == CAV_GAME_OVER_ALL_DONE) go = all_done;
                else if (p.game_over_mode == CAV_GAME_OVER_AGENT0) go = s_done[tid * N] != 0;
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
void tpa_geometry(int N, int &EPB, int &threads) {
    int target = 128;
    if (const char *ev = getenv("CAV_TPA_THREADS")) target = atoi(ev);
    EPB = target / N;
    if (EPB < 1) EPB = 1;
    threads = ((EPB * N + 31) / 32) * 32;  // whole warps: the cooperative copy-out needs all 32 lanes to exist
}

// per-warp shared-memory region of kernel B: sort keys [N-1][32] (8 bytes each), later reused as the output arena
// [32 lanes][CW + 1]
template <typename T>
int tpa_region_bytes(int N, const ObsPlan &, int) {
    const size_t keys = (size_t)8 * (N > 1 ? N - 1 : 1) * 32;
    const size_t arena = sizeof(T) * (size_t)32 * (TpaChunk<T>::CW + 1);
    return (int)a16(keys > arena ? keys : arena);
}

// Row-aligned output chunks of at most cw reals over the compact observation vector [0, CL).
static int build_chunks(const ObsPlan &plan, int K, int cw, ChunkPlan &cp) {
    const int CL = plan.compact_len;
    const int c_oas = plan.c_state[CAV_OBS_OTHER_AGENTS_STATES];
    const int oas_end = (c_oas >= 0) ? c_oas + 7 * K : -1;
    cp.n = 0;
    int c = 0;
    while (c < CL) {
        int e = c + cw < CL ? c + cw : CL;
        if (c_oas >= 0 && e > c_oas && e < oas_end) {
            // do not split a row: pull the end back to the previous row boundary (rows are 7 <= cw wide)
            const int rows = (e - c_oas) / 7;
            const int aligned = c_oas + 7 * rows;
            if (aligned > c) e = aligned;
        }
        if (cp.n >= CAV_MAX_CHUNKS) return -1;
        cp.c0[cp.n] = (short)c;
        cp.c1[cp.n] = (short)e;
        cp.n++;
        c = e;
    }
    return 0;
}

template <typename T>
size_t tpa_smem_pairwise_obs(int N, int EPB, int nt, int region_bytes) {
    const int cap = EPB * N;
    const int nwarps = (nt + 31) / 32;
    return a16(sizeof(R4<T>) * cap) + a16(sizeof(T) * cap) + a16(sizeof(T) * (size_t)N * nt) +
           a16((size_t)nwarps * region_bytes) + a16((size_t)cap) + a16(sizeof(int) * EPB) + 64;
}

template <typename T, typename OT>
cudaError_t run_pairwise_obs_tpa(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan,
                                 const WallMap &wm, const uint8_t *env_mask, int obs_only, int sm_count,
                                 cudaStream_t st) {
    int EPB, threads;
    tpa_geometry(p.N, EPB, threads);
    const int ntiles = (p.E + EPB - 1) / EPB;
    const int rb = tpa_region_bytes<T>(p.N, plan, p.K);
    const size_t smem = tpa_smem_pairwise_obs<T>(p.N, EPB, threads, rb);
    auto kern = k_pairwise_obs_tpa<T, OT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    ChunkPlan cp;
    if (build_chunks(plan, p.K, TpaChunk<T>::CW, cp) != 0) return cudaErrorInvalidValue;
    const int grid = ntiles < sm_count * 16 ? ntiles : sm_count * 16;
    kern<<<grid, threads, smem, st>>>(make_tpa(bufs), p, plan, wm, env_mask, obs_only, EPB, ntiles, rb, cp);
    return cudaGetLastError();
}

size_t tpa_smem_bytes(int N, int precision_f64, const ObsPlan &plan, int K) {
    int EPB, threads;
    tpa_geometry(N, EPB, threads);
    return precision_f64 ? tpa_smem_pairwise_obs<double>(N, EPB, threads, tpa_region_bytes<double>(N, plan, K))
                         : tpa_smem_pairwise_obs<float>(N, EPB, threads, tpa_region_bytes<float>(N, plan, K));
}

template cudaError_t run_pairwise_obs_tpa<float, float>(const cav_buffers_t &, const KParams<float> &, const ObsPlan &,
                                                        const WallMap &, const uint8_t *, int, int, cudaStream_t);
template cudaError_t run_pairwise_obs_tpa<double, double>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                          const WallMap &, const uint8_t *, int, int, cudaStream_t);
template cudaError_t run_pairwise_obs_tpa<double, float>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                         const WallMap &, const uint8_t *, int, int, cudaStream_t);

}  // namespace cav
