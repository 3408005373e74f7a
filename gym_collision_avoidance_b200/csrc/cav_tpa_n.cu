// cav_tpa_n.cu -- pairwise collisions + rewards + sorted-neighbour observation + done, one thread per agent, with
// the agent count N a COMPILE-TIME constant (reference env.py:394-512, OtherAgentsStatesSensor.py:58-144,
// agent.py:323-327, env.py:514-553 -- the same contract as k_pairwise_obs_tpa in cav_tpa.cu).
//
// Why: ncu showed k_pairwise_obs_tpa latency-bound (profiles/r01b_*: ~5 warps per scheduler, "wait" and
// short-scoreboard stalls, ~11 cycles per instruction): its per-thread work is one long dependent chain -- a
// run-time loop over neighbours with an insertion sort through shared memory, then run-time loops over output rows.
// With N known at compile time everything unrolls: the N-1 centre distances (each an IEEE sqrt chain) are independent
// instruction streams, the neighbour keys and distances live in registers and are ordered by a branch-free odd-even
// transposition network (no shared-memory round trips, no divergence), and the N-1 observation rows are independent
// as well.  Used for the default observation layout (6 scalars then the (N-1) x 7 neighbour rows, closest_first,
// K = N-1); every other configuration takes the generic kernel.  Arithmetic is identical to the generic kernel.
#include <type_traits>

#include "cav_sortnet.cuh"
#include "cav_tpa_common.cuh"

namespace cav {

// Resident 128-thread equivalents per SM the compiler must allow.  Round 1 (distances in registers): float64 at N = 10
// wanted 138 registers, capped at 5 (<= 102 registers, ~170 B of spills: 44.8 us vs 51.3 / 49.1 / 48.0 us with 1 / 4 / 6).
// Round 2 sweep at 16384 x 10, float64 state + float32 observation (profiles/r02c_libsweep.txt): the spills were the
// long-scoreboard stalls -- parking the nine centre distances in shared memory (CAV_TPA_N_DC_SMEM) leaves 92 registers
// and no spill, and with 2-warp CTAs 33.8 -> 27.6 us (cap 5; 28.4 / 28.9 us with 6 / 7; 32.4 us with 4-warp CTAs).
#ifndef CAV_TPA_N_MINB_F32
#define CAV_TPA_N_MINB_F32 5
#endif
#ifndef CAV_TPA_N_MINB_F64
#define CAV_TPA_N_MINB_F64 5
#endif
#ifndef CAV_TPA_N_WARPS_F64
#define CAV_TPA_N_WARPS_F64 2
#endif
#ifndef CAV_TPA_N_WARPS_F32
#define CAV_TPA_N_WARPS_F32 2
#endif
// Centre distances of the N-1 neighbours: carried through the sorting network next to the keys in registers (0), or
// parked in a per-warp shared-memory table [slot][lane] and fetched by neighbour index after the sort (1) -- 2 x (N-1)
// fewer live registers and half the selects per compare-exchange for float64 state.
#ifndef CAV_TPA_N_DC_SMEM
// centre distances of an agent to its neighbours: 0 = registers, 1 = shared memory, one private column of NO values per
// lane (default), 2 = shared memory, ONE value per unordered pair of the environment (the distance is symmetric bit for
// bit: (-rx)*(-rx) == rx*rx), each lane computing half of its pairs -- half the IEEE float64 square roots and 1.2 KB less
// shared memory per warp, i.e. a tenth CTA per SM.  Measured (profiles/r02r_ab.txt, parity suite green on it): the
// kernel alone 28.3 vs 27.9 us at 16384 x 10, the step 0.0969 vs 0.0984 ms, but 4096 x 10 and 32768 x 10 noncoop ~8 %
// WORSE (an extra __syncwarp and scattered 8-byte reads in front of the sort): stays off.
#define CAV_TPA_N_DC_SMEM 1
#endif
template <typename T>
struct TpaNBounds {
    static constexpr int MINB = CAV_TPA_N_MINB_F32;
};
template <>
struct TpaNBounds<double> {
    static constexpr int MINB = CAV_TPA_N_MINB_F64;
};

// Thread mapping: environments never straddle a warp.  A warp owns EPW = 32 / N consecutive environments, i.e. the
// LW = EPW * N consecutive agents starting at (first env) * N; lanes >= LW idle (2 of 32 at N = 10).  Everything a lane
// needs from other agents is in its own warp, so there is no block barrier anywhere: neighbour states are exchanged
// through a per-warp shared-memory slice behind __syncwarp, game_over is three ballots.  (ncu on the block-wide
// version: a quarter of all stall samples sat behind the two __syncthreads, and its one partially filled warp per CTA
// spent more instructions in the predicated copy-out than the three full warps together.)
//
// Observation stores: the LW agents' packed rows are ONE contiguous block of LW * CL values in HBM.  Each lane writes
// its own row into the warp's staging slice [LW][WS] (WS odd: conflict-free), then the warp streams the slice out with
// consecutive lanes on consecutive addresses -- 128 B (float32) / 256 B (float64) per store instruction, the whole
// 8 KB block in 65 stores instead of ~100 partial-row ones.  float64 stages the row in two column ranges to keep the
// slice at 8 KB per warp (5 CTAs / SM).
template <typename T, int NT, typename OT = T>
struct TpaNShape {
    static constexpr int NO = NT - 1;
    static constexpr int EPW = 32 / NT;
    static constexpr int LW = EPW * NT;
    static constexpr int CL = 6 + 7 * NO;
    static constexpr int P = (sizeof(OT) == 8 && NO >= 2) ? 2 : 1;  // staging passes
    static constexpr int RA = P == 1 ? NO : NO / 2;                // neighbour rows in pass 0
    static constexpr int W0 = 6 + 7 * RA;
    static constexpr int W1 = 7 * (NO - RA);
    static constexpr int WM = W0 > W1 ? W0 : W1;
    static constexpr int WS = WM | 1;                               // staging row stride, odd
    // warps per CTA, measured at 16384 x 10: float64 38.9 / 40.7 / 45.7 us with 4 / 2 / 1; float32 26.0 / 23.6 / 27.9 us
    static constexpr int WARPS = sizeof(T) == 8 ? CAV_TPA_N_WARPS_F64 : CAV_TPA_N_WARPS_F32;
    static constexpr int EPB = WARPS * EPW;
    // float32, single pass, staging stride == packed row length: the slice is a flat image of the HBM block
    static constexpr bool FLAT4 = sizeof(OT) == 4 && P == 1 && WS == CL;
    static constexpr size_t r16(size_t x) { return (x + 15) & ~(size_t)15; }
    static constexpr int NP = NT * (NT - 1) / 2;  // unordered pairs of one environment
    static constexpr size_t DC_BYTES = CAV_TPA_N_DC_SMEM == 2 ? r16(sizeof(T) * (NP > 0 ? NP : 1) * EPW)
                                       : (CAV_TPA_N_DC_SMEM ? r16(sizeof(T) * (NO > 0 ? NO : 1) * 32) : 0);
    static constexpr size_t WARP_BYTES =
        r16(sizeof(R4<T>) * LW) + r16(sizeof(T) * LW) + r16(sizeof(OT) * (LW * WS + 4)) + DC_BYTES;
};

// T = state precision, OT = precision of the packed observation and of the reward (float when the state is float, or when
// cfg.obs_f32 asks for the reference's DECLARED observation dtype -- np.float32, config.py:93-170 -- next to float64 state)
template <typename T, int NT, typename OT>
__global__ void __launch_bounds__(32 * TpaNShape<T, NT, OT>::WARPS, (TpaNBounds<T>::MINB * 4) / TpaNShape<T, NT, OT>::WARPS)
    k_pairwise_obs_tpa_n(TpaBuffers b, KParams<T> p, WallMap wm,
                                                              const uint8_t *__restrict__ env_mask, int obs_only,
                                                              int tile0, int ntiles) {
    using S = TpaNShape<T, NT, OT>;
    constexpr int NO = S::NO, EPW = S::EPW, LW = S::LW, CL = S::CL, WS = S::WS, RA = S::RA;
    constexpr int KSHIFT = KeyCodec<T>::PB + 6;
    constexpr unsigned LWMASK = LW >= 32 ? 0xffffffffu : ((1u << LW) - 1u);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned char *sp = smem_raw + (size_t)warp * S::WARP_BYTES;
    R4<T> *s_pv = reinterpret_cast<R4<T> *>(sp);
    sp += a16(sizeof(R4<T>) * LW);
    T *s_rad = reinterpret_cast<T *>(sp);
    sp += a16(sizeof(T) * LW);
    OT *const stage0 = reinterpret_cast<OT *>(sp);  // [LW][WS] (+ up to 3 values of phase shift, see FLAT4)
    sp += a16(sizeof(OT) * (LW * WS + 4));
    T *const s_dc = reinterpret_cast<T *>(sp);  // centre distances (CAV_TPA_N_DC_SMEM): [NO][32], or [EPW][NP] per pair
#if CAV_TPA_N_DC_SMEM == 2
    const int dc_env = (lane / NT) * S::NP;
    // pair (a < b) of one environment -> a * NT - a * (a + 1) / 2 + (b - a - 1)
    auto dc_of = [&](int o, int hh) -> T {
        const int a = o < hh ? o : hh, bb = o < hh ? hh : o;
        return s_dc[dc_env + a * NT - ((a * (a + 1)) >> 1) + (bb - a - 1)];
    };
#else
    auto dc_of = [&](int o, int hh) -> T { return s_dc[(o - (o > hh ? 1 : 0)) * 32 + lane]; };
#endif
    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    const R4<T> *g_goal = reinterpret_cast<const R4<T> *>(b.goal);
    const R4<T> *g_ego = reinterpret_cast<const R4<T> *>(b.ego);
    OT *g_obs = reinterpret_cast<OT *>(b.obs);
    OT *g_reward = reinterpret_cast<OT *>(b.reward);
    bool bulk_pending = false;  // lane 0: a bulk copy out of the staging slice may still be reading it
    T *g_mind = reinterpret_cast<T *>(b.min_dist);
    const int el = lane / NT, h = lane - el * NT;  // env within the warp, agent within the env
    const int eb = el * NT;

    for (int tile = tile0 + blockIdx.x; tile < ntiles; tile += gridDim.x) {  // tiles [tile0, ntiles) of this launch
        const int ew0 = (tile * S::WARPS + warp) * EPW;  // first environment of this warp
        const int e = ew0 + el;
        const bool have = lane < LW && e < p.E;
        const size_t wbase = (size_t)ew0 * NT;
        const size_t gi = wbase + lane;
        R4<T> pvh{}, glh{}, egh{};
        uint32_t f = 0, kind = 0;
        float adh = 0.f;
        int n = -1;
        // masked launches (the re-observation of freshly reset environments) touch nothing of the other environments
        const bool take = have && !(env_mask && !env_mask[e]);
        if (__ballot_sync(0xffffffffu, take) == 0u) continue;
        if (take) {
            pvh = ld4(g_pv + gi);
            glh = ld4(g_goal + gi);
            egh = ld4(g_ego + gi);
            f = b.flags[gi];
            kind = b.kinds[gi];
            adh = b.actions[2 * gi + 1];
            n = b.num_active ? b.num_active[e] : NT;
            s_pv[lane] = pvh;
            s_rad[lane] = glh.z;
        }
        __syncwarp();

        const bool active = have && h < n;
#if CAV_TPA_N_DC_SMEM == 2
        // every unordered pair once: lane h takes the partners h + 1 .. h + (NT - 1) / 2 (mod NT), and for even NT the
        // opposite agent h + NT / 2 in the lower half of the lanes only
        if (active) {
#pragma unroll
            for (int k = 1; k <= NT / 2; ++k) {
                if (2 * k == NT && h >= NT / 2) continue;
                int o = h + k;
                if (o >= NT) o -= NT;
                if (o < n) {
                    const R4<T> pvo = s_pv[eb + o];
                    const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                    const int a = o < h ? o : h, bb = o < h ? h : o;
                    s_dc[dc_env + a * NT - ((a * (a + 1)) >> 1) + (bb - a - 1)] = M<T>::sqrt_(rx * rx + ry * ry);
                }
            }
        }
        __syncwarp();
#endif
        const T prx = egh.z, pry = egh.w;
        const T orx = -pry, ory = prx;  // ref_orth, agent.py:347
        unsigned long long key[NO > 0 ? NO : 1];
#if !CAV_TPA_N_DC_SMEM
        T dcv[NO > 0 ? NO : 1];
#endif
        int nc = 0;
        bool done_now = false;
        if (active) {
            T md = M<T>::inf();
            bool coll = false;
            // ---- all neighbours at once: slot j <-> other agent o = j + (j >= h) ----
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                const int o = j + (j >= h ? 1 : 0);
                key[j] = ~0ull;
#if !CAV_TPA_N_DC_SMEM
                dcv[j] = (T)0;
#endif
                if (o < n) {
                    const R4<T> pvo = s_pv[eb + o];
                    const T ro = s_rad[eb + o];
                    const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
#if CAV_TPA_N_DC_SMEM == 2
                    const T dc = dc_of(o, h);  // util.l2norm == util.vec2_l2_norm, computed once per pair above
#else
                    const T dc = M<T>::sqrt_(rx * rx + ry * ry);  // util.l2norm == util.vec2_l2_norm
#endif
                    const T cr = glh.z + ro;                       // env.py:481
                    md = min_keep(md, dc - cr);                    // :482-487
                    coll = coll || (dc <= cr);                     // :490
                    if (!(dc > p.sensing_horizon)) {               // OtherAgentsStatesSensor.py:90
                        const T d2o = dc - glh.z - ro;             // :87
                        const T po = (rx * orx + ry * ory) + (T)0;
                        key[j] = make_key<T>(M<T>::rint_(d2o * (T)100), po, o, false);
#if CAV_TPA_N_DC_SMEM == 1
                        s_dc[j * 32 + lane] = dc;
#elif !CAV_TPA_N_DC_SMEM
                        dcv[j] = dc;
#endif
                        ++nc;
                    }
                }
            }
            // ---- order: branch-free sorting network on (key, dc) in registers (keys are distinct: index bits) ----
            sort_network<NO>([&](int i, int j) {
                const bool sw = key[j] < key[i];
                const unsigned long long ka = key[i], kb = key[j];
                key[i] = sw ? kb : ka;
                key[j] = sw ? ka : kb;
#if !CAV_TPA_N_DC_SMEM
                const T da = dcv[i], db = dcv[j];
                dcv[i] = sw ? db : da;
                dcv[j] = sw ? da : db;
#endif
            });
            // exact re-comparison inside runs of equal kint (truncated p_orth / clamped kint, see KeyCodec): rare
            bool tie = false;
#pragma unroll
            for (int r = 0; r + 1 < NO; ++r) tie = tie || ((key[r] >> KSHIFT) == (key[r + 1] >> KSHIFT) && key[r + 1] != ~0ull);
            if (tie) {
#pragma unroll 1
                for (int pass = 0; pass < NO; ++pass) {
                    bool swapped = false;
#pragma unroll
                    for (int r = 0; r + 1 < NO; ++r) {
                        const unsigned long long ka = key[r], kb = key[r + 1];
                        if ((ka >> KSHIFT) == (kb >> KSHIFT) && kb != ~0ull) {
                            const int oa = (int)(ka & 63ull), ob = (int)(kb & 63ull);
                            const R4<T> pa = s_pv[eb + oa], pb = s_pv[eb + ob];
#if CAV_TPA_N_DC_SMEM
                            const T dca = dc_of(oa, h), dcb = dc_of(ob, h);
#else
                            const T dca = dcv[r], dcb = dcv[r + 1];
#endif
                            const T kia = M<T>::rint_((dca - glh.z - s_rad[eb + oa]) * (T)100);
                            const T kib = M<T>::rint_((dcb - glh.z - s_rad[eb + ob]) * (T)100);
                            const T poa = (pa.x - pvh.x) * orx + (pa.y - pvh.y) * ory;
                            const T pob = (pb.x - pvh.x) * orx + (pb.y - pvh.y) * ory;
                            const bool b_first = (kib < kia) || (kib == kia && (pob < poa || (pob == poa && ob < oa)));
                            if (b_first) {
                                key[r] = kb;
                                key[r + 1] = ka;
#if !CAV_TPA_N_DC_SMEM
                                const T da = dcv[r];
                                dcv[r] = dcv[r + 1];
                                dcv[r + 1] = da;
#endif
                                swapped = true;
                            }
                        }
                    }
                    if (!swapped) break;
                }
            }
            if (nc > p.K) nc = p.K;
            // ---- collisions + reward (env.py:394-456) ----
            if (!obs_only) {
                f &= ~(uint32_t)(CAV_FLAG_COLLIDED_AGENT | CAV_FLAG_COLLIDED_WALL);
                if (coll) f |= CAV_FLAG_COLLIDED_AGENT;
                const bool wall = wm.bits && wall_hit(wm, (double)pvh.x, (double)pvh.y, (double)glh.z, 0, 1);
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
                done_now = (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) != 0;
                f = done_now ? (f | CAV_FLAG_DONE) : (f & ~(uint32_t)CAV_FLAG_DONE);
                b.flags[gi] = f;
                b.done[gi] = done_now ? 1 : 0;
            }
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, active);
        // ---- game_over (env.py:538-551): per environment, from ballots over its N lanes ----
        if (!obs_only) {
            const unsigned dmask = __ballot_sync(0xffffffffu, active && done_now);
            const unsigned lmask = __ballot_sync(0xffffffffu, active && CAV_KIND_POLICY(kind) == CAV_POLICY_LEARNING);
            const unsigned pmask = __ballot_sync(0xffffffffu, take && n >= 0);  // environment taking part in this launch
            if (lane < EPW) {
                const int sh = lane * NT;
                const unsigned em = (NT >= 32 ? 0xffffffffu : ((1u << NT) - 1u)) << sh;
                if ((pmask >> sh) & 1u) {
                    const unsigned open = vmask & ~dmask & em;  // live agents that are not done
                    bool go;
                    if (p.game_over_mode == CAV_GAME_OVER_ALL_DONE) go = open == 0u;
                    else if (p.game_over_mode == CAV_GAME_OVER_AGENT0) go = ((dmask >> sh) & 1u) != 0u;
                    else go = (open & lmask) == 0u;
                    b.game_over[ew0 + lane] = go ? 1 : 0;
                }
            }
        }
        // ---- packed observation through the warp's staging slice ----
        if (vmask != 0u) {
            OT *gwarp = g_obs + wbase * (size_t)CL;
            // FLAT4: shift the slice so that shared and global addresses agree modulo 16 bytes (obs is 16-byte aligned)
            const int mis = S::FLAT4 ? (int)((wbase * (size_t)CL) & 3) : 0;
            OT *stage = stage0 + mis;
            OT *my = stage + lane * WS;
            if (S::FLAT4) {  // the previous tile's bulk copy must have finished READING the slice before it is rewritten
                if (bulk_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                bulk_pending = false;
                __syncwarp();
            }
            auto fill_row = [&](int r, OT *dst) {  // r is a compile-time constant at every call site
                if (r < nc) {
                    const int o = (int)(key[r] & 63ull);
                    const R4<T> pvo = s_pv[eb + o];
                    const T ro = s_rad[eb + o];
                    const T rx = pvo.x - pvh.x, ry = pvo.y - pvh.y;
                    dst[0] = (OT)(rx * prx + ry * pry);
                    dst[1] = (OT)(rx * orx + ry * ory);
                    dst[2] = (OT)(pvo.z * prx + pvo.w * pry);
                    dst[3] = (OT)(pvo.z * orx + pvo.w * ory);
                    dst[4] = (OT)ro;
                    dst[5] = (OT)(glh.z + ro);
#if CAV_TPA_N_DC_SMEM
                    dst[6] = (OT)(dc_of(o, h) - glh.z - ro);  // np.linalg.norm(..) - r - r (:125)
#else
                    dst[6] = (OT)(dcv[r] - glh.z - ro);  // np.linalg.norm(..) - r - r (:125)
#endif
                } else {
#pragma unroll
                    for (int k = 0; k < 7; ++k) dst[k] = (OT)0;  // zero rows (:108)
                }
            };
            // columns [c0, c0 + W) of every agent of the warp: element idx = agent * W + column
            auto stream_out = [&](auto c0_, auto w_) {
                constexpr int c0 = decltype(c0_)::value, W = decltype(w_)::value;
                constexpr int TOT = LW * W;
                if (vmask == LWMASK && S::FLAT4) {
                    // One contiguous block with the same 16-byte phase in shared and global memory: its 16-byte-aligned
                    // body (8 KB at N = 10) leaves as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA
                    // engine; SASS UBLKCP) issued by lane 0 instead of 17 LDS.128 + STG.128 per lane; up to 3 values of
                    // head and tail go through ordinary stores.  Writers fence their shared-memory stores towards the
                    // async proxy, the warp converges, lane 0 issues; the slice is reused only after wait_group.read.
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    const int head = (4 - mis) & 3;
                    const int nvec = (TOT - head) >> 2;
                    if (lane == 0) {
                        const unsigned src = (unsigned)__cvta_generic_to_shared(stage + head);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gwarp + head),
                                     "r"(src), "r"(nvec * 16)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        bulk_pending = true;
                    }
                    if (lane < head) gwarp[lane] = stage[lane];
                    const int t = head + 4 * nvec + lane;
                    if (t < TOT) gwarp[t] = stage[t];
                    return;
                }
                __syncwarp();
                if (vmask == LWMASK) {
                    // element idx = 32 t + lane sits in agent row j = idx / W; W > 32, so j is q_t or q_t + 1 with
                    // compile-time q_t and a per-lane carry (no division in the unrolled code)
#pragma unroll
                    for (int t = 0; t < (TOT + 31) / 32; ++t) {
                        const int idx = 32 * t + lane;
                        int j;
                        if (W > 32) j = (32 * t) / W + ((lane >= W - (32 * t) % W) ? 1 : 0);
                        else j = idx / W;
                        if (idx < TOT) gwarp[idx + c0 + j * (CL - W)] = stage[idx + j * (WS - W)];
                    }
                } else {  // ragged / masked environments: inactive agents keep their old observation
#pragma unroll 1
                    for (int i0 = 0; i0 < TOT; i0 += 32) {
                        const int idx = i0 + lane;
                        if (idx < TOT) {
                            const int j = idx / W, k = idx - j * W;
                            if ((vmask >> j) & 1u) gwarp[j * CL + c0 + k] = stage[j * WS + k];
                        }
                    }
                }
                __syncwarp();
            };
            // pass 0: the six scalars (agent.py:323-327, default STATES_IN_OBS order) + rows [0, RA)
            if (active) {
                my[0] = (CAV_KIND_POLICY(kind) == CAV_POLICY_LEARNING) ? (OT)1 : (OT)0;  // is_learning
                my[1] = (OT)nc;                                                           // num_other_agents
                my[2] = (OT)egh.x;                                                        // dist_to_goal
                my[3] = (OT)egh.y;                                                        // heading_ego_frame
                my[4] = (OT)glh.w;                                                        // pref_speed
                my[5] = (OT)glh.z;                                                        // radius
#pragma unroll
                for (int r = 0; r < RA; ++r) fill_row(r, my + 6 + 7 * r);
            }
            stream_out(std::integral_constant<int, 0>{}, std::integral_constant<int, S::W0>{});
            if (S::P == 2) {  // pass 1: rows [RA, NO)
                if (active) {
#pragma unroll
                    for (int r = RA; r < NO; ++r) fill_row(r, my + 7 * (r - RA));
                }
                stream_out(std::integral_constant<int, S::W0>{}, std::integral_constant<int, (S::W1 > 0 ? S::W1 : 1)>{});
            }
        }
        __syncwarp();
    }
    if (S::FLAT4 && bulk_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// true when the plan is the default layout this kernel hard-codes
static bool standard_layout(const ObsPlan &plan, const KParams<float> *, int N, int K, int obs_len, int sorting) {
    if (K != N - 1 || sorting != CAV_SORT_CLOSEST_FIRST || plan.oas7_goff >= 0) return false;
    if (plan.nseg != 1 || plan.goff[0] != 0 || plan.coff[0] != 0) return false;
    if (plan.compact_len != 6 + 7 * K || obs_len != plan.compact_len) return false;
    const int want[6] = {CAV_OBS_IS_LEARNING, CAV_OBS_NUM_OTHER_AGENTS, CAV_OBS_DIST_TO_GOAL, CAV_OBS_HEADING_EGO_FRAME,
                         CAV_OBS_PREF_SPEED, CAV_OBS_RADIUS};
    for (int i = 0; i < 6; ++i)
        if (plan.c_state[want[i]] != i) return false;
    return plan.c_state[CAV_OBS_OTHER_AGENTS_STATES] == 6;
}

template <typename T, int NT, typename OT>
static cudaError_t launch_n(const cav_buffers_t &bufs, const KParams<T> &p, const WallMap &wm, const uint8_t *env_mask,
                            int obs_only, int sm_count, cudaStream_t st, int env0, int env1, int *tile_envs) {
    using S = TpaNShape<T, NT, OT>;
    if (tile_envs) {  // query only: environments per tile of this instantiation
        *tile_envs = S::EPB;
        return cudaSuccess;
    }
    if (env1 < 0 || env1 > p.E) env1 = p.E;
    const int tile0 = env0 / S::EPB;
    const int ntiles = (env1 + S::EPB - 1) / S::EPB;
    if (ntiles <= tile0) return cudaSuccess;
    const size_t smem = S::WARP_BYTES * S::WARPS;
    if (smem > 48 * 1024) {  // per device, so not cached in a process-wide flag (no current shape needs it: <= 38.4 KB)
        cudaError_t e = cudaFuncSetAttribute(k_pairwise_obs_tpa_n<T, NT, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = (ntiles - tile0) < sm_count * 16 ? (ntiles - tile0) : sm_count * 16;
    k_pairwise_obs_tpa_n<T, NT, OT><<<grid, 32 * S::WARPS, smem, st>>>(make_tpa(bufs), p, wm, env_mask, obs_only, tile0, ntiles);
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when no specialisation applies (the caller then runs the generic kernel).
// env range [env0, env1) (env0 a multiple of the tile; env1 < 0 = all); tile_envs != nullptr only queries the tile size
template <typename T, typename OT>
cudaError_t run_pairwise_obs_tpa_n(const cav_buffers_t &bufs, const KParams<T> &p, const ObsPlan &plan, const WallMap &wm,
                                   const uint8_t *env_mask, int obs_only, int sm_count, cudaStream_t st, int env0, int env1,
                                   int *tile_envs) {
    if (!standard_layout(plan, nullptr, p.N, p.K, p.obs_len, p.sorting)) return cudaErrorNotSupported;
    switch (p.N) {
#define CAV_CASE(NN) \
    case NN: return launch_n<T, NN, OT>(bufs, p, wm, env_mask, obs_only, sm_count, st, env0, env1, tile_envs);
        CAV_CASE(2)
        CAV_CASE(3)
        CAV_CASE(4)
        CAV_CASE(5)
        CAV_CASE(6)
        CAV_CASE(8)
        CAV_CASE(10)
        CAV_CASE(12)
#undef CAV_CASE
        default: return cudaErrorNotSupported;
    }
}

template cudaError_t run_pairwise_obs_tpa_n<float, float>(const cav_buffers_t &, const KParams<float> &, const ObsPlan &,
                                                          const WallMap &, const uint8_t *, int, int, cudaStream_t, int, int, int *);
template cudaError_t run_pairwise_obs_tpa_n<double, double>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                            const WallMap &, const uint8_t *, int, int, cudaStream_t, int, int, int *);
template cudaError_t run_pairwise_obs_tpa_n<double, float>(const cav_buffers_t &, const KParams<double> &, const ObsPlan &,
                                                           const WallMap &, const uint8_t *, int, int, cudaStream_t, int, int, int *);

}  // namespace cav
