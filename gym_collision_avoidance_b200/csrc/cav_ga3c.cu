// cav_ga3c.cu -- GA3C-CADRL policy network on the device (reference policies/GA3CCADRLPolicy.py:49-84 ->
// policies/GA3C_CADRL/network.py:37-41 `sess.run(Softmax:0, {X: vec_obs})`; the graph is spelled out in include/cavoid.h).
//
// One CTA = 128 agents (rows) x 512 threads (CAV_GA3C_ROWS; 64 x 256 measured 18 % slower), float32 on the CUDA cores: the reference computes in float32 and the policy
// takes an ARGMAX, so the arithmetic stays IEEE float32 (TF32 / BF16 tensor-core inputs would move ~1e-3 of the softmax
// and flip near-ties).  Every layer is a [64 rows] x [K] x [256 columns] product with the same register tile: thread
// (cg = tid & 63, rg = tid >> 6) owns rows rg*16 .. +15 and
//   * LSTM step : hidden unit cg, its four gate columns (the blob stores Wl as [k][unit][i j f o], one LDS.128 per k);
//                 the cell state c of its 16 (row, unit) pairs never leaves registers, h goes through a double buffer in
//                 shared memory laid out [row][unit] (it is the next step's K operand: LDS.128 over k, broadcast to the warp);
//   * FC layers : columns 4*cg .. 4*cg+3 (one LDS.128 of W[k][4 cols] per k, lanes consecutive -> conflict-free),
//                 activations [row][k] in shared memory, overwritten in place behind a barrier.
// 64 FMAs per 5 LDS.128 in the inner loops (explicit fmaf: the library is built with -fmad=false for the ORCA solve).  Weights stream from the (L2-resident, 683 KB) blob into shared memory one
// section at a time: LSTM kernel 71 KB, W1 68 KB, then 64-row slabs of the two 256 x 256 layers.
// Inside the environment a first tiny kernel (k_ga3c_compact) lists the live agents whose policy is GA3C-CADRL, so that CTAs
// are full whatever the policy mix and the share of finished agents (rows are independent; their order is immaterial).
// Sequence handling as dynamic_rnn's: step t updates (c, h) of a row only while t < int(X[row][0]); the loop stops at the
// CTA's longest sequence (the remaining steps would copy state through unchanged).
#include <cuda_runtime.h>
#include <stdint.h>

#include "cavoid.h"
#include "cav_kernels.h"

#ifndef CAV_GA3C_ROWS
#define CAV_GA3C_ROWS 128  // agents per CTA: 128 (512 threads, 225 KB shared, 128 registers: 2.77 ms at 163840 agents) or 64
                           // (256 threads, 149 KB, 152 registers: 3.36 ms) -- profiles/r02m_ga3c_variants.txt
#endif

namespace cav {

namespace ga {
constexpr int XL = CAV_GA3C_X_LEN, T = CAV_GA3C_MAX_OTHERS, H = CAV_GA3C_HIDDEN, F = CAV_GA3C_FC, A = CAV_GA3C_ACTIONS;
constexpr int R = CAV_GA3C_ROWS, NT = 4 * R;  // 16 rows x 64 column groups per 64 threads
constexpr int OFF_AVG = 0, OFF_STD = XL, OFF_WL = 2 * XL, OFF_BL = OFF_WL + (7 + H) * H * 4, OFF_W1 = OFF_BL + H * 4,
              OFF_B1 = OFF_W1 + (4 + H) * F, OFF_W2 = OFF_B1 + F, OFF_B2 = OFF_W2 + F * F, OFF_W3 = OFF_B2 + F,
              OFF_B3 = OFF_W3 + F * F, OFF_WO = OFF_B3 + F, OFF_BO = OFF_WO + F * A;
static_assert(OFF_BO + A + 1 == CAV_GA3C_BLOB_FLOATS, "blob layout");
constexpr int XO_STRIDE = T * 8;  // other-agent states of a row: 19 steps padded from 7 to 8 floats
constexpr int SW_FLOATS = (7 + H) * H * 4;                       // 18176: the largest weight section held at once
constexpr int SX_FLOATS = R * XO_STRIDE + R * 4 + 2 * R * H;     // sXo + sXh + sH[2] = 18176 >= the 64 x 256 activations
static_assert(SX_FLOATS >= R * F, "activation buffer aliases the LSTM operands");
constexpr size_t SMEM_BYTES = (size_t)(SW_FLOATS + SX_FLOATS) * 4 + R * 4 + R * 12 * 4 + R * 4;
}  // namespace ga

struct Ga3cEnvIn {  // MODE 1: rows are the agents of a bound handle
    const void *obs;
    const int32_t *list;   // compacted indices of the live GA3C agents (k_ga3c_compact), list[max_rows] = their count
    float *actions;
    int obs_len, pref_off;
    short src[ga::XL];
};

#ifndef CAV_GA3C_FAST_MATH
#define CAV_GA3C_FAST_MATH 0  // 1: MUFU-based exp / reciprocal in the gates (a few ulp, like Eigen's vectorised ones in TF)
#endif
#if CAV_GA3C_FAST_MATH
__device__ __forceinline__ float ga_sigmoid(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float ga_tanh(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf((1.0f - t) * __frcp_rn(1.0f + t), x);
}
#else
__device__ __forceinline__ float ga_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float ga_tanh(float x) { return tanhf(x); }
#endif

__device__ __forceinline__ void ga_copy_to_smem(float *dst, const float *__restrict__ src, int floats, int tid) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = tid; i < floats / 4; i += ga::NT) d4[i] = __ldg(s4 + i);
}

// acc[r][0..3] += sum over 4*k4n consecutive k of in[row][k] * W[k][4 cols]; `in` rows are `in_stride` floats apart,
// W rows `w_stride4` float4 apart (the thread's float4 column already folded into `w`)
template <int K4N>
__device__ __forceinline__ void ga_tile_fma(float (&acc)[16][4], const float *in, int in_stride, const float4 *w, int w_stride4) {
#pragma unroll 2
    for (int k4 = 0; k4 < K4N; ++k4) {
        const float4 w0 = w[(4 * k4 + 0) * w_stride4], w1 = w[(4 * k4 + 1) * w_stride4], w2 = w[(4 * k4 + 2) * w_stride4],
                     w3 = w[(4 * k4 + 3) * w_stride4];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(in + r * in_stride + 4 * k4);
            acc[r][0] = fmaf(a.x, w0.x, acc[r][0]); acc[r][1] = fmaf(a.x, w0.y, acc[r][1]); acc[r][2] = fmaf(a.x, w0.z, acc[r][2]); acc[r][3] = fmaf(a.x, w0.w, acc[r][3]);
            acc[r][0] = fmaf(a.y, w1.x, acc[r][0]); acc[r][1] = fmaf(a.y, w1.y, acc[r][1]); acc[r][2] = fmaf(a.y, w1.z, acc[r][2]); acc[r][3] = fmaf(a.y, w1.w, acc[r][3]);
            acc[r][0] = fmaf(a.z, w2.x, acc[r][0]); acc[r][1] = fmaf(a.z, w2.y, acc[r][1]); acc[r][2] = fmaf(a.z, w2.z, acc[r][2]); acc[r][3] = fmaf(a.z, w2.w, acc[r][3]);
            acc[r][0] = fmaf(a.w, w3.x, acc[r][0]); acc[r][1] = fmaf(a.w, w3.y, acc[r][1]); acc[r][2] = fmaf(a.w, w3.z, acc[r][2]); acc[r][3] = fmaf(a.w, w3.w, acc[r][3]);
        }
    }
}

template <int MODE, typename OT>
__global__ void __launch_bounds__(ga::NT, 1)
    k_ga3c(const float *__restrict__ blob, const float *__restrict__ X, int rows, float *__restrict__ probs,
           int32_t *__restrict__ action_index, Ga3cEnvIn env) {
    using namespace ga;
    extern __shared__ __align__(16) float smem[];
    float *sW = smem;
    float *sXo = sW + SW_FLOATS;          // [R][19][8] normalised other-agent states
    float *sXh = sXo + R * XO_STRIDE;     // [R][4]     normalised host state (X[1:5])
    float *sH = sXh + R * 4;              // [2][R][H]
    float *sAct = sXo;                    // [R][F], aliases sXo / sXh / sH once the LSTM is done
    int *sLen = reinterpret_cast<int *>(sW + SW_FLOATS + SX_FLOATS);
    float *sLogit = reinterpret_cast<float *>(sLen + R);  // [R][12]
    int *sRow = reinterpret_cast<int *>(sLogit + R * 12);  // global row (agent) index of each tile row, -1 = none

    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * R;
    const int cg = tid & 63, rg = tid >> 6, rbase = rg * 16;

    // ---- which rows run (MODE 1: the compacted list of live agents whose policy is GA3C) ----
    if (MODE == 1) {
        const int count = env.list[rows];
        if (row0 >= count) return;
        if (tid < R) sRow[tid] = row0 + tid < count ? env.list[row0 + tid] : -1;
    } else if (tid < R) {
        sRow[tid] = row0 + tid < rows ? row0 + tid : -1;
    }
    // ---- inputs: (X - avg) / std, sequence length, zero initial h ----
    for (int i = tid; i < R * XO_STRIDE / 4; i += NT) reinterpret_cast<float4 *>(sXo)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < R * H / 4; i += NT) reinterpret_cast<float4 *>(sH)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < R * XL; i += NT) {
        const int r = i & (R - 1), k = i / R;
        const int row = sRow[r];
        float x = 0.0f;
        if (row >= 0) {
            if (MODE == 0) {
                x = X[(size_t)row * XL + k];
            } else {
                const int s = env.src[k];
                if (s >= 0) x = (float)reinterpret_cast<const OT *>(env.obs)[(size_t)row * env.obs_len + s];
            }
        }
        const float xn = (x - __ldg(blob + OFF_AVG + k)) / __ldg(blob + OFF_STD + k);
        if (k == 0) {
            int len = row >= 0 ? (int)x : 0;  // ToInt32(X[:, 0])
            sLen[r] = len < 0 ? 0 : (len > T ? T : len);
        } else if (k < 5) {
            sXh[r * 4 + (k - 1)] = xn;
        } else {
            const int j = k - 5, t = j / 7;
            sXo[r * XO_STRIDE + t * 8 + (j - 7 * t)] = xn;
        }
    }
    ga_copy_to_smem(sW, blob + OFF_WL, SW_FLOATS, tid);
    __syncthreads();
    int tmax = 0;
    for (int r = 0; r < R; ++r) tmax = max(tmax, sLen[r]);

    // ---- LSTM over the other agents (rnn/while) ----
    float c[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) c[r] = 0.0f;
    const float4 bl = __ldg(reinterpret_cast<const float4 *>(blob + OFF_BL) + cg);
    const float4 *sW4 = reinterpret_cast<const float4 *>(sW);
    int cur = 0;
    for (int t = 0; t < tmax; ++t) {
        float acc[16][4];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
        {  // x_t: 7 inputs (k = 0..6)
            float4 w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = sW4[k * H + cg];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float4 x0 = *reinterpret_cast<const float4 *>(sXo + (rbase + r) * XO_STRIDE + t * 8);
                const float4 x1 = *reinterpret_cast<const float4 *>(sXo + (rbase + r) * XO_STRIDE + t * 8 + 4);
                const float xs[7] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z};
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    acc[r][0] = fmaf(xs[k], w[k].x, acc[r][0]); acc[r][1] = fmaf(xs[k], w[k].y, acc[r][1]); acc[r][2] = fmaf(xs[k], w[k].z, acc[r][2]); acc[r][3] = fmaf(xs[k], w[k].w, acc[r][3]);
                }
            }
        }
        ga_tile_fma<H / 4>(acc, sH + cur * R * H + rbase * H, H, sW4 + 7 * H + cg, H);  // h: k = 7 .. 70
        float *hn = sH + (cur ^ 1) * R * H;
        const float *hc = sH + cur * R * H;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int row = rbase + r;
            float h;
            if (t < sLen[row]) {
                const float gi = acc[r][0] + bl.x, gj = acc[r][1] + bl.y, gf = acc[r][2] + bl.z, go = acc[r][3] + bl.w;
                c[r] = ga_sigmoid(gf + 1.0f) * c[r] + ga_sigmoid(gi) * ga_tanh(gj);
                h = ga_sigmoid(go) * ga_tanh(c[r]);
            } else {
                h = hc[row * H + cg];  // Select(time >= sequence_length, old, new)
            }
            hn[row * H + cg] = h;
        }
        __syncthreads();
        cur ^= 1;
    }
    // ---- layer1: relu([Xn[1:5], h] @ W1 + b1) ----
    float acc[16][4];
    const float4 *w4 = sW4 + cg;
    ga_copy_to_smem(sW, blob + OFF_W1, (4 + H) * F, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
    ga_tile_fma<1>(acc, sXh + rbase * 4, 4, w4, F / 4);
    ga_tile_fma<H / 4>(acc, sH + cur * R * H + rbase * H, H, w4 + 4 * (F / 4), F / 4);
    __syncthreads();  // every read of sXh / sH is done: the activations may overwrite them
    {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(blob + OFF_B1) + cg);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            *reinterpret_cast<float4 *>(sAct + (rbase + r) * F + 4 * cg) =
                make_float4(fmaxf(acc[r][0] + b.x, 0.f), fmaxf(acc[r][1] + b.y, 0.f), fmaxf(acc[r][2] + b.z, 0.f), fmaxf(acc[r][3] + b.w, 0.f));
    }
    // ---- layer2, fullyconnected1: relu(a @ W + b), W in four 64-row slabs ----
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
        const float *W = blob + (layer == 0 ? OFF_W2 : OFF_W3);
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
#pragma unroll 1
        for (int slab = 0; slab < F / 64; ++slab) {
            __syncthreads();  // previous slab's readers / the activation writers are done
            ga_copy_to_smem(sW, W + slab * 64 * F, 64 * F, tid);
            __syncthreads();
            ga_tile_fma<16>(acc, sAct + rbase * F + slab * 64, F, w4, F / 4);
        }
        __syncthreads();
        const float4 b = __ldg(reinterpret_cast<const float4 *>(blob + (layer == 0 ? OFF_B2 : OFF_B3)) + cg);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            *reinterpret_cast<float4 *>(sAct + (rbase + r) * F + 4 * cg) =
                make_float4(fmaxf(acc[r][0] + b.x, 0.f), fmaxf(acc[r][1] + b.y, 0.f), fmaxf(acc[r][2] + b.z, 0.f), fmaxf(acc[r][3] + b.w, 0.f));
    }
    // ---- logits_p, softmax, argmax ----
    __syncthreads();
    for (int i = tid; i < F * A; i += NT) sW[i] = __ldg(blob + OFF_WO + i);
    __syncthreads();
    if (cg < A) {
        float lo[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) lo[r] = 0.0f;
#pragma unroll 4
        for (int k = 0; k < F; ++k) {
            const float w = sW[k * A + cg];
#pragma unroll
            for (int r = 0; r < 16; ++r) lo[r] = fmaf(sAct[(rbase + r) * F + k], w, lo[r]);
        }
        const float b = __ldg(blob + OFF_BO + cg);
#pragma unroll
        for (int r = 0; r < 16; ++r) sLogit[(rbase + r) * 12 + cg] = lo[r] + b;
    }
    __syncthreads();
    if (tid < R && sRow[tid] >= 0) {
        const int row = sRow[tid];
        const float *lg = sLogit + tid * 12;
        float mx = lg[0];
        int best = 0;
#pragma unroll
        for (int j = 1; j < A; ++j)
            if (lg[j] > mx) {
                mx = lg[j];
                best = j;
            }
        if (probs) {
            float ex[A], sum = 0.0f;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                ex[j] = expf(lg[j] - mx);
                sum += ex[j];
            }
#pragma unroll
            for (int j = 0; j < A; ++j) probs[(size_t)row * A + j] = ex[j] / sum;
        }
        if (action_index) action_index[row] = best;
        if (MODE == 1) {
            // network.py:7-16: full speed x 5 headings, half speed x 3, zero speed x 3; GA3CCADRLPolicy.py:80-84.  The speed
            // factor is 1, 0.5 or 0, so rounding pref_speed to float32 before the product equals the reference's float64
            // product rounded into its float32 all_actions array (env.py:305-307)
            const double pi6 = 3.14159265358979323846 / 6, pi12 = 3.14159265358979323846 / 12;
            double sp, dh;
            if (best < 5) {
                sp = 1.0;
                dh = -pi6 + best * pi12;
            } else if (best < 8) {
                sp = 0.5;
                dh = -pi6 + (best - 5) * pi6;
            } else {
                sp = 0.0;
                dh = -pi6 + (best - 8) * pi6;
            }
            const float ps = (float)reinterpret_cast<const OT *>(env.obs)[(size_t)row * env.obs_len + env.pref_off];
            reinterpret_cast<float2 *>(env.actions)[row] = make_float2((float)((double)ps * sp), (float)dh);
        }
    }
}

// indices of the live agents whose policy is GA3C-CADRL, in any order; list[rows] counts them (zeroed by the caller)
__global__ void __launch_bounds__(256) k_ga3c_compact(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ kinds,
                                                      const int32_t *__restrict__ num_active, int N, int rows,
                                                      int32_t *__restrict__ list) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    bool act = false;
    if (gi < rows) {
        const int e = gi / N, a = gi - e * N;
        const int n = num_active ? num_active[e] : N;
        act = a < n && CAV_KIND_POLICY(kinds[gi]) == CAV_POLICY_GA3C && !(flags[gi] & CAV_FLAG_DONE);
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(list + rows, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (act) list[base + __popc(m & ((1u << lane) - 1u))] = gi;
}

template <int MODE, typename OT>
static cudaError_t launch_ga3c(const float *blob, const float *X, int rows, float *probs, int32_t *action_index,
                               const Ga3cEnvIn &env, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    auto kern = k_ga3c<MODE, OT>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<(rows + ga::R - 1) / ga::R, ga::NT, ga::SMEM_BYTES, st>>>(blob, X, rows, probs, action_index, env);
    return cudaGetLastError();
}

cudaError_t run_ga3c_forward(const float *blob, const float *X, int rows, float *probs, int32_t *action_index, cudaStream_t st) {
    Ga3cEnvIn env{};
    return launch_ga3c<0, float>(blob, X, rows, probs, action_index, env, st);
}

// the flattened observation the policy hands the network (GA3CCADRLPolicy.py:68-75): Config.STATES_IN_OBS in order
// without is_learning, cut / zero-padded to 138 (network.py:24-35)
void ga3c_gather_table(const cav_config_t &cfg, short *src) {
    int order[CAV_OBS_NUM_STATES], n = 0;
    for (int s = 0; s < CAV_OBS_NUM_STATES; ++s)
        if (cfg.obs_offset[s] >= 0) order[n++] = s;
    for (int i = 1; i < n; ++i)
        for (int j = i; j > 0 && cfg.obs_offset[order[j]] < cfg.obs_offset[order[j - 1]]; --j) {
            const int t = order[j];
            order[j] = order[j - 1];
            order[j - 1] = t;
        }
    int k = 0;
    for (int i = 0; i < n && k < ga::XL; ++i) {
        const int s = order[i];
        const int beg = cfg.obs_offset[s], end = (i + 1 < n) ? cfg.obs_offset[order[i + 1]] : cfg.obs_len;
        if (s == CAV_OBS_IS_LEARNING) continue;  // Config.STATES_NOT_USED_IN_POLICY
        for (int o = beg; o < end && k < ga::XL; ++o) src[k++] = (short)o;
    }
    for (; k < ga::XL; ++k) src[k] = -1;
}

cudaError_t run_ga3c_policy(const cav_buffers_t &bufs, const cav_config_t &cfg, const float *blob, float *probs,
                            int32_t *list, cudaStream_t st) {
    const int nrows = cfg.num_envs * cfg.num_agents;
    cudaError_t ce = cudaMemsetAsync(list + nrows, 0, sizeof(int32_t), st);
    if (ce != cudaSuccess) return ce;
    k_ga3c_compact<<<(nrows + 255) / 256, 256, 0, st>>>(bufs.flags, bufs.kinds, bufs.num_active, cfg.num_agents, nrows, list);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return ce;
    Ga3cEnvIn env{};
    env.obs = bufs.obs;
    env.list = list;
    env.actions = bufs.actions;
    env.obs_len = cfg.obs_len;
    env.pref_off = cfg.obs_offset[CAV_OBS_PREF_SPEED];
    ga3c_gather_table(cfg, env.src);
    const int rows = cfg.num_envs * cfg.num_agents;
    const bool obs32 = cfg.precision == CAV_F32 || cfg.obs_f32;
    return obs32 ? launch_ga3c<1, float>(blob, nullptr, rows, probs, nullptr, env, st)
                 : launch_ga3c<1, double>(blob, nullptr, rows, probs, nullptr, env, st);
}

}  // namespace cav
