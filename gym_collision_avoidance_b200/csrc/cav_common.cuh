// cav_common.cuh -- shared device helpers for libcavoid_sm100a.so (sm_100a only; compiled with -fmad=false so
// that float/double expressions round exactly like the reference's NumPy float64 / x86-64 float32 arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "cavoid.h"
#include "cav_hd.cuh"  // CAV_MAX_AGENTS, min_keep / max_keep (shared with the host-compiled parts)

#define CAV_MAX_OBS_SEG 10

namespace cav {

// ---------------------------------------------------------------------------------------------------------
// real4 columns: 16-byte (float) / 2x16-byte (double) vector loads and stores.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct alignas(16) R4 {
    T x, y, z, w;
};

__device__ __forceinline__ R4<float> ld4(const R4<float> *p) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return R4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ R4<double> ld4(const R4<double> *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0];
    const double2 b = reinterpret_cast<const double2 *>(p)[1];
    return R4<double>{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st4(R4<float> *p, const R4<float> &v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void st4(R4<double> *p, const R4<double> &v) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v.z, v.w);
}

// ---------------------------------------------------------------------------------------------------------
// scalar math in `real` (IEEE sqrt/div; libm-quality sin/cos/atan2)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct M;
template <>
struct M<float> {
    static __device__ __forceinline__ float sqrt_(float a) { return sqrtf(a); }
    static __device__ __forceinline__ float sin_(float a) { return sinf(a); }
    static __device__ __forceinline__ float cos_(float a) { return cosf(a); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
    // round-half-even to an integer WITHOUT the XU pipe (FRND): adding and subtracting 1.5 * 2^23 rounds exactly
    // like rintf for |a| < 2^22; larger magnitudes are already integers.  (ncu: the pairwise+obs kernels were bound by
    // the XU pipe -- FRND / F2I / MUFU.RSQ64H -- at 65 % of its peak, profiles/r01b_*.)
    static __device__ __forceinline__ float rint_(float a) {
        const float r = __fsub_rn(__fadd_rn(fabsf(a), 12582912.0f), 12582912.0f);
        return fabsf(a) < 4194304.0f ? copysignf(r, a) : a;
    }
    static __device__ __forceinline__ float fmod_(float a, float b) { return fmodf(a, b); }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
template <>
struct M<double> {
    static __device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
    static __device__ __forceinline__ double sin_(double a) { return sin(a); }
    static __device__ __forceinline__ double cos_(double a) { return cos(a); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double abs_(double a) { return fabs(a); }
    static __device__ __forceinline__ double rint_(double a) {  // see the float version; exact for |a| < 2^51
        const double r = __dsub_rn(__dadd_rn(fabs(a), 6755399441055744.0), 6755399441055744.0);
        return fabs(a) < 2251799813685248.0 ? copysign(r, a) : a;
    }
    static __device__ __forceinline__ double fmod_(double a, double b) { return fmod(a, b); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};

template <typename T>
__device__ __forceinline__ T pi() {
    return (T)3.14159265358979323846;
}
template <typename T>
__device__ __forceinline__ T two_pi() {
    return (T)(2.0 * 3.14159265358979323846);
}

// reference util.wrap (util.py:141-146): [-pi, pi).  The reference's while-loops never terminate on inf/nan;
// the device version is bounded (identical results for every finite |a| < 9*pi, which covers every call site:
// arguments are sums of two angles already in [-2*pi, 2*pi]).
template <typename T>
__device__ __forceinline__ T wrap(T a) {
#pragma unroll 1
    for (int it = 0; it < 4 && a >= pi<T>(); ++it) a -= two_pi<T>();
#pragma unroll 1
    for (int it = 0; it < 4 && a < -pi<T>(); ++it) a += two_pi<T>();
    return a;
}
template <typename T>
__device__ __forceinline__ T sign(T a) {
    return (a > (T)0) ? (T)1 : ((a < (T)0) ? (T)-1 : (T)0);
}
// Python / NumPy float modulo (npy_divmod): result has the sign of the divisor.
template <typename T>
__device__ __forceinline__ T pymod(T a, T b) {
    T m = M<T>::fmod_(a, b);
    if (m != (T)0) {
        if ((b < (T)0) != (m < (T)0)) m += b;
    } else {
        m = (b < (T)0) ? (T)-0.0 : (T)0.0;
    }
    return m;
}

// Dynamics.update_ego_frame (dynamics/Dynamics.py:24-41) + Agent.get_ref (agent.py:329-349).
template <typename T>
__device__ __forceinline__ void update_ego(const R4<T> &pv, const R4<T> &goal, T heading, R4<T> &ego, T &vex, T &vey) {
    const T gdx = goal.x - pv.x, gdy = goal.y - pv.y;
    const T dist = M<T>::sqrt_(gdx * gdx + gdy * gdy);
    T prx, pry;
    if (dist > (T)1e-8) {
        prx = gdx / dist;
        pry = gdy / dist;
    } else {
        prx = gdx;
        pry = gdy;
    }
    const T ang = M<T>::atan2_(pry, prx);
    const T he = wrap<T>(heading - ang);
    const T cur_speed = M<T>::sqrt_(pv.z * pv.z + pv.w * pv.w);
    vex = cur_speed * M<T>::cos_(he);
    vey = cur_speed * M<T>::sin_(he);
    ego.x = dist;
    ego.y = he;
    ego.z = prx;
    ego.w = pry;
}

// One trajectory row (Agent.to_vector, agent.py:275-289): 11 global + 3 ego values, written with 7 vector stores
// (double2 for float64 rows: 112-byte rows are 16-byte aligned; float2 for float32 rows: 56-byte rows, 8-byte aligned).
__device__ __forceinline__ void store_history_row(double *row, const double (&v)[CAV_HISTORY_ROW]) {
#pragma unroll
    for (int k = 0; k < CAV_HISTORY_ROW / 2; ++k) reinterpret_cast<double2 *>(row)[k] = make_double2(v[2 * k], v[2 * k + 1]);
}
__device__ __forceinline__ void store_history_row(float *row, const float (&v)[CAV_HISTORY_ROW]) {
#pragma unroll
    for (int k = 0; k < CAV_HISTORY_ROW / 2; ++k) reinterpret_cast<float2 *>(row)[k] = make_float2(v[2 * k], v[2 * k + 1]);
}

// ---------------------------------------------------------------------------------------------------------
// sub-warp groups: G consecutive lanes (G in {1,2,4,8,16,32}) cooperate on one agent.
// ---------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned base = lane & ~(unsigned)(G - 1);
    return ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << base;
}

template <int G, typename V>
__device__ __forceinline__ V gshfl(unsigned mask, V v, int src) {
    if (G == 1) return v;
    return __shfl_sync(mask, v, src, G);
}
template <int G, typename V>
__device__ __forceinline__ V gshfl_xor(unsigned mask, V v, int lanemask) {
    if (G == 1) return v;
    return __shfl_xor_sync(mask, v, lanemask, G);
}
// ballot restricted to the group, shifted so that bit k = group lane k
template <int G>
__device__ __forceinline__ unsigned gballot(unsigned mask, bool pred) {
    if (G == 1) return pred ? 1u : 0u;
    const unsigned b = __ballot_sync(mask, pred);
    const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
    return (G == 32) ? b : ((b >> base) & ((1u << G) - 1u));
}

// Static-map geometry for the wall-collision test (env.py:494-506, Map.py:26-32,52-56); bits == nullptr disables it.
struct WallMap {
    const uint32_t *bits;  // row-major bitmap of Map.static_map, rows x wpr words
    int rows, cols, wpr;
    double origin0, origin1, cell, inv_cell;
};

// cell_coord (Map.world_coordinates_to_map_indices) lives in cav_hd.cuh: it is also compiled by the CPU tests

// collision_with_wall[i] (env.py:494-506): the agent's centre cell is inside the map and some static cell lies under
// its disc mask (col-gy)^2 + (row-gx)^2 < (radius/cell)^2.  Cells are visited `step` apart starting at `first`
// (1 thread: first=0, step=1; G cooperating lanes: first=lane, step=G -> OR the results).
__device__ __forceinline__ bool wall_hit(const WallMap &wm, double px, double py, double radius, int first, int step) {
    const double fx = cell_coord(wm.origin0, py, wm.cell, wm.inv_cell, true);
    const double fy = cell_coord(wm.origin1, px, wm.cell, wm.inv_cell, false);
    if (!(fx >= 0.0 && fy >= 0.0 && fx < (double)wm.rows && fy < (double)wm.cols)) return false;
    const int gr = (int)fx, gc = (int)fy;
    const double rc = radius / wm.cell;
    const double rr2 = rc * rc;
    double rcd = sqrt(rr2);
    if (!(rcd < (double)(wm.rows + wm.cols))) rcd = (double)(wm.rows + wm.cols);
    const int R = (int)rcd + 1;
    const int r0 = max(0, gr - R), r1 = min(wm.rows - 1, gr + R);
    const int c0 = max(0, gc - R), c1 = min(wm.cols - 1, gc + R);
    const int wbox = c1 - c0 + 1, cells = wbox * (r1 - r0 + 1);
    for (int k = first; k < cells; k += step) {
        const int r = r0 + k / wbox, c = c0 + k % wbox;
        const double dc = (double)(c - gc), dr = (double)(r - gr);
        if (dc * dc + dr * dr < rr2 && ((wm.bits[r * wm.wpr + (c >> 5)] >> (c & 31)) & 1u)) return true;
    }
    return false;
}

// Observation packing plan (host-built): which packed-observation segments the pairwise+obs kernel writes.
struct ObsPlan {
    int nseg;                    // merged contiguous segments staged through shared memory
    int goff[CAV_MAX_OBS_SEG];   // global offset inside one agent's observation vector
    int coff[CAV_MAX_OBS_SEG];   // offset inside the compact staging vector
    int len[CAV_MAX_OBS_SEG];
    int compact_len;             // total staged reals per agent
    int c_state[CAV_OBS_NUM_STATES];  // compact offset of each state, -1 when not staged
    int oas7_goff;               // global offset of the (7,) other_agent_states state, -1 when not observed
};

// Output chunks of the thread-per-agent pairwise+obs kernel: compact ranges [c0, c1) that never split a neighbour
// row, each at most `cw` reals wide (host-built from the ObsPlan).
#define CAV_MAX_CHUNKS 40
struct ChunkPlan {
    int n;
    short c0[CAV_MAX_CHUNKS];
    short c1[CAV_MAX_CHUNKS];
};

// Flat copy of the parts of cav_config_t the kernels need, already converted to `real` where it matters.
template <typename T>
struct KParams {
    int E, N;
    T dt;
    T near_goal_thr;  // squared on device like the reference (thr**2)
    T max_heading_change;
    T reward_at_goal, reward_coll_agent, reward_coll_wall, reward_time_step, reward_wiggly, wiggly_thr, getting_close;
    T reward_min, reward_max;
    T sensing_horizon;
    int K, sorting, game_over_mode;
    // RVO (float32 solver)
    float rvo_inv_time_horizon_num;  // timeHorizon (1/x computed on device in float like Agent.cpp:79)
    float rvo_time_step, rvo_collab, rvo_neighbor_dist;
    T rvo_max_delta_heading, rvo_radius_scale, rvo_inv_dt;  // 1/self.dt (RVOPolicy.py:105)
    int rvo_max_neighbors;
    int store_history, history_len;
    int obs_len;
};

}  // namespace cav
