// cav_math.cuh -- `real` scalar math, angle helpers, the ego-frame update and the column / trajectory-row accessors of
// the step kernels, as host/device source: nvcc compiles it into every kernel, g++ into tests/native/move_host.cpp, where
// tests/test_move_native_cpu.py runs the kernels' Agent.take_action code against the oracle bit for bit on the CPU.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <math_constants.h>
#endif

#include "cavoid.h"
#include "cav_hd.cuh"

namespace cav {

// ---------------------------------------------------------------------------------------------------------
// real4 columns: 16-byte (float) / 2x16-byte (double) vector loads and stores.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct alignas(16) R4 {
    T x, y, z, w;
};

CAV_HDI R4<float> ld4(const R4<float> *p) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return R4<float>{v.x, v.y, v.z, v.w};
}
CAV_HDI R4<double> ld4(const R4<double> *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0];
    const double2 b = reinterpret_cast<const double2 *>(p)[1];
    return R4<double>{a.x, a.y, b.x, b.y};
}
CAV_HDI void st4(R4<float> *p, const R4<float> &v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v.x, v.y, v.z, v.w);
}
CAV_HDI void st4(R4<double> *p, const R4<double> &v) {
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
    static CAV_HDI float sqrt_(float a) { return sqrtf(a); }
    static CAV_HDI float sin_(float a) { return sinf(a); }
    static CAV_HDI float cos_(float a) { return cosf(a); }
    static CAV_HDI float atan2_(float y, float x) { return atan2f(y, x); }
    static CAV_HDI float abs_(float a) { return fabsf(a); }
    // round-half-even to an integer WITHOUT the XU pipe (FRND): adding and subtracting 1.5 * 2^23 rounds exactly
    // like rintf for |a| < 2^22; larger magnitudes are already integers.  (ncu: the pairwise+obs kernels were bound by
    // the XU pipe -- FRND / F2I / MUFU.RSQ64H -- at 65 % of its peak, profiles/r01b_*.)
    static CAV_HDI float rint_(float a) {
#ifdef __CUDA_ARCH__
        const float r = __fsub_rn(__fadd_rn(fabsf(a), 12582912.0f), 12582912.0f);
#else
        const volatile float t = fabsf(a) + 12582912.0f;
        const float r = t - 12582912.0f;
#endif
        return fabsf(a) < 4194304.0f ? copysignf(r, a) : a;
    }
    static CAV_HDI float fmod_(float a, float b) { return fmodf(a, b); }
    static CAV_HDI float inf() {
#ifdef __CUDA_ARCH__
        return CUDART_INF_F;
#else
        return INFINITY;
#endif
    }
};
template <>
struct M<double> {
    static CAV_HDI double sqrt_(double a) { return sqrt(a); }
    static CAV_HDI double sin_(double a) { return sin(a); }
    static CAV_HDI double cos_(double a) { return cos(a); }
    static CAV_HDI double atan2_(double y, double x) { return atan2(y, x); }
    static CAV_HDI double abs_(double a) { return fabs(a); }
    static CAV_HDI double rint_(double a) {  // see the float version; exact for |a| < 2^51
#ifdef __CUDA_ARCH__
        const double r = __dsub_rn(__dadd_rn(fabs(a), 6755399441055744.0), 6755399441055744.0);
#else
        const volatile double t = fabs(a) + 6755399441055744.0;
        const double r = t - 6755399441055744.0;
#endif
        return fabs(a) < 2251799813685248.0 ? copysign(r, a) : a;
    }
    static CAV_HDI double fmod_(double a, double b) { return fmod(a, b); }
    static CAV_HDI double inf() {
#ifdef __CUDA_ARCH__
        return CUDART_INF;
#else
        return (double)INFINITY;
#endif
    }
};

template <typename T>
CAV_HDI T pi() {
    return (T)3.14159265358979323846;
}
template <typename T>
CAV_HDI T two_pi() {
    return (T)(2.0 * 3.14159265358979323846);
}

// reference util.wrap (util.py:141-146): [-pi, pi).  The reference's while-loops never terminate on inf/nan;
// the device version is bounded (identical results for every finite |a| < 9*pi, which covers every call site:
// arguments are sums of two angles already in [-2*pi, 2*pi]).
template <typename T>
CAV_HDI T wrap(T a) {
#pragma unroll 1
    for (int it = 0; it < 4 && a >= pi<T>(); ++it) a -= two_pi<T>();
#pragma unroll 1
    for (int it = 0; it < 4 && a < -pi<T>(); ++it) a += two_pi<T>();
    return a;
}
template <typename T>
CAV_HDI T sign(T a) {
    return (a > (T)0) ? (T)1 : ((a < (T)0) ? (T)-1 : (T)0);
}
// Python / NumPy float modulo (npy_divmod): result has the sign of the divisor.
template <typename T>
CAV_HDI T pymod(T a, T b) {
    T m = M<T>::fmod_(a, b);
    if (m != (T)0) {
        if ((b < (T)0) != (m < (T)0)) m += b;
    } else {
        m = (b < (T)0) ? (T)-0.0 : (T)0.0;
    }
    return m;
}

// Dynamics.update_ego_frame (dynamics/Dynamics.py:24-41) + Agent.get_ref (agent.py:329-349).
// vel_ego_frame (vex, vey) is only ever stored in the optional aux0 column: with want_vel_ego == false its sqrt + sin +
// cos are skipped (vex = vey = 0), everything else is unchanged.
template <typename T>
CAV_HDI void update_ego(const R4<T> &pv, const R4<T> &goal, T heading, R4<T> &ego, T &vex, T &vey,
                        bool want_vel_ego = true) {
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
    vex = (T)0;
    vey = (T)0;
    if (want_vel_ego) {
        const T cur_speed = M<T>::sqrt_(pv.z * pv.z + pv.w * pv.w);
        vex = cur_speed * M<T>::cos_(he);
        vey = cur_speed * M<T>::sin_(he);
    }
    ego.x = dist;
    ego.y = he;
    ego.z = prx;
    ego.w = pry;
}

// One trajectory row (Agent.to_vector, agent.py:275-289): 11 global + 3 ego values, written with 7 vector stores
// (double2 for float64 rows: 112-byte rows are 16-byte aligned; float2 for float32 rows: 56-byte rows, 8-byte aligned).
CAV_HDI void store_history_row(double *row, const double (&v)[CAV_HISTORY_ROW]) {
#pragma unroll
    for (int k = 0; k < CAV_HISTORY_ROW / 2; ++k) reinterpret_cast<double2 *>(row)[k] = make_double2(v[2 * k], v[2 * k + 1]);
}
CAV_HDI void store_history_row(float *row, const float (&v)[CAV_HISTORY_ROW]) {
#pragma unroll
    for (int k = 0; k < CAV_HISTORY_ROW / 2; ++k) reinterpret_cast<float2 *>(row)[k] = make_float2(v[2 * k], v[2 * k + 1]);
}

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
    T rvo_anti_collab_t;      // Config.RVO_ANTI_COLLAB_T (anti-collaborative RVO only)
    float rvo_anti_p;         // |rvo_collab| when negative: probability of the adversarial choice
    unsigned long long rvo_seed;
    int store_history, history_len;
    int obs_len;
};

}  // namespace cav
