// cav_tpa_common.cuh -- pieces shared by the thread-per-agent kernels (cav_tpa.cu: any N <= 25; cav_tpa_n.cu:
// compile-time N, register-resident).
#pragma once
#include "cav_kernels.h"

namespace cav {

struct TpaBuffers {
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
};

static inline TpaBuffers make_tpa(const cav_buffers_t &b) {
    TpaBuffers d;
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
    return d;
}

__host__ __device__ __forceinline__ size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }


__device__ __forceinline__ int obs_gmap(const ObsPlan &plan, int c) {  // compact offset -> offset in the obs vector
    if (plan.nseg == 1) return plan.goff[0] + c - plan.coff[0];
    int g = 0;
    for (int sgi = 0; sgi < plan.nseg; ++sgi)
        if (c >= plan.coff[sgi]) g = plan.goff[sgi] + c - plan.coff[sgi];
    return g;
}

// output chunk width (reals): one store instruction writes 32/CW agents x CW consecutive reals
template <typename T>
struct TpaChunk;
template <>
struct TpaChunk<float> {
    static constexpr int CW = 32;
};
template <>
struct TpaChunk<double> {
    static constexpr int CW = 16;  // measured at 16384 x 10, f64: CW 16 -> 50 us, 8 -> 64 us, 32 -> 60 us
};

template <typename T>
struct KeyCodec;
template <>
struct KeyCodec<float> {
    static constexpr int KB = 23;  // |kint| < 2^22: exact magic-number float -> int conversion
    static constexpr int PB = 32;
    static __device__ __forceinline__ unsigned long long po_bits(float po) {
        unsigned u = __float_as_uint(po);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        return (unsigned long long)u;
    }
};
template <>
struct KeyCodec<double> {
    static constexpr int KB = 24;
    static constexpr int PB = 34;
    static __device__ __forceinline__ unsigned long long po_bits(double po) {
        unsigned long long u = (unsigned long long)__double_as_longlong(po);
        u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
        return u >> 30;
    }
};

// integer value of an integer-valued real with |v| < 2^22, without the XU pipe (F2I): magic-number addition
__device__ __forceinline__ int real_to_int(float v) { return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000; }
__device__ __forceinline__ int real_to_int(double v) { return __double2loint(__dadd_rn(v, 6755399441055744.0)); }

template <typename T>
__device__ __forceinline__ unsigned long long make_key(T kint_d, T po, int o, bool descending_kr) {
    constexpr int KB = KeyCodec<T>::KB, PB = KeyCodec<T>::PB;
    const T kmax = (T)(1 << (KB - 1));
    T kc = kint_d < -kmax ? -kmax : (kint_d > kmax - (T)1 ? kmax - (T)1 : kint_d);
    if (!(kc == kc)) kc = kmax - (T)1;  // NaN sorts last
    unsigned long long kb = (unsigned long long)(unsigned)(real_to_int(kc) + (1 << (KB - 1)));
    if (descending_kr) kb = ((1ull << KB) - 1ull) - kb;
    return (kb << (PB + 6)) | (KeyCodec<T>::po_bits(po) << 6) | (unsigned long long)o;
}


}  // namespace cav
