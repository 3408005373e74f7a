// cav_keys.cuh -- packed 64-bit sort keys of the thread-per-agent neighbour ordering (host/device source).
//
// key = [kint : KB bits | order-preserving bits of p_orth : PB bits | agent index : 6 bits], kint = rint(dist_2_other * 100)
// (NumPy's round(x, 2) is rint(x * 100) / 100), so an unsigned 64-bit compare orders neighbours like the reference's
// stable two-key sort `sorted(key=(round(dist, 2), p_orth))` (OtherAgentsStatesSensor.py:33-35) with ties in agent-index
// order.  float32: 23 + 32 + 6 bits, exact up to the sign of a zero p_orth.  float64: 24 + the TOP 34 bits of p_orth + 6: keys whose kint fields are equal
// are re-compared exactly by the kernels (bounded bubble pass, both precisions), which also covers +0.0 vs -0.0 and a
// kint clamped at the field's range.
// tests/test_keys_native_cpu.py compiles this file with g++ and checks those ordering properties.
#pragma once
#include "cav_hd.cuh"

namespace cav {

CAV_HDI unsigned long long double_bits(double d) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long u;
    std::memcpy(&u, &d, sizeof(u));
    return u;
#endif
}

template <typename T>
struct KeyCodec;
template <>
struct KeyCodec<float> {
    static constexpr int KB = 23;  // |kint| < 2^22: exact magic-number float -> int conversion
    static constexpr int PB = 32;
    static CAV_HDI unsigned long long po_bits(float po) {
        unsigned u = float_bits(po);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        return (unsigned long long)u;
    }
};
template <>
struct KeyCodec<double> {
    static constexpr int KB = 24;
    static constexpr int PB = 34;
    static CAV_HDI unsigned long long po_bits(double po) {
        unsigned long long u = double_bits(po);
        u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
        return u >> 30;
    }
};

// integer value of an integer-valued real with |v| < 2^22, without the XU pipe (F2I): magic-number addition
CAV_HDI int real_to_int(float v) {
#ifdef __CUDA_ARCH__
    return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
#else
    const volatile float s = v + 12582912.0f;
    return (int)float_bits(s) - 0x4B400000;
#endif
}
CAV_HDI int real_to_int(double v) {
#ifdef __CUDA_ARCH__
    return __double2loint(__dadd_rn(v, 6755399441055744.0));
#else
    const volatile double s = v + 6755399441055744.0;
    return (int)(unsigned)(double_bits(s) & 0xffffffffull);
#endif
}

template <typename T>
CAV_HDI unsigned long long make_key(T kint_d, T po, int o, bool descending_kr) {
    constexpr int KB = KeyCodec<T>::KB, PB = KeyCodec<T>::PB;
    const T kmax = (T)(1 << (KB - 1));
    T kc = kint_d < -kmax ? -kmax : (kint_d > kmax - (T)1 ? kmax - (T)1 : kint_d);
    if (!(kc == kc)) kc = kmax - (T)1;  // NaN sorts last
    unsigned long long kb = (unsigned long long)(unsigned)(real_to_int(kc) + (1 << (KB - 1)));
    if (descending_kr) kb = ((1ull << KB) - 1ull) - kb;
    return (kb << (PB + 6)) | (KeyCodec<T>::po_bits(po) << 6) | (unsigned long long)o;
}

}  // namespace cav
