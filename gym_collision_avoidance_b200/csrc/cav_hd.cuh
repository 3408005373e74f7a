// cav_hd.cuh -- host/device plumbing for the parts of the library that are ALSO compiled by g++ into CPU test harnesses
// (tests/native/*.cpp): qualifier macros, a float4 stand-in for the host, and the handful of helpers those parts share
// with the rest of the kernels.  Under nvcc everything here is what it was before (`__device__ __forceinline__` code
// that additionally gets a host twin); the generated SASS is unchanged (tools/sass_all_hist.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CAV_HD __host__ __device__
#define CAV_HDI __host__ __device__ __forceinline__
#define CAV_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CAV_HD
#define CAV_HDI inline
#define CAV_HD_NOINLINE
struct alignas(16) float4 {
    float x, y, z, w;
};
struct alignas(8) float2 {
    float x, y;
};
struct alignas(16) double2 {
    double x, y;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#endif

#ifndef CAV_MAX_AGENTS
#define CAV_MAX_AGENTS 64
#endif

namespace cav {

// std::min / std::max argument-order semantics of the reference C++ (NaN-transparent on the new value)
template <typename T>
CAV_HDI T min_keep(T cur, T v) {
    return (v < cur) ? v : cur;
}
template <typename T>
CAV_HDI T max_keep(T cur, T v) {
    return (cur < v) ? v : cur;
}

CAV_HDI unsigned float_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    unsigned u;
    std::memcpy(&u, &f, sizeof(u));
    return u;
#endif
}

// One uniform [0, 1) draw from Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
// Random123 constants) with counter (c0, c1, c2, c3) and the 64-bit key `seed`: first output word / 2^32.  Used for the
// anti-collaborative RVO agents' policy re-draw (RVOPolicy.py:77-88); oracle/cav_oracle.c carries the same function in C.
CAV_HDI double philox_u01(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (double)c0 * (1.0 / 4294967296.0);
}

// Map.world_coordinates_to_map_indices (Map.py:28-29,36-37): floor(origin -/+ v/cell).  v*inv_cell replaces the division
// unless the result is within 1e-6 of an integer, where the exact quotient decides -- so the cell is the reference's
// for EVERY input (|v*inv_cell - v/cell| <= 2 ulp ~ 1e-13 for map-sized values, far inside the guard band;
// tests/test_geometry_native_cpu.py: random and adversarial values against NumPy's expression).  The floor is two adds
// against 1.5 * 2^52 and one fix-up -- plain float64 pipe work instead of FRND / rint on the quarter-rate XU pipe (ncu,
// round-1 form of this function: 25 % of k_laserscan's instructions).  Valid for |c| < 2^51.
CAV_HDI double floor_small(double c) {
#ifdef __CUDA_ARCH__
    const double r = __dsub_rn(__dadd_rn(c, 6755399441055744.0), 6755399441055744.0);  // round to nearest integer
#else
    const volatile double t = c + 6755399441055744.0;
    const double r = t - 6755399441055744.0;
#endif
    return r > c ? r - 1.0 : r;
}
CAV_HDI double cell_coord(double origin, double v, double cell, double inv_cell, bool negate) {
    const double a = v * inv_cell;
    double c = negate ? (origin - a) : (origin + a);
    double fl = floor_small(c);
    const double fr = c - fl;
    if (!(fr > 1e-6 && fr < 1.0 - 1e-6)) {  // close to a cell border (or not finite / huge): decide with the exact quotient
        const double q = v / cell;
        c = negate ? (origin - q) : (origin + q);
        fl = floor(c);
    }
    return fl;
}

}  // namespace cav
