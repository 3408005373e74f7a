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

// Map.world_coordinates_to_map_indices (Map.py:28-29,36-37): floor(origin -/+ v/cell).  v*inv_cell replaces the division
// unless the result is within 1e-9 of an integer, where the exact quotient decides -- so the cell is the reference's
// for EVERY input (tests/test_geometry_native_cpu.py: random and adversarial values against NumPy's expression).
CAV_HDI double cell_coord(double origin, double v, double cell, double inv_cell, bool negate) {
    const double a = v * inv_cell;
    double c = negate ? (origin - a) : (origin + a);
    if (fabs(c - rint(c)) < 1e-9 * fmax(1.0, fabs(c))) {
        const double q = v / cell;
        c = negate ? (origin - q) : (origin + q);
    }
    return floor(c);
}

}  // namespace cav
