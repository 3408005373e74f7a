// Host build of the kernels' time_to_impact (csrc/cav_tti.cuh).  TEST INFRASTRUCTURE (tests/test_tti_native_cpu.py).
#include "cav_tti.cuh"

extern "C" {
void host_tti_f64(int n, const double *in9, double *out) {
    for (int i = 0; i < n; ++i) {
        const double *v = in9 + 9 * i;
        out[i] = cav::time_to_impact<double>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
    }
}
void host_tti_f32(int n, const float *in9, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *v = in9 + 9 * i;
        out[i] = cav::time_to_impact<float>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
    }
}
}
