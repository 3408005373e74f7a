// Host build of the packed neighbour sort keys (csrc/cav_keys.cuh).  TEST INFRASTRUCTURE (tests/test_keys_native_cpu.py).
#include "cav_keys.cuh"

extern "C" {
void host_make_keys_f64(int n, const double *kint, const double *po, const int *idx, int descending, unsigned long long *out) {
    for (int i = 0; i < n; ++i) out[i] = cav::make_key<double>(kint[i], po[i], idx[i], descending != 0);
}
void host_make_keys_f32(int n, const float *kint, const float *po, const int *idx, int descending, unsigned long long *out) {
    for (int i = 0; i < n; ++i) out[i] = cav::make_key<float>(kint[i], po[i], idx[i], descending != 0);
}
int host_key_layout(int is_double, int *kb, int *pb) {
    *kb = is_double ? cav::KeyCodec<double>::KB : cav::KeyCodec<float>::KB;
    *pb = is_double ? cav::KeyCodec<double>::PB : cav::KeyCodec<float>::PB;
    return 0;
}
}
