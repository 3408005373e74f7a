// cav_scenario.cu -- scenario banks generated on the device.
//
// One thread builds one scenario: the reference's `get_testcase_random` -> `generate_rand_test_case_multi`
// (reference envs/test_cases.py:212-253; policies/CADRL/scripts/multi/gen_rand_testcases.py: mixture :111-142,
// rejection-sampled random case :144-233, swap case :322-377, circle case :380-423, straight-line test :425-444,
// segment distances :53-108), with the same sequence of uniform draws per scenario, but from a counter-based
// Philox4x32-10 stream keyed by (seed, scenario) instead of NumPy's global MT19937 -- so parity with the host
// generator (envs/scenarios.py, which IS bit-exact against the reference) is statistical, not bitwise
// (tests/test_gpu_scenarios.py).  Agents are placed one after the other, each new start / goal tested against all
// earlier agents, exactly like the reference's loops; the rows go straight into the bank in HBM ([S,N,8] doubles,
// thread-private region, re-read through L1 for the pair tests).
//
// The reference's loop can diverge (its arena side grows 1 % per rejected placement; once every pair can drive
// straight, every placement is rejected): like the host generator, a scenario whose side passed 1000x its initial
// value is dropped and redrawn from the continuing stream.
#include <cstdint>

#include "cav_kernels.h"
#include "cav_scenario_gen.cuh"

namespace cav {

namespace {

using scenario_gen::GenParams;

__global__ void __launch_bounds__(64) k_generate_scenarios(double *__restrict__ rows, int32_t *__restrict__ kind_out,
                                                           GenParams p) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.S) return;
    const int kind = scenario_gen::generate_scenario(rows + (size_t)s * p.N * CAV_BANK_ROW, p, (uint64_t)s);
    if (kind_out) kind_out[s] = kind;
}

}  // namespace

cudaError_t run_generate_scenarios(double *rows, int32_t *kind_out, int S, int N, uint64_t seed, double side_lo,
                                   double side_hi, const double speed_bnds[2], const double radius_bnds[2],
                                   double close_range, int random_heading, cudaStream_t st) {
    GenParams p;
    p.S = S;
    p.N = N;
    p.side_lo = side_lo;
    p.side_hi = side_hi;
    p.speed_lo = speed_bnds[0];
    p.speed_hi = speed_bnds[1];
    p.radius_lo = radius_bnds[0];
    p.radius_hi = radius_bnds[1];
    p.close_range = close_range;
    p.seed = seed;
    p.random_heading = random_heading;
    k_generate_scenarios<<<(S + 63) / 64, 64, 0, st>>>(rows, kind_out, p);
    return cudaGetLastError();
}

}  // namespace cav
