// Host build of the device scenario generator (gym_collision_avoidance_b200/csrc/cav_scenario_gen.cuh): the very
// source the kernel k_generate_scenarios compiles, wrapped for ctypes.  TEST INFRASTRUCTURE
// (tests/test_scenario_native_cpu.py builds it with g++ -ffp-contract=off).
#include "cav_scenario_gen.cuh"

using namespace cav::scenario_gen;

extern "C" {

// Philox4x32-10 block for a counter / key (known-answer tests)
void host_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    Philox g(0, 0);
    g.key0 = key[0];
    g.key1 = key[1];
    g.c0 = ctr[0];
    g.c1 = ctr[1];
    g.c2 = ctr[2];
    g.c3 = ctr[3];
    g.have = 0;
    g.refill();
    for (int i = 0; i < 4; ++i) out[i] = g.out[i];
}

// the first n uniforms of stream (seed, scenario)
void host_uniforms(uint64_t seed, uint64_t scenario, int n, double *out) {
    Philox g(seed, scenario);
    for (int i = 0; i < n; ++i) out[i] = g.rand();
}

// rows [S][N][8], kind [S]: what k_generate_scenarios writes
void host_generate_scenarios(double *rows, int32_t *kind, int S, int N, uint64_t seed, double side_lo, double side_hi,
                             double speed_lo, double speed_hi, double radius_lo, double radius_hi, double close_range) {
    GenParams p;
    p.S = S;
    p.N = N;
    p.side_lo = side_lo;
    p.side_hi = side_hi;
    p.speed_lo = speed_lo;
    p.speed_hi = speed_hi;
    p.radius_lo = radius_lo;
    p.radius_hi = radius_hi;
    p.close_range = close_range;
    p.seed = seed;
    p.random_heading = 0;
    for (int s = 0; s < S; ++s) kind[s] = generate_scenario(rows + (size_t)s * N * CAV_BANK_ROW, p, (uint64_t)s);
}
}
