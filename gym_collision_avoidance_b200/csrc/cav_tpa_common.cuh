// cav_tpa_common.cuh -- pieces shared by the thread-per-agent kernels (cav_tpa.cu: any N <= 25; cav_tpa_n.cu:
// compile-time N, register-resident).
#pragma once
#include "cav_keys.cuh"
#include "cav_move.cuh"
#include "cav_kernels.h"

namespace cav {

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
    d.episode = b.episode;
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

// KeyCodec / real_to_int / make_key (the packed 64-bit neighbour sort keys) live in cav_keys.cuh: host/device source,
// also compiled by tests/native/keys_host.cpp

// TpaBuffers and tpa_move<T> (Agent.take_action): cav_move.cuh (host/device source)

}  // namespace cav
