// cav_common.cuh -- shared device helpers for libcavoid_sm100a.so (sm_100a only; compiled with -fmad=false so
// that float/double expressions round exactly like the reference's NumPy float64 / x86-64 float32 arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "cavoid.h"
#include "cav_hd.cuh"  // CAV_MAX_AGENTS, min_keep / max_keep (shared with the host-compiled parts)
#include "cav_math.cuh"

#define CAV_MAX_OBS_SEG 10

namespace cav {

// R4 / ld4 / st4, M<T> scalar math, wrap / sign / pymod, update_ego, store_history_row: cav_math.cuh (host/device source)

// ---------------------------------------------------------------------------------------------------------
// sub-warp groups: G consecutive lanes (G in {1,2,4,8,16,32}) cooperate on one agent.
// ---------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned base = lane & ~(unsigned)(G - 1);
    return ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << base;
}

template <int G, typename V>
__device__ __forceinline__ V gshfl(unsigned mask, V v, int src) {
    if (G == 1) return v;
    return __shfl_sync(mask, v, src, G);
}
template <int G, typename V>
__device__ __forceinline__ V gshfl_xor(unsigned mask, V v, int lanemask) {
    if (G == 1) return v;
    return __shfl_xor_sync(mask, v, lanemask, G);
}
// ballot restricted to the group, shifted so that bit k = group lane k
template <int G>
__device__ __forceinline__ unsigned gballot(unsigned mask, bool pred) {
    if (G == 1) return pred ? 1u : 0u;
    const unsigned b = __ballot_sync(mask, pred);
    const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
    return (G == 32) ? b : ((b >> base) & ((1u << G) - 1u));
}

// Static-map geometry for the wall-collision test (env.py:494-506, Map.py:26-32,52-56); bits == nullptr disables it.
struct WallMap {
    const uint32_t *bits;  // row-major bitmap of Map.static_map, rows x wpr words
    int rows, cols, wpr;
    double origin0, origin1, cell, inv_cell;
};

// cell_coord (Map.world_coordinates_to_map_indices) lives in cav_hd.cuh: it is also compiled by the CPU tests

// collision_with_wall[i] (env.py:494-506): the agent's centre cell is inside the map and some static cell lies under
// its disc mask (col-gy)^2 + (row-gx)^2 < (radius/cell)^2.  Cells are visited `step` apart starting at `first`
// (1 thread: first=0, step=1; G cooperating lanes: first=lane, step=G -> OR the results).
__device__ __forceinline__ bool wall_hit(const WallMap &wm, double px, double py, double radius, int first, int step) {
    const double fx = cell_coord(wm.origin0, py, wm.cell, wm.inv_cell, true);
    const double fy = cell_coord(wm.origin1, px, wm.cell, wm.inv_cell, false);
    if (!(fx >= 0.0 && fy >= 0.0 && fx < (double)wm.rows && fy < (double)wm.cols)) return false;
    const int gr = (int)fx, gc = (int)fy;
    const double rc = radius / wm.cell;
    const double rr2 = rc * rc;
    double rcd = sqrt(rr2);
    if (!(rcd < (double)(wm.rows + wm.cols))) rcd = (double)(wm.rows + wm.cols);
    const int R = (int)rcd + 1;
    const int r0 = max(0, gr - R), r1 = min(wm.rows - 1, gr + R);
    const int c0 = max(0, gc - R), c1 = min(wm.cols - 1, gc + R);
    const int wbox = c1 - c0 + 1, cells = wbox * (r1 - r0 + 1);
    for (int k = first; k < cells; k += step) {
        const int r = r0 + k / wbox, c = c0 + k % wbox;
        const double dc = (double)(c - gc), dr = (double)(r - gr);
        if (dc * dc + dr * dr < rr2 && ((wm.bits[r * wm.wpr + (c >> 5)] >> (c & 31)) & 1u)) return true;
    }
    return false;
}

// Observation packing plan (host-built): which packed-observation segments the pairwise+obs kernel writes.
struct ObsPlan {
    int nseg;                    // merged contiguous segments staged through shared memory
    int goff[CAV_MAX_OBS_SEG];   // global offset inside one agent's observation vector
    int coff[CAV_MAX_OBS_SEG];   // offset inside the compact staging vector
    int len[CAV_MAX_OBS_SEG];
    int compact_len;             // total staged reals per agent
    int c_state[CAV_OBS_NUM_STATES];  // compact offset of each state, -1 when not staged
    int oas7_goff;               // global offset of the (7,) other_agent_states state, -1 when not observed
};

// Output chunks of the thread-per-agent pairwise+obs kernel: compact ranges [c0, c1) that never split a neighbour
// row, each at most `cw` reals wide (host-built from the ObsPlan).
#define CAV_MAX_CHUNKS 40
struct ChunkPlan {
    int n;
    short c0[CAV_MAX_CHUNKS];
    short c1[CAV_MAX_CHUNKS];
};

// KParams<T>: cav_math.cuh

}  // namespace cav
