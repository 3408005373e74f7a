// cav_orca.cuh -- group-parallel ORCA / RVO2 velocity solve in float32 for sm_100a.
//
// Replaces, per ego agent, what the reference does through rvo2.PyRVOSimulator.doStep():
//   neighbour ordering      RVO2 Agent.cpp:424-448 (insertAgentNeighbor: ascending distSq, ties in insertion order)
//   ORCA half-planes        RVO2 Agent.cpp:295-366
//   linearProgram1/2/3      RVO2 Agent.cpp:477-630
// Mapping: the G lanes of a group own one ORCA line each (SL lines per lane when N-1 > 32).  The neighbour order
// is produced by a warp-shuffle bitonic sort on a 64-bit key (distSq bits : agent index); LP2's sequential scan
// over lines becomes "ballot for the first violated line at or after the cursor" (skipping satisfied lines is
// exact because the running result does not change while skipping); LP1's loop over earlier lines is a pure
// min / max / or reduction, evaluated across lanes with shuffles (bit-identical to the sequential loop).
// Every float expression keeps the reference's operand order; the translation unit is built with -fmad=false so
// nvcc cannot contract a*b+c, and division / sqrt are IEEE (-prec-div/-prec-sqrt defaults).
#pragma once
#include "cav_common.cuh"
#include "cav_orca_core.cuh"  // Line, fdet / fdot, CAV_RVO_EPSILON

namespace cav {

// ---- generic bitonic sort over G lanes x SL slots; element index e = s*G + lane, ascending in e -----------------
template <int G, int SL, typename KeyOps, typename Key>
__device__ __forceinline__ void bitonic_sort(unsigned gm, int glane, Key (&key)[SL]) {
    constexpr int Mtot = G * SL;
#pragma unroll
    for (int k = 2; k <= Mtot; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < G) {
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int e = s * G + glane;
                    const Key other = KeyOps::shfl_xor(gm, key[s], j, G);
                    const bool up = ((e & k) == 0);
                    const bool lower = ((e & j) == 0);
                    const bool take_min = (up == lower);
                    const bool other_less = KeyOps::less(other, key[s]);
                    // take_min: keep the smaller of (mine, other); else keep the larger
                    if (take_min == other_less) key[s] = other;
                }
            } else {
                // j == G (only when SL == 2): partner is the other slot of the same lane
                if (SL == 2) {
                    const int e0 = glane;  // slot 0 element index, (e0 & j) == 0
                    const bool up = ((e0 & k) == 0);
                    const bool swap = up ? KeyOps::less(key[SL - 1], key[0]) : KeyOps::less(key[0], key[SL - 1]);
                    if (swap) {
                        const Key t = key[0];
                        key[0] = key[SL - 1];
                        key[SL - 1] = t;
                    }
                }
            }
        }
    }
}

struct U64Key {
    static __device__ __forceinline__ bool less(unsigned long long a, unsigned long long b) { return a < b; }
    static __device__ __forceinline__ unsigned long long shfl_xor(unsigned gm, unsigned long long v, int j, int w) {
        return __shfl_xor_sync(gm, v, j, w);
    }
};

// ---- LP2 (and the LP1 it calls) across the group.  Lines live in registers: ln[s] has rank s*G + glane. ---------
// Returns -1 on success, else the rank of the line on which linearProgram1 failed (result restored).
template <int G, int SL>
__device__ __forceinline__ int lp2_group(unsigned gm, int glane, const Line (&ln)[SL], const bool (&valid)[SL],
                                         int n_total, float radius, float ox, float oy, bool directionOpt, float &rx,
                                         float &ry) {
    // Agent.cpp:552-567
    if (directionOpt) {
        rx = ox * radius;
        ry = oy * radius;
    } else if (fdot(ox, oy, ox, oy) > radius * radius) {
        const float inv = 1.0f / sqrtf(fdot(ox, oy, ox, oy));
        rx = (ox * inv) * radius;
        ry = (oy * inv) * radius;
    } else {
        rx = ox;
        ry = oy;
    }
    int start = 0;
#pragma unroll 1
    for (int iter = 0; iter <= G * SL; ++iter) {
        // first violated line with rank >= start (Agent.cpp:569-570)
        int i = -1;
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int rank = s * G + glane;
            const bool v = valid[s] && rank >= start && rank < n_total &&
                           fdet(ln[s].dx, ln[s].dy, ln[s].px - rx, ln[s].py - ry) > 0.0f;
            const unsigned m = gballot<G>(gm, v);
            if (i < 0 && m != 0u) i = s * G + (__ffs(m) - 1);
        }
        if (i < 0) return -1;
        // broadcast line i
        const int si = i / G, li = i - si * G;
        Line L;
        {
            const Line &src = (SL == 2 && si == 1) ? ln[SL - 1] : ln[0];
            L.px = gshfl<G>(gm, src.px, li);
            L.py = gshfl<G>(gm, src.py, li);
            L.dx = gshfl<G>(gm, src.dx, li);
            L.dy = gshfl<G>(gm, src.dy, li);
        }
        // linearProgram1 (Agent.cpp:477-548)
        const float dotProduct = fdot(L.px, L.py, L.dx, L.dy);
        const float discriminant = dotProduct * dotProduct + radius * radius - fdot(L.px, L.py, L.px, L.py);
        if (discriminant < 0.0f) return i;
        const float sqrtDiscriminant = sqrtf(discriminant);
        float tLeft = -dotProduct - sqrtDiscriminant;
        float tRight = -dotProduct + sqrtDiscriminant;
        bool failp = false;
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int rank = s * G + glane;
            if (valid[s] && rank < i) {
                const float denominator = fdet(L.dx, L.dy, ln[s].dx, ln[s].dy);
                const float numerator = fdet(ln[s].dx, ln[s].dy, L.px - ln[s].px, L.py - ln[s].py);
                if (fabsf(denominator) <= CAV_RVO_EPSILON) {
                    if (numerator < 0.0f) failp = true;
                } else {
                    const float t = numerator / denominator;
                    if (denominator >= 0.0f) tRight = min_keep(tRight, t);
                    else tLeft = max_keep(tLeft, t);
                }
            }
        }
#pragma unroll
        for (int off = G >> 1; off > 0; off >>= 1) {
            tRight = min_keep(tRight, gshfl_xor<G>(gm, tRight, off));
            tLeft = max_keep(tLeft, gshfl_xor<G>(gm, tLeft, off));
        }
        const bool anyfail = gballot<G>(gm, failp) != 0u;
        // make the bounds bit-identical on every lane (signed zeros)
        tLeft = gshfl<G>(gm, tLeft, 0);
        tRight = gshfl<G>(gm, tRight, 0);
        if (anyfail || tLeft > tRight) return i;
        if (directionOpt) {
            if (fdot(ox, oy, L.dx, L.dy) > 0.0f) {
                rx = L.px + tRight * L.dx;
                ry = L.py + tRight * L.dy;
            } else {
                rx = L.px + tLeft * L.dx;
                ry = L.py + tLeft * L.dy;
            }
        } else {
            const float t = fdot(L.dx, L.dy, ox - L.px, oy - L.py);
            if (t < tLeft) {
                rx = L.px + tLeft * L.dx;
                ry = L.py + tLeft * L.dy;
            } else if (t > tRight) {
                rx = L.px + tRight * L.dx;
                ry = L.py + tRight * L.dy;
            } else {
                rx = L.px + t * L.dx;
                ry = L.py + t * L.dy;
            }
        }
        start = i + 1;
    }
    return -1;
}

// ---- LP3 (Agent.cpp:583-630, numObstLines == 0) -----------------------------------------------------------------
template <int G, int SL>
__device__ __forceinline__ void lp3_group(unsigned gm, int glane, const Line (&ln)[SL], const bool (&valid)[SL], int nl,
                                          int beginLine, float radius, float &rx, float &ry) {
    float distance = 0.0f;
    int cur = beginLine;
#pragma unroll 1
    for (int iter = 0; iter <= G * SL; ++iter) {
        int i = -1;
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int rank = s * G + glane;
            const bool v = valid[s] && rank >= cur && rank < nl &&
                           fdet(ln[s].dx, ln[s].dy, ln[s].px - rx, ln[s].py - ry) > distance;
            const unsigned m = gballot<G>(gm, v);
            if (i < 0 && m != 0u) i = s * G + (__ffs(m) - 1);
        }
        if (i < 0) return;
        const int si = i / G, li = i - si * G;
        Line L;
        {
            const Line &src = (SL == 2 && si == 1) ? ln[SL - 1] : ln[0];
            L.px = gshfl<G>(gm, src.px, li);
            L.py = gshfl<G>(gm, src.py, li);
            L.dx = gshfl<G>(gm, src.dx, li);
            L.dy = gshfl<G>(gm, src.dy, li);
        }
        Line proj[SL];
        bool pvalid[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int rank = s * G + glane;
            pvalid[s] = valid[s] && rank < i;
            proj[s] = Line{0.f, 0.f, 1.f, 0.f};
            if (pvalid[s]) {
                const float determinant = fdet(L.dx, L.dy, ln[s].dx, ln[s].dy);
                if (fabsf(determinant) <= CAV_RVO_EPSILON) {
                    if (fdot(L.dx, L.dy, ln[s].dx, ln[s].dy) > 0.0f) {
                        pvalid[s] = false;  // same direction: skipped (Agent.cpp:600-603)
                    } else {
                        proj[s].px = 0.5f * (L.px + ln[s].px);
                        proj[s].py = 0.5f * (L.py + ln[s].py);
                    }
                } else {
                    const float sc = fdet(ln[s].dx, ln[s].dy, L.px - ln[s].px, L.py - ln[s].py) / determinant;
                    proj[s].px = L.px + sc * L.dx;
                    proj[s].py = L.py + sc * L.dy;
                }
                const float ddx = ln[s].dx - L.dx, ddy = ln[s].dy - L.dy;
                const float inv = 1.0f / sqrtf(fdot(ddx, ddy, ddx, ddy));
                proj[s].dx = ddx * inv;
                proj[s].dy = ddy * inv;
            }
        }
        const float tx = rx, ty = ry;
        if (lp2_group<G, SL>(gm, glane, proj, pvalid, i, radius, -L.dy, L.dx, true, rx, ry) >= 0) {
            rx = tx;
            ry = ty;
        }
        distance = fdet(L.dx, L.dy, L.px - rx, L.py - ry);
        cur = i + 1;
    }
}

// ---- one ego solve --------------------------------------------------------------------------------------------
// s_f4[a] = (pos.x, pos.y, vel.x, vel.y) as float32, s_frad[a] = float32((1+5e-2)*radius) for the env's agents
// (exactly what RVOPolicy.py:57-74 pushes through the Cython float arguments).
template <int G, int SL>
__device__ __forceinline__ void orca_group(unsigned gm, int glane, int n, int ego, const float4 *s_f4,
                                           const float *s_frad, float prefx, float prefy, float maxSpeed,
                                           float timeHorizon, float timeStep, float collab, float neighborDist,
                                           int maxNeighbors, const KdTree *kd, float &nvx, float &nvy) {
    const float4 me = s_f4[ego];
    const float rad_me = s_frad[ego];
    const float rangeSq = neighborDist * neighborDist;
    unsigned long long key[SL];
    // tie-break of equal distSq = insertion order (Agent.cpp:424-448): the agent index inside a single kd-tree leaf
    // (n <= 10), else the rank of the agent in the tree's visiting order for this ego (KdTree.cpp:266-298); every lane
    // walks the (tiny) tree itself and keeps the ranks of the neighbours it owns
    int ins[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * G + glane;
        ins[s] = (e < ego) ? e : e + 1;
    }
    if (kd) {
        int seq = 0;
        int mine[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) mine[s] = ins[s];
        kd_traverse(*kd, me.x, me.y, [&](int ag) {
#pragma unroll
            for (int s = 0; s < SL; ++s)
                if (ag == mine[s]) ins[s] = seq;
            ++seq;
        });
    }
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * G + glane;
        const int o = (e < ego) ? e : e + 1;
        key[s] = ~0ull;
        if (e < n - 1 && maxNeighbors > 0 && collab != 0.0f) {
            const float4 ot = s_f4[o];
            const float dx = me.x - ot.x, dy = me.y - ot.y;  // absSq(position_ - agent->position_), Agent.cpp:427
            const float distSq = fdot(dx, dy, dx, dy);
            if (distSq < rangeSq)
                key[s] = ((unsigned long long)__float_as_uint(distSq) << 32) | ((unsigned)ins[s] << 8) | (unsigned)o;
        }
    }
    bitonic_sort<G, SL, U64Key>(gm, glane, key);
    Line ln[SL];
    bool valid[SL];
    int nl = 0;
    const float invTimeHorizon = 1.0f / timeHorizon;
    if (collab < 0.0f) {
        // Anti-collaborative (Agent.cpp:368-415): the two legs of the CLOSEST neighbour's velocity obstacle facing inward
        // (ranks 0 and 1; needs G * SL >= 2, see pick_shape), preferred velocity v_other + (-collab) * normalize(relPos).
        const unsigned long long k0 = gshfl<G>(gm, key[0], 0);  // rank 0 = the closest neighbour
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            valid[s] = false;
            ln[s] = Line{0.f, 0.f, 1.f, 0.f};
        }
        if (k0 != ~0ull && maxNeighbors > 0) {
            const int o = (int)(k0 & 0xffull);
            const float4 ot = s_f4[o];
            const float rpx = ot.x - me.x, rpy = ot.y - me.y;
            const float distSq = fdot(rpx, rpy, rpx, rpy);
            const float combinedRadius = rad_me + s_frad[o];
            const float combinedRadiusSq = combinedRadius * combinedRadius;
            const float leg = sqrtf(distSq - combinedRadiusSq);
            const float inv = 1.0f / distSq;
            if (glane < 2) {  // rank = glane (slot 0)
                valid[0] = true;
                ln[0].px = ot.z + rpx;
                ln[0].py = ot.w + rpy;
                if (glane == 0) {
                    ln[0].dx = (-(rpx * leg - rpy * combinedRadius)) * inv;
                    ln[0].dy = (-(rpx * combinedRadius + rpy * leg)) * inv;
                } else {
                    ln[0].dx = (rpx * leg + rpy * combinedRadius) * inv;
                    ln[0].dy = (-rpx * combinedRadius + rpy * leg) * inv;
                }
            }
            nl = 2;
            const float invn = 1.0f / sqrtf(distSq);
            const float sc = -collab;
            prefx = ot.z + sc * (rpx * invn);
            prefy = ot.w + sc * (rpy * invn);
        }
    } else {
    #pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int rank = s * G + glane;
            valid[s] = (key[s] != ~0ull) && rank < maxNeighbors;
            ln[s] = Line{0.f, 0.f, 1.f, 0.f};
            if (valid[s]) {
                const int o = (int)(key[s] & 0xffull);
                const float4 ot = s_f4[o];
                const float rpx = ot.x - me.x, rpy = ot.y - me.y;
                const float rvx = me.z - ot.z, rvy = me.w - ot.w;
                const float distSq = fdot(rpx, rpy, rpx, rpy);
                const float combinedRadius = rad_me + s_frad[o];
                const float combinedRadiusSq = combinedRadius * combinedRadius;
                float ux, uy;
                if (distSq > combinedRadiusSq) {
                    const float wx = rvx - invTimeHorizon * rpx, wy = rvy - invTimeHorizon * rpy;
                    const float wLengthSq = fdot(wx, wy, wx, wy);
                    const float dotProduct1 = fdot(wx, wy, rpx, rpy);
                    if (dotProduct1 < 0.0f && dotProduct1 * dotProduct1 > combinedRadiusSq * wLengthSq) {
                        const float wLength = sqrtf(wLengthSq);
                        const float inv = 1.0f / wLength;
                        const float uwx = wx * inv, uwy = wy * inv;
                        ln[s].dx = uwy;
                        ln[s].dy = -uwx;
                        const float sc = combinedRadius * invTimeHorizon - wLength;
                        ux = sc * uwx;
                        uy = sc * uwy;
                    } else {
                        const float leg = sqrtf(distSq - combinedRadiusSq);
                        const float inv = 1.0f / distSq;
                        if (fdet(rpx, rpy, wx, wy) > 0.0f) {
                            ln[s].dx = (rpx * leg - rpy * combinedRadius) * inv;
                            ln[s].dy = (rpx * combinedRadius + rpy * leg) * inv;
                        } else {
                            ln[s].dx = (-(rpx * leg + rpy * combinedRadius)) * inv;
                            ln[s].dy = (-(-rpx * combinedRadius + rpy * leg)) * inv;
                        }
                        const float dotProduct2 = fdot(rvx, rvy, ln[s].dx, ln[s].dy);
                        ux = dotProduct2 * ln[s].dx - rvx;
                        uy = dotProduct2 * ln[s].dy - rvy;
                    }
                } else {
                    const float invTimeStep = 1.0f / timeStep;
                    const float wx = rvx - invTimeStep * rpx, wy = rvy - invTimeStep * rpy;
                    const float wLength = sqrtf(fdot(wx, wy, wx, wy));
                    const float inv = 1.0f / wLength;
                    const float uwx = wx * inv, uwy = wy * inv;
                    ln[s].dx = uwy;
                    ln[s].dy = -uwx;
                    const float sc = combinedRadius * invTimeStep - wLength;
                    ux = sc * uwx;
                    uy = sc * uwy;
                }
                ln[s].px = me.z + collab * ux;
                ln[s].py = me.w + collab * uy;
            }
            nl += __popc(gballot<G>(gm, valid[s]));
        }
    }
    float rx, ry;
    const int fail = lp2_group<G, SL>(gm, glane, ln, valid, nl, maxSpeed, prefx, prefy, false, rx, ry);
    if (fail >= 0) lp3_group<G, SL>(gm, glane, ln, valid, nl, fail, maxSpeed, rx, ry);
    // every lane holds the same result; pin it to lane 0's bits for group-uniform control flow downstream
    nvx = gshfl<G>(gm, rx, 0);
    nvy = gshfl<G>(gm, ry, 0);
}

}  // namespace cav
