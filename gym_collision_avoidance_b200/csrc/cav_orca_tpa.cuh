// cav_orca_tpa.cuh -- one-thread-per-agent ORCA / RVO2 velocity solve in float32 (sm_100a).
//
// Same arithmetic as cav_orca.cuh (the sub-warp-group solver used for large N) and as the reference
// (RVO2 Agent.cpp:295-366 lines, :424-448 neighbour order, :477-630 LP1/LP2/LP3), but one thread owns one ego
// agent and walks its N-1 lines sequentially.  Lines live in a shared-memory array laid out [line][thread]
// (float4 per entry: conflict-free, consecutive threads -> consecutive 16-byte words); LP3's projected lines,
// which are touched only on the rare infeasible solves, live in thread-local memory.
// For small N this mapping needs ~10x fewer issue slots per agent than the group mapping (measured, see
// profiles/ and DESIGN.md): no lane sits idle and nothing is replicated across lanes.
#pragma once
#include "cav_orca.cuh"

namespace cav {

struct SmemLines {  // line i of this thread at base[i * stride]
    float4 *base;
    int stride;
    __device__ __forceinline__ Line get(int i) const {
        const float4 v = base[i * stride];
        return Line{v.x, v.y, v.z, v.w};
    }
    __device__ __forceinline__ void set(int i, const Line &l) { base[i * stride] = make_float4(l.px, l.py, l.dx, l.dy); }
};
struct LocalLines {
    Line *p;
    __device__ __forceinline__ Line get(int i) const { return p[i]; }
    __device__ __forceinline__ void set(int i, const Line &l) { p[i] = l; }
};

// RVO2 Agent.cpp:477-548
template <typename Lines>
__device__ __forceinline__ bool tpa_lp1(const Lines &lines, int lineNo, float radius, float ox, float oy,
                                        bool directionOpt, float &rx, float &ry) {
    const Line L = lines.get(lineNo);
    const float dotProduct = fdot(L.px, L.py, L.dx, L.dy);
    const float discriminant = dotProduct * dotProduct + radius * radius - fdot(L.px, L.py, L.px, L.py);
    if (discriminant < 0.0f) return false;
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
#pragma unroll 1
    for (int i = 0; i < lineNo; ++i) {
        const Line li = lines.get(i);
        const float denominator = fdet(L.dx, L.dy, li.dx, li.dy);
        const float numerator = fdet(li.dx, li.dy, L.px - li.px, L.py - li.py);
        if (fabsf(denominator) <= CAV_RVO_EPSILON) {
            if (numerator < 0.0f) return false;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = min_keep(tRight, t);
        else tLeft = max_keep(tLeft, t);
        if (tLeft > tRight) return false;
    }
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
    return true;
}

// RVO2 Agent.cpp:550-581; returns n on success, else the index of the failing line
template <typename Lines>
__device__ __forceinline__ int tpa_lp2(const Lines &lines, int n, float radius, float ox, float oy, bool directionOpt,
                                       float &rx, float &ry) {
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
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const Line li = lines.get(i);
        if (fdet(li.dx, li.dy, li.px - rx, li.py - ry) > 0.0f) {
            const float tx = rx, ty = ry;
            if (!tpa_lp1(lines, i, radius, ox, oy, directionOpt, rx, ry)) {
                rx = tx;
                ry = ty;
                return i;
            }
        }
    }
    return n;
}

// RVO2 Agent.cpp:583-630 (numObstLines == 0)
template <typename Lines>
__device__ __noinline__ void tpa_lp3(const Lines &lines, int n, int beginLine, float radius, float &rx, float &ry) {
    float distance = 0.0f;
    Line projbuf[CAV_MAX_AGENTS];
    LocalLines proj{projbuf};
#pragma unroll 1
    for (int i = beginLine; i < n; ++i) {
        const Line Li = lines.get(i);
        if (fdet(Li.dx, Li.dy, Li.px - rx, Li.py - ry) > distance) {
            int np = 0;
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
                const Line Lj = lines.get(j);
                Line ln;
                const float determinant = fdet(Li.dx, Li.dy, Lj.dx, Lj.dy);
                if (fabsf(determinant) <= CAV_RVO_EPSILON) {
                    if (fdot(Li.dx, Li.dy, Lj.dx, Lj.dy) > 0.0f) continue;
                    ln.px = 0.5f * (Li.px + Lj.px);
                    ln.py = 0.5f * (Li.py + Lj.py);
                } else {
                    const float sc = fdet(Lj.dx, Lj.dy, Li.px - Lj.px, Li.py - Lj.py) / determinant;
                    ln.px = Li.px + sc * Li.dx;
                    ln.py = Li.py + sc * Li.dy;
                }
                const float ddx = Lj.dx - Li.dx, ddy = Lj.dy - Li.dy;
                const float inv = 1.0f / sqrtf(fdot(ddx, ddy, ddx, ddy));
                ln.dx = ddx * inv;
                ln.dy = ddy * inv;
                proj.set(np++, ln);
            }
            const float tx = rx, ty = ry;
            if (tpa_lp2(proj, np, radius, -Li.dy, Li.dx, true, rx, ry) < np) {
                rx = tx;
                ry = ty;
            }
            distance = fdet(Li.dx, Li.dy, Li.px - rx, Li.py - ry);
        }
    }
}

// ORCA half-plane for neighbour `ot` seen from `me` (RVO2 Agent.cpp:299-365); f4 = (pos.xy, vel.xy)
__device__ __forceinline__ Line orca_line(const float4 &me, float rad_me, const float4 &ot, float rad_ot,
                                          float invTimeHorizon, float timeStep, float collab) {
    Line ln;
    const float rpx = ot.x - me.x, rpy = ot.y - me.y;
    const float rvx = me.z - ot.z, rvy = me.w - ot.w;
    const float distSq = fdot(rpx, rpy, rpx, rpy);
    const float combinedRadius = rad_me + rad_ot;
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
            ln.dx = uwy;
            ln.dy = -uwx;
            const float sc = combinedRadius * invTimeHorizon - wLength;
            ux = sc * uwx;
            uy = sc * uwy;
        } else {
            const float leg = sqrtf(distSq - combinedRadiusSq);
            const float inv = 1.0f / distSq;
            if (fdet(rpx, rpy, wx, wy) > 0.0f) {
                ln.dx = (rpx * leg - rpy * combinedRadius) * inv;
                ln.dy = (rpx * combinedRadius + rpy * leg) * inv;
            } else {
                ln.dx = (-(rpx * leg + rpy * combinedRadius)) * inv;
                ln.dy = (-(-rpx * combinedRadius + rpy * leg)) * inv;
            }
            const float dotProduct2 = fdot(rvx, rvy, ln.dx, ln.dy);
            ux = dotProduct2 * ln.dx - rvx;
            uy = dotProduct2 * ln.dy - rvy;
        }
    } else {
        const float invTimeStep = 1.0f / timeStep;
        const float wx = rvx - invTimeStep * rpx, wy = rvy - invTimeStep * rpy;
        const float wLength = sqrtf(fdot(wx, wy, wx, wy));
        const float inv = 1.0f / wLength;
        const float uwx = wx * inv, uwy = wy * inv;
        ln.dx = uwy;
        ln.dy = -uwx;
        const float sc = combinedRadius * invTimeStep - wLength;
        ux = sc * uwx;
        uy = sc * uwy;
    }
    ln.px = me.z + collab * ux;
    ln.py = me.w + collab * uy;
    return ln;
}

// One ego solve by one thread.  s_f4 / s_frad: the env's agents (shared memory).  keys: this thread's sort-key
// slots ([slot * kstride]); lines: this thread's line slots.
__device__ __forceinline__ void orca_tpa(int n, int ego, const float4 *s_f4, const float *s_frad, float prefx,
                                         float prefy, float maxSpeed, float timeHorizon, float timeStep, float collab,
                                         float neighborDist, int maxNeighbors, unsigned long long *keys, int kstride,
                                         SmemLines lines, float &nvx, float &nvy) {
    const float4 me = s_f4[ego];
    const float rad_me = s_frad[ego];
    int nn = 0;
    if (maxNeighbors > 0 && collab > 0.0f) {
        const float rangeSq = neighborDist * neighborDist;
        // neighbour order: ascending distSq, ties keep agent-index order (insertAgentNeighbor, Agent.cpp:424-448)
#pragma unroll 1
        for (int o = 0; o < n; ++o) {
            if (o == ego) continue;
            const float4 ot = s_f4[o];
            const float dx = me.x - ot.x, dy = me.y - ot.y;
            const float distSq = fdot(dx, dy, dx, dy);
            if (!(distSq < rangeSq)) continue;
            const unsigned long long key = ((unsigned long long)__float_as_uint(distSq) << 32) | (unsigned)o;
            int i = nn;
#pragma unroll 1
            while (i > 0) {
                const unsigned long long prev = keys[(i - 1) * kstride];
                if (!(key < prev)) break;
                keys[i * kstride] = prev;
                --i;
            }
            keys[i * kstride] = key;
            ++nn;
        }
        if (nn > maxNeighbors) nn = maxNeighbors;
    }
    const float invTimeHorizon = 1.0f / timeHorizon;
#pragma unroll 1
    for (int k = 0; k < nn; ++k) {
        const int o = (int)(keys[k * kstride] & 0xffffffffull);
        lines.set(k, orca_line(me, rad_me, s_f4[o], s_frad[o], invTimeHorizon, timeStep, collab));
    }
    float rx, ry;
    const int fail = tpa_lp2(lines, nn, maxSpeed, prefx, prefy, false, rx, ry);
    if (fail < nn) tpa_lp3(lines, nn, fail, maxSpeed, rx, ry);
    nvx = rx;
    nvy = ry;
}

}  // namespace cav
