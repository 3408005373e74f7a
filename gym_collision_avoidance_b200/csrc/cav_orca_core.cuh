// cav_orca_core.cuh -- the float32 ORCA / RVO2 velocity solve of the thread-per-agent kernels, as host/device source.
//
// One thread owns one ego agent (RVO2 Agent.cpp:295-366 half-planes, :424-448 neighbour order, :477-630
// linearProgram1/2/3).  This header is what k_policy_move_tpa (generic, run-time N, lines in shared memory laid out
// [line][thread]) and k_policy_move_tpa_n (compile-time N, lines in registers) compile -- and what g++ compiles into
// tests/native/orca_host.cpp, where tests/test_orca_native_cpu.py checks it BIT-EXACTLY against the reference's own
// RVO2 library on the CPU.  Every float expression keeps the reference's operand order; nvcc builds it with
// -fmad=false (g++ with -ffp-contract=off), division and sqrt are IEEE.
#pragma once
#include "cav_hd.cuh"
#include "cav_sortnet.cuh"

namespace cav {

struct Line {
    float px, py, dx, dy;
};

#define CAV_RVO_EPSILON 0.00001f  // RVO2 Definitions.h:52

CAV_HDI float fdet(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
CAV_HDI float fdot(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }

struct SmemLines {  // line i of this thread at base[i * stride]
    float4 *base;
    int stride;
    CAV_HDI Line get(int i) const {
        const float4 v = base[i * stride];
        return Line{v.x, v.y, v.z, v.w};
    }
    CAV_HDI void set(int i, const Line &l) { base[i * stride] = make_float4(l.px, l.py, l.dx, l.dy); }
};
struct LocalLines {
    Line *p;
    CAV_HDI Line get(int i) const { return p[i]; }
    CAV_HDI void set(int i, const Line &l) { p[i] = l; }
};

// RVO2 Agent.cpp:477-548
template <typename Lines>
CAV_HDI bool tpa_lp1(const Lines &lines, int lineNo, float radius, float ox, float oy,
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
CAV_HDI int tpa_lp2(const Lines &lines, int n, float radius, float ox, float oy, bool directionOpt,
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
CAV_HD_NOINLINE void tpa_lp3(const Lines &lines, int n, int beginLine, float radius, float &rx, float &ry) {
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
CAV_HDI Line orca_line(const float4 &me, float rad_me, const float4 &ot, float rad_ot,
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
CAV_HDI void orca_tpa(int n, int ego, const float4 *s_f4, const float *s_frad, float prefx,
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
            const unsigned long long key = ((unsigned long long)float_bits(distSq) << 32) | (unsigned)o;
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

// ORCA with everything unrolled: the N-1 neighbour keys are ordered by an odd-even transposition network in
// registers, the N-1 half-planes (each with its own sqrt / divide chains) are built as independent instruction
// streams into registers, and linearProgram2 / linearProgram1 (RVO2 Agent.cpp:477-581) run as straight-line
// predicated code over those registers.  Only the rare infeasible case (linearProgram3, Agent.cpp:583-630) spills the
// lines to shared memory and takes the generic routine.
template <int NO>
CAV_HDI void orca_tpa_n(int n, int ego, const float4 *s_f4, const float *s_frad, float prefx,
                                           float prefy, float maxSpeed, float timeHorizon, float timeStep, float collab,
                                           float neighborDist, int maxNeighbors, SmemLines slines, float &nvx, float &nvy) {
    const float4 me = s_f4[ego];
    const float rad_me = s_frad[ego];
    unsigned long long key[NO > 0 ? NO : 1];
    int nn = 0;
    const bool any = maxNeighbors > 0 && collab > 0.0f;
    const float rangeSq = neighborDist * neighborDist;
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        const int o = j + (j >= ego ? 1 : 0);
        key[j] = ~0ull;
        if (any && o < n) {
            const float4 ot = s_f4[o];
            const float dx = me.x - ot.x, dy = me.y - ot.y;  // absSq(position_ - agent->position_), Agent.cpp:427
            const float distSq = fdot(dx, dy, dx, dy);
            if (distSq < rangeSq) {
                key[j] = ((unsigned long long)float_bits(distSq) << 32) | (unsigned)o;
                ++nn;
            }
        }
    }
    sort_network<NO>([&](int i, int j) {
        const unsigned long long ka = key[i], kb = key[j];
        const bool sw = kb < ka;
        key[i] = sw ? kb : ka;
        key[j] = sw ? ka : kb;
    });
    if (nn > maxNeighbors) nn = maxNeighbors;
    const float invTimeHorizon = 1.0f / timeHorizon;
    Line ln[NO > 0 ? NO : 1];
#pragma unroll
    for (int k = 0; k < NO; ++k) {
        ln[k] = Line{0.f, 0.f, 1.f, 0.f};
        if (k < nn) {
            const int o = (int)(key[k] & 0xffffffffull);
            ln[k] = orca_line(me, rad_me, s_f4[o], s_frad[o], invTimeHorizon, timeStep, collab);
        }
    }
    // linearProgram2 (Agent.cpp:550-581), closest-point mode
    float rx, ry;
    if (fdot(prefx, prefy, prefx, prefy) > maxSpeed * maxSpeed) {
        const float inv = 1.0f / sqrtf(fdot(prefx, prefy, prefx, prefy));
        rx = (prefx * inv) * maxSpeed;
        ry = (prefy * inv) * maxSpeed;
    } else {
        rx = prefx;
        ry = prefy;
    }
    int fail = -1;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
        if (fail < 0 && i < nn && fdet(ln[i].dx, ln[i].dy, ln[i].px - rx, ln[i].py - ry) > 0.0f) {
            // linearProgram1 on line i (Agent.cpp:477-548)
            const Line L = ln[i];
            const float dotProduct = fdot(L.px, L.py, L.dx, L.dy);
            const float discriminant = dotProduct * dotProduct + maxSpeed * maxSpeed - fdot(L.px, L.py, L.px, L.py);
            bool ok = !(discriminant < 0.0f);
            const float sqrtDiscriminant = sqrtf(discriminant);
            float tLeft = -dotProduct - sqrtDiscriminant;
            float tRight = -dotProduct + sqrtDiscriminant;
#pragma unroll
            for (int j = 0; j < i; ++j) {
                const float denominator = fdet(L.dx, L.dy, ln[j].dx, ln[j].dy);
                const float numerator = fdet(ln[j].dx, ln[j].dy, L.px - ln[j].px, L.py - ln[j].py);
                if (fabsf(denominator) <= CAV_RVO_EPSILON) {
                    if (numerator < 0.0f) ok = false;
                } else {
                    const float t = numerator / denominator;
                    if (denominator >= 0.0f) tRight = min_keep(tRight, t);
                    else tLeft = max_keep(tLeft, t);
                }
            }
            // the reference returns false as soon as tLeft > tRight inside the loop; the bounds only tighten, so the
            // final comparison is equivalent
            if (ok && !(tLeft > tRight)) {
                const float t = fdot(L.dx, L.dy, prefx - L.px, prefy - L.py);
                const float tc = (t < tLeft) ? tLeft : ((t > tRight) ? tRight : t);
                rx = L.px + tc * L.dx;
                ry = L.py + tc * L.dy;
            } else {
                fail = i;
            }
        }
    }
    if (fail >= 0) {  // rare: infeasible -> linearProgram3 through shared memory (generic routine)
#pragma unroll
        for (int k = 0; k < NO; ++k)
            if (k < nn) slines.set(k, ln[k]);
        tpa_lp3(slines, nn, fail, maxSpeed, rx, ry);
    }
    nvx = rx;
    nvy = ry;
}

}  // namespace cav
