/*
 * cav_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's per-step hot path (mit-acl/gym-collision-avoidance,
 * CollisionAvoidanceEnv.step / reset), used ONLY by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the reported CPU baseline.  The product path
 * (gym_collision_avoidance_b200 + libcavoid_sm100a.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the reference's own tests hold no golden vectors for this path, so this
 * restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
 *   - tests/golden/*.npz, minted by tools/make_golden.py from the unmodified Python reference
 *     (tests/test_oracle_golden.py), and
 *   - oracle/_ref/librvo2_ref.so, the unmodified reference RVO2 C++ (tests/test_oracle_rvo2.py, bit-exact).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/gym_collision_avoidance/envs/; R/ = policies/Python-RVO2/src/).
 *
 * Compiled with -ffp-contract=off: the reference is NumPy float64 (no FMA) and x86-64 g++ without -mfma.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cavoid.h"

#define CAVO_MAX_AGENTS 64

static int g_threads = 0; /* 0 = OpenMP default */
int cavo_get_max_threads(void);
static int g_faithful_dostep = 0;

/* ------------------------------------------------------------------------------------------------------------
 * ORCA / RVO2 in float32 (R/Agent.cpp:287-366 lines, :417-421 solve, :424-448 neighbours, :470-475 update,
 * :477-630 LP1/LP2/LP3; R/Vector2.h:95-126,285-328 arithmetic; R/Definitions.h:52 RVO_EPSILON).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    float px, py, dx, dy;
} orca_line_t;
#define RVO_EPS 0.00001f

static inline float o_det(float ax, float ay, float bx, float by) { return ax * by - ay * bx; } /* Vector2.h:313-316 */
static inline float o_dot(float ax, float ay, float bx, float by) { return ax * bx + ay * by; } /* Vector2.h:95-98 */

/* R/Agent.cpp:477-548 */
static int orca_lp1(const orca_line_t *lines, int lineNo, float radius, float ox, float oy, int directionOpt,
                    float *rx, float *ry) {
    const orca_line_t *L = &lines[lineNo];
    const float dotProduct = o_dot(L->px, L->py, L->dx, L->dy);
    const float discriminant = dotProduct * dotProduct + radius * radius - o_dot(L->px, L->py, L->px, L->py);
    if (discriminant < 0.0f) return 0;
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
    for (int i = 0; i < lineNo; ++i) {
        const float denominator = o_det(L->dx, L->dy, lines[i].dx, lines[i].dy);
        const float numerator = o_det(lines[i].dx, lines[i].dy, L->px - lines[i].px, L->py - lines[i].py);
        if (fabsf(denominator) <= RVO_EPS) {
            if (numerator < 0.0f) return 0;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) {
            tRight = (t < tRight) ? t : tRight; /* std::min(tRight, t) */
        } else {
            tLeft = (tLeft < t) ? t : tLeft; /* std::max(tLeft, t) */
        }
        if (tLeft > tRight) return 0;
    }
    if (directionOpt) {
        if (o_dot(ox, oy, L->dx, L->dy) > 0.0f) {
            *rx = L->px + tRight * L->dx;
            *ry = L->py + tRight * L->dy;
        } else {
            *rx = L->px + tLeft * L->dx;
            *ry = L->py + tLeft * L->dy;
        }
    } else {
        const float t = o_dot(L->dx, L->dy, ox - L->px, oy - L->py);
        if (t < tLeft) {
            *rx = L->px + tLeft * L->dx;
            *ry = L->py + tLeft * L->dy;
        } else if (t > tRight) {
            *rx = L->px + tRight * L->dx;
            *ry = L->py + tRight * L->dy;
        } else {
            *rx = L->px + t * L->dx;
            *ry = L->py + t * L->dy;
        }
    }
    return 1;
}

/* R/Agent.cpp:550-581 */
static int orca_lp2(const orca_line_t *lines, int n, float radius, float ox, float oy, int directionOpt, float *rx,
                    float *ry) {
    if (directionOpt) {
        *rx = ox * radius;
        *ry = oy * radius;
    } else if (o_dot(ox, oy, ox, oy) > radius * radius) {
        const float inv = 1.0f / sqrtf(o_dot(ox, oy, ox, oy)); /* normalize: v / abs(v) = v * (1/abs) */
        *rx = (ox * inv) * radius;
        *ry = (oy * inv) * radius;
    } else {
        *rx = ox;
        *ry = oy;
    }
    for (int i = 0; i < n; ++i) {
        if (o_det(lines[i].dx, lines[i].dy, lines[i].px - *rx, lines[i].py - *ry) > 0.0f) {
            const float tx = *rx, ty = *ry;
            if (!orca_lp1(lines, i, radius, ox, oy, directionOpt, rx, ry)) {
                *rx = tx;
                *ry = ty;
                return i;
            }
        }
    }
    return n;
}

/* R/Agent.cpp:583-630 (numObstLines == 0 on this path: RVOPolicy never adds obstacles) */
static void orca_lp3(const orca_line_t *lines, int n, int beginLine, float radius, float *rx, float *ry) {
    float distance = 0.0f;
    orca_line_t proj[CAVO_MAX_AGENTS];
    for (int i = beginLine; i < n; ++i) {
        if (o_det(lines[i].dx, lines[i].dy, lines[i].px - *rx, lines[i].py - *ry) > distance) {
            int np = 0;
            for (int j = 0; j < i; ++j) {
                orca_line_t ln;
                const float determinant = o_det(lines[i].dx, lines[i].dy, lines[j].dx, lines[j].dy);
                if (fabsf(determinant) <= RVO_EPS) {
                    if (o_dot(lines[i].dx, lines[i].dy, lines[j].dx, lines[j].dy) > 0.0f) continue;
                    ln.px = 0.5f * (lines[i].px + lines[j].px);
                    ln.py = 0.5f * (lines[i].py + lines[j].py);
                } else {
                    const float s = o_det(lines[j].dx, lines[j].dy, lines[i].px - lines[j].px, lines[i].py - lines[j].py) /
                                    determinant;
                    ln.px = lines[i].px + s * lines[i].dx;
                    ln.py = lines[i].py + s * lines[i].dy;
                }
                const float ddx = lines[j].dx - lines[i].dx, ddy = lines[j].dy - lines[i].dy;
                const float inv = 1.0f / sqrtf(o_dot(ddx, ddy, ddx, ddy));
                ln.dx = ddx * inv;
                ln.dy = ddy * inv;
                proj[np++] = ln;
            }
            const float tx = *rx, ty = *ry;
            if (orca_lp2(proj, np, radius, -lines[i].dy, lines[i].dx, 1, rx, ry) < np) {
                *rx = tx;
                *ry = ty;
            }
            distance = o_det(lines[i].dx, lines[i].dy, lines[i].px - *rx, lines[i].py - *ry);
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Agent kd-tree (R/KdTree.cpp:47-111 build, :266-298 query; MAX_LEAF_SIZE R/KdTree.h:48 = 10).
 * `perm` is the simulator's persistent KdTree::agents_ permutation: buildAgentTree re-partitions it IN PLACE every
 * doStep (R/RVOSimulator.cpp:163-168), it is never reset, so the leaf order -- and with it the order in which
 * insertAgentNeighbor sees agents at EXACTLY equal float distances -- depends on the whole position history since
 * the simulator was created (RVOPolicy.__init__, i.e. once per Agent / episode).  Every RVO agent of an environment
 * owns its own simulator, but all of them are fed the same positions on the same steps from identity, so one
 * permutation per environment reproduces them all.  For n <= 10 the root is a leaf: no swap ever happens.
 * ---------------------------------------------------------------------------------------------------------- */
#define CAVO_KD_MAX_LEAF 10
typedef struct {
    int begin, end, left, right;
    float minX, maxX, minY, maxY;
} kd_node_t;
typedef struct {
    kd_node_t node[2 * CAVO_MAX_AGENTS];
    const uint8_t *perm;
    const float *pos;
} kd_tree_t;

static void kd_build_rec(kd_tree_t *t, uint8_t *perm, int begin, int end, int node) { /* KdTree.cpp:62-111 */
    const float *pos = t->pos;
    kd_node_t *nd = &t->node[node];
    nd->begin = begin;
    nd->end = end;
    nd->minX = nd->maxX = pos[2 * perm[begin]];
    nd->minY = nd->maxY = pos[2 * perm[begin] + 1];
    for (int i = begin + 1; i < end; ++i) {
        const float x = pos[2 * perm[i]], y = pos[2 * perm[i] + 1];
        nd->maxX = (nd->maxX < x) ? x : nd->maxX; /* std::max(a, b) = (a < b) ? b : a */
        nd->minX = (x < nd->minX) ? x : nd->minX; /* std::min(a, b) = (b < a) ? b : a */
        nd->maxY = (nd->maxY < y) ? y : nd->maxY;
        nd->minY = (y < nd->minY) ? y : nd->minY;
    }
    if (end - begin > CAVO_KD_MAX_LEAF) {
        const int isVertical = (nd->maxX - nd->minX > nd->maxY - nd->minY);
        const float splitValue = isVertical ? 0.5f * (nd->maxX + nd->minX) : 0.5f * (nd->maxY + nd->minY);
        int left = begin, right = end;
        while (left < right) {
            while (left < right && pos[2 * perm[left] + (isVertical ? 0 : 1)] < splitValue) ++left;
            while (right > left && pos[2 * perm[right - 1] + (isVertical ? 0 : 1)] >= splitValue) --right;
            if (left < right) {
                const uint8_t tmp = perm[left];
                perm[left] = perm[right - 1];
                perm[right - 1] = tmp;
                ++left;
                --right;
            }
        }
        if (left == begin) {
            ++left;
            ++right;
        }
        nd->left = node + 1;
        nd->right = node + 2 * (left - begin);
        kd_build_rec(t, perm, begin, left, nd->left);
        kd_build_rec(t, perm, left, end, nd->right);
    }
}
/* KdTree::buildAgentTree (KdTree.cpp:47-60): pos [n][2] float32, perm [n] updated in place */
void cavo_kd_build(kd_tree_t *t, const float *pos, uint8_t *perm, int n) {
    t->pos = pos;
    t->perm = perm;
    if (n > 0) kd_build_rec(t, perm, 0, n, 0);
}
/* Replays the persistent permutation over a sequence of position sets (test helper: the teacher-forced golden
 * tests need the permutation the reference's simulators had reached at step t). pos_seq [steps][n][2] float32. */
void cavo_kd_replay(const float *pos_seq, int steps, int n, uint8_t *perm) {
    kd_tree_t t;
    for (int i = 0; i < n; ++i) perm[i] = (uint8_t)i;
    for (int s = 0; s < steps; ++s) cavo_kd_build(&t, pos_seq + (size_t)s * n * 2, perm, n);
}

typedef struct {
    int ego, nn, maxNeighbors;
    const float *pos;
    int *nb_idx;
    float *nb_d;
} kd_query_t;
/* Agent::insertAgentNeighbor, R/Agent.cpp:424-448 */
static void orca_insert_neighbor(kd_query_t *q, int j, float *rangeSq) {
    if (j == q->ego) return;
    const float dx = q->pos[2 * q->ego] - q->pos[2 * j], dy = q->pos[2 * q->ego + 1] - q->pos[2 * j + 1];
    const float distSq = o_dot(dx, dy, dx, dy);
    if (distSq < *rangeSq) {
        if (q->nn < q->maxNeighbors) ++q->nn;
        int i = q->nn - 1;
        while (i != 0 && distSq < q->nb_d[i - 1]) {
            q->nb_d[i] = q->nb_d[i - 1];
            q->nb_idx[i] = q->nb_idx[i - 1];
            --i;
        }
        q->nb_d[i] = distSq;
        q->nb_idx[i] = j;
        if (q->nn == q->maxNeighbors) *rangeSq = q->nb_d[q->nn - 1];
    }
}
static inline float o_sqr(float a) { return a * a; }
static inline float o_max0(float a) { return (0.0f < a) ? a : 0.0f; } /* std::max(0.0f, a) */
/* KdTree::queryAgentTreeRecursive, KdTree.cpp:266-298 */
static void kd_query_rec(const kd_tree_t *t, kd_query_t *q, float *rangeSq, int node) {
    const kd_node_t *nd = &t->node[node];
    if (nd->end - nd->begin <= CAVO_KD_MAX_LEAF) {
        for (int i = nd->begin; i < nd->end; ++i) orca_insert_neighbor(q, t->perm[i], rangeSq);
        return;
    }
    const float px = q->pos[2 * q->ego], py = q->pos[2 * q->ego + 1];
    const kd_node_t *L = &t->node[nd->left], *R = &t->node[nd->right];
    const float distSqLeft = o_sqr(o_max0(L->minX - px)) + o_sqr(o_max0(px - L->maxX)) + o_sqr(o_max0(L->minY - py)) +
                             o_sqr(o_max0(py - L->maxY));
    const float distSqRight = o_sqr(o_max0(R->minX - px)) + o_sqr(o_max0(px - R->maxX)) + o_sqr(o_max0(R->minY - py)) +
                              o_sqr(o_max0(py - R->maxY));
    if (distSqLeft < distSqRight) {
        if (distSqLeft < *rangeSq) {
            kd_query_rec(t, q, rangeSq, nd->left);
            if (distSqRight < *rangeSq) kd_query_rec(t, q, rangeSq, nd->right);
        }
    } else {
        if (distSqRight < *rangeSq) {
            kd_query_rec(t, q, rangeSq, nd->right);
            if (distSqLeft < *rangeSq) kd_query_rec(t, q, rangeSq, nd->left);
        }
    }
}

/* One ego solve: neighbours (R/Agent.cpp:44-56 computeNeighbors -> kd-tree query -> insertAgentNeighbor :424-448; with
 * tree == NULL agents are inserted in index order, which is what the single-leaf tree of n <= 10 agents does), ORCA
 * lines (R/Agent.cpp:295-366), LP2 then LP3 (:417-421).
 * pos/vel [n][2] float32, rad [n] (already scaled), returns newVelocity. */
void cavo_orca_solve_kd(int n, int ego, const float *pos, const float *vel, const float *rad, float maxSpeed,
                        float prefx, float prefy, float timeHorizon, float timeStep, float collab,
                        float neighborDist, int maxNeighbors, const kd_tree_t *tree, float *out_vx, float *out_vy,
                        orca_line_t *out_lines, int *out_nlines) {
    int nb_idx[CAVO_MAX_AGENTS];
    float nb_d[CAVO_MAX_AGENTS];
    int nn = 0;
    if (maxNeighbors > 0) {
        float rangeSq = neighborDist * neighborDist;
        kd_query_t q = {ego, 0, maxNeighbors, pos, nb_idx, nb_d};
        if (tree) {
            kd_query_rec(tree, &q, &rangeSq, 0);
        } else {
            for (int j = 0; j < n; ++j) orca_insert_neighbor(&q, j, &rangeSq);
        }
        nn = q.nn;
    }
    orca_line_t lines[CAVO_MAX_AGENTS];
    int nl = 0;
    const float invTimeHorizon = 1.0f / timeHorizon;
    const float vx = vel[2 * ego], vy = vel[2 * ego + 1];
    if (collab > 0.0f) {
        for (int k = 0; k < nn; ++k) {
            const int o = nb_idx[k];
            const float rpx = pos[2 * o] - pos[2 * ego], rpy = pos[2 * o + 1] - pos[2 * ego + 1];
            const float rvx = vx - vel[2 * o], rvy = vy - vel[2 * o + 1];
            const float distSq = o_dot(rpx, rpy, rpx, rpy);
            const float combinedRadius = rad[ego] + rad[o];
            const float combinedRadiusSq = combinedRadius * combinedRadius;
            orca_line_t ln;
            float ux, uy;
            if (distSq > combinedRadiusSq) {
                const float wx = rvx - invTimeHorizon * rpx, wy = rvy - invTimeHorizon * rpy;
                const float wLengthSq = o_dot(wx, wy, wx, wy);
                const float dotProduct1 = o_dot(wx, wy, rpx, rpy);
                if (dotProduct1 < 0.0f && dotProduct1 * dotProduct1 > combinedRadiusSq * wLengthSq) {
                    const float wLength = sqrtf(wLengthSq);
                    const float inv = 1.0f / wLength;
                    const float uwx = wx * inv, uwy = wy * inv;
                    ln.dx = uwy;
                    ln.dy = -uwx;
                    const float s = combinedRadius * invTimeHorizon - wLength;
                    ux = s * uwx;
                    uy = s * uwy;
                } else {
                    const float leg = sqrtf(distSq - combinedRadiusSq);
                    const float inv = 1.0f / distSq;
                    if (o_det(rpx, rpy, wx, wy) > 0.0f) {
                        ln.dx = (rpx * leg - rpy * combinedRadius) * inv;
                        ln.dy = (rpx * combinedRadius + rpy * leg) * inv;
                    } else {
                        ln.dx = (-(rpx * leg + rpy * combinedRadius)) * inv;
                        ln.dy = (-(-rpx * combinedRadius + rpy * leg)) * inv;
                    }
                    const float dotProduct2 = o_dot(rvx, rvy, ln.dx, ln.dy);
                    ux = dotProduct2 * ln.dx - rvx;
                    uy = dotProduct2 * ln.dy - rvy;
                }
            } else {
                const float invTimeStep = 1.0f / timeStep;
                const float wx = rvx - invTimeStep * rpx, wy = rvy - invTimeStep * rpy;
                const float wLength = sqrtf(o_dot(wx, wy, wx, wy));
                const float inv = 1.0f / wLength;
                const float uwx = wx * inv, uwy = wy * inv;
                ln.dx = uwy;
                ln.dy = -uwx;
                const float s = combinedRadius * invTimeStep - wLength;
                ux = s * uwx;
                uy = s * uwy;
            }
            ln.px = vx + collab * ux;
            ln.py = vy + collab * uy;
            lines[nl++] = ln;
        }
    } else if (collab < 0.0f && nn > 0) {
        /* R/Agent.cpp:368-415: negative collaboration -- steer INTO the closest neighbour.  The closest one is the first
         * of the list (ascending distSq; the strict `<` of the reference's scan keeps the first of equals).  Two
         * half-planes: the legs of its velocity obstacle facing INWARD, both through v_other + relativePosition; the
         * preferred velocity becomes v_other + (-collab) * normalize(relativePosition). */
        const int o = nb_idx[0];
        const float rpx = pos[2 * o] - pos[2 * ego], rpy = pos[2 * o + 1] - pos[2 * ego + 1];
        const float distSq = o_dot(rpx, rpy, rpx, rpy);
        const float combinedRadius = rad[ego] + rad[o];
        const float combinedRadiusSq = combinedRadius * combinedRadius;
        const float leg = sqrtf(distSq - combinedRadiusSq);
        const float inv = 1.0f / distSq; /* Vector2 / float = multiply by the reciprocal, R/Vector2.h:115-120 */
        orca_line_t ln;
        ln.px = vel[2 * o] + rpx;
        ln.py = vel[2 * o + 1] + rpy;
        ln.dx = (-(rpx * leg - rpy * combinedRadius)) * inv; /* left leg */
        ln.dy = (-(rpx * combinedRadius + rpy * leg)) * inv;
        lines[nl++] = ln;
        ln.dx = (rpx * leg + rpy * combinedRadius) * inv; /* right leg */
        ln.dy = (-rpx * combinedRadius + rpy * leg) * inv;
        lines[nl++] = ln;
        const float invn = 1.0f / sqrtf(distSq); /* normalize(v) = v / abs(v) */
        const float s_ = -collab;
        prefx = vel[2 * o] + s_ * (rpx * invn);
        prefy = vel[2 * o + 1] + s_ * (rpy * invn);
    }
    float rx, ry;
    const int fail = orca_lp2(lines, nl, maxSpeed, prefx, prefy, 0, &rx, &ry);
    if (fail < nl) orca_lp3(lines, nl, fail, maxSpeed, &rx, &ry);
    *out_vx = rx;
    *out_vy = ry;
    if (out_lines) memcpy(out_lines, lines, sizeof(orca_line_t) * (size_t)nl);
    if (out_nlines) *out_nlines = nl;
}

void cavo_orca_solve(int n, int ego, const float *pos, const float *vel, const float *rad, float maxSpeed,
                     float prefx, float prefy, float timeHorizon, float timeStep, float collab,
                     float neighborDist, int maxNeighbors, float *out_vx, float *out_vy, orca_line_t *out_lines,
                     int *out_nlines) {
    cavo_orca_solve_kd(n, ego, pos, vel, rad, maxSpeed, prefx, prefy, timeHorizon, timeStep, collab, neighborDist,
                       maxNeighbors, NULL, out_vx, out_vy, out_lines, out_nlines);
}
/* One solve on a caller-held permutation (tests: bit-exact pin against the reference library over several doSteps).
 * perm [n] is updated in place exactly like one buildAgentTree. */
void cavo_orca_solve_perm(int n, int ego, const float *pos, const float *vel, const float *rad, float maxSpeed,
                          float prefx, float prefy, float timeHorizon, float timeStep, float collab,
                          float neighborDist, int maxNeighbors, uint8_t *perm, int rebuild, float *out_vx,
                          float *out_vy) {
    static __thread kd_tree_t t;
    if (rebuild || t.pos != pos || t.perm != perm) cavo_kd_build(&t, pos, perm, n);
    cavo_orca_solve_kd(n, ego, pos, vel, rad, maxSpeed, prefx, prefy, timeHorizon, timeStep, collab, neighborDist,
                       maxNeighbors, &t, out_vx, out_vy, NULL, NULL);
}

/* Philox4x32-10 (Salmon et al., SC'11; Random123 constants), first output word / 2^32: the anti-collaborative RVO agents'
 * policy re-draw.  The reference draws np.random.choice from NumPy's global stream (RVOPolicy.py:84), which a batched
 * environment cannot share; the product (csrc/cav_hd.cuh philox_u01) and this restatement use the same counter-based
 * draw keyed by (seed; env, agent, episode, step_num), so they agree draw for draw, and the distribution is the reference's. */
static double cavo_philox_u01(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (double)c0 * (1.0 / 4294967296.0);
}
double cavo_philox_u01_export(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return cavo_philox_u01(seed, c0, c1, c2, c3);
}

/* ------------------------------------------------------------------------------------------------------------
 * Map + laserscan, float64 geometry in both precisions (Map.py:6-64, LaserScanSensor.py:24-101).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int rows, cols;
    double origin0, origin1, cell;
} cavo_map_t;

static cavo_map_t cavo_map_dims(const cav_config_t *c) {
    cavo_map_t m;
    m.cell = c->map_grid_cell_size;
    m.rows = (int)(c->map_x_width / m.cell); /* Map.py:13 dims */
    m.cols = (int)(c->map_y_width / m.cell);
    m.origin0 = (c->map_x_width / 2.) / m.cell; /* Map.py:23 */
    m.origin1 = (c->map_y_width / 2.) / m.cell;
    return m;
}
/* Map.py:26-32 */
static int cavo_world_to_cell(const cavo_map_t *m, double px, double py, int *gx, int *gy) {
    const double fx = floor(m->origin0 - py / m->cell);
    const double fy = floor(m->origin1 + px / m->cell);
    if (!(fx >= 0 && fy >= 0 && fx < m->rows && fy < m->cols)) {
        *gx = -1;
        *gy = -1;
        return 0;
    }
    *gx = (int)fx;
    *gy = (int)fy;
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * GA3C-CADRL network forward pass (policies/GA3CCADRLPolicy.py:49-84 -> policies/GA3C_CADRL/network.py:37-41
 * sess.run('Softmax:0')).  TensorFlow is not in this image, so the restatement follows the reference's own GRAPH
 * (network_01900000.meta, node for node: sub / div / strided_slice* / Reshape / transpose / dynamic_rnn LSTMCell while
 * loop with sequence_length Select / layer1_input concat / layer1 / layer2 / fullyconnected1 / logits_p / Softmax) and
 * is pinned against oracle/tf1_graph.py, a numpy interpreter that EXECUTES that graph's nodes with the checkpoint's
 * variables (tests/test_ga3c_cpu.py, tests/golden/ga3c/ga3c_iros18.npz).  float32 throughout, sums in ascending k.
 * Weight blob layout: include/cavoid.h (CAV_GA3C_*).
 * ---------------------------------------------------------------------------------------------------------- */
static const float *g_ga3c_blob = NULL;
void cavo_set_ga3c(const float *blob) { g_ga3c_blob = blob; }

static float ga3c_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

void cavo_ga3c_forward(const float *blob, const float *X, int rows, float *probs, int32_t *action_index) {
    enum { XL = CAV_GA3C_X_LEN, T = CAV_GA3C_MAX_OTHERS, H = CAV_GA3C_HIDDEN, F = CAV_GA3C_FC, A = CAV_GA3C_ACTIONS };
    const float *avg = blob, *std_ = avg + XL, *wl = std_ + XL, *bl = wl + (7 + H) * H * 4, *w1 = bl + H * 4,
                *b1 = w1 + (4 + H) * F, *w2 = b1 + F, *b2 = w2 + F * F, *w3 = b2 + F, *b3 = w3 + F * F, *wo = b3 + F,
                *bo = wo + F * A;
    for (int r = 0; r < rows; ++r) {
        const float *x = X + (size_t)r * XL;
        float xn[XL];
        for (int k = 0; k < XL; ++k) xn[k] = (x[k] - avg[k]) / std_[k]; /* sub, div */
        const int len = (int)x[0];                                       /* ToInt32(X[:, 0]) */
        float c[H], h[H];
        memset(c, 0, sizeof(c));
        memset(h, 0, sizeof(h));
        for (int t = 0; t < T; ++t) {       /* rnn/while: time < 19 */
            if (t >= len) continue;         /* Select(time >= sequence_length, old state, new state) */
            const float *xt = xn + 5 + 7 * t;
            float cn[H], hn[H];
            for (int u = 0; u < H; ++u) {
                float g[4];
                for (int q = 0; q < 4; ++q) {
                    float acc = 0.0f;
                    for (int k = 0; k < 7; ++k) acc += xt[k] * wl[(k * H + u) * 4 + q];
                    for (int k = 0; k < H; ++k) acc += h[k] * wl[((7 + k) * H + u) * 4 + q];
                    g[q] = acc + bl[u * 4 + q]; /* BiasAdd */
                }
                cn[u] = ga3c_sigmoid(g[2] + 1.0f) * c[u] + ga3c_sigmoid(g[0]) * tanhf(g[1]);
                hn[u] = ga3c_sigmoid(g[3]) * tanhf(cn[u]);
            }
            memcpy(c, cn, sizeof(c));
            memcpy(h, hn, sizeof(h));
        }
        float in1[4 + H], a1[F], a2[F], a3[F], logit[A];
        for (int k = 0; k < 4; ++k) in1[k] = xn[1 + k]; /* layer1_input = concat(Xn[:, 1:5], h) */
        for (int k = 0; k < H; ++k) in1[4 + k] = h[k];
        for (int j = 0; j < F; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 4 + H; ++k) acc += in1[k] * w1[k * F + j];
            acc += b1[j];
            a1[j] = acc > 0.0f ? acc : 0.0f;
        }
        for (int j = 0; j < F; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < F; ++k) acc += a1[k] * w2[k * F + j];
            acc += b2[j];
            a2[j] = acc > 0.0f ? acc : 0.0f;
        }
        for (int j = 0; j < F; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < F; ++k) acc += a2[k] * w3[k * F + j];
            acc += b3[j];
            a3[j] = acc > 0.0f ? acc : 0.0f;
        }
        float mx = -INFINITY;
        int best = 0;
        for (int j = 0; j < A; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < F; ++k) acc += a3[k] * wo[k * A + j];
            logit[j] = acc + bo[j];
            if (logit[j] > mx) {
                mx = logit[j];
                best = j; /* np.argmax: first maximum (softmax is monotone) */
            }
        }
        if (probs) {
            float e[A], sum = 0.0f;
            for (int j = 0; j < A; ++j) {
                e[j] = expf(logit[j] - mx);
                sum += e[j];
            }
            for (int j = 0; j < A; ++j) probs[(size_t)r * A + j] = e[j] / sum;
        }
        if (action_index) action_index[r] = best;
    }
}

/* the flattened observation the policy hands the network (GA3CCADRLPolicy.py:68-75): Config.STATES_IN_OBS in order
 * without is_learning, cut / zero-padded to 138 (network.py:24-35).  src[k] = float index in the packed row, -1 = 0. */
void cavo_ga3c_gather(const cav_config_t *cfg, int16_t *src) {
    int order[CAV_OBS_NUM_STATES], n = 0;
    for (int s = 0; s < CAV_OBS_NUM_STATES; ++s)
        if (cfg->obs_offset[s] >= 0) order[n++] = s;
    for (int i = 1; i < n; ++i) /* by offset = STATES_IN_OBS order */
        for (int j = i; j > 0 && cfg->obs_offset[order[j]] < cfg->obs_offset[order[j - 1]]; --j) {
            const int t = order[j];
            order[j] = order[j - 1];
            order[j - 1] = t;
        }
    int k = 0;
    for (int i = 0; i < n && k < CAV_GA3C_X_LEN; ++i) {
        const int s = order[i];
        const int beg = cfg->obs_offset[s], end = (i + 1 < n) ? cfg->obs_offset[order[i + 1]] : cfg->obs_len;
        if (s == CAV_OBS_IS_LEARNING) continue; /* Config.STATES_NOT_USED_IN_POLICY */
        for (int o = beg; o < end && k < CAV_GA3C_X_LEN; ++o) src[k++] = (int16_t)o;
    }
    for (; k < CAV_GA3C_X_LEN; ++k) src[k] = -1;
}

/* speed fraction / heading change of action j (network.py:7-16) */
static void ga3c_action(int j, double *speed, double *dheading) {
    const double pi6 = 3.14159265358979323846 / 6, pi12 = 3.14159265358979323846 / 12;
    if (j < 5) {
        *speed = 1.0;
        *dheading = -pi6 + j * pi12;
    } else if (j < 8) {
        *speed = 0.5;
        *dheading = -pi6 + (j - 5) * pi6;
    } else {
        *speed = 0.0;
        *dheading = -pi6 + (j - 8) * pi6;
    }
}

#define REAL double
#define SFX(n) n##_f64
#define R_SIN sin
#define R_COS cos
#define R_ATAN2 atan2
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMOD fmod
#define R_RINT rint
#include "cav_oracle_impl.h"
#undef REAL
#undef SFX
#undef R_SIN
#undef R_COS
#undef R_ATAN2
#undef R_SQRT
#undef R_FABS
#undef R_FMOD
#undef R_RINT

#define REAL float
#define SFX(n) n##_f32
#define R_SIN sinf
#define R_COS cosf
#define R_ATAN2 atan2f
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FMOD fmodf
#define R_RINT rintf
#include "cav_oracle_impl.h"

/* ---------------------------------------------------------------------------------------------------------- */
void cavo_set_threads(int n) { g_threads = n; }
int cavo_get_max_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
/* When non-zero, RVO agents redo the reference's full N-agent doStep per ego (RVOPolicy.py:93 ->
 * R/RVOSimulator.cpp:163-183): N solves of which only the ego's is used.  Results are identical; only the
 * CPU cost changes.  Used to report a "faithful" CPU baseline next to the efficient one. */
void cavo_set_faithful_dostep(int on) { g_faithful_dostep = on; }

#define DISPATCH(fn, ...) (cfg->precision == CAV_F64 ? fn##_f64(__VA_ARGS__) : fn##_f32(__VA_ARGS__))

/* cfg.obs_f32 (float64 state, float32 `obs` / `reward` columns): the restatement computes in REAL; the float32 columns
 * are widened into scratch doubles before a phase and rounded back after it.  Values that merely persist (laserscan
 * history rows, rows of absent agents) survive the round trip unchanged: float(double(float x)) == float x. */
typedef struct {
    cav_config_t cfg;
    cav_buffers_t b;
    double *obs, *rew;
    float *fobs, *frew;
    size_t no, nr;
} widen_t;
static void widen_begin(const cav_config_t **cfg, const cav_buffers_t **b, widen_t *w) {
    w->obs = NULL;
    if (!((*cfg)->obs_f32 && (*cfg)->precision == CAV_F64)) return;
    w->cfg = **cfg;
    w->cfg.obs_f32 = 0;
    w->b = **b;
    w->no = (size_t)w->cfg.num_envs * w->cfg.num_agents * w->cfg.obs_len;
    w->nr = (size_t)w->cfg.num_envs * w->cfg.num_agents;
    w->fobs = (float *)w->b.obs;
    w->frew = (float *)w->b.reward;
    w->obs = (double *)malloc(sizeof(double) * w->no);
    w->rew = (double *)malloc(sizeof(double) * w->nr);
    for (size_t i = 0; i < w->no; ++i) w->obs[i] = (double)w->fobs[i];
    for (size_t i = 0; i < w->nr; ++i) w->rew[i] = (double)w->frew[i];
    w->b.obs = w->obs;
    w->b.reward = w->rew;
    *cfg = &w->cfg;
    *b = &w->b;
}
static void widen_end(widen_t *w) {
    if (!w->obs) return;
    for (size_t i = 0; i < w->no; ++i) w->fobs[i] = (float)w->obs[i];
    for (size_t i = 0; i < w->nr; ++i) w->frew[i] = (float)w->rew[i];
    free(w->obs);
    free(w->rew);
}

int cavo_policy(const cav_config_t *cfg, const cav_buffers_t *b, const float *ext) {
    if (!g_ga3c_blob) return DISPATCH(policy, cfg, b, ext);
    widen_t w; /* GA3C-CADRL agents read their observation row */
    widen_begin(&cfg, &b, &w);
    const int rc = DISPATCH(policy, cfg, b, ext);
    widen_end(&w);
    return rc;
}
int cavo_dynamics(const cav_config_t *cfg, const cav_buffers_t *b, double dt) { return DISPATCH(dynamics, cfg, b, dt); }
int cavo_pairwise_obs(const cav_config_t *cfg, const cav_buffers_t *b) {
    widen_t w;
    widen_begin(&cfg, &b, &w);
    const int rc = DISPATCH(pairwise_obs, cfg, b);
    widen_end(&w);
    return rc;
}
int cavo_laserscan(const cav_config_t *cfg, const cav_buffers_t *b) {
    widen_t w;
    widen_begin(&cfg, &b, &w);
    const int rc = DISPATCH(laserscan, cfg, b);
    widen_end(&w);
    return rc;
}
int cavo_reset(const cav_config_t *cfg, const cav_buffers_t *b, const uint8_t *mask, const cav_scenario_bank_t *bank,
               const int32_t *scenario_idx, int accumulate_stats) {
    widen_t w;
    widen_begin(&cfg, &b, &w);
    const int rc = DISPATCH(reset, cfg, b, mask, bank, scenario_idx, accumulate_stats);
    widen_end(&w);
    return rc;
}
/* CollisionAvoidanceEnv.step, collision_avoidance_env.py:156-234 */
int cavo_step(const cav_config_t *cfg, const cav_buffers_t *b, const float *ext, double dt) {
    if (dt <= 0) dt = cfg->dt;
    int rc;
    if ((rc = cavo_policy(cfg, b, ext))) return rc;
    if ((rc = cavo_dynamics(cfg, b, dt))) return rc;
    if ((rc = cavo_pairwise_obs(cfg, b))) return rc;
    if (cfg->obs_offset[CAV_OBS_LASERSCAN] >= 0)
        if ((rc = cavo_laserscan(cfg, b))) return rc;
    return 0;
}

/* util.compute_time_to_impact as restated in cav_oracle_impl.h, exported for the unit tests that pin it against values
 * computed by the reference's own Python function (tests/test_tti_native_cpu.py). */
double cavo_time_to_impact_f64(double px, double py, double a, double b, double hvx, double hvy, double ovx, double ovy,
                               double r) {
    return time_to_impact_f64(px, py, a, b, hvx, hvy, ovx, ovy, r);
}
float cavo_time_to_impact_f32(float px, float py, float a, float b, float hvx, float hvy, float ovx, float ovy, float r) {
    return time_to_impact_f32(px, py, a, b, hvx, hvy, ovx, ovy, r);
}
