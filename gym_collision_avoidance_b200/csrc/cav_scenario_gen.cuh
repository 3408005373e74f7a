// cav_scenario_gen.cuh -- the scenario generator proper (see cav_scenario.cu), written as host/device code so that
// the SAME source is compiled into the kernel by nvcc and into a host test harness by g++
// (tests/native/scenario_host.cpp, tests/test_scenario_native_cpu.py: Philox known answers, construction invariants
// and distribution on the CPU).  Reference: policies/CADRL/scripts/multi/gen_rand_testcases.py (see cav_scenario.cu).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define CAV_HD __host__ __device__
#else
#define CAV_HD
#endif

#ifndef CAV_BANK_ROW
#define CAV_BANK_ROW 8
#endif

namespace cav {
namespace scenario_gen {

CAV_HD inline uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct Philox {
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;
    uint32_t out[4];
    int have;

    CAV_HD Philox(uint64_t seed, uint64_t stream) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        c0 = 0;
        c1 = 0;
        c2 = (uint32_t)stream;
        c3 = (uint32_t)(stream >> 32);
        have = 0;
    }
    CAV_HD void round_(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t k0, uint32_t k1) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
        a = hi1 ^ b ^ k0;
        b = lo1;
        c = hi0 ^ d ^ k1;
        d = lo0;
    }
    CAV_HD void refill() {
        uint32_t a = c0, b = c1, c = c2, d = c3, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round_(a, b, c, d, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = a;
        out[1] = b;
        out[2] = c;
        out[3] = d;
        if (++c0 == 0) ++c1;
        have = 4;
    }
    CAV_HD uint32_t next32() {
        if (have == 0) refill();
        return out[4 - (have--)];
    }
    // uniform double in [0, 1) with 53 random bits (the construction of NumPy's random_sample)
    CAV_HD double rand() {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
};

struct GenParams {
    int S, N;
    double side_lo, side_hi, speed_lo, speed_hi, radius_lo, radius_hi, close_range;
    uint64_t seed;
    int random_heading;  // training mode: heading ~ U(-pi, pi) (test_cases.py:558-559) instead of toward the goal
};

struct V2 {
    double x, y;
};
CAV_HD inline V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
CAV_HD inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
CAV_HD inline V2 operator*(double s, V2 a) { return V2{s * a.x, s * a.y}; }
CAV_HD inline double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
CAV_HD inline double norm(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }

// distance from p3 to the segment p1->p2 (gen_rand_testcases.py:93-108)
CAV_HD double point_to_segment(V2 p1, V2 p2, V2 p3) {
    const V2 d = p2 - p1;
    const double nd = norm(d);
    double u = (nd < 1e-5) ? 0.0 : dot(d, p3 - p1) / (nd * nd);
    u = fmax(0.0, fmin(u, 1.0));
    return norm(p3 - (p1 + u * d));
}

// closest approach of two points moving x1->x2 and y1->y2 over the same unit time, start excluded (:53-90)
CAV_HD double segments_min_dist(V2 x1, V2 x2, V2 y1, V2 y2) {
    const double end_dist = norm(x2 - y2);
    double critical = end_dist;
    const V2 z = (x2 - x1) - (y2 - y1);
    if (norm(z) > 0.0) {
        const double t = -dot(x1 - y1, z) / dot(z, z);
        if (t > 0.0 && t < 1.0) critical = norm((x1 + t * (x2 - x1)) - y1 - t * (y2 - y1));
    }
    return fmin(end_dist, critical);
}

// both agents can drive straight at their preferred speeds without coming within `radius` (:425-444)
CAV_HD bool straight_lines_clear(V2 x1, V2 x2, double s1, V2 y1, V2 y2, double s2, double radius) {
    const double t1 = norm(x2 - x1) / s1, t2 = norm(y2 - y1) / s2;
    V2 xc, yc;
    if (t1 < t2) {
        xc = x2;
        yc = y1 + (1.0 / t2) * (t1 * (y2 - y1));
        if (point_to_segment(yc, y2, xc) < radius) return false;
    } else {
        xc = x1 + (1.0 / t1) * (t2 * (x2 - x1));
        yc = y2;
        if (point_to_segment(xc, x2, yc) < radius) return false;
    }
    const double d = fmin(fmin(norm(x1 - y1), norm(xc - yc)), segments_min_dist(x1, xc, y1, yc));
    return !(d < radius);
}

struct Rows {
    double *r;  // [N][CAV_BANK_ROW] of this scenario
    CAV_HD V2 start(int i) const { return V2{r[i * CAV_BANK_ROW + 0], r[i * CAV_BANK_ROW + 1]}; }
    CAV_HD V2 goal(int i) const { return V2{r[i * CAV_BANK_ROW + 2], r[i * CAV_BANK_ROW + 3]}; }
    CAV_HD double speed(int i) const { return r[i * CAV_BANK_ROW + 4]; }
    CAV_HD double radius(int i) const { return r[i * CAV_BANK_ROW + 5]; }
    CAV_HD void place(int i, V2 s, V2 g) {
        r[i * CAV_BANK_ROW + 0] = s.x;
        r[i * CAV_BANK_ROW + 1] = s.y;
        r[i * CAV_BANK_ROW + 2] = g.x;
        r[i * CAV_BANK_ROW + 3] = g.y;
    }
};

CAV_HD void draw_radius_speed(Philox &g, const GenParams &p, Rows &tc, int i) {
    tc.r[i * CAV_BANK_ROW + 5] = (p.radius_hi - p.radius_lo) * g.rand() + p.radius_lo;
    const double s1 = (p.speed_hi - p.speed_lo) * g.rand() + p.speed_lo;
    const double s2 = (p.speed_hi - p.speed_lo) * g.rand() + p.speed_lo;
    tc.r[i * CAV_BANK_ROW + 4] = fmax(s1, s2);
}

CAV_HD bool too_close(const Rows &tc, const GenParams &p, int i, V2 start, V2 end) {
    const double ri = tc.radius(i);
    for (int j = 0; j < i; ++j) {
        const double gap = tc.radius(j) + ri + p.close_range;
        if (norm(start - tc.start(j)) < gap || norm(end - tc.goal(j)) < gap) return true;
    }
    return false;
}

// returns false when the rejection loop diverged
CAV_HD bool rand_case(Philox &g, const GenParams &p, Rows &tc, double side) {
    const double side_limit = 1e3 * side;
    for (int i = 0; i < p.N; ++i) {
        draw_radius_speed(g, p, tc, i);
        V2 start, end;
        while (true) {
            side *= 1.01;
            if (side > side_limit) return false;
            start.x = side * 2 * g.rand() - side;
            start.y = side * 2 * g.rand() - side;
            end.x = side * 2 * g.rand() - side;
            end.y = side * 2 * g.rand() - side;
            if (too_close(tc, p, i, start, end)) continue;
            if (i >= 1) {  // agents must interact: reject when every earlier pair could simply drive straight
                bool all_clear = true;
                for (int j = 0; j < i && all_clear; ++j)
                    all_clear = straight_lines_clear(tc.start(j), tc.goal(j), tc.speed(j), start, end, tc.speed(i),
                                                     tc.radius(j) + tc.radius(i) + p.close_range);
                if (all_clear) continue;
            }
            if (norm(start - end) > side * 0.5) break;
        }
        tc.place(i, start, end);
    }
    return true;
}

CAV_HD void ring_case(Philox &g, const GenParams &p, Rows &tc, bool swap) {
    const double kPi = 3.141592653589793;
    const double r_min = p.N / 2.0;
    double r = g.rand() * 2.0 + r_min;
    V2 offset{0.0, 0.0};
    double r_swap = 0.0;
    if (swap) {
        r_swap = 1.5 + g.rand() * 2.0;
        offset.y = 1.0 + r_min + g.rand() * 2.0;
        if (g.rand() > 0.5) offset.y = -offset.y;
    }
    for (int i = 0; i < p.N; ++i) {
        draw_radius_speed(g, p, tc, i);
        V2 start, end;
        if (swap && i == 0) {
            start = V2{-r_swap, 0.0};
            end = V2{r_swap, 0.0};
        } else if (swap && i == 1) {
            start = V2{r_swap, 0.0};
            end = V2{-r_swap, 0.0};
        } else {
            int tries = 0;
            while (true) {
                if (tries > 10) {
                    r *= 1.01;
                    tries = 0;
                }
                const double a0 = g.rand() * 2 * kPi - kPi, a1 = kPi + a0;
                start = V2{r * cos(a0), r * sin(a0)} + offset;
                end = V2{r * cos(a1), r * sin(a1)} + offset;
                if (too_close(tc, p, i, start, end)) {
                    ++tries;
                    continue;
                }
                break;
            }
        }
        tc.place(i, start, end);
    }
}

// One scenario: side draw + mixture (get_testcase_random -> generate_rand_test_case_multi), rows[N][CAV_BANK_ROW],
// heading toward the goal (EVALUATE_MODE, test_cases.py:553-557) or, in training mode, drawn from U(-pi, pi) per agent
// after the case like cadrl_test_case_to_agents does (:558-559).  Returns the case kind (0 swap, 1 circle, 2 random).
CAV_HD inline int generate_scenario(double *rows, const GenParams &p, uint64_t scenario) {
    Philox g(p.seed, scenario);
    Rows tc{rows};
    int kind;
    while (true) {
        const double side = p.side_lo + (p.side_hi - p.side_lo) * g.rand();  // get_testcase_random's side draw
        const double u = g.rand();
        if (u < 0.15) {
            ring_case(g, p, tc, true);
            kind = 0;
            break;
        }
        if (u > 0.15 && u < 0.3) {
            ring_case(g, p, tc, false);
            kind = 1;
            break;
        }
        kind = 2;
        if (rand_case(g, p, tc, side)) break;
    }
    for (int i = 0; i < p.N; ++i) {
        double *r = tc.r + i * CAV_BANK_ROW;
        r[6] = p.random_heading ? (-3.14159265358979323846 + 2.0 * 3.14159265358979323846 * g.rand())
                                : atan2(r[3] - r[1], r[2] - r[0]);
        r[7] = 0.0;
    }
    return kind;
}

}  // namespace scenario_gen
}  // namespace cav
