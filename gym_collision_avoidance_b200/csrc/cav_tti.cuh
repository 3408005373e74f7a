// cav_tti.cuh -- util.compute_time_to_impact for the 'time_to_impact' neighbour ordering (host/device source: compiled
// into the sub-warp-group pairwise+obs kernel by nvcc and into tests/native/tti_host.cpp by g++;
// tests/test_tti_native_cpu.py pins it against values computed by the reference's own Python function).
#pragma once
#include "cav_hd.cuh"

namespace cav {

CAV_HDI float hd_sqrt(float a) { return sqrtf(a); }
CAV_HDI double hd_sqrt(double a) { return sqrt(a); }
CAV_HDI float hd_abs(float a) { return fabsf(a); }
CAV_HDI double hd_abs(double a) { return fabs(a); }
template <typename T>
CAV_HDI T hd_inf() {
    return (T)INFINITY;
}

// util.compute_time_to_impact (util.py:23-79) with tangent_vecs_from_external_pt (:81-112), operation for operation
// (same expression order as the oracle): host at (px,py) with velocity hv, other at (a,b) with velocity ov, r = the
// combined radius.  Only + - * / sqrt, so the result is the reference's to the last bit in float64.
template <typename T>
CAV_HD T time_to_impact(T px, T py, T a, T b, T hvx, T hvy, T ovx, T ovy, T r) {
    const T v0 = hvx - ovx, v1 = hvy - ovy;
    const T sq = (px - a) * (px - a) + (py - b) * (py - b) - r * r;
    if (sq < (T)0) return (T)0;  // already inside the collision zone
    const T sqrt_term = hd_sqrt(sq);
    const T xnum1 = (r * r) * (px - a), xnum2 = r * (py - b) * sqrt_term;
    const T ynum1 = (r * r) * (py - b), ynum2 = r * (px - a) * sqrt_term;
    const T den = (px - a) * (px - a) + (py - b) * (py - b);
    const T c1x = ((xnum1 + xnum2) / den + a) - px, c1y = ((ynum1 - ynum2) / den + b) - py;
    const T c2x = ((xnum1 - xnum2) / den + a) - px, c2y = ((ynum1 + ynum2) / den + b) - py;
    const T k1v = c1x * v1 - c1y * v0, k12 = c1x * c2y - c1y * c2x;
    const T k2v = c2x * v1 - c2y * v0, k21 = c2x * c1y - c2y * c1x;
    if (!(k1v * k12 >= (T)0 && k2v * k21 >= (T)0)) return hd_inf<T>();  // v_rel outside the collision cone
    if (hd_abs(v0) < (T)1e-5 && hd_abs(v1) < (T)1e-5) return hd_inf<T>();
    T x1, x2, y1, y2;
    if (hd_abs(v0) < (T)1e-5) {  // vertical v_rel
        const T A = (T)1, B = (T)-2 * b, C = b * b + (px - a) * (px - a) - r * r;
        x1 = x2 = px;
        y1 = (-B + hd_sqrt(B * B - (T)4 * A * C)) / ((T)2 * A);
        y2 = (-B - hd_sqrt(B * B - (T)4 * A * C)) / ((T)2 * A);
    } else {
        const T m = v1 / v0;
        const T A = (T)1 + m * m;
        const T B = (T)-2 * a + (T)2 * m * (py - b - m * px);
        const T C = a * a - r * r + (m * px - (py - b)) * (m * px - (py - b));
        x1 = (-B + hd_sqrt(B * B - (T)4 * A * C)) / ((T)2 * A);
        x2 = (-B - hd_sqrt(B * B - (T)4 * A * C)) / ((T)2 * A);
        y1 = m * (x1 - px) + py;
        y2 = m * (x2 - px) + py;
    }
    const T d1 = hd_sqrt((x1 - px) * (x1 - px) + (y1 - py) * (y1 - py));
    const T d2 = hd_sqrt((x2 - px) * (x2 - px) + (y2 - py) * (y2 - py));
    const T d = d2 < d1 ? d2 : d1;  // Python min(d1, d2)
    return d / hd_sqrt(v0 * v0 + v1 * v1);
}

}  // namespace cav
