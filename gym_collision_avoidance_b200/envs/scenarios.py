"""Synthetic random-start / random-goal scenario banks (host side, NumPy, seeded).

The reference's own generator restated draw for draw: ``generate_rand_test_case_multi`` (reference
``policies/CADRL/scripts/multi/gen_rand_testcases.py:111-142``) draws 15 % swap cases (:322-377), 15 % circle cases
(:380-423) and 70 % rejection-sampled random cases (:144-233) whose starts (and goals) keep
``r_i + r_j + GETTING_CLOSE_RANGE`` apart, whose start-goal distance exceeds half the side length and in which no
pair can simply drive straight.  Rows are the reference's "cadrl test case" rows
``[px, py, gx, gy, pref_speed, radius]`` (``test_cases.py:508``).
"""
import numpy as np

GETTING_CLOSE_RANGE = 0.2  # reference CADRL/scripts/multi/global_var.py:8


# ---------------------------------------------------------------------------------------------------------------
# The reference's generator, draw for draw.  ``reference_test_case(rng, ...)`` consumes a ``np.random.RandomState`` in
# exactly the order ``generate_rand_test_case_multi`` consumes the global ``np.random`` stream, and evaluates the
# accept / reject tests with the same NumPy expressions, so that the same seed gives the same rows bit for bit
# (pinned by tests/golden/generator/scenario_cases.npz, minted from the running reference by tools/make_scenario_golden.py).
# Reference: policies/CADRL/scripts/multi/gen_rand_testcases.py -- mixture :111-142, random case :144-233, swap case
# :322-377, circle case :380-423, straight-line test :425-444, segment distances :53-108.
# ---------------------------------------------------------------------------------------------------------------
_EPS = 1e-5  # reference global_var.py:9 (gb.EPS)


def _norm(v):
    return np.linalg.norm(v)


def _too_close(tc, i, start, end):
    """start (goal) closer than r_i + r_j + GETTING_CLOSE_RANGE to an earlier agent's start (goal)."""
    for j in range(i):
        gap = tc[j, 5] + tc[i, 5] + GETTING_CLOSE_RANGE
        if _norm(start - tc[j, 0:2]) < gap or _norm(end - tc[j, 2:4]) < gap:
            return True
    return False


def _draw_radius_speed(rng, tc, i, speed_bnds, radius_bnds):
    tc[i, 5] = (radius_bnds[1] - radius_bnds[0]) * rng.rand() + radius_bnds[0]
    s1 = (speed_bnds[1] - speed_bnds[0]) * rng.rand() + speed_bnds[0]
    s2 = (speed_bnds[1] - speed_bnds[0]) * rng.rand() + speed_bnds[0]
    tc[i, 4] = max(s1, s2)


def _point_to_segment(p1, p2, p3):
    """distance from p3 to the segment p1->p2 (:93-108)"""
    d = p2 - p1
    if _norm(d) < _EPS:
        u = 0.0
    else:
        u = np.dot(d, (p3 - p1)) / (_norm(d) ** 2.0)
    u = max(0.0, min(u, 1.0))
    return _norm(p3 - (p1 + u * d))


def _segments_min_dist(x1, x2, y1, y2):
    """closest approach of two points moving x1->x2 and y1->y2 over the same unit time, start excluded (:53-90, the
    single-end-point branch)"""
    x2 = x2.reshape((1, 2))
    y2 = y2.reshape((1, 2))
    end_dist = np.linalg.norm(x2 - y2, axis=1)
    critical = end_dist.copy()
    z_bar = (x2 - x1) - (y2 - y1)
    inds = np.where((np.linalg.norm(z_bar, axis=1) > 0))[0]
    t_bar = -np.sum((x1 - y1) * z_bar[inds, :], axis=1) / np.sum(z_bar[inds, :] * z_bar[inds, :], axis=1)
    t_rep = np.tile(t_bar, (2, 1)).transpose()
    dist_bar = np.linalg.norm(x1 + (x2[inds, :] - x1) * t_rep - y1 - (y2[inds, :] - y1) * t_rep, axis=1)
    inside = np.where((t_bar > 0) & (t_bar < 1.0))
    critical[inds[inside]] = dist_bar[inside]
    return np.amin(np.vstack((end_dist, critical)), axis=0)[0]


def _straight_lines_clear(x1, x2, s1, y1, y2, s2, radius):
    """True when both agents can drive straight to their goals at their preferred speeds without coming within
    `radius` (:425-444)"""
    t1 = _norm(x2 - x1) / s1
    t2 = _norm(y2 - y1) / s2
    if t1 < t2:
        x_crit = x2
        y_crit = y1 + t1 * (y2 - y1) / t2
        if _point_to_segment(y_crit, y2, x_crit) < radius:
            return False
    else:
        x_crit = x1 + t2 * (x2 - x1) / t1
        y_crit = y2
        if _point_to_segment(x_crit, x2, y_crit) < radius:
            return False
    if min(_norm(x1 - y1), _norm(x_crit - y_crit), _segments_min_dist(x1, x_crit, y1, y_crit)) < radius:
        return False
    return True


def _reference_rand_case(rng, n, side, speed_bnds, radius_bnds):
    tc = np.zeros((n, 6))
    side_limit = 1e3 * side
    for i in range(n):
        _draw_radius_speed(rng, tc, i, speed_bnds, radius_bnds)
        while True:
            side *= 1.01
            if side > side_limit:
                raise ScenarioRunaway("side length grew past %g" % side_limit)
            start = side * 2 * rng.rand(2,) - side
            end = side * 2 * rng.rand(2,) - side
            if _too_close(tc, i, start, end):
                continue
            # agents must interact: a pair that can simply drive straight is rejected (:209-222)
            if i >= 1 and all(_straight_lines_clear(tc[j, 0:2], tc[j, 2:4], tc[j, 4], start, end, tc[i, 4],
                                                    tc[j, 5] + tc[i, 5] + GETTING_CLOSE_RANGE) for j in range(i)):
                continue
            if _norm(start - end) > side * 0.5:
                break
        tc[i, 0:2], tc[i, 2:4] = start, end
    return tc


def _on_circle(r, angle, offset):
    return np.array([r * np.cos(angle), r * np.sin(angle)]) + offset


def _reference_ring_case(rng, n, speed_bnds, radius_bnds, swap):
    """circle case (antipodal start / goal on a ring that grows when crowded); the swap case adds a head-on pair on
    the x axis and shifts the ring sideways"""
    r_min = n / 2.0
    r = rng.rand() * 2.0 + r_min
    tc = np.zeros((n, 6))
    offset = np.zeros(2)
    if swap:
        r_swap = 1.5 + rng.rand() * 2.0
        offset = np.array([0, 1.0 + r_min + rng.rand() * 2.0])
        if rng.rand() > 0.5:
            offset = -offset
    for i in range(n):
        _draw_radius_speed(rng, tc, i, speed_bnds, radius_bnds)
        if swap and i == 0:
            start, end = np.array([-r_swap, 0.0]), np.array([r_swap, 0.0])
        elif swap and i == 1:
            start, end = np.array([r_swap, 0.0]), np.array([-r_swap, 0.0])
        else:
            tries = 0
            while True:
                if tries > 10:
                    r *= 1.01
                    tries = 0
                a0 = rng.rand() * 2 * np.pi - np.pi
                a1 = np.pi + a0
                if swap:
                    start, end = _on_circle(r, a0, offset), _on_circle(r, a1, offset)
                else:
                    start = np.array([r * np.cos(a0), r * np.sin(a0)])
                    end = np.array([r * np.cos(a1), r * np.sin(a1)])
                if _too_close(tc, i, start, end):
                    tries += 1
                    continue
                break
        tc[i, 0:2], tc[i, 2:4] = start, end
    return tc


def reference_test_case(rng, num_agents, side_length, speed_bnds=(0.5, 2.0), radius_bnds=(0.2, 0.8)):
    """One `[N, 6]` cadrl test case = the reference's ``generate_rand_test_case_multi(num_agents, side_length, ...)``
    for the same random stream: 15 % swap, 15 % circle, 70 % rejection-sampled random."""
    u = rng.rand()
    if u < 0.15:
        return _reference_ring_case(rng, num_agents, speed_bnds, radius_bnds, swap=True)
    if u > 0.15 and u < 0.3:
        return _reference_ring_case(rng, num_agents, speed_bnds, radius_bnds, swap=False)
    return _reference_rand_case(rng, num_agents, float(side_length), speed_bnds, radius_bnds)


def reference_side_length(rng, num_agents):
    """``get_testcase_random``'s side length draw for the default ``Config.TEST_CASE_ARGS`` (test_cases.py:230-238,
    config.py:56-59): U(4,5) below 5 agents, U(6,8) from 5 on.  (Crowded cases enlarge the arena by themselves: every
    rejected placement multiplies the side by 1.01, gen_rand_testcases.py:161.)"""
    lo, hi = (4, 5) if num_agents < 5 else (6, 8)
    return rng.uniform(lo, hi)


class ScenarioRunaway(RuntimeError):
    """The reference's rejection loop diverged: its side length grows 1 % per rejected placement, and once the arena is
    so large that every new pair can drive straight, every placement is rejected for good (the reference would spin
    forever)."""


def default_side_length(num_agents, rng):
    return reference_side_length(rng, num_agents)


def random_scenario_bank(num_scenarios, num_agents, seed=0, side_length=None, speed_bnds=(0.5, 2.0),
                         radius_bnds=(0.2, 0.8), random_heading=False):
    """[S, N, 8] float64 bank rows ``[px, py, gx, gy, pref_speed, radius, heading, 0]``: scenario s is what the
    reference's ``get_testcase_random`` builds from the stream ``RandomState(seed)`` after s earlier scenarios (side
    length draw, then ``generate_rand_test_case_multi``), heading toward the goal (EVALUATE_MODE, reference
    test_cases.py:553-557) or -- ``random_heading``, the reference's training mode (:558-559) -- one ``uniform(-pi, pi)``
    draw per agent after the case.  ``side_length`` fixes the side instead of drawing it (dense test banks)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((num_scenarios, num_agents, 8))
    s = 0
    while s < num_scenarios:
        side = reference_side_length(rng, num_agents) if side_length is None else float(side_length)
        try:
            tc = reference_test_case(rng, num_agents, side, speed_bnds, radius_bnds)
        except ScenarioRunaway:
            continue  # drop the diverged case, keep consuming the same stream
        out[s, :, :6] = tc
        out[s, :, 6] = np.arctan2(tc[:, 3] - tc[:, 1], tc[:, 2] - tc[:, 0])
        if random_heading:
            out[s, :, 6] = [rng.uniform(-np.pi, np.pi) for _ in range(num_agents)]
        s += 1
    return out
