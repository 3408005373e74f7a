"""Synthetic random-start / random-goal scenario banks (host side, NumPy, seeded).

Follows the reference's generator in distribution, not in its RNG stream: ``generate_rand_test_case_multi``
(reference ``policies/CADRL/scripts/multi/gen_rand_testcases.py:111-142``) draws 15 % swap cases (:322-377),
15 % circle cases (:380-423) and 70 % rejection-sampled random cases (:144-233) whose starts (and goals) keep
``r_i + r_j + GETTING_CLOSE_RANGE`` apart and whose start-goal distance exceeds half the side length.  Rows are
the reference's "cadrl test case" rows ``[px, py, gx, gy, pref_speed, radius]`` (``test_cases.py:508``).
"""
import numpy as np

GETTING_CLOSE_RANGE = 0.2  # reference CADRL/scripts/multi/global_var.py:8


def _rand_case(rng, n, side, speed_bnds, radius_bnds):
    tc = np.zeros((n, 6))
    for i in range(n):
        tc[i, 5] = rng.uniform(*radius_bnds)
        tc[i, 4] = max(rng.uniform(*speed_bnds), rng.uniform(*speed_bnds))
        while True:
            side *= 1.01
            start = rng.uniform(-side, side, 2)
            end = rng.uniform(-side, side, 2)
            if i > 0:
                rr = tc[:i, 5] + tc[i, 5] + GETTING_CLOSE_RANGE
                if (np.linalg.norm(tc[:i, 0:2] - start, axis=1) < rr).any():
                    continue
                if (np.linalg.norm(tc[:i, 2:4] - end, axis=1) < rr).any():
                    continue
            if np.linalg.norm(start - end) > side * 0.5:
                break
        tc[i, 0:2], tc[i, 2:4] = start, end
    return tc


def _swap_case(rng, n, side, speed_bnds, radius_bnds):
    tc = _rand_case(rng, n, side, speed_bnds, radius_bnds)
    for i in range(0, n - 1, 2):  # pairs swap positions
        tc[i + 1, 0:2], tc[i + 1, 2:4] = tc[i, 2:4].copy(), tc[i, 0:2].copy()
    return tc


def _circle_case(rng, n, side, speed_bnds, radius_bnds):
    tc = np.zeros((n, 6))
    tc[:, 5] = rng.uniform(radius_bnds[0], radius_bnds[1], n)
    tc[:, 4] = np.maximum(rng.uniform(speed_bnds[0], speed_bnds[1], n), rng.uniform(speed_bnds[0], speed_bnds[1], n))
    need = np.sum(2 * tc[:, 5] + GETTING_CLOSE_RANGE) / (2 * np.pi) * 1.15
    R = max(side, need)
    th0 = rng.uniform(0, 2 * np.pi)
    th = th0 + 2 * np.pi * np.arange(n) / n + rng.uniform(-0.05, 0.05, n)
    tc[:, 0], tc[:, 1] = R * np.cos(th), R * np.sin(th)
    tc[:, 2], tc[:, 3] = -tc[:, 0], -tc[:, 1]
    return tc


def default_side_length(num_agents, rng):
    """Config.TEST_CASE_ARGS['side_length'] (reference config.py:56-59), extended ~sqrt(N) for N > 10 to keep the
    agent density of the 10-agent cases."""
    if num_agents < 5:
        return rng.uniform(4, 5)
    if num_agents <= 10:
        return rng.uniform(6, 8)
    return rng.uniform(6, 8) * np.sqrt(num_agents / 10.0)


def random_scenario_bank(num_scenarios, num_agents, seed=0, side_length=None, speed_bnds=(0.5, 2.0),
                         radius_bnds=(0.2, 0.8), mix=(0.15, 0.15)):
    """[S, N, 8] float64 bank rows (heading toward the goal, reference test_cases.py:553-557)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((num_scenarios, num_agents, 8))
    for s in range(num_scenarios):
        side = default_side_length(num_agents, rng) if side_length is None else float(side_length)
        u = rng.rand()
        if u < mix[0]:
            tc = _swap_case(rng, num_agents, side, speed_bnds, radius_bnds)
        elif u < mix[0] + mix[1]:
            tc = _circle_case(rng, num_agents, side, speed_bnds, radius_bnds)
        else:
            tc = _rand_case(rng, num_agents, side, speed_bnds, radius_bnds)
        out[s, :, :6] = tc
        out[s, :, 6] = np.arctan2(tc[:, 3] - tc[:, 1], tc[:, 2] - tc[:, 0])
    return out
