"""Host scenario generator == the reference's generator, draw for draw.

tests/golden/generator/scenario_cases.npz holds 360 cases minted from the RUNNING reference (tools/make_scenario_golden.py:
``np.random.seed(s)``, the ``get_testcase_random`` side-length draw, ``generate_rand_test_case_multi``); the restatement
in gym_collision_avoidance_b200/envs/scenarios.py must reproduce every row bit for bit from ``RandomState(s)``."""
import os

import numpy as np

from gym_collision_avoidance_b200.envs import scenarios

GOLD = os.path.join(os.path.dirname(__file__), "golden", "generator", "scenario_cases.npz")


def test_restatement_reproduces_reference_rows_bit_for_bit():
    g = np.load(GOLD)
    kinds = set()
    for seed, n, side, kind, rows in zip(g["seeds"], g["num_agents"], g["side"], g["kind"], g["rows"]):
        rng = np.random.RandomState(int(seed))
        mine_side = scenarios.reference_side_length(rng, int(n))
        assert mine_side == side
        tc = scenarios.reference_test_case(rng, int(n), mine_side)
        assert np.array_equal(tc, rows[:n]), (seed, n, kind)
        kinds.add(int(kind))
    assert kinds == {0, 1, 2}  # swap, circle and random cases all covered


def test_bank_rows_separated_and_heading_toward_goal():
    for n in (2, 5, 10, 20):
        b = scenarios.random_scenario_bank(48, n, seed=3)
        for s in range(b.shape[0]):
            p, q, r = b[s, :, 0:2], b[s, :, 2:4], b[s, :, 5]
            gap = r[:, None] + r[None] + scenarios.GETTING_CLOSE_RANGE - 1e-12
            dp = np.linalg.norm(p[:, None] - p[None], axis=-1) + np.eye(n) * 1e9
            dq = np.linalg.norm(q[:, None] - q[None], axis=-1) + np.eye(n) * 1e9
            assert (dp >= gap).all() and (dq >= gap).all()
        assert np.allclose(b[..., 6], np.arctan2(b[..., 3] - b[..., 1], b[..., 2] - b[..., 0]))
        assert (b[..., 4] >= 0.5).all() and (b[..., 4] <= 2.0).all() and (b[..., 5] >= 0.2).all() and (b[..., 5] <= 0.8).all()


def test_bank_is_the_continued_stream():
    """scenario s of the bank = the (s+1)-th get_testcase_random call on one seeded stream"""
    b = scenarios.random_scenario_bank(5, 4, seed=11)
    rng = np.random.RandomState(11)
    for s in range(5):
        tc = scenarios.reference_test_case(rng, 4, scenarios.reference_side_length(rng, 4))
        assert np.array_equal(b[s, :, :6], tc)


def test_global_stream_entry_point_matches_golden():
    """envs.test_cases.generate_rand_test_case_multi consumes np.random's global stream like the reference does"""
    from gym_collision_avoidance_b200.envs import test_cases
    g = np.load(GOLD)
    for k in (0, 50, 170, 300, 359):
        seed, n, side = int(g["seeds"][k]), int(g["num_agents"][k]), float(g["side"][k])
        np.random.seed(seed)
        lo, hi = (4, 5) if n < 5 else (6, 8)
        assert np.random.uniform(lo, hi) == side
        tc = test_cases.generate_rand_test_case_multi(n, side, [0.5, 2.0], [0.2, 0.8])
        assert np.array_equal(tc, g["rows"][k][:n])
