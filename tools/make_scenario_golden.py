#!/usr/bin/env python
"""Mint tests/golden/generator/scenario_cases.npz from the RUNNING reference generator (this container only).

For each (seed, num_agents): ``np.random.seed(seed)``; side = the ``get_testcase_random`` draw for the default
``Config.TEST_CASE_ARGS`` (reference test_cases.py:230-238); rows = ``generate_rand_test_case_multi(num_agents, side,
[0.5, 2.0], [0.2, 0.8])`` (reference policies/CADRL/scripts/multi/gen_rand_testcases.py:111-142).  The committed
fixture pins ``gym_collision_avoidance_b200.envs.scenarios.reference_test_case`` bit for bit (tests/test_scenarios_cpu.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.load_reference()
from gym_collision_avoidance.envs.policies.CADRL.scripts.multi import gen_rand_testcases as ref_tc  # noqa: E402

from gym_collision_avoidance_b200.envs import scenarios  # noqa: E402


def main():
    seeds, ns, sides, rows, kinds = [], [], [], [], []
    mism = 0
    for n in (2, 3, 4, 5, 6, 8, 10, 16, 20):
        for seed in range(40):
            np.random.seed(1000 * n + seed)
            lo, hi = (4, 5) if n < 5 else (6, 8)
            side = np.random.uniform(lo, hi)
            st = np.random.get_state()
            u = np.random.rand()
            np.random.set_state(st)
            tc = ref_tc.generate_rand_test_case_multi(n, side, [0.5, 2.0], [0.2, 0.8])
            rng = np.random.RandomState(1000 * n + seed)
            side2 = rng.uniform(lo, hi)
            mine = scenarios.reference_test_case(rng, n, side2)
            if side2 != side or not np.array_equal(mine, tc):
                mism += 1
                print("MISMATCH n=%d seed=%d u=%.3f max|d|=%g" % (n, seed, u, np.abs(mine - tc).max()))
            seeds.append(1000 * n + seed)
            ns.append(n)
            sides.append(side)
            kinds.append(0 if u < 0.15 else (1 if u < 0.3 else 2))
            pad = np.zeros((20, 6))
            pad[:n] = tc
            rows.append(pad)
    out = os.path.join(ROOT, "tests", "golden", "generator", "scenario_cases.npz")
    np.savez_compressed(out, seeds=np.array(seeds), num_agents=np.array(ns), side=np.array(sides),
                        kind=np.array(kinds), rows=np.array(rows))
    print("wrote %s: %d cases (%d swap, %d circle, %d random), %d mismatches vs the restatement" % (
        out, len(seeds), kinds.count(0), kinds.count(1), kinds.count(2), mism))
    return mism


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
