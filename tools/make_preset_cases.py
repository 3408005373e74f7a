#!/usr/bin/env python
"""Mint gym_collision_avoidance_b200/envs/data/preset_cases.npz from the UNMODIFIED reference (build container only):
the hard-coded small test suite of ``preset_testCases(num_agents)`` (reference ``envs/test_cases.py:626-897``, rows
``[px, py, gx, gy, pref_speed, radius]``) for every agent count it defines.  Data, not code: the same route as the golden
fixtures (tools/make_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.load_reference()
from gym_collision_avoidance.envs import test_cases as tc  # noqa: E402

out = {}
for n in (1, 2, 3, 4, 5, 6, 10, 20):
    cases = tc.preset_testCases(n)
    for k, c in enumerate(cases):
        out["n%d_c%d" % (n, k)] = np.asarray(c, dtype=np.float64)
    print(n, len(cases))
d = os.path.join(ROOT, "gym_collision_avoidance_b200", "envs", "data")
os.makedirs(d, exist_ok=True)
np.savez_compressed(os.path.join(d, "preset_cases.npz"), **out)
