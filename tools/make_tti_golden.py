#!/usr/bin/env python
"""Mint tests/golden/generator/time_to_impact.npz from the reference's own util.compute_time_to_impact (this container
only): 6000 random host / other configurations -- approaching, receding, grazing the collision cone, overlapping discs,
near-vertical and near-zero relative velocities -- with the value the reference computes for each."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.load_reference()
from gym_collision_avoidance.envs import util as ref_util  # noqa: E402


def main():
    rng = np.random.RandomState(2024)
    rows = []
    for k in range(6000):
        hp = rng.uniform(-6, 6, 2)
        op = hp + rng.uniform(-5, 5, 2)
        r = rng.uniform(0.4, 1.6)
        mode = k % 6
        hv = rng.uniform(-2, 2, 2)
        ov = rng.uniform(-2, 2, 2)
        if mode == 1:      # aimed at the other agent (+ noise): inside the cone
            d = op - hp
            hv = d / np.linalg.norm(d) * rng.uniform(0.2, 2.0) + rng.normal(0, 0.05, 2)
            ov = rng.normal(0, 0.05, 2)
        elif mode == 2:    # overlapping discs -> 0.0
            op = hp + rng.uniform(-0.3, 0.3, 2)
        elif mode == 3:    # (near-)vertical relative velocity
            ov = hv + np.array([rng.choice([0.0, 1e-7, -3e-6]), rng.uniform(-2, 2)])
        elif mode == 4:    # (near-)zero relative velocity
            ov = hv + rng.choice([0.0, 1e-6, 5e-6], 2)
        elif mode == 5:    # grazing: relative velocity along a tangent of the collision cone
            v1, v2 = ref_util.tangent_vecs_from_external_pt(hp[0], hp[1], op[0], op[1], r)
            if v1 is not None:
                t = v1 if rng.rand() < 0.5 else v2
                hv = t / np.linalg.norm(t) * rng.uniform(0.3, 1.5) * (1 + rng.choice([0.0, 1e-12, -1e-12]))
                ov = np.zeros(2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            devnull = open(os.devnull, "w")
            old = sys.stdout
            sys.stdout = devnull  # the reference prints warnings for det <= 0
            try:
                tti = ref_util.compute_time_to_impact(hp, op, hv, ov, r)
            finally:
                sys.stdout = old
                devnull.close()
        rows.append([hp[0], hp[1], op[0], op[1], hv[0], hv[1], ov[0], ov[1], r, float(tti)])
    a = np.array(rows)
    out = os.path.join(ROOT, "tests", "golden", "generator", "time_to_impact.npz")
    np.savez_compressed(out, cases=a)
    t = a[:, 9]
    print("wrote %s: %d cases; inf %d, zero %d, finite %d, nan %d" % (out, len(a), np.isinf(t).sum(), (t == 0).sum(),
                                                                        np.isfinite(t).sum() - (t == 0).sum(),
                                                                        np.isnan(t).sum()))


if __name__ == "__main__":
    main()
