#!/usr/bin/env python
"""Mints tests/golden/ga3c/ga3c_iros18.npz in the build container (where /root/reference exists):

  blob      float32 [170784]  the IROS18 checkpoint's variables + the graph's normalisation constants, packed by the
                              product loader (envs/policies/GA3C_CADRL/network.py: load_blob, CRC-checked bundle read)
  X         float32 [R, 138]  network inputs: every sequence length 0..19, lengths beyond 19, garbage after the sequence
  probs     float32 [R, 11]   Softmax:0 of the reference's graph on X, evaluated by oracle/tf1_graph.py (numpy execution of
                              the meta-graph's own nodes with the checkpoint's variables; TensorFlow is not installed)
  value     float32 [R]       Squeeze:0 (the value head), same way -- not used by the policy, kept as a second fetch
  actions   float64 [11, 2]   the action table exactly as the reference builds it (GA3C_CADRL/network.py:12-16: np.mgrid)
  crc       per-variable CRC-32C check happens inside the bundle reader (a mismatch raises)

The other two checkpoints the reference ships (run-20190727_*) have the same graph; their blobs are compared against the
same interpreter by tests/test_ga3c_cpu.py when /root/reference is present."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tf1_graph  # noqa: E402
from gym_collision_avoidance_b200.envs.policies.GA3C_CADRL import network  # noqa: E402

REF = "/root/reference/gym_collision_avoidance/envs/policies/GA3C_CADRL/checkpoints/"


def sample_inputs(rows, seed):
    rng = np.random.RandomState(seed)
    X = np.zeros((rows, 138), np.float32)
    for r in range(rows):
        n = r % 23  # 0..19 and three lengths beyond the 19 time steps
        X[r, 0] = n
        X[r, 1] = rng.uniform(0.05, 14.0)          # dist_to_goal
        X[r, 2] = rng.uniform(-np.pi, np.pi)       # heading_ego_frame
        X[r, 3] = rng.uniform(0.2, 2.0)            # pref_speed
        X[r, 4] = rng.uniform(0.2, 1.5)            # radius
        m = min(n, 19)
        o = X[r, 5:5 + 7 * m].reshape(m, 7)
        o[:, 0:2] = rng.uniform(-6, 6, (m, 2))     # p_parallel, p_orthog
        o[:, 2:4] = rng.uniform(-2, 2, (m, 2))     # v_parallel, v_orthog
        o[:, 4] = rng.uniform(0.2, 1.5, m)         # radius
        o[:, 5] = o[:, 4] + X[r, 4]                # combined radius
        o[:, 6] = rng.uniform(-0.2, 8.0, m)        # dist_2_other
        if r % 5 == 0:
            X[r, 5 + 7 * m:] = rng.uniform(-5, 5, 138 - 5 - 7 * m)  # stale rows behind the sequence must not matter
    return X


if __name__ == "__main__":
    prefix = REF + "IROS18/network_01900000"
    blob = network.load_blob(prefix)
    it = tf1_graph.load(prefix)
    X = sample_inputs(276, seed=2018)
    probs = it.run("Softmax", {"X": X})
    assert it.iterations == 19
    value = it.run("Squeeze", {"X": X})
    actions = np.mgrid[1.0:1.1:0.5, -np.pi / 6:np.pi / 6 + 0.01:np.pi / 12].reshape(2, -1).T
    for speed in (0.5, 0.0):
        actions = np.vstack([actions, np.mgrid[speed:speed + 0.1:0.5, -np.pi / 6:np.pi / 6 + 0.01:np.pi / 6].reshape(2, -1).T])
    out = os.path.join(ROOT, "tests", "golden", "ga3c", "ga3c_iros18.npz")
    np.savez_compressed(out, blob=blob, X=X, probs=probs, value=value.astype(np.float32), actions=actions)
    print("wrote", out, os.path.getsize(out), "bytes; argmax histogram", np.bincount(probs.argmax(1), minlength=11))
