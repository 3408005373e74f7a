"""Agent.take_action as the thread-per-agent KERNELS implement it (csrc/cav_move.cuh::tpa_move with cav_math.cuh --
host/device source, compiled here with g++) against the oracle's dynamics step, on the CPU, bit for bit: from every
recorded state of every reference-minted golden episode (and their float32 twins), after the oracle chose the actions.
Covers UnicycleDynamics / MaxTurnRate / ExternalDynamics, the done-agent latch, turning_dir, the ego frame, timeouts,
at-goal flags, past actions / velocities and the trajectory rows (reference agent.py:192-289, dynamics/*.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import common
from cav_oracle import Oracle
from gym_collision_avoidance_b200.state import HostState

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
CSRC = os.path.join(ROOT, "gym_collision_avoidance_b200", "csrc")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libmove_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-fPIC",
                           "-shared", "-x", "c++", os.path.join(HERE, "native", "move_host.cpp"), "-I", CSRC, "-I",
                           os.path.join(ROOT, "include"), "-o", out])
    L = ctypes.CDLL(out)
    L.host_move.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    L.host_move.restype = ctypes.c_int
    return L


NAMES = common.golden_names() + common.golden_names(cpu_only=True)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", NAMES)
def test_kernel_take_action_source_equals_oracle_bitwise(host, name, prec):
    g = common.Golden(name)
    T = g.T
    hist = int(np.max(g.z["st_step_num"])) + 2
    c = g.cav_config(T, prec, track_aux=True, history_len=hist)
    st = HostState(c)
    for t in range(T):
        g.inject(st, t, t)
    ext = np.ascontiguousarray(g.z["ext_actions"][:T], dtype=np.float32)
    orc = Oracle(c)
    orc.policy(st, ext)                      # the oracle chooses the actions (float32 column)
    a, b = st.copy(), st.copy()
    orc.dynamics(a)                          # oracle: Agent.take_action for every agent
    bufs = b.buffers()
    assert host.host_move(ctypes.addressof(c), ctypes.addressof(bufs), 0.0) == 0
    moved = 0
    for col in ("pv", "ht", "ego", "flags", "step_num", "aux0", "past_actions", "past_vel", "history"):
        x, y = a.arrays[col], b.arrays[col]
        assert x.tobytes() == y.tobytes(), (col, np.argwhere(x != y)[:4].tolist())
        moved += int((x != st.arrays[col]).sum())
    assert moved > 0
