"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

ctypes wrapper over ``oracle/libcav_oracle.so`` (the C restatement in ``oracle/cav_oracle.c``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs import this.
It reuses the product's ABI structs (``gym_collision_avoidance_b200.abi``) so the oracle and the CUDA library
are fed byte-identical configs and buffers.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gym_collision_avoidance_b200 import abi  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libcav_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cav_oracle.c", "cav_oracle_impl.h")] + \
           [os.path.join(_ROOT, "include", "cavoid.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        cfgp, bufp, vp = ctypes.POINTER(abi.CavConfig), ctypes.POINTER(abi.CavBuffers), ctypes.c_void_p
        L.cavo_policy.argtypes = [cfgp, bufp, vp]
        L.cavo_dynamics.argtypes = [cfgp, bufp, ctypes.c_double]
        L.cavo_pairwise_obs.argtypes = [cfgp, bufp]
        L.cavo_laserscan.argtypes = [cfgp, bufp]
        L.cavo_reset.argtypes = [cfgp, bufp, vp, ctypes.POINTER(abi.CavScenarioBank), vp, ctypes.c_int]
        L.cavo_step.argtypes = [cfgp, bufp, vp, ctypes.c_double]
        L.cavo_set_threads.argtypes = [ctypes.c_int]
        L.cavo_set_faithful_dostep.argtypes = [ctypes.c_int]
        L.cavo_get_max_threads.restype = ctypes.c_int
        fp = ctypes.POINTER(ctypes.c_float)
        L.cavo_orca_solve.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_int, fp, fp, vp, ctypes.POINTER(ctypes.c_int)]
        L.cavo_kd_build.argtypes = [vp, vp, vp, ctypes.c_int]
        L.cavo_kd_build.restype = None
        L.cavo_ga3c_forward.argtypes = [vp, vp, ctypes.c_int, vp, vp]
        L.cavo_ga3c_forward.restype = None
        L.cavo_set_ga3c.argtypes = [vp]
        L.cavo_set_ga3c.restype = None
        L.cavo_ga3c_gather.argtypes = [cfgp, vp]
        L.cavo_ga3c_gather.restype = None
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data


_GA3C_BLOB = None  # keeps the array behind cavo_set_ga3c's pointer alive


def set_ga3c(blob):
    """Weights for CAV_POLICY_GA3C agents in Oracle.policy / Oracle.step (float32 blob, include/cavoid.h layout)."""
    global _GA3C_BLOB
    import numpy as np
    _GA3C_BLOB = None if blob is None else np.ascontiguousarray(blob, dtype=np.float32)
    lib().cavo_set_ga3c(_ptr(_GA3C_BLOB))


def ga3c_forward(blob, X):
    """(softmax [rows, 11] float32, argmax [rows] int32) of the C restatement of the GA3C-CADRL network."""
    import numpy as np
    blob = np.ascontiguousarray(blob, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    assert X.ndim == 2 and X.shape[1] == 138
    probs = np.zeros((X.shape[0], 11), np.float32)
    act = np.zeros(X.shape[0], np.int32)
    lib().cavo_ga3c_forward(_ptr(blob), _ptr(X), X.shape[0], _ptr(probs), _ptr(act))
    return probs, act


def ga3c_gather(c):
    """int16 [138]: packed-observation index feeding each network input (-1 = zero padding)."""
    import numpy as np
    src = np.zeros(138, np.int16)
    lib().cavo_ga3c_gather(ctypes.byref(c), _ptr(src))
    return src


class Oracle(object):
    """Runs the C restatement on a HostState (NumPy SoA) in place."""

    def __init__(self, c, threads=0):
        self.c = c
        self.L = lib()
        self.threads = threads

    def _t(self):
        self.L.cavo_set_threads(self.threads)

    def policy(self, st, ext_actions=None):
        self._t()
        b = st.buffers()
        assert self.L.cavo_policy(ctypes.byref(self.c), ctypes.byref(b), _ptr(ext_actions)) == 0

    def dynamics(self, st, dt=None):
        self._t()
        b = st.buffers()
        assert self.L.cavo_dynamics(ctypes.byref(self.c), ctypes.byref(b), self.c.dt if dt is None else dt) == 0

    def pairwise_obs(self, st):
        self._t()
        b = st.buffers()
        assert self.L.cavo_pairwise_obs(ctypes.byref(self.c), ctypes.byref(b)) == 0

    def laserscan(self, st):
        self._t()
        b = st.buffers()
        assert self.L.cavo_laserscan(ctypes.byref(self.c), ctypes.byref(b)) == 0

    def step(self, st, ext_actions=None, dt=0.0):
        self._t()
        b = st.buffers()
        assert self.L.cavo_step(ctypes.byref(self.c), ctypes.byref(b), _ptr(ext_actions), dt) == 0

    def reset(self, st, bank, env_mask=None, scenario_idx=None, accumulate_stats=False):
        self._t()
        b = st.buffers()
        bs = bank.struct()
        assert self.L.cavo_reset(ctypes.byref(self.c), ctypes.byref(b), _ptr(env_mask), ctypes.byref(bs),
                                 _ptr(scenario_idx), int(accumulate_stats)) == 0
