"""-m gpu: the C ABI with memory that does NOT come from torch -- plain cudaMalloc through cuda-python's runtime bindings,
host<->device copies with cudaMemcpy, the legacy default stream -- to show that `cav_buffers_t` really is "borrowed device
pointers from any CUDA allocator" (include/cavoid.h).  Reset + steps are compared with the oracle."""
import ctypes

import numpy as np
import pytest

import common
from cav_oracle import Oracle
from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank
from gym_collision_avoidance_b200.envs.vec_env import kinds_array
from gym_collision_avoidance_b200.state import HostBank, HostState

pytestmark = pytest.mark.gpu


def _cudart():
    try:
        from cuda.bindings import runtime as rt
    except ImportError:
        from cuda import cudart as rt
    return rt


def _ok(ret):
    err = ret[0]
    assert int(err) == 0, "CUDA runtime error %s" % (err,)
    return ret[1] if len(ret) == 2 else ret[1:]


@pytest.mark.parametrize("N,prec,obs_f32", [(6, "f64", False), (10, "f64", True), (12, "f32", False)])
def test_c_abi_on_cudamalloc_memory(N, prec, obs_f32):
    rt = _cudart()
    _ok(rt.cudaSetDevice(0))
    E = 48
    cfg = make_config(N, evaluate=True)
    c = to_cav_config(cfg, E, N, precision=prec, track_aux=True, history_len=32, store_history=True, obs_f32=obs_f32)
    rows = random_scenario_bank(64, N, seed=21)
    bank = HostBank(rows, kinds=kinds_array(64, N, "RVO"))
    hs = HostState(c)
    orc = Oracle(c)

    L = abi.load_library()
    h = ctypes.c_void_p()
    assert L.cav_create(ctypes.byref(c), ctypes.byref(h)) == 0
    dev = {}
    bufs = abi.CavBuffers()
    for name, arr in hs.arrays.items():
        p = _ok(rt.cudaMalloc(arr.nbytes))
        _ok(rt.cudaMemcpy(p, arr.ctypes.data, arr.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
        dev[name] = p
        setattr(bufs, name, int(p))
    assert L.cav_bind(h, ctypes.byref(bufs)) == 0, L.cav_last_error(h)
    d_rows = _ok(rt.cudaMalloc(bank.rows.nbytes))
    d_kinds = _ok(rt.cudaMalloc(bank.kinds.nbytes))
    _ok(rt.cudaMemcpy(d_rows, bank.rows.ctypes.data, bank.rows.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
    _ok(rt.cudaMemcpy(d_kinds, bank.kinds.ctypes.data, bank.kinds.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
    sb = abi.CavScenarioBank()
    sb.rows, sb.kinds, sb.num_active, sb.num_scenarios = int(d_rows), int(d_kinds), None, 64

    def download():
        _ok(rt.cudaDeviceSynchronize())
        out = HostState(c)
        for name, arr in out.arrays.items():
            _ok(rt.cudaMemcpy(arr.ctypes.data, dev[name], arr.nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost))
        return out

    assert L.cav_reset(h, None, ctypes.byref(sb), None, 0, None) == 0, L.cav_last_error(h)  # stream NULL = legacy default
    orc.reset(hs, bank)
    tol = 1e-12 if prec == "f64" else 3e-6
    common.compare_states(download(), hs, N, tol, tol)
    for s in range(4):
        for name, arr in hs.arrays.items():  # every step starts from the oracle's state on both sides
            _ok(rt.cudaMemcpy(dev[name], arr.ctypes.data, arr.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
        assert L.cav_step(h, None, 0.0, None) == 0, L.cav_last_error(h)
        orc.step(hs)
        want = hs
        got = download()
        cols = [k for k in want.arrays if k not in ("stats", "episode", "kinds", "num_active") and
                not (obs_f32 and k in ("obs", "reward"))]
        common.compare_states(got, want, N, max(tol, 1e-9), max(tol, 1e-9), cols=cols)
        if obs_f32:
            common.assert_close("obs", got.obs, want.obs, 2e-7, 1e-7)
    assert L.cav_last_launch_count(h) >= 2
    assert L.cav_destroy(h) == 0
    for p in list(dev.values()) + [d_rows, d_kinds]:
        _ok(rt.cudaFree(p))
