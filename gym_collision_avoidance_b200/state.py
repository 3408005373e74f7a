"""Struct-of-arrays environment state: column specs shared by the device container (torch CUDA tensors, used
by the product) and the host container (NumPy, used by tests to feed the oracle identical bytes).

Every per-agent column is [E, N, ...] env-major (``include/cavoid.h``): one environment's agent block is
contiguous in every column.  Field meaning follows the reference Agent attributes (``envs/agent.py:59-138``).
"""
import ctypes

import numpy as np

from . import abi


def column_specs(c):
    """name -> (shape, kind) for a cav_config_t ``c``; kind in {'real','f32','u32','i32','u8','u64'}.  With ``c.obs_f32``
    (float64 state, float32 observation) the ``obs`` and ``reward`` columns are float32."""
    E, N = c.num_envs, c.num_agents
    ok = "f32" if (c.obs_f32 and c.precision == abi.CAV_F64) else "real"
    specs = {
        "pv": ((E, N, 4), "real"), "goal": ((E, N, 4), "real"), "ht": ((E, N, 4), "real"),
        "ego": ((E, N, 4), "real"),
        "flags": ((E, N), "u32"), "kinds": ((E, N), "u32"), "step_num": ((E, N), "i32"),
        "num_active": ((E,), "i32"),
        "actions": ((E, N, 2), "f32"), "reward": ((E, N), ok), "min_dist": ((E, N), "real"),
        "done": ((E, N), "u8"), "game_over": ((E,), "u8"),
        "obs": ((E, N, c.obs_len), ok),
        "episode": ((E,), "i32"), "stats": ((abi.CAV_STAT_NUM,), "u64"),
    }
    if N > abi.CAV_KD_MAX_LEAF:  # RVO2's persistent kd-tree agent permutation (cavoid.h: kd_order)
        specs["kd_order"] = ((E, N), "u8")
    if c.obs_offset[abi.CAV_OBS_LASERSCAN] >= 0:
        specs["laser_count"] = ((E, N), "u32")
    if c.store_history and c.history_len > 0:
        specs["history"] = ((E, N, c.history_len, abi.CAV_HISTORY_ROW), "real")
    if c.track_aux:
        specs["aux0"] = ((E, N, 4), "real")
        specs["past_actions"] = ((E, N, 4), "real")
        specs["past_vel"] = ((E, N, 4), "real")
    return specs


_NP_KIND = {"f32": np.float32, "u32": np.uint32, "i32": np.int32, "u8": np.uint8, "u64": np.uint64}


class HostState(object):
    """NumPy SoA state (tests / oracle side)."""

    def __init__(self, c):
        self.c = c
        self.real = np.float64 if c.precision == abi.CAV_F64 else np.float32
        self.arrays = {}
        for name, (shape, kind) in column_specs(c).items():
            dt = self.real if kind == "real" else _NP_KIND[kind]
            self.arrays[name] = np.zeros(shape, dtype=dt)
        self.arrays["num_active"][:] = c.num_agents
        if "kd_order" in self.arrays:
            self.arrays["kd_order"][:] = np.arange(c.num_agents, dtype=np.uint8)
        self.static_map_bits = None

    def __getattr__(self, k):
        arrs = self.__dict__.get("arrays", {})
        if k in arrs:
            return arrs[k]
        raise AttributeError(k)

    def buffers(self):
        b = abi.CavBuffers()
        for n in abi.BUFFER_FIELDS:
            a = self.arrays.get(n)
            setattr(b, n, a.ctypes.data if a is not None else None)
        if self.static_map_bits is not None:
            b.static_map_bits = self.static_map_bits.ctypes.data
        return b

    def copy(self):
        o = HostState(self.c)
        for k, v in self.arrays.items():
            o.arrays[k][...] = v
        o.static_map_bits = self.static_map_bits
        return o


def make_bank(rows, kinds=None, num_active=None):
    """Host scenario bank: rows [S, N, 6|7|8] -> float64 [S, N, 8] (heading toward goal when absent,
    reference ``test_cases.py:553-557``)."""
    rows = np.asarray(rows, dtype=np.float64)
    S, N, W = rows.shape
    out = np.zeros((S, N, abi.CAV_BANK_ROW), dtype=np.float64)
    out[:, :, :min(W, 8)] = rows[:, :, :min(W, 8)]
    if W < 7:
        vec = out[:, :, 2:4] - out[:, :, 0:2]
        out[:, :, 6] = np.arctan2(vec[:, :, 1], vec[:, :, 0])
    k = None if kinds is None else np.ascontiguousarray(kinds, dtype=np.uint32).reshape(S, N)
    na = None if num_active is None else np.ascontiguousarray(num_active, dtype=np.int32).reshape(S)
    return out, k, na


class HostBank(object):
    def __init__(self, rows, kinds=None, num_active=None):
        self.rows, self.kinds, self.num_active = make_bank(rows, kinds, num_active)

    def struct(self):
        s = abi.CavScenarioBank()
        s.rows = self.rows.ctypes.data
        s.kinds = self.kinds.ctypes.data if self.kinds is not None else None
        s.num_active = self.num_active.ctypes.data if self.num_active is not None else None
        s.num_scenarios = self.rows.shape[0]
        return s
