"""Device-side containers and the handle wrapper over ``libcavoid_sm100a.so``.

PyTorch is used only as the CUDA memory container / stream provider (``tensor.data_ptr()``,
``torch.cuda.current_stream()``); every computation on the step path happens inside the library's sm_100a kernels.
"""
import ctypes

import numpy as np
import torch

from . import abi
from .state import HostState, column_specs, make_bank

_TORCH_KIND = {"f32": torch.float32, "u32": torch.int32, "i32": torch.int32, "u8": torch.uint8, "u64": torch.int64}
_NP_VIEW = {"u32": np.uint32, "u64": np.uint64}


class DeviceState(object):
    """SoA state as torch CUDA tensors ([E, N, ...] env-major; see ``state.column_specs``)."""

    def __init__(self, c, device):
        self.c = c
        self.device = torch.device(device)
        self.real = torch.float64 if c.precision == abi.CAV_F64 else torch.float32
        self.kinds_of = {}
        self.t = {}
        for name, (shape, kind) in column_specs(c).items():
            dt = self.real if kind == "real" else _TORCH_KIND[kind]
            self.t[name] = torch.zeros(shape, dtype=dt, device=self.device)
            self.kinds_of[name] = kind
        self.t["num_active"].fill_(c.num_agents)
        if "kd_order" in self.t:
            self.t["kd_order"].copy_(torch.arange(c.num_agents, dtype=torch.uint8, device=self.device).expand(
                c.num_envs, c.num_agents))
        self.static_map_bits = None

    def __getattr__(self, k):
        t = self.__dict__.get("t", {})
        if k in t:
            return t[k]
        raise AttributeError(k)

    def buffers(self):
        b = abi.CavBuffers()
        for n in abi.BUFFER_FIELDS:
            a = self.t.get(n)
            setattr(b, n, a.data_ptr() if a is not None else None)
        if self.static_map_bits is not None:
            b.static_map_bits = self.static_map_bits.data_ptr()
        return b

    def set_static_map(self, bits_u32):
        """bits: uint32 array [rows * ceil(cols/32)] (row-major bitmap of Map.static_map)."""
        a = np.ascontiguousarray(bits_u32, dtype=np.uint32).view(np.int32)
        self.static_map_bits = torch.from_numpy(a.copy()).to(self.device)

    def load_host(self, hs, cols=None):
        """Host columns -> device (all of them, or only the named ones)."""
        for name, arr in hs.arrays.items():
            if cols is not None and name not in cols:
                continue
            a = arr
            if a.dtype == np.uint32:
                a = a.view(np.int32)
            elif a.dtype == np.uint64:
                a = a.view(np.int64)
            self.t[name].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        if hs.static_map_bits is not None and self.static_map_bits is None:
            self.set_static_map(hs.static_map_bits)

    def to_host(self):
        hs = HostState(self.c)
        for name, t in self.t.items():
            a = t.detach().cpu().numpy()
            k = self.kinds_of[name]
            if k in _NP_VIEW:
                a = a.view(_NP_VIEW[k])
            hs.arrays[name][...] = a
        return hs


class DeviceBank(object):
    """Scenario bank resident in HBM (rows float64 [S, N, 8], optional kinds / num_active)."""

    def __init__(self, rows, kinds=None, num_active=None, device="cuda"):
        r, k, na = make_bank(rows, kinds, num_active)
        self.rows = torch.from_numpy(r).to(device)
        self.kinds = None if k is None else torch.from_numpy(k.view(np.int32).copy()).to(device)
        self.num_active = None if na is None else torch.from_numpy(na).to(device)
        self.num_scenarios = r.shape[0]

    @classmethod
    def empty(cls, num_scenarios, num_agents, device="cuda"):
        """Uninitialised bank for `Cavoid.generate_scenarios` to fill."""
        b = cls.__new__(cls)
        b.rows = torch.zeros((num_scenarios, num_agents, abi.CAV_BANK_ROW), dtype=torch.float64, device=device)
        b.kinds, b.num_active, b.num_scenarios = None, None, int(num_scenarios)
        b.case_kind = torch.zeros((num_scenarios,), dtype=torch.int32, device=device)
        return b

    def struct(self):
        s = abi.CavScenarioBank()
        s.rows = self.rows.data_ptr()
        s.kinds = self.kinds.data_ptr() if self.kinds is not None else None
        s.num_active = self.num_active.data_ptr() if self.num_active is not None else None
        s.num_scenarios = self.num_scenarios
        return s


def _ptr(t):
    return None if t is None else t.data_ptr()


class Cavoid(object):
    """One handle bound to one device + its SoA state.  Thin: every method is one C-ABI call."""

    def __init__(self, cav_config, device="cuda:0", state=None):
        self.L = abi.load_library()
        self.c = cav_config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise abi.CavoidError("the step path runs on CUDA only (no CPU fallback)")
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.L.cav_create(ctypes.byref(self.c), ctypes.byref(self._h))
        if rc != 0:
            raise abi.CavoidError("cav_create failed (%s): %s" % (
                abi.STATUS_NAMES.get(rc, rc), self.L.cav_last_error(None).decode()))
        self.state = state if state is not None else DeviceState(self.c, self.device)
        self.bind()
        self.launches = 0

    def bind(self):
        b = self.state.buffers()
        self._check(self.L.cav_bind(self._h, ctypes.byref(b)), "cav_bind")

    def _check(self, rc, what):
        if rc != 0:
            raise abi.CavoidError("%s failed (%s): %s" % (what, abi.STATUS_NAMES.get(rc, rc),
                                                          self.L.cav_last_error(self._h).decode()))

    def _stream(self, stream):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(s.cuda_stream)

    def step(self, ext_actions=None, dt=0.0, stream=None):
        self._check(self.L.cav_step(self._h, _ptr(ext_actions), float(dt), self._stream(stream)), "cav_step")
        self.launches += self.L.cav_last_launch_count(self._h)

    def reset(self, bank, env_mask=None, scenario_idx=None, accumulate_stats=False, stream=None):
        bs = bank.struct()
        self._check(self.L.cav_reset(self._h, _ptr(env_mask), ctypes.byref(bs), _ptr(scenario_idx),
                                     int(accumulate_stats), self._stream(stream)), "cav_reset")
        self.launches += self.L.cav_last_launch_count(self._h)

    def bind_ga3c(self, blob):
        """Attach the GA3C-CADRL network (cav_bind_ga3c): ``blob`` is the float32 weight blob (numpy array or CUDA tensor,
        ``envs/policies/GA3C_CADRL/network.py: load_blob``); None detaches.  The tensor is kept alive by this object."""
        if blob is None:
            self._ga3c_blob = None
            self._check(self.L.cav_bind_ga3c(self._h, None), "cav_bind_ga3c")
            return
        if not isinstance(blob, torch.Tensor):
            import numpy as np
            blob = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32))
        blob = blob.to(device=self.device, dtype=torch.float32).contiguous()
        if blob.numel() != abi.CAV_GA3C_BLOB_FLOATS:
            raise abi.CavoidError("GA3C-CADRL weight blob must hold %d floats" % abi.CAV_GA3C_BLOB_FLOATS)
        self._ga3c_blob = blob
        self._check(self.L.cav_bind_ga3c(self._h, blob.data_ptr()), "cav_bind_ga3c")

    def k_ga3c(self, probs=None, stream=None):
        """The network for every live CAV_POLICY_GA3C agent -> state.actions; ``probs`` (optional float32 [E*N*11] CUDA
        tensor) receives the softmax rows of those agents."""
        self._check(self.L.cav_k_ga3c(self._h, _ptr(probs), self._stream(stream)), "cav_k_ga3c")
        self.launches += self.L.cav_last_launch_count(self._h)

    def env_schedule(self, stream=None):
        """(order, cost) of the policy + move kernel's environment schedule (cav_env_schedule) as int32 tensors, or None
        when the handle has no schedule (fewer than 2048 environments, CAV_ENV_ORDER=0, sub-warp-group policy kernel)."""
        E = self.c.num_envs
        order = torch.empty(E, dtype=torch.int32, device=self.device)
        cost = torch.empty(E, dtype=torch.int16, device=self.device)
        rc = self.L.cav_env_schedule(self._h, order.data_ptr(), cost.data_ptr(), self._stream(stream))
        if rc == abi.CAV_ERR_BAD_CONFIG:  # no schedule on this handle
            return None
        self._check(rc, "cav_env_schedule")
        return order, cost.to(torch.int32) & 0xFFFF

    def generate_scenarios(self, bank, seed, side_length=None, speed_bnds=(0.5, 2.0), radius_bnds=(0.2, 0.8),
                           stream=None, random_heading=False):
        """Fill `bank.rows` on the device with reference-distributed random scenarios (cav_generate_scenarios).
        side_length = (lo, hi) of the uniform side draw; default: the reference's Config.TEST_CASE_ARGS."""
        n = self.c.num_agents
        if side_length is None:
            side_length = (4.0, 5.0) if n < 5 else (6.0, 8.0)
        g = abi.CavScenarioGen()
        g.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        g.side_length[0], g.side_length[1] = float(side_length[0]), float(side_length[1])
        g.speed_bnds[0], g.speed_bnds[1] = float(speed_bnds[0]), float(speed_bnds[1])
        g.radius_bnds[0], g.radius_bnds[1] = float(radius_bnds[0]), float(radius_bnds[1])
        g.random_heading = int(bool(random_heading))
        if tuple(bank.rows.shape[1:]) != (n, abi.CAV_BANK_ROW) or bank.rows.dtype != torch.float64:
            raise abi.CavoidError("bank rows must be float64 [S, %d, %d]" % (n, abi.CAV_BANK_ROW))
        ck = getattr(bank, "case_kind", None)
        self._check(self.L.cav_generate_scenarios(self._h, ctypes.byref(g), bank.rows.data_ptr(), _ptr(ck),
                                                  int(bank.num_scenarios), self._stream(stream)),
                    "cav_generate_scenarios")
        self.launches += self.L.cav_last_launch_count(self._h)
        return bank

    def k_policy(self, ext_actions=None, stream=None):
        self._check(self.L.cav_k_policy(self._h, _ptr(ext_actions), self._stream(stream)), "cav_k_policy")
        self.launches += self.L.cav_last_launch_count(self._h)

    def k_dynamics(self, dt=0.0, stream=None):
        self._check(self.L.cav_k_dynamics(self._h, float(dt), self._stream(stream)), "cav_k_dynamics")
        self.launches += self.L.cav_last_launch_count(self._h)

    def k_pairwise_obs(self, stream=None):
        self._check(self.L.cav_k_pairwise_obs(self._h, self._stream(stream)), "cav_k_pairwise_obs")
        self.launches += self.L.cav_last_launch_count(self._h)

    def k_laserscan(self, stream=None):
        self._check(self.L.cav_k_laserscan(self._h, self._stream(stream)), "cav_k_laserscan")
        self.launches += self.L.cav_last_launch_count(self._h)

    def k_observe(self, env_mask=None, stream=None):
        self._check(self.L.cav_k_observe(self._h, _ptr(env_mask), self._stream(stream)), "cav_k_observe")
        self.launches += self.L.cav_last_launch_count(self._h)

    def stats(self):
        raw = self.state.stats.detach().cpu().numpy().view(np.uint64)
        out = {}
        for i, name in enumerate(abi.STAT_NAMES):
            out[name] = float(raw[i:i + 1].view(np.float64)[0]) if i in abi.STAT_IS_DOUBLE else int(raw[i])
        return out

    def close(self):
        if self._h:
            self.L.cav_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
