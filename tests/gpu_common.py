"""Helpers for the -m gpu parity tests: run the CUDA library on a HostState and bring the result back."""
import numpy as np
import torch

from gym_collision_avoidance_b200.device import Cavoid, DeviceBank, DeviceState


def to_dev(a, dev="cuda:0"):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(dev)


class GpuRunner(object):
    """Creates a handle for config `c`, uploads a HostState, runs entry points, downloads the result."""

    def __init__(self, c, dev="cuda:0"):
        self.c = c
        self.dev = dev
        self.cav = Cavoid(c, dev)

    def upload(self, hs):
        had = self.cav.state.static_map_bits is not None
        self.cav.state.load_host(hs)
        if hs.static_map_bits is not None and not had:
            self.cav.bind()  # the static-map pointer is part of the bound buffer set

    def set_static_map(self, bits):
        if bits is not None:
            self.cav.state.set_static_map(bits)
            self.cav.bind()

    def download(self):
        torch.cuda.synchronize()
        return self.cav.state.to_host()

    def step(self, hs, ext=None, dt=0.0):
        self.upload(hs)
        self.cav.step(to_dev(ext, self.dev), dt)
        return self.download()

    def bank(self, host_bank):
        return DeviceBank(host_bank.rows, host_bank.kinds, host_bank.num_active, device=self.dev)
