"""Dynamics plugins (reference ``envs/dynamics/*.py``): descriptors with a ``kernel_id``; the integration runs
on the device inside the move phase (``cav_k_dynamics``)."""
from ... import abi


class Dynamics(object):
    """Reference ``dynamics/Dynamics.py:15-41``."""
    kernel_id = None

    def __init__(self, agent):
        self.agent = agent

    def step(self, action, dt):
        raise NotImplementedError("dynamics are integrated on the device by cav_k_dynamics")

    def update_ego_frame(self):
        """Host mirror of ``Dynamics.update_ego_frame`` (``Dynamics.py:31-41``), used only when an Agent object is
        (re)initialised on the host; per-step updates happen on the device."""
        import math
        import numpy as np
        a = self.agent
        a.ref_prll, a.ref_orth = a.get_ref()
        ang = np.arctan2(a.ref_prll[1], a.ref_prll[0])
        h = a.heading_global_frame - ang
        while h >= np.pi:
            h -= 2 * np.pi
        while h < -np.pi:
            h += 2 * np.pi
        a.heading_ego_frame = h
        cur_speed = math.sqrt(a.vel_global_frame[0] ** 2 + a.vel_global_frame[1] ** 2)
        a.vel_ego_frame = np.array([cur_speed * np.cos(h), cur_speed * np.sin(h)])


class UnicycleDynamics(Dynamics):
    """Reference ``dynamics/UnicycleDynamics.py:27-47``."""
    kernel_id = abi.CAV_DYN_UNICYCLE


class UnicycleDynamicsMaxTurnRate(Dynamics):
    """Reference ``dynamics/UnicycleDynamicsMaxTurnRate.py:30-43`` (3 rad/s clip)."""
    kernel_id = abi.CAV_DYN_UNICYCLE_MAX_TURN_RATE

    def __init__(self, agent):
        Dynamics.__init__(self, agent)
        self.max_turn_rate = 3.0


class ExternalDynamics(Dynamics):
    """Reference ``dynamics/ExternalDynamics.py:10-13``: state is set from outside (``Agent.set_state``)."""
    kernel_id = abi.CAV_DYN_EXTERNAL

    def step(self, action, dt):
        return
