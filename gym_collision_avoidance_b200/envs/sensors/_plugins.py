"""Sensor plugins (reference ``envs/sensors/*.py``): descriptors; sensing runs on the device
(``cav_k_pairwise_obs`` for OtherAgentsStatesSensor, ``cav_k_laserscan`` for LaserScanSensor)."""


class Sensor(object):
    """Reference ``sensors/Sensor.py:5-23``."""
    name = None

    def __init__(self):
        pass

    def sense(self, agents, agent_index, top_down_map):
        raise NotImplementedError("sensing runs on the device")

    def set_args(self, args):
        for arg, value in args.items():
            setattr(self, arg, value)


class OtherAgentsStatesSensor(Sensor):
    """Reference ``sensors/OtherAgentsStatesSensor.py:58-144``: (K,7) rows of the closest other agents in the ego
    frame, sorted by (round(dist,2), p_orth)."""

    def __init__(self, max_num_other_agents_observed=None, agent_sorting_method=None):
        Sensor.__init__(self)
        self.name = "other_agents_states"
        self.max_num_other_agents_observed = max_num_other_agents_observed
        self.agent_sorting_method = agent_sorting_method


class LaserScanSensor(Sensor):
    """Reference ``sensors/LaserScanSensor.py:24-101``: 512 beams x 60 range samples against the 160x160 grid."""

    def __init__(self):
        Sensor.__init__(self)
        self.name = "laserscan"
        self.range_resolution = 0.1
        self.max_range = 6
        self.min_range = 0
