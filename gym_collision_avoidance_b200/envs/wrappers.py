"""Observation packing helpers (reference ``envs/wrappers.py:143-173``,
``MultiagentDictToMultiagentArrayWrapper``): the device already produces the packed array; these convert between
the packed ``[num_agents, obs_len]`` layout and the reference's dict-of-dicts."""
import numpy as np

from . import config as _config


def packed_to_dict(cfg, packed, num_slots=None):
    """[N, obs_len] -> {agent_index: {state: ndarray}} in ``cfg.STATES_IN_OBS`` order."""
    off, total = _config.obs_layout(cfg)
    packed = np.asarray(packed)
    out = {}
    for i in range(packed.shape[0] if num_slots is None else num_slots):
        d = {}
        for s in cfg.STATES_IN_OBS:
            size = cfg.STATE_INFO_DICT[s]["size"]
            n = _config.state_size(cfg, s)
            v = packed[i, off[s]:off[s] + n]
            d[s] = np.array(v.reshape(size) if isinstance(size, tuple) else (v[0] if n == 1 else v))
        out[i] = d
    return out


def dict_to_packed(cfg, obs):
    """Inverse of :func:`packed_to_dict` (what MultiagentDictToMultiagentArrayWrapper.observation builds)."""
    off, total = _config.obs_layout(cfg)
    arr = np.zeros((len(obs), total))
    for i in range(len(obs)):
        for s in cfg.STATES_IN_OBS:
            n = _config.state_size(cfg, s)
            arr[i, off[s]:off[s] + n] = np.asarray(obs[i][s], dtype=np.float64).reshape(-1)
    return arr
