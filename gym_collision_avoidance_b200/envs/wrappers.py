"""Observation wrappers with the reference's names (reference ``envs/wrappers.py``: ``MultiagentFlattenDictWrapper``
:11-134, ``FlattenDictWrapper`` :136-141, ``MultiagentDictToMultiagentArrayWrapper`` :143-173) over the single-env
drop-in, plus the packing helpers they share with the batched environment (the device already produces the
``[num_agents, obs_len]`` array of ``MultiagentDictToMultiagentArrayWrapper``; the wrappers only re-view it)."""
import numpy as np

from . import config as _config

__all__ = ["FlattenDictWrapper", "MultiagentFlattenDictWrapper", "MultiagentDictToMultiagentArrayWrapper",
           "packed_to_dict", "dict_to_packed"]


class _Box(object):
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class _ObservationWrapper(object):
    """gym.ObservationWrapper without gym: forwards everything, converts the observation of reset / step."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, k):
        return getattr(self.__dict__["env"], k)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, **kw):
        obs, info = self.env.reset(**kw)
        return self.observation(obs), info

    def step(self, action, **kw):
        out = self.env.step(action, **kw)
        return (self.observation(out[0]),) + tuple(out[1:])


class MultiagentFlattenDictWrapper(_ObservationWrapper):
    """Dict observation -> one long vector, agents then states (reference ``wrappers.py:11-134``)."""

    def __init__(self, env, dict_keys, max_num_agents):
        _ObservationWrapper.__init__(self, env)
        self.dict_keys = list(dict_keys)
        self.max_num_agents = int(max_num_agents)
        self.dict_observation_space = env.observation_space
        self.setup_obs(self.max_num_agents, self.dict_keys)
        self.observation_space = _Box(-np.inf, np.inf, self.obs_shape, "float32")
        lo, hi = self.observation_indices[0]["BOUNDS"]
        self.single_agent_observation_space = _Box(-np.inf, np.inf, (hi - lo,), "float32")

    def _key_shape(self, key):
        sp = self.dict_observation_space
        sub = sp.spaces[0] if 0 in sp.spaces else sp  # our space is {agent: {state: box}}
        return np.shape(sub.spaces[key].low) or (1,)

    def setup_obs(self, max_num_agents, dict_keys):
        self.observation_indices, size = {}, 0
        for agent in range(max_num_agents):
            self.observation_indices[agent] = {}
            lower = size
            for key in dict_keys:
                n = int(np.prod(self._key_shape(key)))
                self.observation_indices[agent][key] = [size, size + n]
                size += n
            self.observation_indices[agent]["BOUNDS"] = [lower, size]
        self.obs_shape = (size,)

    def observation(self, observation):
        assert isinstance(observation, dict)
        return np.concatenate([np.asarray(observation[a][k]).ravel() for a in range(self.max_num_agents)
                               for k in self.dict_keys])

    def observationArrayToDict(self, observation_array):
        assert isinstance(observation_array, np.ndarray) and observation_array.shape == self.observation_space.shape
        return {a: {k: observation_array[slice(*self.observation_indices[a][k])] for k in self.dict_keys}
                for a in range(self.max_num_agents)}

    def singleAgentObservationArrayToDict(self, observation_array, agent):
        return [{k: observation_array[e, slice(*self.observation_indices[agent][k])].reshape(self._key_shape(k))
                 for k in self.dict_keys} for e in range(observation_array.shape[0])]

    def keyToArrayInds(self, key):
        return [self.observation_indices[a][key] for a in range(self.max_num_agents)]

    def singleAgentObservationArray(self, observation_array, agent):
        return observation_array[slice(*self.observation_indices[agent]["BOUNDS"])]

    def singleAgentObservationInds(self, agent):
        return self.observation_indices[agent]["BOUNDS"]


class FlattenDictWrapper(MultiagentFlattenDictWrapper):
    """Agent 0's observation only (reference ``wrappers.py:136-141``; used with ``Config.TRAIN_SINGLE_AGENT``)."""

    def __init__(self, env, dict_keys):
        MultiagentFlattenDictWrapper.__init__(self, env, dict_keys, max_num_agents=1)


class MultiagentDictToMultiagentArrayWrapper(MultiagentFlattenDictWrapper):
    """Dict observation -> ``(max_num_agents, num_states_per_agent)`` array (reference ``wrappers.py:143-173``): the
    layout ``VecCollisionAvoidanceEnv`` produces on the device for every environment at once."""

    def setup_obs(self, max_num_agents, dict_keys):
        self.observation_indices = {}
        size = 0
        for agent in range(max_num_agents):
            self.observation_indices[agent] = {}
            size = 0
            for key in dict_keys:
                n = int(np.prod(self._key_shape(key)))
                self.observation_indices[agent][key] = [size, size + n]
                size += n
            self.observation_indices[agent]["BOUNDS"] = [0, size]
        self.obs_shape = (max_num_agents, size)

    def observation(self, observation):
        assert isinstance(observation, dict)
        obs = np.zeros(shape=self.obs_shape)
        for agent in range(self.max_num_agents):
            for key in self.dict_keys:
                low, high = self.observation_indices[agent][key]
                obs[agent][low:high] = np.asarray(observation[agent][key]).ravel()
        return obs


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
