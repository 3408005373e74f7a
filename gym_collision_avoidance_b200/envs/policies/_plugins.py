"""Policy plugins with the reference's class names and flags (reference ``envs/policies/*.py``).

On this framework the action of every built-in policy is computed on the GPU (``cav_k_policy``); the classes
below are *descriptors*: they carry the reference's public attributes (``str``, ``is_external``,
``is_still_learning``) plus a ``kernel_id`` that the environment writes into the per-agent ``kinds`` column.
A user subclass without a ``kernel_id`` is rejected when the environment is built -- there is no CPU fallback.
"""
import numpy as np

from ... import abi


class Policy(object):
    """Reference ``policies/Policy.py:11-14``."""
    kernel_id = None

    def __init__(self, str="NoPolicy"):
        self.str = str
        self.is_still_learning = False
        self.is_external = False

    def find_next_action(self, obs, agents, i):
        raise NotImplementedError("actions of %s are computed on the device by cav_k_policy" % type(self).__name__)


class InternalPolicy(Policy):
    """Reference ``policies/InternalPolicy.py``: computes its action inside the environment."""

    def __init__(self, str="Internal"):
        Policy.__init__(self, str=str)


class ExternalPolicy(Policy):
    """Reference ``policies/ExternalPolicy.py:14-16``: the caller's ``[speed, delta_heading]`` is used as is."""
    kernel_id = abi.CAV_POLICY_EXTERNAL

    def __init__(self, str="External"):
        Policy.__init__(self, str=str)
        self.is_external = True

    def external_action_to_action(self, agent, external_action):
        return external_action

    def find_next_action(self, obs, agents, i):
        return None


class LearningPolicy(ExternalPolicy):
    """Reference ``policies/LearningPolicy.py:30-32``: ``[pref_speed*a0, max_heading_change*(2*a1-1)]``."""
    kernel_id = abi.CAV_POLICY_LEARNING

    def __init__(self):
        ExternalPolicy.__init__(self, str="learning")
        self.is_still_learning = True
        self.ppo_or_learning_policy = True

    def external_action_to_action(self, agent, external_action):
        heading_change = agent.max_heading_change * (2. * external_action[1] - 1.)
        speed = agent.pref_speed * external_action[0]
        return np.array([speed, heading_change])


class NonCooperativePolicy(InternalPolicy):
    """Reference ``policies/NonCooperativePolicy.py:21``: ``[pref_speed, -heading_ego_frame]``."""
    kernel_id = abi.CAV_POLICY_NONCOOP

    def __init__(self):
        InternalPolicy.__init__(self, str="NonCooperativePolicy")


class StaticPolicy(InternalPolicy):
    """Reference ``policies/StaticPolicy.py:21-22``: goal := position, action ``[0, 0]``."""
    kernel_id = abi.CAV_POLICY_STATIC

    def __init__(self):
        InternalPolicy.__init__(self, str="Static")


class RVOPolicy(InternalPolicy):
    """Reference ``policies/RVOPolicy.py:10-122`` (+ the RVO2 C++ solver): ORCA half-planes + LP in float32 on
    the device, turn clip at pi/6.  The constants the reference reads from Config at construction are taken from
    the environment's config when the handle is created."""
    kernel_id = abi.CAV_POLICY_RVO

    def __init__(self):
        InternalPolicy.__init__(self, str="RVO")
        self.max_delta_heading = np.pi / 6
        self.has_fixed_speed = False
        self.heading_noise = False


class GA3CCADRLPolicy(InternalPolicy):
    """Reference ``policies/GA3CCADRLPolicy.py:10-84``: the pre-trained GA3C-CADRL LSTM network (Everett et al., IROS 2018)
    picks one of 11 discrete ``[speed, heading change]`` actions from the agent's own observation.  Here the network runs
    on the GPU inside the step (``cav_ga3c.cu``); ``initialize_network`` reads the reference's TF1 checkpoint files
    without TensorFlow (``GA3C_CADRL/network.py``) and the environment binds the weights when it is built.

    The checkpoints are the reference's data and are not redistributed with this package: ``checkpt_dir`` is an absolute
    directory, or a name below ``$GA3C_CADRL_CHECKPOINTS`` (default: the reference's own
    ``envs/policies/GA3C_CADRL/checkpoints`` if ``gym_collision_avoidance`` is importable)."""
    kernel_id = abi.CAV_POLICY_GA3C

    def __init__(self):
        InternalPolicy.__init__(self, str="GA3C_CADRL")
        from .GA3C_CADRL import network
        self.possible_actions = network.Actions()
        self.device = "cuda:0"
        self.nn = network.NetworkVP_rnn(self.device, "network", self.possible_actions.num_actions)

    @staticmethod
    def _checkpoint_root():
        import os
        root = os.environ.get("GA3C_CADRL_CHECKPOINTS")
        if root:
            return root
        try:
            import importlib.util
            spec = importlib.util.find_spec("gym_collision_avoidance")
            if spec and spec.submodule_search_locations:
                return os.path.join(list(spec.submodule_search_locations)[0], "envs", "policies", "GA3C_CADRL", "checkpoints")
        except (ImportError, ValueError):
            pass
        raise abi.CavoidError("GA3C-CADRL checkpoints: set GA3C_CADRL_CHECKPOINTS or pass an absolute checkpt_dir")

    def initialize_network(self, **kwargs):
        """Same keywords as the reference (GA3CCADRLPolicy.py:24-47): ``checkpt_name`` (default ``network_01900000``),
        ``checkpt_dir`` (absolute, or relative to the checkpoint root; default ``IROS18``)."""
        import os
        name = kwargs.get("checkpt_name", "network_01900000")
        d = kwargs.get("checkpt_dir", "IROS18")
        if not os.path.isabs(d):
            d = os.path.join(self._checkpoint_root(), d)
        self.nn.simple_load(os.path.join(d, name))

    @property
    def weight_blob(self):
        return self.nn.blob
