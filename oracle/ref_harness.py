"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Imports the UNMODIFIED Python reference from /root/reference in THIS container
(SURVEY.md Appendix C).  The reference's import-time third-party imports that are
absent here (gym, matplotlib, imageio, scipy.misc, tensorflow, moviepy) are
satisfied with inert stubs -- none of the stubbed code runs on the step path --
and ``rvo2`` is served by ``oracle/rvo2_ref.py`` (ctypes over the reference's own
C++ RVO2 compiled into ``oracle/_ref``).

Only golden-vector generation (``tools/make_golden.py``) and the optional
``reference_available()``-gated CPU tests use this module; /root/reference does
not exist on the GPU box, so nothing in ``-m gpu`` tests, smoke() or bench.py
imports it.
"""
import importlib
import os
import sys
import types
from unittest import mock

REF_ROOT = os.environ.get("CAV_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "gym_collision_avoidance")) and os.path.exists(
        os.path.join(_HERE, "_ref", "librvo2_ref.so"))


def _stub(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: mock.MagicMock(name=name + "." + attr)  # type: ignore
    m.__path__ = []  # behave as a package
    sys.modules[name] = m
    return m


def _install_gym_stub():
    gym = types.ModuleType("gym")
    gym.__path__ = []

    class Env(object):
        metadata = {}

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, k):
            return getattr(self.env, k)

    class ObservationWrapper(Wrapper):
        pass

    spaces = types.ModuleType("gym.spaces")

    class Box(object):
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(object):
        def __init__(self, spaces=None):
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def __setitem__(self, k, v):
            self.spaces[k] = v

    class Discrete(object):
        def __init__(self, n, dtype=None):
            self.n = n

    spaces.Box, spaces.Dict, spaces.Discrete = Box, Dict, Discrete
    envs = types.ModuleType("gym.envs")
    envs.__path__ = []
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    logger = types.ModuleType("gym.logger")
    logger.set_level = lambda *a, **k: None
    gym.Env, gym.Wrapper, gym.ObservationWrapper = Env, Wrapper, ObservationWrapper
    gym.spaces, gym.envs, gym.logger = spaces, envs, logger
    gym.make = lambda *a, **k: None
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                        "gym.envs.registration": registration, "gym.logger": logger})


def load_reference(config_path=None, config_class=None):
    """Import (or re-import) the reference with the given config; returns the envs package."""
    if not reference_available():
        raise RuntimeError("reference not available (needs /root/reference and oracle/_ref/librvo2_ref.so)")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.cm",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "scipy.misc", "tensorflow", "tensorflow.compat",
                 "tensorflow.compat.v1", "moviepy", "moviepy.editor"):
        if name not in sys.modules or not hasattr(sys.modules[name], "__file__"):
            _stub(name)
    if "gym" not in sys.modules:
        _install_gym_stub()
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import rvo2_ref
    sys.modules["rvo2"] = rvo2_ref
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for k in [k for k in sys.modules if k.startswith("gym_collision_avoidance")
              and not k.startswith("gym_collision_avoidance_b200")]:
        del sys.modules[k]
    if config_path is not None:
        os.environ["GYM_CONFIG_PATH"] = config_path
    else:
        os.environ.pop("GYM_CONFIG_PATH", None)
    if config_class is not None:
        os.environ["GYM_CONFIG_CLASS"] = config_class
    else:
        os.environ.pop("GYM_CONFIG_CLASS", None)
    return importlib.import_module("gym_collision_avoidance.envs")
