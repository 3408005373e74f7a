"""``gym_collision_avoidance_b200.envs``: the reference's ``envs`` package surface on top of the B200 library.

Like the reference (``envs/__init__.py:4-18``) a ``Config`` singleton is instantiated at import time from the class
named by the environment variables ``GYM_CONFIG_CLASS`` / ``GYM_CONFIG_PATH`` (default: :class:`config.Config`).
"""
import importlib.util
import os

from . import config as _config_module

_path = os.environ.get("GYM_CONFIG_PATH")
_cls = os.environ.get("GYM_CONFIG_CLASS", "Config")
if _path:
    _spec = importlib.util.spec_from_file_location(_cls, _path)
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    _config_class = getattr(_mod, _cls, None)
else:
    _config_class = getattr(_config_module, _cls, None)
assert callable(_config_class), "GYM_CONFIG_CLASS=%r not found" % (_cls,)
Config = _config_class()


def set_config(cfg):
    """Replace the module-level Config singleton (tests / notebooks; the reference needs a re-import for this)."""
    global Config
    Config = cfg
    return cfg
