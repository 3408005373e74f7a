"""Reference module path ``envs/policies/GA3CCADRLPolicy.py``."""
from ._plugins import GA3CCADRLPolicy  # noqa: F401
