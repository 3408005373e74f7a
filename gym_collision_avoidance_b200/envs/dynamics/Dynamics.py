from ._plugins import Dynamics  # noqa: F401  (reference module path envs/dynamics/Dynamics.py)
