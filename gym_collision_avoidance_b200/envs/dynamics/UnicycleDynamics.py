from ._plugins import UnicycleDynamics  # noqa: F401  (reference module path envs/dynamics/UnicycleDynamics.py)
