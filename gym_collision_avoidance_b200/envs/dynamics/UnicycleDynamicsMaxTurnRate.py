from ._plugins import UnicycleDynamicsMaxTurnRate  # noqa: F401  (reference module path envs/dynamics/UnicycleDynamicsMaxTurnRate.py)
