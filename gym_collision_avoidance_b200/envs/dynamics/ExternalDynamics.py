from ._plugins import ExternalDynamics  # noqa: F401  (reference module path envs/dynamics/ExternalDynamics.py)
