from ._plugins import StaticPolicy  # noqa: F401  (reference module path envs/policies/StaticPolicy.py)
