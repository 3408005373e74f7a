from ._plugins import NonCooperativePolicy  # noqa: F401  (reference module path envs/policies/NonCooperativePolicy.py)
