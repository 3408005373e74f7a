from ._plugins import RVOPolicy  # noqa: F401  (reference module path envs/policies/RVOPolicy.py)
