from ._plugins import Policy  # noqa: F401  (reference module path envs/policies/Policy.py)
