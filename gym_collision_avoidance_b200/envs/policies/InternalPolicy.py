from ._plugins import InternalPolicy  # noqa: F401  (reference module path envs/policies/InternalPolicy.py)
