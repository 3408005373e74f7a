from ._plugins import ExternalPolicy  # noqa: F401  (reference module path envs/policies/ExternalPolicy.py)
