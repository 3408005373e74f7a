from ._plugins import LearningPolicy  # noqa: F401  (reference module path envs/policies/LearningPolicy.py)
