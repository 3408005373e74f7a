from ._plugins import OtherAgentsStatesSensor  # noqa: F401  (reference module path envs/sensors/OtherAgentsStatesSensor.py)
