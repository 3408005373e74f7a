from ._plugins import Sensor  # noqa: F401  (reference module path envs/sensors/Sensor.py)
