from ._plugins import LaserScanSensor  # noqa: F401  (reference module path envs/sensors/LaserScanSensor.py)
