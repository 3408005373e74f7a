"""Reference-compatible module layout: one module per class (e.g. ``envs.sensors.X`` exposes class ``X``)."""
