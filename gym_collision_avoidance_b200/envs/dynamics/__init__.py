"""Reference-compatible module layout: one module per class (e.g. ``envs.dynamics.X`` exposes class ``X``)."""
