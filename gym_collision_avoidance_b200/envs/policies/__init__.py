"""Reference-compatible module layout: one module per class (e.g. ``envs.policies.X`` exposes class ``X``)."""
