"""B200-native batched collision-avoidance step path.

Registers ``CollisionAvoidance-v0`` like the reference package does (reference ``gym_collision_avoidance/__init__.py:6-9``)
when ``gym`` or ``gymnasium`` is installed; neither is required."""


def _register():
    entry = "gym_collision_avoidance_b200.envs.collision_avoidance_env:CollisionAvoidanceEnv"
    for mod in ("gym", "gymnasium"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"])
            reg.register(id="CollisionAvoidance-v0", entry_point=entry)
            return mod
        except Exception:
            continue
    return None


REGISTERED_WITH = _register()
