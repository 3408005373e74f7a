"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

ctypes stand-in for the reference's Cython module ``rvo2`` (reference
``Python-RVO2/src/rvo2.pyx:80-230``) on top of ``oracle/_ref/librvo2_ref.so``
(the unmodified reference RVO2 C++ built by ``oracle/Makefile``).  Only the
methods ``RVOPolicy.py:25-96`` calls are provided.  Arguments cross as C
``float`` exactly like the Cython binding (``rvo2.pyx:212-228``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "librvo2_ref.so")
        if not os.path.exists(path):
            raise ImportError("oracle/_ref/librvo2_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(path)
        f, sz, vp = ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
        fp = ctypes.POINTER(ctypes.c_float)
        L.rvo2ref_new.restype = vp
        L.rvo2ref_new.argtypes = [f, f, sz, f, f, f, f, f, f]
        L.rvo2ref_delete.argtypes = [vp]
        L.rvo2ref_add_agent.restype = sz
        L.rvo2ref_add_agent.argtypes = [vp, f, f]
        L.rvo2ref_do_step.argtypes = [vp]
        for name in ("set_position", "set_velocity", "set_pref_velocity"):
            getattr(L, "rvo2ref_" + name).argtypes = [vp, sz, f, f]
        for name in ("set_radius", "set_max_speed", "set_collab_coeff", "set_neighbor_dist", "set_time_horizon"):
            getattr(L, "rvo2ref_" + name).argtypes = [vp, sz, f]
        L.rvo2ref_set_max_neighbors.argtypes = [vp, sz, sz]
        L.rvo2ref_set_time_step.argtypes = [vp, f]
        for name in ("get_position", "get_velocity", "get_pref_velocity"):
            getattr(L, "rvo2ref_" + name).argtypes = [vp, sz, fp]
        L.rvo2ref_num_agents.restype = sz
        L.rvo2ref_num_agents.argtypes = [vp]
        L.rvo2ref_num_orca_lines.restype = sz
        L.rvo2ref_num_orca_lines.argtypes = [vp, sz]
        L.rvo2ref_get_orca_line.argtypes = [vp, sz, sz, fp]
        L.rvo2ref_num_agent_neighbors.restype = sz
        L.rvo2ref_num_agent_neighbors.argtypes = [vp, sz]
        L.rvo2ref_agent_neighbor.restype = sz
        L.rvo2ref_agent_neighbor.argtypes = [vp, sz, sz]
        L.rvo2ref_solve_ego.argtypes = [vp, sz, sz, fp, fp, fp, fp, fp, f, fp, fp]
        _LIB = L
    return _LIB


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0, 0)):
        self._L = lib()
        self._p = self._L.rvo2ref_new(timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius,
                                      maxSpeed, velocity[0], velocity[1])
        self._buf = (ctypes.c_float * 4)()

    def addAgent(self, pos):
        return self._L.rvo2ref_add_agent(self._p, pos[0], pos[1])

    def doStep(self):
        self._L.rvo2ref_do_step(self._p)

    def setAgentPosition(self, i, p):
        self._L.rvo2ref_set_position(self._p, i, p[0], p[1])

    def setAgentVelocity(self, i, p):
        self._L.rvo2ref_set_velocity(self._p, i, p[0], p[1])

    def setAgentPrefVelocity(self, i, p):
        self._L.rvo2ref_set_pref_velocity(self._p, i, p[0], p[1])

    def setAgentRadius(self, i, r):
        self._L.rvo2ref_set_radius(self._p, i, r)

    def setAgentMaxSpeed(self, i, v):
        self._L.rvo2ref_set_max_speed(self._p, i, v)

    def setAgentCollabCoeff(self, i, c):
        self._L.rvo2ref_set_collab_coeff(self._p, i, c)

    def setAgentNeighborDist(self, i, d):
        self._L.rvo2ref_set_neighbor_dist(self._p, i, d)

    def setAgentMaxNeighbors(self, i, n):
        self._L.rvo2ref_set_max_neighbors(self._p, i, n)

    def setTimeStep(self, t):
        self._L.rvo2ref_set_time_step(self._p, t)

    def _get2(self, fn, i):
        fn(self._p, i, self._buf)
        return (self._buf[0], self._buf[1])

    def getAgentPosition(self, i):
        return self._get2(self._L.rvo2ref_get_position, i)

    def getAgentVelocity(self, i):
        return self._get2(self._L.rvo2ref_get_velocity, i)

    def getAgentPrefVelocity(self, i):
        return self._get2(self._L.rvo2ref_get_pref_velocity, i)

    def getNumAgents(self):
        return self._L.rvo2ref_num_agents(self._p)

    def getAgentNumORCALines(self, i):
        return self._L.rvo2ref_num_orca_lines(self._p, i)

    def getAgentORCALine(self, i, l):
        self._L.rvo2ref_get_orca_line(self._p, i, l, self._buf)
        return tuple(self._buf)

    def getAgentNumAgentNeighbors(self, i):
        return self._L.rvo2ref_num_agent_neighbors(self._p, i)

    def getAgentAgentNeighbor(self, i, k):
        return self._L.rvo2ref_agent_neighbor(self._p, i, k)
