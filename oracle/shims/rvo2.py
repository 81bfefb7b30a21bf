"""TEST INFRASTRUCTURE -- stand-in for the Python-RVO2 Cython module (`import rvo2` at reference
scripts/policy/orca.py:2 and crowd_sim.py:6). Exposes exactly the PyRVOSimulator surface used at
orca.py:80-114 and forwards to the C restatement in oracle/rvo2_ref.c (fp32; Python floats are
rounded to float at this boundary like Cython does)."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_here, "..", "liboracle.so"))
_f, _i, _p = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
_lib.rvo_sim_create.restype = _p
_lib.rvo_sim_create.argtypes = [_f, _f, _i, _f, _f, _f, _f, _f, _f]
_lib.rvo_sim_destroy.argtypes = [_p]
_lib.rvo_sim_add_agent.restype = _i
_lib.rvo_sim_add_agent.argtypes = [_p, _f, _f, _f, _i, _f, _f, _f, _f, _f, _f]
_lib.rvo_sim_num_agents.restype = _i
_lib.rvo_sim_num_agents.argtypes = [_p]
for _n in ("rvo_sim_set_position", "rvo_sim_set_velocity", "rvo_sim_set_pref_velocity"):
    getattr(_lib, _n).argtypes = [_p, _i, _f, _f]
_lib.rvo_sim_get_velocity.argtypes = [_p, _i, ctypes.POINTER(_f)]
_lib.rvo_sim_get_position.argtypes = [_p, _i, ctypes.POINTER(_f)]
_lib.rvo_sim_do_step.argtypes = [_p]


class PyRVOSimulator:
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius,
                 maxSpeed, velocity=(0, 0)):
        self._defaults = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
        self._h = _lib.rvo_sim_create(timeStep, neighborDist, int(maxNeighbors), timeHorizon, timeHorizonObst,
                                      radius, maxSpeed, velocity[0], velocity[1])

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rvo_sim_destroy(self._h)
            self._h = None

    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        d = self._defaults
        nd = d[0] if neighborDist is None else neighborDist
        mn = d[1] if maxNeighbors is None else maxNeighbors
        th = d[2] if timeHorizon is None else timeHorizon
        tho = d[3] if timeHorizonObst is None else timeHorizonObst
        r = d[4] if radius is None else radius
        ms = d[5] if maxSpeed is None else maxSpeed
        v = d[6] if velocity is None else velocity
        return _lib.rvo_sim_add_agent(self._h, pos[0], pos[1], nd, int(mn), th, tho, r, ms, v[0], v[1])

    def getNumAgents(self):
        return _lib.rvo_sim_num_agents(self._h)

    def setAgentPosition(self, i, pos):
        _lib.rvo_sim_set_position(self._h, i, pos[0], pos[1])

    def setAgentVelocity(self, i, vel):
        _lib.rvo_sim_set_velocity(self._h, i, vel[0], vel[1])

    def setAgentPrefVelocity(self, i, vel):
        _lib.rvo_sim_set_pref_velocity(self._h, i, vel[0], vel[1])

    def doStep(self):
        _lib.rvo_sim_do_step(self._h)

    def getAgentVelocity(self, i):
        out = (_f * 2)()
        _lib.rvo_sim_get_velocity(self._h, i, out)
        return (float(out[0]), float(out[1]))

    def getAgentPosition(self, i):
        out = (_f * 2)()
        _lib.rvo_sim_get_position(self._h, i, out)
        return (float(out[0]), float(out[1]))
