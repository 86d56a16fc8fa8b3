"""`rvo2.PyRVOSimulator` stand-in backed by the C oracle (oracle/ebc_oracle.c).

Test infrastructure only: lets the UNMODIFIED reference (simulator/policy/orca.py:105-154)
run in this image, where Python-RVO2 is neither installed nor installable.  Doubles are
narrowed to float on entry exactly as Cython does for `float` arguments."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_here, "..", "_ref", "libebc_oracle.so"))
_f = ctypes.c_float
_lib.ebc_rvo_create.restype = ctypes.c_void_p
_lib.ebc_rvo_create.argtypes = [_f]
_lib.ebc_rvo_destroy.argtypes = [ctypes.c_void_p]
_lib.ebc_rvo_add_agent.argtypes = [ctypes.c_void_p, _f, _f, _f, ctypes.c_int, _f, _f, _f, _f, _f]
_lib.ebc_rvo_add_agent.restype = ctypes.c_int
_lib.ebc_rvo_num_agents.argtypes = [ctypes.c_void_p]
_lib.ebc_rvo_num_agents.restype = ctypes.c_int
for _n in ("ebc_rvo_set_pos", "ebc_rvo_set_vel", "ebc_rvo_set_pref"):
    getattr(_lib, _n).argtypes = [ctypes.c_void_p, ctypes.c_int, _f, _f]
for _n in ("ebc_rvo_get_vel", "ebc_rvo_get_pos"):
    getattr(_lib, _n).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_f), ctypes.POINTER(_f)]
_lib.ebc_rvo_do_step.argtypes = [ctypes.c_void_p]
_lib.ebc_rvo_set_time_horizon_obst.argtypes = [ctypes.c_void_p, ctypes.c_int, _f]
_lib.ebc_rvo_add_obstacle.argtypes = [ctypes.c_void_p, ctypes.POINTER(_f), ctypes.c_int]
_lib.ebc_rvo_add_obstacle.restype = ctypes.c_int
_lib.ebc_rvo_process_obstacles.argtypes = [ctypes.c_void_p]
_lib.ebc_rvo_process_obstacles.restype = ctypes.c_int


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst,
                 radius, maxSpeed, velocity=(0, 0)):
        self._h = _lib.ebc_rvo_create(timeStep)
        self._defaults = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.ebc_rvo_destroy(self._h)
            self._h = None

    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None,
                 timeHorizonObst=None, radius=None, maxSpeed=None, velocity=(0, 0)):
        d = self._defaults
        neighborDist = d[0] if neighborDist is None else neighborDist
        maxNeighbors = d[1] if maxNeighbors is None else maxNeighbors
        timeHorizon = d[2] if timeHorizon is None else timeHorizon
        radius = d[4] if radius is None else radius
        maxSpeed = d[5] if maxSpeed is None else maxSpeed
        i = _lib.ebc_rvo_add_agent(self._h, pos[0], pos[1], neighborDist, int(maxNeighbors),
                                   timeHorizon, radius, maxSpeed, velocity[0], velocity[1])
        _lib.ebc_rvo_set_time_horizon_obst(self._h, i, d[3] if timeHorizonObst is None else timeHorizonObst)
        return i

    def getNumAgents(self):
        return _lib.ebc_rvo_num_agents(self._h)

    def setAgentPosition(self, i, pos):
        _lib.ebc_rvo_set_pos(self._h, i, pos[0], pos[1])

    def setAgentVelocity(self, i, vel):
        _lib.ebc_rvo_set_vel(self._h, i, vel[0], vel[1])

    def setAgentPrefVelocity(self, i, vel):
        _lib.ebc_rvo_set_pref(self._h, i, vel[0], vel[1])

    def doStep(self):
        _lib.ebc_rvo_do_step(self._h)

    def getAgentVelocity(self, i):
        x, y = _f(), _f()
        _lib.ebc_rvo_get_vel(self._h, i, ctypes.byref(x), ctypes.byref(y))
        return (x.value, y.value)

    def getAgentPosition(self, i):
        x, y = _f(), _f()
        _lib.ebc_rvo_get_pos(self._h, i, ctypes.byref(x), ctypes.byref(y))
        return (x.value, y.value)

    def addObstacle(self, vertices):
        """RVOSimulator::addObstacle (simulator/policy/orca_obstacles.py:105-106): counter-clockwise polygon."""
        flat = (_f * (2 * len(vertices)))(*[c for v in vertices for c in v[:2]])
        r = _lib.ebc_rvo_add_obstacle(self._h, flat, len(vertices))
        if r < 0:
            raise RuntimeError("addObstacle: more than 64 obstacle vertices")
        return r

    def processObstacles(self):
        if _lib.ebc_rvo_process_obstacles(self._h) != 0:
            raise RuntimeError("processObstacles: the kd-tree needs more than 64 obstacle vertices")
