"""Value types of the hot path — same names, fields and tuple order as the reference
(simulator/utils/state.py:1-92): a joint row is `full + observable` =
(px, py, vx, vy, radius, gx, gy, v_pref, theta | px1, py1, vx1, vy1, radius1, type1)."""


class FullState(object):
    FIELDS = ("px", "py", "vx", "vy", "radius", "gx", "gy", "v_pref", "theta")

    def __init__(self, px, py, vx, vy, radius, gx, gy, v_pref, theta, obj_type=None):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.gx, self.gy, self.v_pref, self.theta = gx, gy, v_pref, theta
        self.position = (px, py)
        self.goal_position = (gx, gy)
        self.velocity = (vx, vy)
        self.obj_type = obj_type

    def _tuple(self):
        return tuple(getattr(self, f) for f in self.FIELDS)

    def __add__(self, other):
        return other + self._tuple()

    def __str__(self):
        return " ".join(str(x) for x in self._tuple() + (self.obj_type,))


class ObservableState(object):
    FIELDS = ("px", "py", "vx", "vy", "radius", "obj_type")

    def __init__(self, px, py, vx, vy, radius, obj_type=None):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.position = (px, py)
        self.velocity = (vx, vy)
        self.obj_type = obj_type

    def _tuple(self):
        return tuple(getattr(self, f) for f in self.FIELDS)

    def __add__(self, other):
        return other + self._tuple()

    def __str__(self):
        return " ".join(str(x) for x in self._tuple())


class JointState(object):
    def __init__(self, self_state, agent_states):
        assert isinstance(self_state, FullState)
        for agent_state in agent_states:
            assert isinstance(agent_state, ObservableState)
        self.self_state = self_state
        self.agent_states = agent_states
