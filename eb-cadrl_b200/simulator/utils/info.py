"""Event classes returned by env.step / onestep_lookahead (simulator/utils/info.py:1-152).
The device reports an event code (ebc.abi.EVENT_NAMES); `from_code` rebuilds the object."""


class Info(object):
    NAME = ""

    def __init__(self, dist_to_goal=None, dmin_adult=None, dmin_bicycle=None, dmin_child=None):
        self.dist_to_goal = dist_to_goal
        self.dmin_adult = dmin_adult
        self.dmin_bicycle = dmin_bicycle
        self.dmin_child = dmin_child

    def __str__(self):
        return self.NAME


class Timeout(Info):
    NAME = "Timeout"


class ReachGoal(Info):
    NAME = "Reaching goal"


class Danger(Info):
    NAME = "Too close"

    def __init__(self, min_dist, dist_to_goal=None, dmin_adult=None, dmin_bicycle=None, dmin_child=None):
        super().__init__(dist_to_goal, dmin_adult, dmin_bicycle, dmin_child)
        self.min_dist = min_dist


class CollisionAdult(Info):
    NAME = "CollisionAdult"


class CollisionBicycle(Info):
    NAME = "CollisionBicycle"


class CollisionObstacle(Info):
    NAME = "CollisionObstacle"


class CollisionChild(Info):
    NAME = "CollisionChild"


class Collision(Info):
    NAME = "Collision"


class CollisionOtherAgent(Info):
    NAME = "Collision from other agent"


class Nothing(Info):
    NAME = ""

    def __init__(self, dmin_adult=None, dmin_bicycle=None, dmin_child=None):
        super().__init__(None, dmin_adult, dmin_bicycle, dmin_child)


_BY_CODE = [Nothing, Danger, ReachGoal, CollisionAdult, CollisionBicycle, CollisionChild, CollisionObstacle, Timeout]


def from_code(code, dist_to_goal, dmin, discomfort=None):
    """code: device event (include/ebcadrl.h EBC_EV_*); dmin: (adult, bicycle, child)."""
    cls = _BY_CODE[int(code)]
    da, db, dc = (float(x) for x in dmin)
    if cls is Nothing:
        return Nothing(da, db, dc)
    if cls is Danger:
        # reward.py:138-167: the first type (child, bicycle, adult) below its discomfort distance
        dd = discomfort or (float("inf"),) * 3
        min_dist = dc if dc < dd[2] else (db if db < dd[1] else da)
        return Danger(min_dist, float(dist_to_goal), da, db, dc)
    return cls(float(dist_to_goal), da, db, dc)
