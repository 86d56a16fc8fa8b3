"""Entity-type enum (simulator/utils/utils.py:9-14).  The plotting helper of the reference
(plot_value_heatmap) is out of scope."""
from enum import IntEnum


class AgentType(IntEnum):
    ADULT = 0
    BICYCLE = 1
    CHILD = 2
    ADULT_STATIC = 3
    ROBOT = 4
