"""Human entity classes (simulator/agents/agents.py:8-105).  `BicycleRectangle` is unreachable in the
reference (no config sets bicycle_type = rectangle) and is not provided."""
from simulator.agents.agent import Agent
from simulator.utils.state import JointState
from simulator.utils.utils import AgentType


class _Human(Agent):
    TYPE = None

    def __init__(self, config, section):
        super().__init__(config, section)
        self.agent_type = self.TYPE

    def act(self, ob=None, global_map=None, local_map=None):
        """Full state + the other agents' observable states -> policy (agents.py:13-30)."""
        if ob is None:
            return self.policy.predict(self)
        return self.policy.predict(JointState(self.get_full_state(), ob))


class Adult(_Human):
    TYPE = AgentType.ADULT


class Bicycle(_Human):
    TYPE = AgentType.BICYCLE


class Child(_Human):
    TYPE = AgentType.CHILD
