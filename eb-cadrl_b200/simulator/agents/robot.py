"""Robot (simulator/agents/robot.py:7-25)."""
from simulator.agents.agent import Agent
from simulator.utils.state import JointState
from simulator.utils.utils import AgentType


class Robot(Agent):
    def __init__(self, config, section):
        super().__init__(config, section)
        self.action_index = None
        self.attention_weights = None
        self.last_state = None
        self.adults_in_FOV = None
        self.agent_type = AgentType.ROBOT

    def act(self, ob, local_map=None, env=None):
        if self.policy is None:
            raise AttributeError("Policy attribute has to be set!")
        return self.policy.predict(JointState(self.get_full_state(), ob), env)
