"""Robust value-iteration agent on the device engine.  Drop-in for
rl_agents.agents.dynamic_programming.robust_value_iteration.RobustValueIterationAgent
(robust_value_iteration.py:6-73): the MDP models come from config["models"]."""
import numpy as np

from rl_agents_b200.agents.common.abstract import AbstractAgent, register_with_reference


@register_with_reference
class RobustValueIterationAgent(AbstractAgent):
    def __init__(self, env, config=None):
        super(RobustValueIterationAgent, self).__init__(config)
        self.env = env
        if not self.config.get("models", None):
            raise ValueError("No finite MDP model provided in agent configuration")
        self.mode = self.config["models"][0]["mode"]          # all modes are assumed equal (:25)
        self.transitions = np.array([mdp["transition"] for mdp in self.config["models"]])
        self.rewards = np.array([mdp["reward"] for mdp in self.config["models"]])
        self.sweeps = 0
        self._q = None

    @classmethod
    def default_config(cls):
        return dict(gamma=1.0, iterations=100, models=[])

    def get_state_action_value(self):
        from rl_agents_b200.engine.vi import RobustVIEngine
        eng = RobustVIEngine(self.mode, self.transitions, self.rewards, gamma=self.config["gamma"])
        q, self.sweeps = eng.solve(self.config["iterations"])
        self._q = q.cpu().numpy()
        return self._q

    def get_state_value(self):
        return self.get_state_action_value().max(axis=-1)

    def act(self, state):
        return np.argmax(self.get_state_action_value()[state, :])    # re-solved on every act, like the reference (:29-30)

    def record(self, state, action, reward, next_state, done, info):
        pass

    def reset(self):
        pass

    def seed(self, seed=None):
        pass

    def save(self, filename):
        return False

    def load(self, filename):
        return False
