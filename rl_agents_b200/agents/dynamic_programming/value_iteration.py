"""Value-iteration agent on the device engine.  Drop-in for
rl_agents.agents.dynamic_programming.value_iteration.ValueIterationAgent
(value_iteration.py:9-111)."""
import numpy as np

from rl_agents_b200.agents.common.abstract import AbstractAgent, register_with_reference


@register_with_reference
class ValueIterationAgent(AbstractAgent):
    def __init__(self, env, config=None):
        super(ValueIterationAgent, self).__init__(config)
        self.finite_mdp = self.is_finite_mdp(env)
        if self.finite_mdp:
            self.mdp = env.unwrapped.mdp
        else:
            try:
                self.mdp = env.unwrapped.to_finite_mdp()
            except AttributeError:
                raise TypeError("Environment must be of type finite_mdp.envs.finite_mdp.FiniteMDPEnv or handle a "
                                "conversion method called 'to_finite_mdp' to such a type.")
        self.env = env
        self.sweeps = 0
        self.state_action_value = self.get_state_action_value()

    @classmethod
    def default_config(cls):
        return dict(gamma=1.0, iterations=100)

    def act(self, state):
        if not self.finite_mdp:                       # value_iteration.py:31-34: re-solve on every act
            self.mdp = self.env.unwrapped.to_finite_mdp()
            state = self.mdp.state
            self.state_action_value = self.get_state_action_value()
        return np.argmax(self.state_action_value[state, :])

    def get_state_action_value(self):
        """fixed_point_iteration on Q (value_iteration.py:42-45,65-73) by b2_vi_solve."""
        from rl_agents_b200.engine.vi import VIEngine
        m = self.mdp
        eng = VIEngine(m.mode, m.transition, m.reward, m.terminal, nxt=getattr(m, "next", None),
                       gamma=self.config["gamma"])
        q, self.sweeps = eng.solve(self.config["iterations"])
        return q.cpu().numpy()

    def get_state_value(self):
        return self.state_action_value.max(axis=-1)

    @staticmethod
    def best_action_value(action_values):
        return action_values.max(axis=-1)

    @staticmethod
    def is_finite_mdp(env):
        u = getattr(env, "unwrapped", env)
        if getattr(u, "b2_env_kind", None) == "finite":
            return True
        try:
            finite_mdp = __import__("finite_mdp.envs.finite_mdp_env")
            return isinstance(u, finite_mdp.envs.finite_mdp_env.FiniteMDPEnv)
        except (ModuleNotFoundError, TypeError):
            return False

    def plan_trajectory(self, state, horizon=10):
        action_value = self.state_action_value
        states, actions = [], []
        for _ in range(horizon):
            action = np.argmax(action_value[state])
            states.append(state)
            actions.append(action)
            state = self.mdp.next_state(state, action)
            if self.mdp.terminal[state]:
                states.append(state)
                actions.append(None)
                break
        return states, actions

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
