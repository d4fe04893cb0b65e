"""Value-iteration agent on the device engine.

Drop-in for the reference's ValueIterationAgent
(rl_agents/agents/dynamic_programming/value_iteration.py:9-111): the constructor
solves the MDP once, `act` is an arg-max over the state's row of Q.  The fixed
point iteration itself (Bellman operator, max over actions, allclose early exit
that returns the previous iterate) runs in `b2_vi_solve`; nothing is iterated on
the host.
"""
import numpy as np

from rl_agents_b200.agents.common.abstract import AbstractAgent, register_with_reference

_NOT_AN_MDP = ("Environment must be of type finite_mdp.envs.finite_mdp.FiniteMDPEnv or handle a "
               "conversion method called 'to_finite_mdp' to such a type.")


def _third_party_finite_mdp_env():
    """The `finite_mdp` package's env class when it is installed, else None."""
    try:
        return __import__("finite_mdp.envs.finite_mdp_env").envs.finite_mdp_env.FiniteMDPEnv
    except (ModuleNotFoundError, AttributeError, TypeError):
        return None


@register_with_reference
class ValueIterationAgent(AbstractAgent):
    def __init__(self, env, config=None):
        super(ValueIterationAgent, self).__init__(config)
        self.env = env
        self.finite_mdp = self.is_finite_mdp(env)
        self.mdp = self._current_mdp()
        self.sweeps = 0                      # sweeps the last solve performed (diagnostics)
        self.state_action_value = self.get_state_action_value()

    @classmethod
    def default_config(cls):
        return {"gamma": 1.0, "iterations": 100}       # value_iteration.py:24-27

    def _scene_words(self):
        """The 136-word scene when the env is a HighwayLite scene model and the conversion may run on the device
        (config["conversion"]: "device" (default) | "host"), else None."""
        unwrapped = getattr(self.env, "unwrapped", self.env)
        if getattr(unwrapped, "b2_env_kind", None) == "highway" and self.config.get("conversion", "device") == "device":
            return unwrapped.words
        return None

    # -- MDP hand-off ---------------------------------------------------------
    @staticmethod
    def is_finite_mdp(env):
        unwrapped = getattr(env, "unwrapped", env)
        if getattr(unwrapped, "b2_env_kind", None) == "finite":
            return True
        third_party = _third_party_finite_mdp_env()
        return third_party is not None and isinstance(unwrapped, third_party)

    def _current_mdp(self):
        if self.finite_mdp:
            return self.env.unwrapped.mdp
        converter = getattr(self.env.unwrapped, "to_finite_mdp", None)
        if converter is None:
            raise TypeError(_NOT_AN_MDP)
        return converter()

    # -- solving ----------------------------------------------------------------
    def get_state_action_value(self):
        """Q of shape [S, A] (numpy, host copy of the device result)."""
        words = self._scene_words()
        if words is not None:
            # HighwayLite scene: TTC-grid conversion + fixed point in one kernel (b2_highway_ttc_vi); self.mdp is the
            # host statement of the same MDP (envs/highway_lite.py::ttc_finite_mdp), kept for plan_trajectory()
            from rl_agents_b200.engine.ttc_vi import HighwayTTCVI
            out = HighwayTTCVI(self.config["gamma"], self.config["iterations"]).solve(words.reshape(1, -1))
            self.sweeps = int(out["sweeps"][0].item())
            return out["q"][0].cpu().numpy()
        from rl_agents_b200.engine.vi import VIEngine
        mdp = self.mdp
        engine = VIEngine(mdp.mode, mdp.transition, mdp.reward, mdp.terminal, nxt=getattr(mdp, "next", None),
                          gamma=self.config["gamma"])
        q, self.sweeps = engine.solve(self.config["iterations"])
        return q.cpu().numpy()

    def get_state_value(self):
        """value_iteration.py:37-40: the reference runs a SEPARATE fixed-point iteration on V (allclose tested on V,
        which can stop at an earlier iterate than the iteration on Q): V' = max_a B(V) with the device sweep, the
        allclose test of fixed_point_iteration (:65-73) on the host, returning the previous iterate."""
        from rl_agents_b200.engine.vi import VIEngine
        mdp = self.mdp
        eng = VIEngine(mdp.mode, mdp.transition, mdp.reward, mdp.terminal, nxt=getattr(mdp, "next", None),
                       gamma=self.config["gamma"], rtol=0.0, atol=-1.0)      # the kernel's own Q test never fires
        iterations = int(self.config["iterations"])
        eng.reset(iterations)
        v = np.zeros(eng.n_states)
        for k in range(iterations):
            eng.sweep(k)
            v_next = eng.v[(k + 1) & 1].cpu().numpy()
            if np.allclose(v, v_next):
                break
            v = v_next
        return v

    @staticmethod
    def best_action_value(action_values):
        return action_values.max(axis=-1)

    # -- acting -----------------------------------------------------------------
    def act(self, state):
        if not self.finite_mdp:
            # envs that are only convertible are re-converted and re-solved at every decision (:31-34)
            self.mdp = self._current_mdp()
            self.state_action_value = self.get_state_action_value()
            state = self.mdp.state
        return np.argmax(self.state_action_value[state, :])

    def plan_trajectory(self, state, horizon=10):
        """Greedy roll-out of the solved policy through the (deterministic) transition table."""
        states, actions = [], []
        for _ in range(horizon):
            states.append(state)
            actions.append(np.argmax(self.state_action_value[state]))
            state = self.mdp.next_state(state, actions[-1])
            if self.mdp.terminal[state]:
                return states + [state], actions + [None]
        return states, actions

    # -- nothing to learn, nothing to checkpoint ----------------------------------
    def record(self, state, action, reward, next_state, done, info):
        return None

    def reset(self):
        return None

    def seed(self, seed=None):
        return None

    def save(self, filename):
        return False

    def load(self, filename):
        return False
