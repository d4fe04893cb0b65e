"""Finite-MDP env with the attribute surface the reference reads from the
third-party `finite_mdp` package (value_iteration.py:52-63,91-92:
env.mdp.{mode,transition,reward,terminal,next,state}) and a 5-tuple step.
Host-side table container: the planners upload the tables once and run the
transitions on the device."""
import numpy as np


class _Space(object):
    def __init__(self, n):
        self.n = int(n)


class FiniteMDP(object):
    def __init__(self, mode, transition, reward, terminal=None, nxt=None, state=0):
        reward = np.asarray(reward, dtype=np.float64)
        if terminal is None:
            terminal = np.zeros(reward.shape[0], dtype=bool)
        self.mode = mode
        self.transition = np.asarray(transition, dtype=np.int64 if mode == "deterministic" else np.float64)
        self.reward = reward
        self.terminal = np.asarray(terminal).astype(bool)
        self.next = None if nxt is None else np.asarray(nxt, dtype=np.int64)
        self.state = int(state)

    def next_state(self, state, action):
        return int(self.transition[state, action])


class FiniteMDPEnv(object):
    b2_env_kind = "finite"

    def __init__(self, transition, reward, terminal=None, mode="deterministic", nxt=None, state=0, seed=None):
        self.mdp = FiniteMDP(mode, transition, reward, terminal, nxt, state)
        self.action_space = _Space(self.mdp.reward.shape[1])
        self.np_random = np.random.default_rng(seed)
        self._initial_state = int(state)

    @classmethod
    def from_config(cls, config):
        """A `finite-mdp-v0` JSON (scripts/configs/FiniteMDPEnv/*.json)."""
        return cls(config["transition"], config["reward"], config.get("terminal"),
                   mode=config.get("mode", "deterministic"), nxt=config.get("next"))

    @property
    def unwrapped(self):
        return self

    def to_finite_mdp(self):
        return self.mdp

    def seed(self, seed=None):
        self.np_random = np.random.default_rng(seed)
        return [seed]

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.seed(seed)
        self.mdp.state = self._initial_state
        return self.mdp.state, {}

    def step(self, action):
        m = self.mdp
        s = m.state
        r = float(m.reward[s, action])
        if m.mode == "deterministic":
            s2 = int(m.transition[s, action])
        else:
            p = m.transition[s, action]
            k = int(self.np_random.choice(p.size, p=p))
            s2 = k if m.mode == "stochastic" else int(m.next[s, action, k])
        m.state = s2
        # the `finite_mdp` package's MDP.step evaluates done on the state the action was taken IN
        # (consistent with value_iteration.py:62, which zeroes the next-value of terminal SOURCE states)
        return s2, r, bool(m.terminal[s]), False, {}
