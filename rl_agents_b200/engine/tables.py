"""Host-computed tables the kernels read so that their floating point matches
the reference's Python float arithmetic bit for bit."""
import numpy as np


def gamma_tables(gamma, n):
    """gamma**d and gamma**d/(1-gamma) with Python float `**` (deterministic.py:52-53)."""
    gamma = float(gamma)
    gp = np.array([gamma ** d for d in range(n)], dtype=np.float64)
    with np.errstate(divide="ignore"):
        gd = np.array([(gamma ** d) / (1 - gamma) if gamma != 1 else np.inf for d in range(n)], dtype=np.float64)
    return gp, gd


def terminal_bonus_table(terminal_reward, gamma, n):
    """(terminal_reward * gamma**d) / (1 - gamma), the reference's association (deterministic.py:60-63)."""
    gamma, tr = float(gamma), float(terminal_reward)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.array([(tr * gamma ** d) / (1 - gamma) if gamma != 1 else np.inf * tr for d in range(n)],
                        dtype=np.float64)


def uniform_cdf_table(n_actions):
    """Row n: cumsum(ones(n)/n)/cumsum[-1] -- the cdf Generator.choice(a, 1, p=p)
    searches for a uniform p over n actions (mcts.py:60-72,172)."""
    t = np.ones((n_actions + 1, n_actions), dtype=np.float64)
    for n in range(1, n_actions + 1):
        cdf = (np.ones(n) / n).cumsum()
        cdf /= cdf[-1]
        t[n, :n] = cdf
    return t


def preference_tables(n_actions, ratio):
    """MCTSAgent.preference_policy (mcts.py:76-97) for every (number of available actions n, position k-1 of
    the preferred action among them; k = 0: not available -> uniform): the probabilities exactly as numpy
    computes them there, and the cdf Generator.choice(actions, 1, p=p) searches.  Shape [(A+1), (A+1), A]."""
    A = int(n_actions)
    prior = np.ones((A + 1, A + 1, A), dtype=np.float64)
    cdf = np.ones((A + 1, A + 1, A), dtype=np.float64)
    for n in range(1, A + 1):
        for k in range(0, n + 1):
            if k == 0:
                p = np.ones(n) / n
            else:
                p = np.ones(n) / (n - 1 + ratio)
                p[k - 1] *= ratio
            c = p.cumsum()
            c /= c[-1]
            prior[n, k, :n] = p
            cdf[n, k, :n] = c
    return prior, cdf


class FiniteTables(object):
    """Device copy of a deterministic finite MDP (int32 transitions)."""

    def __init__(self, mdp, device):
        import torch
        if mdp.mode != "deterministic":
            raise ValueError("tree search on a finite MDP needs mode == 'deterministic' (got %r)" % mdp.mode)
        self.n_states, self.n_actions = mdp.reward.shape
        self.transition = torch.as_tensor(np.ascontiguousarray(mdp.transition, dtype=np.int32), device=device)
        self.reward = torch.as_tensor(np.ascontiguousarray(mdp.reward, dtype=np.float64), device=device)
        self.terminal = torch.as_tensor(np.ascontiguousarray(mdp.terminal, dtype=np.uint8), device=device)

    def struct(self):
        from rl_agents_b200 import _lib
        return _lib.FiniteMDP(self.n_states, self.n_actions, self.transition.data_ptr(),
                              self.reward.data_ptr(), self.terminal.data_ptr())
