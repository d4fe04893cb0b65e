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


def garnet_slab(n_states, n_actions, n_next, row_begin, row_end, seed=0, device="cuda", reward_sparsity=0.5):
    """Rows [row_begin, row_end) of a garnet-style sparse MDP, generated ON the device, slab by
    slab (rows are seeded independently, so any partition of the rows yields the same MDP): successors uniform
    over the states, P = normalised U(0,1), R ~ U(0,1) with `reward_sparsity` of the entries zeroed, no
    terminal states.  Returns (P f64 [rows,A,B], N i32 [rows,A,B], R f64 [rows,A], terminal u8 [rows])."""
    import torch
    dev = torch.device(device)
    rows = int(row_end) - int(row_begin)
    chunk = 65536                                    # rows per seeded block: slabs need not align with blocks
    P = torch.empty((rows, n_actions, n_next), dtype=torch.float64, device=dev)
    N = torch.empty((rows, n_actions, n_next), dtype=torch.int32, device=dev)
    R = torch.empty((rows, n_actions), dtype=torch.float64, device=dev)
    first = (int(row_begin) // chunk) * chunk
    for b0 in range(first, int(row_end), chunk):
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed) * 1000003 + b0 // chunk)
        n = min(chunk, int(n_states) - b0)
        p = torch.rand((n, n_actions, n_next), dtype=torch.float64, device=dev, generator=g)
        nx = torch.randint(0, int(n_states), (n, n_actions, n_next), dtype=torch.int32, device=dev, generator=g)
        r = torch.rand((n, n_actions), dtype=torch.float64, device=dev, generator=g)
        keep = torch.rand((n, n_actions), dtype=torch.float64, device=dev, generator=g) >= reward_sparsity
        lo, hi = max(b0, int(row_begin)), min(b0 + n, int(row_end))
        if lo >= hi:
            continue
        src, dst = slice(lo - b0, hi - b0), slice(lo - int(row_begin), hi - int(row_begin))
        P[dst] = p[src] / p[src].sum(dim=-1, keepdim=True)
        N[dst] = nx[src]
        R[dst] = r[src] * keep[src]
    return P, N, R, torch.zeros(rows, dtype=torch.uint8, device=dev)
