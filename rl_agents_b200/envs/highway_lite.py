"""HighwayLite env (docs/HIGHWAY_LITE_SPEC.md): host-side container of one
scene (136 32-bit words, the layout the kernels read) with a gymnasium-style
API.  `step` runs the CUDA transition (b2_highway_step) -- the same device code
the planners expand nodes with; there is no CPU implementation in the product.
"""
import copy

import numpy as np

from rl_agents_b200 import _lib

V_SLOTS = 16
N_LANES = 4
N_ACTIONS = 5
ACTIONS = {0: "LANE_LEFT", 1: "IDLE", 2: "LANE_RIGHT", 3: "FASTER", 4: "SLOWER"}
_ORDER = (1, 0, 2, 3, 4)   # IDLE, LEFT, RIGHT, FASTER, SLOWER: get_available_actions order


class _Space(object):
    def __init__(self, n):
        self.n = int(n)


def make_scene(seed, n_vehicles=V_SLOTS):
    """Synthetic highway-v0-like scene: per-lane cumulative gaps U(40,80) m,
    speeds U(21,24), ego (slot 0, 25 m/s) = the vehicle nearest x = 0."""
    rng = np.random.default_rng(seed)
    n = int(n_vehicles)
    lanes = rng.integers(0, N_LANES, size=n)
    next_x = -160.0 + rng.uniform(0.0, 40.0, size=N_LANES)
    xs = np.zeros(n)
    for k in range(n):
        xs[k] = next_x[lanes[k]] + rng.uniform(40.0, 80.0)
        next_x[lanes[k]] = xs[k]
    speeds = rng.uniform(21.0, 24.0, size=n)
    timers = rng.uniform(0.0, 1.0, size=n)
    ego = int(np.argmin(np.abs(xs)))
    order = [ego] + [k for k in range(n) if k != ego]
    w = np.zeros(_lib.HW_STATE_WORDS, dtype=np.int32)
    f = w[:96].view(np.float32)
    for slot, k in enumerate(order):
        f[0 * 16 + slot] = xs[k]
        f[1 * 16 + slot] = 4.0 * lanes[k]
        f[3 * 16 + slot] = speeds[k]
        f[4 * 16 + slot] = speeds[k]
        f[5 * 16 + slot] = timers[k]
        w[96 + slot] = lanes[k]
        w[112 + slot] = 1
    f[3 * 16] = 25.0
    f[4 * 16] = 25.0
    w[128] = 0
    w[129] = 1
    return w


def available_actions(words):
    y0 = words[16:17].view(np.float32)[0]
    cur = int(np.clip(np.rint(y0 / np.float32(4.0)), 0, N_LANES - 1))
    si = int(words[129])
    mask = {1: True, 0: cur > 0, 2: cur < N_LANES - 1, 3: si < 2, 4: si > 0}
    return [a for a in _ORDER if mask[a]]


# ---- time-to-collision grid MDP (docs/HIGHWAY_LITE_SPEC.md section 9) ----------------------------------------------
TTC_SPEEDS = np.array([20.0, 25.0, 30.0])
TTC_STEPS = 10                       # horizon 10 s / time quantization 1 s
TTC_REWARDS = {"collision": -1.0, "right_lane": 0.1, "high_speed": 0.4, "lane_change": 0.0}
_COS_C = np.array([-0.5, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600]).astype(np.float32)


def _cos_p(x):
    """The spec's fp32 cosine polynomial (docs/HIGHWAY_LITE_SPEC.md section 3), one rounding per operation."""
    half_pi = np.float32(1.5707963705062866)
    x = np.minimum(np.maximum(x.astype(np.float32), -half_pi), half_pi)
    z = x * x
    acc = np.full_like(z, _COS_C[-1])
    for c in _COS_C[-2::-1]:
        acc = c + z * acc
    return np.float32(1.0) + z * acc


def ttc_grid(words):
    """[3 speeds, 4 lanes, 10 s] costs in {0, 0.5, 1}, all (speed, vehicle, collision point) triples at once."""
    words = np.asarray(words, dtype=np.int32)
    f = words[:96].view(np.float32).reshape(6, V_SLOTS)
    x, y, h, v = f[0], f[1], f[2], f[3]
    others = np.flatnonzero((words[112:128] & 1) != 0)
    others = others[others > 0]
    grid = np.zeros((TTC_SPEEDS.size, N_LANES, TTC_STEPS))
    if others.size == 0:
        return grid
    proj = v[others].astype(np.float64) * _cos_p(h[others] - h[0]).astype(np.float64)             # [K]
    lane = np.clip(np.rint(y[others] / np.float32(4.0)), 0, N_LANES - 1).astype(np.int64)           # [K]
    diff = TTC_SPEEDS[:, None] - proj[None, :]                                                      # [H, K]
    nz = np.where(np.abs(diff) > 0.01, diff, np.where(diff >= 0, 0.01, -0.01))
    offs = np.array([0.0, -5.0, 5.0])                          # collision points: centre, both bumpers (LENGTH/2 + LENGTH/2)
    cost = np.array([1.0, 0.5, 0.5])
    dist = (x[others].astype(np.float64) - np.float64(x[0]))[None, :, None] + offs[None, None, :]   # [1, K, P]
    ttc = dist / nz[:, :, None]                                                                     # [H, K, P]
    ok = (ttc >= 0) & (TTC_SPEEDS[:, None] != v[others].astype(np.float64)[None, :])[:, :, None]
    ttc = np.where(ok, ttc, -1.0)
    for t in (np.trunc(ttc), np.ceil(ttc)):
        hh, kk, pp = np.nonzero(ok & (t >= 0) & (t < TTC_STEPS))
        np.maximum.at(grid, (hh, lane[kk], t[hh, kk, pp].astype(np.int64)), cost[pp])
    return grid


def ttc_finite_mdp(words):
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    words = np.asarray(words, dtype=np.int32)
    grid = ttc_grid(words)
    n_h, n_l, n_t = grid.shape
    hh, ii, jj = np.meshgrid(np.arange(n_h), np.arange(n_l), np.arange(n_t), indexing="ij")

    def cell(a, b, c):
        return (np.clip(a, 0, n_h - 1) * n_l + np.clip(b, 0, n_l - 1)) * n_t + np.clip(c, 0, n_t - 1)

    first = jj == 0
    nxt = [cell(hh, ii - 1, jj + 1), cell(hh, ii, jj + 1), cell(hh, ii + 1, jj + 1),
           cell(hh + first, ii, jj + 1), cell(hh - first, ii, jj + 1)]            # LEFT, IDLE, RIGHT, FASTER, SLOWER
    transition = np.stack([t.reshape(-1) for t in nxt], axis=1)
    rw = TTC_REWARDS
    state_reward = (rw["collision"] * grid + rw["right_lane"] * (ii / max(n_l - 1, 1))
                    + rw["high_speed"] * (hh / max(n_h - 1, 1))).reshape(-1)
    action_reward = np.array([rw["lane_change"], 0.0, rw["lane_change"], 0.0, 0.0])
    reward = state_reward[:, None] + action_reward[None, :]
    terminal = ((grid == 1) | (jj == n_t - 1)).reshape(-1)
    y0 = words[16:17].view(np.float32)[0]
    ego_lane = int(np.clip(np.rint(y0 / np.float32(4.0)), 0, N_LANES - 1))
    mdp = FiniteMDP("deterministic", transition, reward, terminal, state=int(cell(int(words[129]), ego_lane, 0)))
    mdp.original_shape = grid.shape
    return mdp


class HighwayLiteEnv(object):
    b2_env_kind = "highway"

    def __init__(self, words=None, seed=0, config=None):
        self.config = dict(config or {})
        self.words = np.array(words, dtype=np.int32) if words is not None else make_scene(seed)
        assert self.words.shape == (_lib.HW_STATE_WORDS,)
        self.action_space = _Space(N_ACTIONS)
        self._seed = seed

    @property
    def unwrapped(self):
        return self

    def configure(self, config):
        self.config.update(config)

    def seed(self, seed=None):
        self._seed = seed
        return [seed]

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._seed = seed
        self.words = make_scene(self._seed if self._seed is not None else 0)
        return self.observation(), {}

    def simplify(self):
        """`env_preprocessors: [{"method": "simplify"}]`: a planning copy."""
        return copy.deepcopy(self)

    def get_available_actions(self):
        return available_actions(self.words)

    def to_finite_mdp(self):
        """`env.unwrapped.to_finite_mdp()` of the reference's ValueIterationAgent (value_iteration.py:17,32): the
        time-to-collision grid MDP of this scene (docs/HIGHWAY_LITE_SPEC.md section 9)."""
        return ttc_finite_mdp(self.words)

    def assume_traffic(self, args):
        """Model variant for robust planning (`"models": [[{"method": "assume_traffic", "args": {...}}], ...]`):
        a planning copy in which the other vehicles' target speeds are shifted by `target_speed_offset` m/s
        and / or the vehicle in slot `cut_in_slot` intends to move `cut_in_direction` lanes (-1 left, +1 right)."""
        env = copy.deepcopy(self)
        f = env.words[:96].view(np.float32)
        present = (env.words[112:128] & 1) != 0
        off = np.float32(args.get("target_speed_offset", 0.0))
        for k in range(1, V_SLOTS):
            if present[k]:
                f[4 * 16 + k] = np.float32(f[4 * 16 + k] + off)
        slot = args.get("cut_in_slot")
        if slot is not None and 0 < int(slot) < V_SLOTS and present[int(slot)]:
            env.words[96 + int(slot)] = int(np.clip(env.words[96 + int(slot)] + int(args.get("cut_in_direction", -1)),
                                                    0, N_LANES - 1))
        return env

    def observation(self):
        f = self.words[:64].view(np.float32).reshape(4, V_SLOTS)
        present = (self.words[112:128] & 1).astype(np.float32)
        return np.stack([present, f[0], f[1], f[3] * np.cos(f[2]), f[3] * np.sin(f[2])], axis=1)

    def step(self, action):
        import torch
        lib = _lib.load()
        dev = torch.device("cuda")
        st = torch.from_numpy(self.words.reshape(1, -1)).to(dev)
        act = torch.tensor([int(action)], dtype=torch.int32, device=dev)
        rew = torch.empty(1, dtype=torch.float32, device=dev)
        flg = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(lib.b2_highway_step(_lib.ptr(st), _lib.ptr(act), _lib.ptr(rew), _lib.ptr(flg), None, 1,
                                       _lib.current_stream()))
        self.words = st.cpu().numpy().reshape(-1)
        flags = int(flg.item())
        return self.observation(), float(rew.item()), bool(flags & 1), bool(flags & 2), {}
