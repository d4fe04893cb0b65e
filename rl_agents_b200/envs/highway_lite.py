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
