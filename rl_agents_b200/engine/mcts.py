"""Batched MCTS engine (device side of MCTSAgent)."""
import numpy as np

from rl_agents_b200 import _lib
from rl_agents_b200.engine.tables import FiniteTables, gamma_tables, preference_tables, uniform_cdf_table

POLICIES = {"random_available": 0, "random": 1, "preference": 2}


def policy_spec(policy):
    """"random" | "random_available" | ("preference", action, ratio) -> (kernel policy id, action, ratio)"""
    if isinstance(policy, (tuple, list)):
        kind, action, ratio = policy
        if kind != "preference":
            raise ValueError("Unknown policy type")
        return POLICIES[kind], int(action), ratio
    if policy not in ("random_available", "random"):
        raise ValueError("Unknown policy type")
    return POLICIES[policy], -1, 2


def pcg64_words(gen):
    """numpy Generator(PCG64) state -> the 6 x uint64 the kernel advances."""
    st = gen.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise ValueError("the planner RNG must be a numpy PCG64 Generator")
    s, inc = st["state"]["state"], st["state"]["inc"]
    m = (1 << 64) - 1
    return np.array([s >> 64, s & m, inc >> 64, inc & m, st["has_uint32"], st["uinteger"]], dtype=np.uint64)


def set_pcg64_words(gen, w):
    st = gen.bit_generator.state
    w = [int(x) for x in w]
    st["state"]["state"] = (w[0] << 64) | w[1]
    st["state"]["inc"] = (w[2] << 64) | w[3]
    st["has_uint32"], st["uinteger"] = w[4], w[5]
    gen.bit_generator.state = st


def reroot_arrays(arrays, n, action):
    """step_by_subtree (abstract.py:195-206): keep the sub-tree under the root's child `action`, re-indexed
    breadth first (children of a node stay contiguous and in order) with that child as node 0.
    arrays: dict of numpy arrays parent / first_child / count / meta / value / prior (first n entries valid).
    Returns (new arrays dict, number of kept nodes) -- (None, 0) when the action was never expanded."""
    fc, meta = arrays["first_child"], arrays["meta"]
    root_child = -1
    if fc[0] >= 0:
        for c in range(fc[0], fc[0] + ((meta[0] >> 8) & 0xff)):
            if (meta[c] & 0xff) == action:
                root_child = c
    if root_child < 0:
        return None, 0
    order, new_parent = [root_child], [-1]
    head = 0
    new_first = []
    while head < len(order):
        old = order[head]
        k = (meta[old] >> 8) & 0xff if fc[old] >= 0 else 0
        new_first.append(len(order) if k else -1)
        for c in range(fc[old], fc[old] + k):
            order.append(c)
            new_parent.append(head)
        head += 1
    order = np.array(order)
    out = {"parent": np.array(new_parent, dtype=np.int32), "first_child": np.array(new_first, dtype=np.int32),
           "count": arrays["count"][order].copy(), "meta": arrays["meta"][order].copy(),
           "value": arrays["value"][order].copy(), "prior": arrays["prior"][order].copy()}
    out["meta"][0] = (out["meta"][0] & ~0xff) | 0xff          # the new root has no incoming action
    return out, len(order)


class MCTSEngine(object):
    """n_trees independent MCTS decisions per launch; every tree consumes its
    own numpy PCG64 stream exactly as the reference planner would."""

    def __init__(self, env_kind, n_trees, n_actions, episodes, horizon, gamma, temperature, mdp=None,
                 rollout_policy="random_available", prior_policy="random_available", device="cuda", capacity=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        rollout_id, rollout_action, rollout_ratio = policy_spec(rollout_policy)
        prior_id, prior_action, prior_ratio = policy_spec(prior_policy)
        self.n_trees, self.n_actions = int(n_trees), int(n_actions)
        self.episodes, self.horizon = int(episodes), int(horizon)
        self.capacity = max(int(capacity or 0), 1 + self.episodes * self.n_actions)
        gp, _ = gamma_tables(gamma, self.horizon + 1)
        self.gamma_pow = torch.as_tensor(gp, device=self.device)
        self.cdf = torch.as_tensor(uniform_cdf_table(self.n_actions), device=self.device)
        self.tables = FiniteTables(mdp, self.device) if env_kind == _lib.ENV_FINITE else None
        shape = (self.n_trees, self.capacity)
        i32, f64 = torch.int32, torch.float64
        self.parent = torch.empty(shape, dtype=i32, device=self.device)
        self.first_child = torch.empty(shape, dtype=i32, device=self.device)
        self.count = torch.empty(shape, dtype=i32, device=self.device)
        self.meta = torch.empty(shape, dtype=i32, device=self.device)
        self.value = torch.empty(shape, dtype=f64, device=self.device)
        self.prior = torch.empty(shape, dtype=f64, device=self.device)
        self.pref_prior = torch.as_tensor(preference_tables(self.n_actions, prior_ratio)[0], device=self.device)
        self.pref_cdf = torch.as_tensor(preference_tables(self.n_actions, rollout_ratio)[1], device=self.device)
        self.cfg = _lib.MCTSConfig(env_kind, self.n_trees, self.n_actions, self.episodes, self.horizon, self.capacity,
                                   rollout_id, prior_id, float(temperature),
                                   self.gamma_pow.data_ptr(), self.cdf.data_ptr(),
                                   self.tables.struct() if self.tables else _lib.FiniteMDP(),
                                   prior_action, rollout_action, self.pref_prior.data_ptr(), self.pref_cdf.data_ptr(), None)
        self.resume = torch.zeros(self.n_trees, dtype=i32, device=self.device)
        self.tree = _lib.MCTSTree(*[t.data_ptr() for t in (self.parent, self.first_child, self.count, self.meta,
                                                           self.value, self.prior)])
        self.plan_buf = torch.empty((self.n_trees, max(self.horizon, 1)), dtype=torch.int8, device=self.device)
        self.result = torch.empty((self.n_trees, _lib.MCTS_RESULT_WORDS), dtype=i32, device=self.device)
        self.rng = torch.empty((self.n_trees, _lib.PCG64_STATE_WORDS), dtype=torch.int64, device=self.device)

    def plan(self, root_states, rng_words, resume_nodes=None):
        """rng_words: uint64 [n_trees, 6] numpy (pcg64_words per tree).  resume_nodes: per-tree node counts of
        re-rooted sub-trees already in the arrays (see reroot), or None for fresh trees."""
        self.rng.copy_(self.torch.from_numpy(np.ascontiguousarray(rng_words).view(np.int64)))
        if resume_nodes is None:
            self.cfg.resume_nodes = None
        else:
            self.resume.copy_(self.torch.as_tensor(np.asarray(resume_nodes, dtype=np.int32)))
            self.cfg.resume_nodes = self.resume.data_ptr()
        _lib.check(self.lib.b2_mcts_plan(self.cfg, _lib.ptr(root_states), self.tree, _lib.ptr(self.rng),
                                         _lib.ptr(self.plan_buf), _lib.ptr(self.result), _lib.current_stream()))

    def finish(self):
        """Synchronise; returns (plans, result array, advanced rng words)."""
        res = self.result.cpu().numpy()
        plans_dev = self.plan_buf.cpu().numpy()
        plans = [plans_dev[i, :res[i, 1]].astype(int).tolist() for i in range(self.n_trees)]
        return plans, res, self.rng.cpu().numpy().view(np.uint64)

    def reroot(self, tree, action):
        """Keep the sub-tree under root child `action` of `tree` (host-side compaction, once per decision).
        Returns the number of nodes kept (0: the action was never expanded -> start a new tree)."""
        n = int(self.result[tree, 0].item())
        names = ("parent", "first_child", "count", "meta", "value", "prior")
        arrays = {k: getattr(self, k)[tree, :n].cpu().numpy() for k in names}
        out, kept = reroot_arrays(arrays, n, int(action))
        if kept + self.episodes * self.n_actions > self.capacity:
            return 0                                   # would not fit: fall back to a fresh tree
        for k in names if kept else ():
            getattr(self, k)[tree, :kept] = self.torch.as_tensor(out[k], device=self.device)
        return kept

    def tree_dict(self, tree=0):
        n = int(self.result[tree, 0].item())
        meta = self.meta[tree, :n].cpu().numpy()
        action = (meta & 0xff).astype(int)
        action[action == 0xff] = -1
        return {"parent": self.parent[tree, :n].cpu().numpy(), "action": action,
                "count": self.count[tree, :n].cpu().numpy(), "value": self.value[tree, :n].cpu().numpy(),
                "prior": self.prior[tree, :n].cpu().numpy(),
                "first_child": self.first_child[tree, :n].cpu().numpy(), "n_children": (meta >> 8) & 0xff}


class MCTSWaveEngine(object):
    """ONE MCTS decision searched by the whole GPU in waves of `width` episodes (b2_mcts_plan_wave);
    bit-identical with the specification oracle/planners.py::mcts_plan_wavefront."""

    def __init__(self, env_kind, n_actions, episodes, horizon, gamma, temperature, width, mdp=None, device="cuda",
                 max_ctas=0):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.n_actions, self.episodes, self.horizon = int(n_actions), int(episodes), int(horizon)
        self.width = max(1, min(int(width), 1024))
        self.capacity = 1 + self.episodes * self.n_actions
        gp, _ = gamma_tables(gamma, self.horizon + 1)
        self.gamma_pow = torch.as_tensor(gp, device=self.device)
        self.tables = FiniteTables(mdp, self.device) if env_kind == _lib.ENV_FINITE else None
        i32 = torch.int32
        self.parent = torch.empty(self.capacity, dtype=i32, device=self.device)
        self.first_child = torch.empty(self.capacity, dtype=i32, device=self.device)
        self.count = torch.empty(self.capacity, dtype=i32, device=self.device)
        self.meta = torch.empty(self.capacity, dtype=i32, device=self.device)
        self.vsum = torch.empty(self.capacity, dtype=torch.int64, device=self.device)
        self.value = torch.empty(self.capacity, dtype=torch.float64, device=self.device)
        self.cfg = _lib.MCTSWaveConfig(env_kind, self.n_actions, self.episodes, self.horizon, self.capacity, self.width,
                                       0, 0, float(temperature), 0, self.gamma_pow.data_ptr(),
                                       self.tables.struct() if self.tables else _lib.FiniteMDP(), int(max_ctas), 0)
        self.tree = _lib.MCTSWaveTree(*[t.data_ptr() for t in (self.parent, self.first_child, self.count, self.meta,
                                                               self.vsum, self.value)])
        ws = self.lib.b2_mcts_wave_workspace_bytes(self.cfg)
        if ws < 0:
            raise _lib.B2Error("unsupported wavefront MCTS configuration")
        self.workspace = torch.empty(int(ws), dtype=torch.uint8, device=self.device)
        self.plan_buf = torch.empty(max(self.horizon, 1), dtype=torch.int8, device=self.device)
        self.result = torch.empty(_lib.MCTS_RESULT_WORDS, dtype=i32, device=self.device)

    def plan(self, root_state, seed):
        """root_state: int32 device tensor [1] (finite) or [136]; seed: the counter-based generator's seed."""
        assert root_state.dtype == self.torch.int32 and root_state.is_cuda and root_state.is_contiguous()
        self.cfg.seed = int(seed) & ((1 << 64) - 1)
        _lib.check(self.lib.b2_mcts_plan_wave(self.cfg, _lib.ptr(root_state), self.tree, _lib.ptr(self.workspace),
                                              _lib.ptr(self.plan_buf), _lib.ptr(self.result), _lib.current_stream()))

    def finish(self):
        res = self.result.cpu().numpy()
        return self.plan_buf.cpu().numpy()[:res[1]].astype(int).tolist(), res

    def tree_dict(self):
        meta = self.meta.cpu().numpy()
        action = (meta & 0xff).astype(int)
        action[action == 0xff] = -1
        return {"parent": self.parent.cpu().numpy(), "action": action, "count": self.count.cpu().numpy(),
                "first_child": self.first_child.cpu().numpy(), "n_children": (meta >> 8) & 0xff,
                "vsum": self.vsum.cpu().numpy(), "value": self.value.cpu().numpy()}

    def root_statistics(self):
        d = self.tree_dict()
        fc, n = int(d["first_child"][0]), int(d["n_children"][0])
        counts, values = np.zeros(self.n_actions), np.zeros(self.n_actions)
        for c in range(fc, fc + n):
            counts[d["action"][c]], values[d["action"][c]] = d["count"][c], d["value"][c]
        return counts, values
