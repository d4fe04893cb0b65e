"""Batched GBOP-T engine (device side of StateAwarePlannerAgent)."""
import numpy as np

from rl_agents_b200 import _lib
from rl_agents_b200.engine.tables import FiniteTables, gamma_tables, terminal_bonus_table


class GBOPEngine(object):
    """n_trees independent GBOP-T decisions per launch (one warp per tree) on a deterministic finite MDP."""

    def __init__(self, n_trees, n_actions, budget, gamma, mdp, terminal_reward=0.0, backup_aggregated_nodes=True,
                 prune_suboptimal_leaves=True, accuracy=0, device="cuda", queue_factor=64):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.n_trees, self.n_actions = int(n_trees), int(n_actions)
        self.n_expansions = int(budget) // self.n_actions
        self.capacity = 1 + self.n_expansions * self.n_actions
        self.plan_capacity = self.n_expansions + 1
        gamma = float(gamma)
        gp, _ = gamma_tables(gamma, self.n_expansions + 2)
        self.gamma_pow = torch.as_tensor(gp, device=self.device)
        self.terminal_bonus = torch.as_tensor(terminal_bonus_table(terminal_reward, gamma, self.n_expansions + 2),
                                              device=self.device)
        self.tables = FiniteTables(mdp, self.device)
        self.n_states = self.tables.n_states
        shape = (self.n_trees, self.capacity)
        i32, f64 = torch.int32, torch.float64
        names_i = ("parent", "first_child", "depth", "count", "meta", "obs")
        for n in names_i:
            setattr(self, n, torch.empty(shape, dtype=i32, device=self.device))
        self.reward = torch.empty(shape, dtype=f64, device=self.device)
        self.lower = torch.empty(shape, dtype=f64, device=self.device)
        self.cfg = _lib.GBOPConfig(self.n_trees, self.n_actions, self.n_expansions, self.capacity, self.plan_capacity,
                                   int(queue_factor) * self.capacity, 1 if backup_aggregated_nodes else 0,
                                   1 if prune_suboptimal_leaves else 0, gamma, 1 / (1 - gamma), accuracy * (1 - gamma),
                                   self.gamma_pow.data_ptr(), self.terminal_bonus.data_ptr(), self.tables.struct())
        self.tree = _lib.GBOPTree(*[t.data_ptr() for t in (self.parent, self.first_child, self.depth, self.count, self.meta,
                                                           self.reward, self.lower, self.obs)])
        ws = self.lib.b2_gbop_workspace_bytes(self.cfg)
        if ws < 0:
            raise _lib.B2Error("unsupported GBOP configuration")
        self.ws_per_tree = int(ws) // self.n_trees
        self.workspace = torch.empty(int(ws), dtype=torch.uint8, device=self.device)
        self.plan_buf = torch.empty((self.n_trees, self.plan_capacity), dtype=torch.int8, device=self.device)
        self.result = torch.empty((self.n_trees, _lib.OPD_RESULT_WORDS), dtype=i32, device=self.device)

    def plan(self, root_states):
        assert root_states.dtype == self.torch.int32 and root_states.is_cuda and root_states.is_contiguous()
        _lib.check(self.lib.b2_gbop_plan(self.cfg, _lib.ptr(root_states), self.tree, _lib.ptr(self.workspace),
                                         _lib.ptr(self.plan_buf), _lib.ptr(self.result), _lib.current_stream()))

    def finish(self, np_randoms=None):
        """Synchronise; returns (plans, result).  StateAwarePlanner.plan runs get_plan() twice (state_aware.py:124,
        :130; the first inside super().plan()): a tie consumes the planner RNG on both walks, the second is returned."""
        res = self.result.cpu().numpy()
        if (res[:, 4] != 0).any():
            raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
        if (res[:, 7] != 0).any():
            raise _lib.B2Error("GBOP backup queue overflow: raise queue_factor")
        plans_dev = self.plan_buf.cpu().numpy()
        plans = []
        for i in range(self.n_trees):
            head = plans_dev[i, :res[i, 5]].astype(int).tolist()
            if res[i, 6] >= 0:
                rng = np_randoms[i] if np_randoms is not None else np.random.default_rng()
                self._host_plan_from(i, int(res[i, 6]), rng)
                plans.append(head + self._host_plan_from(i, int(res[i, 6]), rng))
            else:
                plans.append(head)
        return plans, res

    def _host_plan_from(self, tree, node, rng):
        fc = self.first_child[tree].cpu().numpy()
        meta = self.meta[tree].cpu().numpy()
        lower = self.lower[tree].cpu().numpy()
        plan = []
        while fc[node] >= 0:
            n = (meta[node] >> 8) & 0xff
            x = lower[fc[node]:fc[node] + n]
            indices = np.nonzero(x == np.amax(x))[0]
            node = fc[node] + int(rng.choice(indices))
            plan.append(int(meta[node] & 0xff))
        return plan

    def state_values(self, tree=0):
        off = tree * self.ws_per_tree
        return self.workspace[off:off + 8 * self.n_states].view(self.torch.float64).cpu().numpy()

    def tree_dict(self, tree=0):
        n = int(self.result[tree, 0].item())
        meta = self.meta[tree, :n].cpu().numpy()
        action = (meta & 0xff).astype(int)
        action[action == 0xff] = -1
        return {"parent": self.parent[tree, :n].cpu().numpy(), "action": action,
                "count": self.count[tree, :n].cpu().numpy(), "depth": self.depth[tree, :n].cpu().numpy(),
                "first_child": self.first_child[tree, :n].cpu().numpy(), "n_children": (meta >> 8) & 0xff,
                "done": ((meta >> 16) & 1).astype(bool), "leaf": ((meta >> 17) & 1).astype(bool),
                "reward": self.reward[tree, :n].cpu().numpy(), "lower": self.lower[tree, :n].cpu().numpy(),
                "obs": self.obs[tree, :n].cpu().numpy()}


class GBOPDEngine(object):
    """n_trees independent GBOP-D decisions per launch (GraphBasedPlanner, one warp per decision)."""

    def __init__(self, n_trees, n_actions, budget, gamma, mdp, accuracy=1e-2, sampling_timeout=100, device="cuda",
                 queue_factor=256):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.n_trees, self.n_actions = int(n_trees), int(n_actions)
        self.tables = FiniteTables(mdp, self.device)
        S = self.n_states = self.tables.n_states
        T = np.asarray(mdp.transition, dtype=np.int64)
        # reverse transitions (the potential parents of every state), ascending and unique
        pairs = np.unique(np.stack([T.reshape(-1), np.repeat(np.arange(S), self.n_actions)], axis=1), axis=0)
        ptr = np.zeros(S + 1, dtype=np.int32)
        np.add.at(ptr, pairs[:, 0] + 1, 1)
        self.rev_ptr = torch.as_tensor(np.cumsum(ptr).astype(np.int32), device=self.device)
        self.rev_idx = torch.as_tensor(pairs[:, 1].astype(np.int32), device=self.device)
        self.timeout = int(sampling_timeout)
        gamma = float(gamma)
        self.queue_capacity = int(queue_factor) * S
        self.cfg = _lib.GBOPDConfig(self.n_trees, self.n_actions, int(budget) // self.n_actions, self.timeout, self.timeout,
                                    self.queue_capacity, gamma, 1 / (1 - gamma), float(accuracy), self.tables.struct(),
                                    self.rev_ptr.data_ptr(), self.rev_idx.data_ptr())
        self.lower = torch.empty((self.n_trees, S), dtype=torch.float64, device=self.device)
        self.upper = torch.empty((self.n_trees, S), dtype=torch.float64, device=self.device)
        self.flags = torch.empty((self.n_trees, S), dtype=torch.uint8, device=self.device)
        self.queue = torch.empty((self.n_trees, self.queue_capacity), dtype=torch.int32, device=self.device)
        self.rng = torch.empty((self.n_trees, _lib.PCG64_STATE_WORDS), dtype=torch.int64, device=self.device)
        self.plan_buf = torch.empty((self.n_trees, self.timeout), dtype=torch.int8, device=self.device)
        self.result = torch.empty((self.n_trees, _lib.OPD_RESULT_WORDS), dtype=torch.int32, device=self.device)

    def plan(self, root_states, rng_words):
        self.rng.copy_(self.torch.from_numpy(np.ascontiguousarray(rng_words).view(np.int64)))
        _lib.check(self.lib.b2_gbopd_plan(self.cfg, _lib.ptr(root_states), _lib.ptr(self.lower), _lib.ptr(self.upper),
                                          _lib.ptr(self.flags), _lib.ptr(self.queue), _lib.ptr(self.rng),
                                          _lib.ptr(self.plan_buf), _lib.ptr(self.result), _lib.current_stream()))

    def finish(self):
        res = self.result.cpu().numpy()
        if (res[:, 7] != 0).any():
            raise _lib.B2Error("GBOP-D backup queue overflow: raise queue_factor")
        plans_dev = self.plan_buf.cpu().numpy()
        return [plans_dev[i, :res[i, 5]].astype(int).tolist() for i in range(self.n_trees)], res, \
            self.rng.cpu().numpy().view(np.uint64)

    def nodes(self, tree=0):
        fl = self.flags[tree].cpu().numpy()
        lo, up = self.lower[tree].cpu().numpy(), self.upper[tree].cpu().numpy()
        return {int(s): dict(lower=float(lo[s]), upper=float(up[s]), expanded=bool(fl[s] & 2)) for s in np.nonzero(fl & 1)[0]}
