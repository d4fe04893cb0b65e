"""GBOP-D agent on the device engine.  Drop-in for
rl_agents.agents.tree_search.graph_based.GraphBasedPlannerAgent (graph_based.py:84-150) on deterministic finite MDPs."""
from rl_agents_b200 import _lib
from rl_agents_b200.agents.common.abstract import register_with_reference
from rl_agents_b200.agents.tree_search.abstract import AbstractPlanner, AbstractTreeSearchAgent
from rl_agents_b200.envs.adapters import describe, mdp_fingerprint


class GraphBasedPlanner(AbstractPlanner):
    def __init__(self, env, config=None):
        super(GraphBasedPlanner, self).__init__(config)
        self.env = env

    def plan(self, state, observation):
        import torch
        from rl_agents_b200.engine.gbop import GBOPDEngine
        from rl_agents_b200.engine.mcts import pcg64_words, set_pcg64_words
        d = describe(state)
        if d.kind != _lib.ENV_FINITE:
            raise TypeError("the device GBOP-D planner builds a graph over state ids: it needs a finite-MDP env")
        key = (d.n_actions, self.config["budget"], self.config["gamma"], self.config["accuracy"],
               self.config["sampling_timeout"], mdp_fingerprint(d.mdp))
        if key != self._engine_key:
            self.engine = GBOPDEngine(1, d.n_actions, self.config["budget"], self.config["gamma"], d.mdp,
                                      self.config["accuracy"], self.config["sampling_timeout"])
            self._engine_key = key
        eng = self.engine
        eng.plan(torch.from_numpy(d.root).to(eng.device).contiguous(), pcg64_words(self.np_random).reshape(1, -1))
        plans, _, words = eng.finish()
        set_pcg64_words(self.np_random, words[0])          # the tie-breaks consumed the planner's stream
        self.last_tree = eng
        return plans[0]


@register_with_reference
class GraphBasedPlannerAgent(AbstractTreeSearchAgent):
    PLANNER_TYPE = GraphBasedPlanner

    @classmethod
    def default_config(cls):
        cfg = super(GraphBasedPlannerAgent, cls).default_config()
        cfg.update({"sampling_timeout": 100, "accuracy": 1e-2})        # graph_based.py:143-150
        return cfg
