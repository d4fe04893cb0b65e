"""OLOP / KL-OLOP agent on the device engine.  Drop-in for
rl_agents.agents.tree_search.olop.OLOPAgent (olop.py:11-200)."""
from rl_agents_b200.agents.common.abstract import register_with_reference
from rl_agents_b200.agents.tree_search.abstract import AbstractPlanner, AbstractTreeSearchAgent
from rl_agents_b200.agents.tree_search.mcts import allocation
from rl_agents_b200.envs.adapters import describe, mdp_fingerprint


class OLOP(AbstractPlanner):
    def __init__(self, env, config=None):
        self.env = env
        super(OLOP, self).__init__(config)

    @classmethod
    def default_config(cls):
        cfg = super(OLOP, cls).default_config()
        cfg.update({"upper_bound": {"type": "hoeffding", "time": "global", "threshold": "4*np.log(time)"},
                    "continuation_type": "zeros"})          # olop.py:20-34
        return cfg

    def reset(self):
        if "horizon" not in self.config:                     # olop.py:36-48
            budget = max(self.env.action_space.n, self.config["budget"])
            self.config["episodes"], self.config["horizon"] = allocation(budget, self.config["gamma"])
        super(OLOP, self).reset()

    def _engine_for(self, d):
        from rl_agents_b200.engine.olop import OLOPEngine
        ub = self.config["upper_bound"]
        key = (d.kind, d.n_actions, self.config["episodes"], self.config["horizon"], self.config["gamma"],
               ub["type"], ub["time"], ub["threshold"], self.config["continuation_type"], mdp_fingerprint(d.mdp))
        if key != self._engine_key:
            self.engine = OLOPEngine(d.kind, 1, d.n_actions, self.config["episodes"], self.config["horizon"],
                                     self.config["gamma"], ub, self.config["continuation_type"], mdp=d.mdp)
            self._engine_key = key
        return self.engine

    def plan(self, state, observation):
        import torch
        from rl_agents_b200.engine.mcts import pcg64_words, set_pcg64_words
        d = describe(state)
        eng = self._engine_for(d)
        root = torch.from_numpy(d.root.reshape(1, -1) if d.root.size > 1 else d.root).to(eng.device)
        eng.plan(root.contiguous(), pcg64_words(self.np_random).reshape(1, -1))
        plans, res, rng_words = eng.finish()
        set_pcg64_words(self.np_random, rng_words[0])
        self.last_tree = eng
        return plans[0]


@register_with_reference
class OLOPAgent(AbstractTreeSearchAgent):
    """An agent that uses Open Loop Optimistic Planning to plan a sequence of actions in an MDP."""
    PLANNER_TYPE = OLOP
