"""OPD agent on the device engine.  Drop-in for
rl_agents.agents.tree_search.deterministic.DeterministicPlannerAgent
(deterministic.py:91-139): same constructor, config keys and defaults, same
plan()/act() results on identical seeds."""
from rl_agents_b200.agents.common.abstract import register_with_reference
from rl_agents_b200.agents.tree_search.abstract import AbstractPlanner, AbstractTreeSearchAgent
from rl_agents_b200 import _lib
from rl_agents_b200.envs.adapters import describe, mdp_fingerprint


class OptimisticDeterministicPlanner(AbstractPlanner):
    """plan(): one OPD decision, searched by libb2planner (b2_opd_plan)."""

    def __init__(self, env, config=None):
        super(OptimisticDeterministicPlanner, self).__init__(config)
        self.env = env

    def _engine_for(self, d):
        """Default: the reference's strict best-first order (one CTA, bit-exact with deterministic.py).
        Extension `"wavefront": K` (K >= 1): the whole GPU searches the ONE decision in waves of K leaves
        (b2_opd_plan_wave; K = 1 is again the strict order) -- 30-60x lower latency at K = 64..128."""
        from rl_agents_b200.engine.opd import OPDEngine, OPDWaveEngine
        width = int(self.config.get("wavefront", 0) or 0)
        if d.kind == _lib.ENV_INTERSECTION:
            width = max(width, 1)          # IntersectionLite lives in the wavefront kernel (width 1 = strict order)
        key = (d.kind, d.n_actions, self.config["budget"], self.config["gamma"],
               self.config.get("terminal_reward", 0), width, mdp_fingerprint(d.mdp))
        if key != self._engine_key:
            if width > 0:
                self.engine = OPDWaveEngine(d.kind, d.n_actions, self.config["budget"], self.config["gamma"], width,
                                            self.config.get("terminal_reward", 0), mdp=d.mdp)
            else:
                self.engine = OPDEngine(d.kind, 1, d.n_actions, self.config["budget"], self.config["gamma"],
                                        self.config.get("terminal_reward", 0), mdp=d.mdp,
                                        keys_in_smem=self.config.get("keys_in_smem", True))
            self._engine_key = key
        return self.engine

    def plan(self, state, observation):
        import torch
        d = describe(state)
        eng = self._engine_for(d)
        root = torch.from_numpy(d.root.reshape(1, -1) if d.root.size > 1 else d.root).to(eng.device)
        eng.plan(root.contiguous())
        plans, res = eng.finish([self.np_random])
        self.last_tree = eng
        return plans[0]


@register_with_reference
class DeterministicPlannerAgent(AbstractTreeSearchAgent):
    """An agent that performs optimistic planning in deterministic MDPs."""
    PLANNER_TYPE = OptimisticDeterministicPlanner
