"""MCTS / UCT agent on the device engine.  Drop-in for
rl_agents.agents.tree_search.mcts.MCTSAgent (mcts.py:12-305, open loop)."""
import numpy as np

from rl_agents_b200.agents.common.abstract import register_with_reference
from rl_agents_b200.agents.tree_search.abstract import AbstractPlanner, AbstractTreeSearchAgent
from rl_agents_b200.envs.adapters import describe


def horizon_for(episodes, gamma):
    # OLOP.horizon (olop.py:42-44)
    return max(int(np.ceil(np.log(episodes) / (2 * np.log(1 / gamma)))), 1)


def allocation(budget, gamma):
    """OLOP.allocation (olop.py:50-62): budget -> (episodes, horizon)."""
    for episodes in range(1, int(budget)):
        if episodes * horizon_for(episodes, gamma) > budget:
            episodes = max(episodes - 1, 1)
            horizon = horizon_for(episodes, gamma)
            break
    else:
        raise ValueError("Could not split budget {} with gamma {}".format(budget, gamma))
    return episodes, horizon


class MCTS(AbstractPlanner):
    def __init__(self, env, prior_policy, rollout_policy, config=None):
        super(MCTS, self).__init__(config)
        self.env = env
        self.prior_policy = prior_policy
        self.rollout_policy = rollout_policy
        if not self.config["horizon"]:                                   # mcts.py:116-118
            self.config["episodes"], self.config["horizon"] = allocation(self.config["budget"], self.config["gamma"])
        if self.config.get("closed_loop"):
            raise NotImplementedError("closed_loop MCTS is outside the device planner's scope (DESIGN.md)")

    @classmethod
    def default_config(cls):
        cfg = super(MCTS, cls).default_config()
        cfg.update({"temperature": 2 / (1 - cfg["gamma"]), "closed_loop": False})    # mcts.py:120-127
        return cfg

    def _engine_for(self, d):
        from rl_agents_b200.engine.mcts import MCTSEngine
        key = (d.kind, d.n_actions, self.config["episodes"], self.config["horizon"], self.config["gamma"],
               self.config["temperature"], id(d.mdp))
        if key != self._engine_key:
            self.engine = MCTSEngine(d.kind, 1, d.n_actions, self.config["episodes"], self.config["horizon"],
                                     self.config["gamma"], self.config["temperature"], mdp=d.mdp,
                                     rollout_policy=self.rollout_policy, prior_policy=self.prior_policy)
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
        set_pcg64_words(self.np_random, rng_words[0])     # the device consumed the planner's stream
        self.last_tree = eng
        return plans[0]


@register_with_reference
class MCTSAgent(AbstractTreeSearchAgent):
    """An agent that uses Monte Carlo Tree Search to plan a sequence of actions in an MDP."""

    def make_planner(self):
        return MCTS(self.env, MCTSAgent.policy_factory(self.config["prior_policy"]),
                    MCTSAgent.policy_factory(self.config["rollout_policy"]), self.config)

    @classmethod
    def default_config(cls):
        config = super(MCTSAgent, cls).default_config()
        config.update({"budget": 100, "horizon": None,
                       "prior_policy": {"type": "random_available"},
                       "rollout_policy": {"type": "random_available"},
                       "env_preprocessors": []})
        return config

    @staticmethod
    def policy_factory(policy_config):
        """mcts.py:34-44 -> the policy id the kernel implements."""
        kind = policy_config["type"]
        if kind in ("random", "random_available"):
            return kind
        if kind == "preference":
            raise NotImplementedError("the 'preference' policy is not implemented by the device planner")
        raise ValueError("Unknown policy type")
