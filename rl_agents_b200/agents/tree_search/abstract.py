"""Agent / planner shells shared by the device tree-search planners.

Mirrors the behaviour of rl_agents/agents/tree_search/abstract.py
(AbstractTreeSearchAgent :15-106, AbstractPlanner :109-206): receding-horizon
bookkeeping, env pre-processing, planner seeding through a numpy PCG64
Generator (what gymnasium's seeding.np_random builds), `reset` tree-step
strategy.  The search itself lives on the device (rl_agents_b200.engine).
"""
import logging
from collections import defaultdict

import numpy as np

from rl_agents_b200.agents.common.abstract import AbstractAgent
from rl_agents_b200.agents.common.factory import preprocess_env
from rl_agents_b200.configuration import Configurable

logger = logging.getLogger(__name__)


def np_random(seed=None):
    """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
    seed_seq = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(seed_seq)), seed_seq.entropy


class AbstractTreeSearchAgent(AbstractAgent):
    PLANNER_TYPE = None

    def __init__(self, env, config=None):
        super(AbstractTreeSearchAgent, self).__init__(config)
        self.env = env
        self.planner = self.make_planner()
        self.previous_actions = []
        self.remaining_horizon = 0
        self.steps = 0

    @classmethod
    def default_config(cls):
        return {"env_preprocessors": [], "display_tree": False, "receding_horizon": 1, "terminal_reward": 0}

    def make_planner(self):
        if self.PLANNER_TYPE:
            return self.PLANNER_TYPE(self.env, self.config)
        raise NotImplementedError()

    def plan(self, observation):
        self.steps += 1
        replanning_required = self.step(self.previous_actions)
        if replanning_required:
            env = preprocess_env(self.env, self.config["env_preprocessors"])
            actions = self.planner.plan(state=env, observation=observation)
        else:
            actions = self.previous_actions[1:]
        self.previous_actions = actions
        return actions

    def step(self, actions):
        replanning_required = self.remaining_horizon == 0 or len(actions) <= 1
        if replanning_required:
            self.remaining_horizon = self.config["receding_horizon"] - 1
        else:
            self.remaining_horizon -= 1
        self.planner.step_tree(actions)
        return replanning_required

    def reset(self):
        self.planner.step_by_reset()
        self.remaining_horizon = 0
        self.steps = 0

    def seed(self, seed=None):
        return self.planner.seed(seed)

    def record(self, state, action, reward, next_state, done, info):
        pass

    def act(self, state):
        return self.plan(state)[0]

    def save(self, filename):
        return False

    def load(self, filename):
        return False


class AbstractPlanner(Configurable):
    def __init__(self, config=None):
        super(AbstractPlanner, self).__init__(config)
        self.np_random = None
        self.engine = None
        self._engine_key = None
        self.last_tree = None
        self.reset()
        self.seed()

    @classmethod
    def default_config(cls):
        return dict(budget=500, gamma=0.8, step_strategy="reset")

    def seed(self, seed=None):
        self.np_random, seed = np_random(seed)
        return [seed]

    def plan(self, state, observation):
        raise NotImplementedError()

    def get_visits(self):
        return defaultdict(int)

    def get_updates(self):
        return defaultdict(int)

    def step_tree(self, actions):
        # the device planners rebuild the tree per decision: only "reset" applies
        # (OPD "subtree" crashes in the reference)
        if self.config["step_strategy"] != "reset":
            logger.warning("step strategy {} is not supported by the device planners: "
                           "resetting".format(self.config["step_strategy"]))
        self.step_by_reset()

    def step_by_reset(self):
        self.reset()

    def reset(self):
        self.last_tree = None
