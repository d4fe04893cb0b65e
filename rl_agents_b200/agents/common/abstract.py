"""The agent plugin contract (reference: rl_agents/agents/common/abstract.py:6-98).

`Evaluation` (rl_agents/trainer/evaluation.py) calls plan / record / seed /
reset / set_writer / set_directory / save / load / eval on whatever class
agent_factory instantiates; when the reference package is importable the
classes here are also registered as virtual subclasses of its AbstractAgent.
"""
from abc import ABC, abstractmethod

from rl_agents_b200.configuration import Configurable


class AbstractAgent(Configurable, ABC):
    def __init__(self, config=None):
        super(AbstractAgent, self).__init__(config)
        self.writer = None
        self.directory = None
        register_with_reference(type(self))

    @abstractmethod
    def record(self, state, action, reward, next_state, done, info):
        raise NotImplementedError()

    @abstractmethod
    def act(self, state):
        raise NotImplementedError()

    def plan(self, state):
        return [self.act(state)]

    @abstractmethod
    def reset(self):
        raise NotImplementedError()

    @abstractmethod
    def seed(self, seed=None):
        raise NotImplementedError()

    @abstractmethod
    def save(self, filename):
        raise NotImplementedError()

    @abstractmethod
    def load(self, filename):
        raise NotImplementedError()

    def eval(self):
        pass

    def set_writer(self, writer):
        self.writer = writer

    def set_directory(self, directory):
        self.directory = directory

    def set_time(self, time):
        pass


def register_with_reference(cls):
    """isinstance(agent, rl_agents...AbstractAgent) holds when the reference is installed."""
    import sys
    ref = sys.modules.get("rl_agents.agents.common.abstract")
    if ref is None and "rl_agents" not in sys.modules:
        return cls          # reference not loaded in this process: nothing to register with
    try:
        if ref is None:
            from rl_agents.agents.common import abstract as ref
        ref.AbstractAgent.register(cls)
    except Exception:
        pass
    return cls
