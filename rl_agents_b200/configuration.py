"""Configuration container with the contract of the reference's
rl_agents/configuration.py:5-44: defaults are recursively overridden by the
user's dict, and the completed configuration is written back INTO the user's
dict (so that `serialize`/metadata dumps see every key)."""
from collections.abc import Mapping


def rec_update(d, u):
    for k, v in u.items():
        if isinstance(v, Mapping):
            d[k] = rec_update(d.get(k, {}), v)
        else:
            d[k] = v
    return d


class Configurable(object):
    def __init__(self, config=None):
        self.config = self.default_config()
        if config:
            rec_update(self.config, config)
            rec_update(config, self.config)

    def update_config(self, config):
        rec_update(self.config, config)

    @classmethod
    def default_config(cls):
        return {}

    rec_update = staticmethod(rec_update)
