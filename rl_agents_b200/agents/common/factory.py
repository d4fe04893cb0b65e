"""Env pre-processing with the semantics of rl_agents/agents/common/factory.py:97-116."""
import logging

logger = logging.getLogger(__name__)


def preprocess_env(env, preprocessor_configs):
    for preprocessor_config in preprocessor_configs:
        if "method" in preprocessor_config:
            try:
                preprocessor = getattr(env.unwrapped, preprocessor_config["method"])
                if "args" in preprocessor_config:
                    env = preprocessor(preprocessor_config["args"])
                else:
                    env = preprocessor()
            except AttributeError:
                logger.warning("The environment does not have a {} method".format(preprocessor_config["method"]))
        else:
            logger.error("The method is not specified in {}".format(preprocessor_config))
    return env
