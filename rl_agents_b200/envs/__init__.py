from rl_agents_b200.envs.finite_mdp import FiniteMDPEnv  # noqa: F401
from rl_agents_b200.envs.highway_lite import HighwayLiteEnv  # noqa: F401
from rl_agents_b200.envs.intersection_lite import IntersectionLiteEnv  # noqa: F401
