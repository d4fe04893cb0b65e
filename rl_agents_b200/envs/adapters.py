"""The one place the planners touch an env's object model ("env
hand-off"): turn the (pre-processed) env object `agent.plan` receives into the
arrays the kernels read."""
import numpy as np

from rl_agents_b200 import _lib


class EnvDescription(object):
    __slots__ = ("kind", "n_actions", "mdp", "root")


def describe(env):
    """-> EnvDescription(kind, action_space.n, finite MDP tables or None, root state int32 array)."""
    u = getattr(env, "unwrapped", env)
    d = EnvDescription()
    d.n_actions = int(env.action_space.n)
    kind = getattr(u, "b2_env_kind", None)
    if kind == "highway" or (kind is None and hasattr(u, "words")):
        d.kind, d.mdp = _lib.ENV_HIGHWAY, None
        d.root = np.ascontiguousarray(u.words, dtype=np.int32)
        return d
    mdp = getattr(u, "mdp", None)
    if mdp is not None and hasattr(mdp, "transition") and hasattr(mdp, "reward"):
        # rl_agents_b200.envs.FiniteMDPEnv or the `finite_mdp` package's FiniteMDPEnv
        d.kind, d.mdp = _lib.ENV_FINITE, mdp
        d.root = np.array([int(mdp.state)], dtype=np.int32)
        return d
    raise TypeError("rl_agents_b200 planners need a HighwayLiteEnv or a finite-MDP env "
                    "(got %r); see INTEGRATION.md for the env hand-off" % type(u).__name__)
