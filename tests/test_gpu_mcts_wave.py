"""GPU parity tests of the wavefront MCTS (b2_mcts_plan_wave): bit-exact against its specification
(oracle/planners.py::mcts_plan_wavefront and the same statement in C) -- node ids, counts, fixed-point
value sums, recommended plan, env-step count."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_mdps

pytestmark = pytest.mark.gpu
M = load_mdps()


def check(eng, plan, t):
    got_plan, res = eng.finish()
    d = eng.tree_dict()
    assert d["parent"].tolist() == list(t.parent if hasattr(t, "parent") else t["parent"])
    return got_plan, res, d


@pytest.mark.parametrize("width,episodes,horizon", [(1, 60, 6), (7, 100, 9), (64, 357, 28), (256, 300, 12), (1024, 1500, 5)])
def test_mcts_wave_finite_matches_the_specification(width, episodes, horizon):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.mcts import MCTSWaveEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    term = M["large1_term"].copy()
    term[[3, 17, 66, 91]] = True
    mdp = FiniteMDP("deterministic", M["large1_T"], M["large1_R"], term)
    eng = MCTSWaveEngine(_lib.ENV_FINITE, 5, episodes, horizon, 0.9, 10.0, width, mdp=mdp)
    for seed in (0, 12345678901234567):
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"), seed)
        plan, t = planners.mcts_plan_wavefront(oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], term), episodes, horizon,
                                               0.9, 10.0, width, seed)
        got_plan, res = eng.finish()
        d = eng.tree_dict()
        assert d["parent"].tolist() == t.parent
        assert d["first_child"].tolist() == t.first_child and d["n_children"].tolist() == t.n_children
        used = np.array(t.parent) != -2
        assert d["action"][used].tolist() == np.array(t.action)[used].tolist()
        assert d["count"].tolist() == t.count
        assert d["vsum"].tolist() == t.vsum
        assert np.array_equal(d["value"], np.array(t.value))
        assert got_plan == plan
        assert int(res[2]) == t.env_steps and int(res[3]) == -(-episodes // width)


@pytest.mark.parametrize("width,episodes,horizon,scene", [(16, 64, 6, 0), (5, 40, 8, 3), (200, 150, 5, 7)])
def test_mcts_wave_highway_matches_the_specification(width, episodes, horizon, scene):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.mcts import MCTSWaveEngine
    eng = MCTSWaveEngine(_lib.ENV_HIGHWAY, 5, episodes, horizon, 0.8, 10.0, width)
    eng.plan(torch.tensor(oenvs.make_highway_state(scene).pack(), dtype=torch.int32, device="cuda"), 3)
    plan, t = planners.mcts_plan_wavefront(oenvs.HighwayLite(seed=scene), episodes, horizon, 0.8, 10.0, width, 3)
    got_plan, res = eng.finish()
    d = eng.tree_dict()
    assert d["parent"].tolist() == t.parent and d["count"].tolist() == t.count and d["vsum"].tolist() == t.vsum
    used = np.array(t.parent) != -2
    assert d["action"][used].tolist() == np.array(t.action)[used].tolist()
    assert got_plan == plan and int(res[2]) == t.env_steps


@pytest.mark.parametrize("width", [256, 512])
def test_mcts_wave_highway_c3_full_size_vs_c_specification(width):
    """BASELINE C3: 4096 episodes x horizon 20 as ONE decision."""
    import torch
    from oracle import c_oracle
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.mcts import MCTSWaveEngine
    eng = MCTSWaveEngine(_lib.ENV_HIGHWAY, 5, 4096, 20, 0.8, 10.0, width)
    for scene, seed in ((0, 0), (4, 99)):
        words = oenvs.make_highway_state(scene).pack()
        eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"), seed)
        c = c_oracle.mcts_plan_wave(words, 4096, 20, 0.8, 10.0, width, seed)
        got_plan, res = eng.finish()
        d = eng.tree_dict()
        for k in ("parent", "first_child", "n_children", "count", "vsum"):
            assert np.array_equal(d[k], c[k]), k
        assert np.array_equal(d["value"], c["value"])
        used = c["parent"] != -2
        assert np.array_equal(d["action"][used], c["action"][used])
        assert int(res[2]) == c["env_steps"] and d["count"][0] == 4096


def test_mcts_agent_wavefront_option():
    """`"wavefront": W` in the MCTSAgent config: same plan as the specification seeded from the planner RNG."""
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    from rl_agents_b200.envs import HighwayLiteEnv
    env = HighwayLiteEnv(seed=2)
    agent = MCTSAgent(env, {"episodes": 96, "horizon": 6, "gamma": 0.8, "wavefront": 32})
    agent.seed(5)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(5)))
    seed = int(rng.integers(0, 2 ** 63 - 1))
    plan, _ = planners.mcts_plan_wavefront(oenvs.HighwayLite(seed=2), 96, 6, 0.8, 10.0, 32, seed)
    assert agent.plan(env.observation()) == plan
    assert agent.planner.root_statistics["counts"].sum() == 96 - 32      # the first wave only expands the root
