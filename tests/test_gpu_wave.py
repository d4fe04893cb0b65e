"""GPU parity tests of the wavefront OPD (b2_opd_plan_wave): bit-exact against its specification
(oracle/planners.py::opd_plan_wavefront and the same statement in C), width 1 against the reference's
strict algorithm (oracle.planners.opd_plan, pinned to the reference's golden trees)."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_mdps

pytestmark = pytest.mark.gpu
M = load_mdps()


def np_random(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def finite(name="large1", terminal=None):
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    term = M[name + "_term"] if terminal is None else terminal
    return FiniteMDP("deterministic", M[name + "_T"], M[name + "_R"], term), \
        lambda: oenvs.FiniteMDPLite(M[name + "_T"], M[name + "_R"], term)


def check(eng, plan, tree):
    plans, res = eng.finish([np_random(0)])
    d = eng.tree_dict(0)
    assert d["parent"].tolist() == list(tree.parent)
    assert d["action"].tolist() == list(tree.action)
    assert d["depth"].tolist() == list(tree.depth)
    assert d["count"].tolist() == list(tree.count)
    assert d["first_child"].tolist() == list(tree.first_child)
    assert d["n_children"].tolist() == list(tree.n_children)
    assert d["done"].tolist() == [bool(x) for x in tree.done]
    assert np.array_equal(d["reward"], np.array(tree.reward))
    assert np.array_equal(d["lower"], np.array(tree.lower))
    assert np.array_equal(d["upper"], np.array(tree.upper))
    assert plans[0] == plan
    assert int(res[0, 1]) == tree.n_leaves
    return res


@pytest.mark.parametrize("width", [1, 2, 5, 16, 64, 300, 5000])
def test_wave_finite_matches_the_specification(width):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    mdp, make = finite()
    eng = OPDWaveEngine(_lib.ENV_FINITE, 5, 500, 0.9, width, mdp=mdp)
    eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
    plan, tree = planners.opd_plan_wavefront(make(), 500, 0.9, width, np_random=np_random(0))
    res = check(eng, plan, tree)
    assert int(res[0, 7]) == len(tree.waves)
    if width == 1:      # the reference's own algorithm
        plan1, t1 = planners.opd_plan(make(), 500, 0.9, np_random=np_random(0))
        assert plan1 == plan and t1.parent == tree.parent and t1.count == tree.count and t1.upper == tree.upper


def test_wave_finite_terminal_states_and_terminal_reward():
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    term = M["large1_term"].copy()
    term[[3, 17, 66, 91]] = True
    mdp, make = finite(terminal=term)
    for width, tr in ((1, 0.3), (8, 0.3), (32, -0.7), (8, 0.1)):
        eng = OPDWaveEngine(_lib.ENV_FINITE, 5, 400, 0.8, width, terminal_reward=tr, mdp=mdp)
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        plan, tree = planners.opd_plan_wavefront(make(), 400, 0.8, width, terminal_reward=tr, np_random=np_random(0))
        res = check(eng, plan, tree)
        assert int(res[0, 3]) == tree.terminal_expansions


def test_wave_finite_many_exact_ties():
    """All rewards equal: every frontier key of a depth ties; the wave must take the lowest node ids."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    rng = np.random.default_rng(3)
    T = rng.integers(0, 30, size=(30, 4)).astype(np.int32)
    R = np.full((30, 4), 0.5)
    R[:, 2] = 1.0               # reward 1: children tie with their parent's key as well
    term = np.zeros(30, bool)
    for width in (1, 3, 7, 50):
        eng = OPDWaveEngine(_lib.ENV_FINITE, 4, 600, 0.8, width, mdp=FiniteMDP("deterministic", T, R, term))
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        plan, tree = planners.opd_plan_wavefront(oenvs.FiniteMDPLite(T, R, term), 600, 0.8, width, np_random=np_random(0))
        check(eng, plan, tree)


def test_wave_finite_reward_out_of_range_raises():
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    R = M["large1_R"].copy()
    R[5, 2] = 1.5
    eng = OPDWaveEngine(_lib.ENV_FINITE, 5, 2000, 0.9, 16, mdp=FiniteMDP("deterministic", M["large1_T"], R, M["large1_term"]))
    eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        eng.finish([np_random(0)])


@pytest.mark.parametrize("width,seed", [(1, 0), (4, 1), (32, 2), (32, 7)])
def test_wave_highway_matches_the_specification(width, seed):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    words = oenvs.make_highway_state(seed).pack()
    eng = OPDWaveEngine(_lib.ENV_HIGHWAY, 5, 400, 0.8, width)
    eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"))
    plan, tree = planners.opd_plan_wavefront(oenvs.HighwayLite(seed=seed), 400, 0.8, width, np_random=np_random(0))
    check(eng, plan, tree)


def c_check(eng, c):
    plans, res = eng.finish([np_random(0)])
    d = eng.tree_dict(0)
    for k in ("parent", "action", "depth", "count", "first_child", "n_children", "reward", "lower", "upper"):
        assert np.array_equal(d[k], c[k]), k
    assert np.array_equal(d["done"], c["done"].astype(bool))
    assert int(res[0, 1]) == c["n_leaves"]
    return res


@pytest.mark.parametrize("width", [1, 64, 128])
def test_wave_highway_c2_full_size_vs_c_specification(width):
    """BASELINE C2 (budget 10 000, gamma 0.8) as ONE decision."""
    import torch
    from oracle import c_oracle
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    eng = OPDWaveEngine(_lib.ENV_HIGHWAY, 5, 10000, 0.8, width)
    for seed in (0, 5):
        words = oenvs.make_highway_state(seed).pack()
        eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"))
        c = c_oracle.opd_plan_wave(words, 10000, 0.8, width)
        res = c_check(eng, c)
        assert int(res[0, 7]) == c["n_waves"]
        if width == 1:
            s = c_oracle.opd_plan(words, 10000, 0.8)
            assert np.array_equal(s["parent"], c["parent"]) and np.array_equal(s["upper"], c["upper"])


def test_wave_highway_large_budget_multi_tile_vs_c_specification():
    """200 000 child nodes: the frontier no longer fits one shared-memory tile."""
    import torch
    from oracle import c_oracle
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    eng = OPDWaveEngine(_lib.ENV_HIGHWAY, 5, 200000, 0.8, 1024)
    words = oenvs.make_highway_state(2).pack()
    eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"))
    c = c_oracle.opd_plan_wave(words, 200000, 0.8, 1024)
    res = c_check(eng, c)
    assert int(res[0, 7]) == c["n_waves"]


def test_agent_wavefront_option_matches_the_specification_and_reuses_tables():
    """`"wavefront": K` in the agent config: same plan as the specification; a copying preprocessor (a new
    but identical mdp object per decision) does not rebuild the device tables."""
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.envs import FiniteMDPEnv, HighwayLiteEnv
    env = FiniteMDPEnv(M["large1_T"], M["large1_R"], M["large1_term"])
    agent = DeterministicPlannerAgent(env, {"budget": 500, "gamma": 0.9, "wavefront": 16})
    agent.seed(0)
    plan, _ = planners.opd_plan_wavefront(oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"]), 500, 0.9,
                                          16, np_random=np_random(0))
    assert agent.plan(None) == plan
    eng = agent.planner.engine
    env2 = FiniteMDPEnv(M["large1_T"].copy(), M["large1_R"].copy(), M["large1_term"].copy())
    agent.env = env2
    agent.plan(None)
    assert agent.planner.engine is eng
    env3 = FiniteMDPEnv(M["large1_T"].copy(), M["large1_R"] * 0.5, M["large1_term"].copy())
    agent.env = env3
    agent.plan(None)
    assert agent.planner.engine is not eng
    # HighwayLite through the plugin surface
    henv = HighwayLiteEnv(seed=4)
    agent = DeterministicPlannerAgent(henv, {"budget": 400, "gamma": 0.8, "wavefront": 32,
                                             "env_preprocessors": [{"method": "simplify"}]})
    agent.seed(0)
    plan, _ = planners.opd_plan_wavefront(oenvs.HighwayLite(seed=4), 400, 0.8, 32, np_random=np_random(0))
    assert agent.plan(henv.observation()) == plan


def test_wave_finite_large_tree_distributed_selection_with_ties():
    """A tree larger than one shared-memory tile: the selection runs on all CTAs (id slices, global
    reductions); many exact ties at the threshold must still resolve to the lowest node ids."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    rng = np.random.default_rng(5)
    T = rng.integers(0, 40, size=(40, 4)).astype(np.int32)
    R = rng.choice([0.25, 0.5, 1.0], size=(40, 4))
    term = np.zeros(40, bool)
    for width in (700, 4096):
        eng = OPDWaveEngine(_lib.ENV_FINITE, 4, 120000, 0.9, width, mdp=FiniteMDP("deterministic", T, R, term))
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        plan, tree = planners.opd_plan_wavefront(oenvs.FiniteMDPLite(T, R, term), 120000, 0.9, width, np_random=np_random(0))
        res = check(eng, plan, tree)
        assert int(res[0, 7]) == len(tree.waves)


# ------------------------------------------------------------------ DROP (robust.py) ----
def _models(names):
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    termx = M["large1_term"].copy()
    termx[[3, 17, 66, 91]] = True
    table = {"large1": ("large1", M["large1_term"]), "large1t": ("large1", termx), "large2": ("large2", M["large2_term"])}
    prod = [FiniteMDP("deterministic", M[table[n][0] + "_T"], M[table[n][0] + "_R"], table[n][1]) for n in names]
    orac = [oenvs.FiniteMDPLite(M[table[n][0] + "_T"], M[table[n][0] + "_R"], table[n][1]) for n in names]
    return prod, orac


def test_drop_finite_matches_the_reference_goldens():
    """DiscreteRobustPlanner (robust.py:28-47) on joint envs of 2 and 3 finite-MDP models: plan, node order,
    counts and robust bounds equal the UNMODIFIED reference's (tests/golden, 'drop')."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    from tests.util import load_golden
    G = load_golden("golden_finite.json")
    for key, g in G["drop"].items():
        prod, _ = _models(g["models"])
        eng = OPDWaveEngine(_lib.ENV_FINITE, 5, g["budget"], g["gamma"], 1, terminal_reward=g["terminal_reward"],
                            n_models=len(prod), model_mdps=prod)
        eng.plan(torch.zeros(len(prod), dtype=torch.int32, device="cuda"))
        plans, _ = eng.finish([np_random(0)])
        d = eng.tree_dict(0)
        t = g["tree"]
        assert plans[0] == g["plan"], key
        assert d["parent"].tolist() == t["parent"] and d["action"].tolist() == t["action"]
        assert d["count"].tolist() == t["count"]
        assert np.array_equal(d["lower"], np.array(t["lower"])) and np.array_equal(d["upper"], np.array(t["upper"]))


@pytest.mark.parametrize("width", [1, 8])
def test_drop_highway_matches_the_oracle(width):
    """Joint env of three HighwayLite models (different assumed traffic) against oracle.planners.robust_plan
    (itself pinned to the reference on the finite goldens)."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    states = []
    for off in (0.0, -3.0, 2.0):
        st = oenvs.make_highway_state(5)
        st.tgt_speed[1:] = (st.tgt_speed[1:] + np.float32(off)).astype(np.float32)
        states.append(st)
    eng = OPDWaveEngine(_lib.ENV_HIGHWAY, 5, 300, 0.8, width, n_models=3)
    eng.plan(torch.tensor(np.stack([s.pack() for s in states]), dtype=torch.int32, device="cuda"))
    plans, _ = eng.finish([np_random(0)])
    plan, t = planners.robust_plan([oenvs.HighwayLite(s.copy()) for s in states], 300, 0.8, np_random=np_random(0),
                                   width=width)
    d = eng.tree_dict(0)
    assert plans[0] == plan
    assert d["parent"].tolist() == t.parent and d["action"].tolist() == t.action and d["count"].tolist() == t.count
    assert np.array_equal(d["lower"], np.array(t.lower)) and np.array_equal(d["upper"], np.array(t.upper))


def test_drop_agent_plugin_surface():
    """DiscreteRobustPlannerAgent: `models` = one preprocessor chain per model, like the reference's configs."""
    from rl_agents_b200.agents.robust.robust import DiscreteRobustPlannerAgent
    from rl_agents_b200.envs import HighwayLiteEnv
    env = HighwayLiteEnv(seed=5)
    models = [[{"method": "simplify"}],
              [{"method": "assume_traffic", "args": {"target_speed_offset": -3.0}}],
              [{"method": "assume_traffic", "args": {"target_speed_offset": 2.0}}]]
    agent = DiscreteRobustPlannerAgent(env, {"budget": 300, "gamma": 0.8, "models": models})
    agent.seed(0)
    states = []
    for off in (0.0, -3.0, 2.0):
        st = oenvs.make_highway_state(5)
        st.tgt_speed[1:] = (st.tgt_speed[1:] + np.float32(off)).astype(np.float32)
        states.append(oenvs.HighwayLite(st))
    plan, _ = planners.robust_plan(states, 300, 0.8, np_random=np_random(0))
    assert agent.plan(env.observation()) == plan
    assert agent.config["models"] == models
