"""GPU parity tests of the speculative strict OPD (b2_opd_plan_spec): whatever the candidate width, the tree is
the reference's own strict best-first tree (oracle.planners.opd_plan, pinned to the reference's golden trees),
bit for bit -- node ids, counts, fp64 bounds, plan."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.test_gpu_wave import check, finite, np_random, M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("width", [1, 2, 7, 64, 256])
def test_spec_finite_is_the_strict_tree(width):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    mdp, make = finite()
    eng = OPDSpeculativeEngine(_lib.ENV_FINITE, 5, 1500, 0.9, width, mdp=mdp)
    eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
    plan, tree = planners.opd_plan(make(), 1500, 0.9, np_random=np_random(0))
    res = check(eng, plan, tree)
    assert 1 <= int(res[0, 7]) <= 300
    if width == 1:
        assert int(res[0, 7]) == 300


def test_spec_finite_terminal_states_and_terminal_reward():
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    term = M["large1_term"].copy()
    term[[3, 17, 66, 91]] = True
    mdp, make = finite(terminal=term)
    for width, tr in ((4, 0.3), (32, -0.7), (128, 0.1)):
        eng = OPDSpeculativeEngine(_lib.ENV_FINITE, 5, 400, 0.8, width, terminal_reward=tr, mdp=mdp)
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        plan, tree = planners.opd_plan(make(), 400, 0.8, terminal_reward=tr, np_random=np_random(0))
        res = check(eng, plan, tree)
        assert int(res[0, 3]) == tree.terminal_expansions


def test_spec_finite_many_exact_ties():
    """All rewards equal / reward 1: children tie with their parents and with each other; the strict order takes
    the lowest node id, so a tying child never overtakes a candidate."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    rng = np.random.default_rng(3)
    T = rng.integers(0, 30, size=(30, 4)).astype(np.int32)
    R = np.full((30, 4), 0.5)
    R[:, 2] = 1.0
    term = np.zeros(30, bool)
    for width in (1, 3, 50, 256):
        eng = OPDSpeculativeEngine(_lib.ENV_FINITE, 4, 600, 0.8, width, mdp=FiniteMDP("deterministic", T, R, term))
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        plan, tree = planners.opd_plan(oenvs.FiniteMDPLite(T, R, term), 600, 0.8, np_random=np_random(0))
        check(eng, plan, tree)


def test_spec_reward_out_of_range_raises_only_when_the_strict_search_reaches_it():
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    R = M["large1_R"].copy()
    R[5, 2] = 1.5
    mdp = FiniteMDP("deterministic", M["large1_T"], R, M["large1_term"])
    make = lambda: oenvs.FiniteMDPLite(M["large1_T"], R, M["large1_term"])
    for budget in (20, 60, 200, 2000):
        try:
            plan, tree = planners.opd_plan(make(), budget, 0.9, np_random=np_random(0))
            raised = False
        except ValueError:
            raised = True
        eng = OPDSpeculativeEngine(_lib.ENV_FINITE, 5, budget, 0.9, 64, mdp=mdp)
        eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
        if raised:
            with pytest.raises(ValueError):
                eng.finish([np_random(0)])
        else:
            check(eng, plan, tree)


@pytest.mark.parametrize("width,seed,budget,gamma", [(1, 0, 200, 0.8), (16, 1, 400, 0.8), (64, 2, 1000, 0.8),
                                                     (64, 7, 3125, 0.8), (256, 3, 3125, 0.95)])
def test_spec_highway_is_the_strict_tree(width, seed, budget, gamma):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    words = oenvs.make_highway_state(seed).pack()
    eng = OPDSpeculativeEngine(_lib.ENV_HIGHWAY, 5, budget, gamma, width)
    eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"))
    plan, tree = planners.opd_plan(oenvs.HighwayLite(seed=seed), budget, gamma, np_random=np_random(0))
    check(eng, plan, tree)


def test_spec_highway_equals_the_batch_kernel():
    """Same decision through b2_opd_plan (one tree per group, strict) and b2_opd_plan_spec."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDEngine, OPDSpeculativeEngine
    words = torch.tensor(oenvs.make_highway_state(11).pack(), dtype=torch.int32, device="cuda")
    a = OPDEngine(_lib.ENV_HIGHWAY, 1, 5, 3125, 0.8)
    a.plan(words.reshape(1, -1).contiguous())
    pa, ra = a.finish([np_random(0)])
    b = OPDSpeculativeEngine(_lib.ENV_HIGHWAY, 5, 3125, 0.8, 64)
    b.plan(words)
    pb, rb = b.finish([np_random(0)])
    da, db = a.tree_dict(0), b.tree_dict(0)
    assert pa == pb
    for k in da:
        assert np.array_equal(da[k], db[k]), k


@pytest.mark.parametrize("width,seed", [(16, 0), (64, 5)])
def test_spec_intersection_is_the_strict_tree(width, seed):
    import torch
    from oracle import intersection as oint
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    env = oint.IntersectionLite(seed=seed)
    words = env.state.pack()
    eng = OPDSpeculativeEngine(_lib.ENV_INTERSECTION, 3, 600, 0.9, width)
    eng.plan(torch.tensor(words, dtype=torch.int32, device="cuda"))
    plan, tree = planners.opd_plan(oint.IntersectionLite(seed=seed), 600, 0.9, np_random=np_random(0))
    check(eng, plan, tree)


def test_spec_rejects_what_it_cannot_hold():
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDSpeculativeEngine
    with pytest.raises(_lib.B2Error):
        OPDSpeculativeEngine(_lib.ENV_HIGHWAY, 5, 200000, 0.8, 64)      # tree larger than one shared-memory tile
    with pytest.raises(_lib.B2Error):
        OPDSpeculativeEngine(_lib.ENV_HIGHWAY, 5, 1000, 0.8, 512)       # more candidates than threads


def test_agent_picks_the_speculative_kernel_where_it_pays_and_plans_do_not_change():
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.engine.opd import OPDEngine, OPDSpeculativeEngine, OPDWaveEngine
    from rl_agents_b200.envs import HighwayLiteEnv, IntersectionLiteEnv
    from tests.util import load_golden
    g = load_golden("golden_highway.json")["opd"]["s0_b75_g0.7"]
    plans = {}
    for spec in ("auto", 0, 16):
        agent = DeterministicPlannerAgent(HighwayLiteEnv(seed=0), {"budget": 75, "gamma": 0.7, "speculative": spec})
        agent.seed(0)
        plans[spec] = agent.plan(None)
        kind = type(agent.planner.engine)
        assert kind is (OPDEngine if spec == 0 else OPDSpeculativeEngine)
    assert plans["auto"] == plans[0] == plans[16] == g["plan"]
    # a discount above 0.9 keeps the one-CTA kernel; a tree beyond one tile as well; "wavefront" wins over auto
    for cfg, kind in (({"budget": 200, "gamma": 0.95}, OPDEngine), ({"budget": 200000, "gamma": 0.8}, OPDEngine),
                      ({"budget": 200, "gamma": 0.8, "wavefront": 8}, OPDWaveEngine)):
        agent = DeterministicPlannerAgent(HighwayLiteEnv(seed=1), cfg)
        agent.seed(0)
        agent.plan(None)
        assert type(agent.planner.engine) is kind
    agent = DeterministicPlannerAgent(IntersectionLiteEnv(seed=2), {"budget": 200, "gamma": 0.9})
    agent.seed(0)
    from oracle import intersection as oint
    plan, _ = planners.opd_plan(oint.IntersectionLite(seed=2), 200, 0.9, np_random=np_random(0))
    assert agent.plan(None) == plan and type(agent.planner.engine) is OPDSpeculativeEngine
    with pytest.raises(ValueError):
        DeterministicPlannerAgent(HighwayLiteEnv(seed=1), {"budget": 200000, "speculative": 64}).plan(None)
