"""GPU parity tests of IntersectionLite (BASELINE C5's env model): the CUDA transition and the wavefront OPD on it
against the numpy statement of the same spec (oracle/intersection.py), bit for bit."""
import numpy as np
import pytest

from oracle import intersection as oit
from oracle import planners

pytestmark = pytest.mark.gpu


def np_random(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def test_intersection_step_batched_vs_oracle():
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.envs.intersection_lite import make_scene
    lib = _lib.load()
    n = 29
    envs = [oit.IntersectionLite(seed=200 + i) for i in range(n)]
    for i in range(n):
        assert make_scene(200 + i).tolist() == envs[i].state.pack().tolist()
    st = torch.tensor(np.stack([e.state.pack() for e in envs]), dtype=torch.int32, device="cuda")
    rew = torch.empty(n, dtype=torch.float32, device="cuda")
    flg = torch.empty(n, dtype=torch.int32, device="cuda")
    avail = torch.empty(n, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(1)
    crashes = arrivals = 0
    for step in range(13):
        acts = []
        for e in envs:
            av = e.get_available_actions()
            acts.append(int(av[rng.integers(len(av))]))
        _lib.check(lib.b2_intersection_step(_lib.ptr(st), _lib.ptr(torch.tensor(acts, dtype=torch.int32, device="cuda")),
                                            _lib.ptr(rew), _lib.ptr(flg), _lib.ptr(avail), n, _lib.current_stream()))
        got = st.cpu().numpy()
        for i, e in enumerate(envs):
            _, r, term, trunc, _ = e.step(acts[i])
            assert got[i].tolist() == e.state.pack().tolist(), (step, i)
            assert float(rew[i].item()) == np.float32(r)
            assert int(flg[i].item()) == (1 if term else 0) | (2 if trunc else 0)
            mask = int(avail[i].item())
            assert sorted(a for a in range(3) if mask >> a & 1) == sorted(e.get_available_actions())
        crashes += sum(int(e.state.flags[0] & 2 != 0) for e in envs)
        arrivals += sum(int(e.state.arrived) for e in envs)
    assert crashes > 0          # the random policy does crash in some scenes: the collision path is exercised


@pytest.mark.parametrize("width,seed,budget", [(1, 0, 240), (1, 3, 150), (16, 1, 300), (64, 5, 600)])
def test_intersection_wave_opd_matches_the_specification(width, seed, budget):
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDWaveEngine
    eng = OPDWaveEngine(_lib.ENV_INTERSECTION, 3, budget, 0.9, width)
    eng.plan(torch.tensor(oit.make_intersection_state(seed).pack(), dtype=torch.int32, device="cuda"))
    plans, res = eng.finish([np_random(0)])
    plan, t = planners.opd_plan_wavefront(oit.IntersectionLite(seed=seed), budget, 0.9, width, np_random=np_random(0))
    d = eng.tree_dict(0)
    assert plans[0] == plan
    assert d["parent"].tolist() == t.parent and d["action"].tolist() == t.action and d["count"].tolist() == t.count
    assert np.array_equal(d["lower"], np.array(t.lower)) and np.array_equal(d["upper"], np.array(t.upper))
    assert d["done"].tolist() == [bool(x) for x in t.done]
    if width == 1:
        plan1, t1 = planners.opd_plan(oit.IntersectionLite(seed=seed), budget, 0.9, np_random=np_random(0))
        assert plan1 == plan and t1.parent == t.parent and t1.upper == t.upper


def test_intersection_agent_and_c5_sized_decision():
    """The plugin surface on IntersectionLite, and one C5-sized decision (budget 1e6 = 333 333 expansions)
    searched by the whole GPU: structural invariants of the tree."""
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.engine.opd import OPDWaveEngine
    from rl_agents_b200.envs import IntersectionLiteEnv
    env = IntersectionLiteEnv(seed=2)
    agent = DeterministicPlannerAgent(env, {"budget": 200, "gamma": 0.9, "env_preprocessors": [{"method": "simplify"}]})
    agent.seed(0)
    plan, _ = planners.opd_plan(oit.IntersectionLite(seed=2), 200, 0.9, np_random=np_random(0))     # the shipped budget
    assert agent.plan(env.observation()) == plan
    eng = OPDWaveEngine(_lib.ENV_INTERSECTION, 3, 1000000, 0.9, 2048)
    eng.plan(torch.tensor(oit.make_intersection_state(0).pack(), dtype=torch.int32, device="cuda"))
    plans, res = eng.finish([np_random(0)])
    n = int(res[0, 0])
    d = eng.tree_dict(0)
    assert (d["n_children"] > 0).sum() == 1000000 // 3
    assert d["count"][0] == n and (d["parent"][1:] < np.arange(1, n)).all()
    assert (d["lower"] <= d["upper"] + 1e-12).all() and d["upper"][0] <= 1 / (1 - 0.9) + 1e-9
    kids = d["parent"][1:]
    assert np.array_equal(np.bincount(kids, minlength=n)[:n], d["n_children"])
