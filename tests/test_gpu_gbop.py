"""GPU parity tests of GBOP-T (b2_gbop_plan): the device planner against the reference's StateAwarePlanner
(golden vectors from the unmodified reference) and the oracle restatement -- plan, node order, leaves left
after pruning, state value table, bit for bit."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_golden, load_mdps

pytestmark = pytest.mark.gpu
G = load_golden("golden_finite.json")
M = load_mdps()


def np_random(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def mdp(name="large1", terminal=None):
    from rl_agents_b200.envs.finite_mdp import FiniteMDP
    term = M[name + "_term"] if terminal is None else terminal
    return FiniteMDP("deterministic", M[name + "_T"], M[name + "_R"], term)


@pytest.mark.parametrize("key", sorted(G["gbopt"]))
def test_gbopt_matches_the_reference_goldens(key):
    import torch
    from rl_agents_b200.engine.gbop import GBOPEngine
    g = G["gbopt"][key]
    eng = GBOPEngine(1, 5, g["budget"], g["gamma"], mdp())
    eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"))
    plans, res = eng.finish([np_random(g["seed"])])
    assert plans[0] == g["plan"]
    d = eng.tree_dict(0)
    sv = eng.state_values(0)
    assert all(sv[int(k)] == v for k, v in g["state_values"].items())
    leaves = np.nonzero(d["leaf"])[0]
    assert len(leaves) == g["n_leaves"] == int(res[0, 1]) and len(set(d["obs"].tolist())) == g["n_states"]
    assert int(d["depth"][leaves].sum()) == g["leaf_depth_sum"]
    assert sum(float(d["lower"][l]) for l in leaves) == g["leaf_lower_sum"]      # python floats: 3.12's sum() compensates


@pytest.mark.parametrize("variant", [dict(), dict(backup_aggregated_nodes=False), dict(prune_suboptimal_leaves=False),
                                     dict(accuracy=0.05), dict(terminal_reward=0.3, terminal=True)])
def test_gbopt_batch_vs_oracle(variant):
    """A batch of roots, the config switches of state_aware.py:79-86 and terminal states, against the oracle."""
    import torch
    from rl_agents_b200.engine.gbop import GBOPEngine
    variant = dict(variant)
    term = None
    if variant.pop("terminal", False):
        term = M["large1_term"].copy()
        term[[3, 17, 66, 91]] = True
    tr = variant.pop("terminal_reward", 0.0)
    roots = [0, 7, 42, 99, 3]
    eng = GBOPEngine(len(roots), 5, 400, 0.85, mdp(terminal=term), terminal_reward=tr, **variant)
    eng.plan(torch.tensor(roots, dtype=torch.int32, device="cuda"))
    plans, res = eng.finish([np_random(1) for _ in roots])
    for i, r in enumerate(roots):
        env = oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"] if term is None else term, state=r)
        plan, t, state_values, leaves = planners.state_aware_plan(env, r, 400, 0.85, np_random(1), terminal_reward=tr,
                                                                  **variant)
        d = eng.tree_dict(i)
        assert plans[i] == plan, (i, r)
        assert d["parent"].tolist() == t.parent and d["action"].tolist() == t.action and d["count"].tolist() == t.count
        assert d["obs"].tolist() == t.obs and np.array_equal(d["lower"], np.array(t.lower))
        assert sorted(np.nonzero(d["leaf"])[0].tolist()) == sorted(leaves)
        sv = eng.state_values(i)
        default = 1 / (1 - 0.85)
        assert all(sv[s] == state_values.get(s, default) for s in range(100))


def test_gbopt_agent_plugin_surface():
    from rl_agents_b200.agents.tree_search.state_aware import StateAwarePlannerAgent
    from rl_agents_b200.envs import FiniteMDPEnv
    g = G["gbopt"]["large1_b500_g0.9"]
    agent = StateAwarePlannerAgent(FiniteMDPEnv(M["large1_T"], M["large1_R"], M["large1_term"]),
                                   {"budget": g["budget"], "gamma": g["gamma"]})
    agent.seed(g["seed"])
    assert agent.plan(0) == g["plan"]
    assert agent.config["prune_suboptimal_leaves"] is True and agent.config["accuracy"] == 0


# ------------------------------------------------------------------ GBOP-D (graph_based.py) ----
@pytest.mark.parametrize("key", sorted(G["gbopd"]))
def test_gbopd_matches_the_reference_and_the_oracle(key):
    """accuracy = 0: plan, node set and both bounds of every node equal the UNMODIFIED reference's (the fixed
    point does not depend on the order in which the reference's parent SETS are iterated); default accuracy: equal
    to the oracle restatement (parents in ascending state id), the reference's plan, its bounds within `accuracy`."""
    import torch
    from rl_agents_b200.engine.gbop import GBOPDEngine
    from rl_agents_b200.engine.mcts import pcg64_words
    g = G["gbopd"][key]
    eng = GBOPDEngine(1, 5, g["budget"], g["gamma"], mdp(), g["accuracy"], g["sampling_timeout"])
    rng = np_random(g["seed"])
    eng.plan(torch.tensor([0], dtype=torch.int32, device="cuda"), pcg64_words(rng).reshape(1, -1))
    plans, res, words = eng.finish()
    nodes = eng.nodes(0)
    ref = {int(k): v for k, v in g["nodes"].items()}
    assert plans[0] == g["plan"] and set(nodes) == set(ref)
    assert all(nodes[s]["expanded"] == ref[s][2] for s in ref)
    if g["accuracy"] == 0:
        assert all(nodes[s]["lower"] == ref[s][0] and nodes[s]["upper"] == ref[s][1] for s in ref)
    else:
        assert all(abs(nodes[s]["lower"] - ref[s][0]) <= 10 * g["accuracy"] and
                   abs(nodes[s]["upper"] - ref[s][1]) <= 10 * g["accuracy"] for s in ref)
    orng = np_random(g["seed"])
    env = oenvs.LegacyStepEnv(oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"]))
    plan, onodes = planners.graph_based_plan(env, 0, g["budget"], g["gamma"], orng, g["accuracy"], g["sampling_timeout"])
    assert plans[0] == plan and nodes == onodes
    from rl_agents_b200.engine.mcts import set_pcg64_words
    set_pcg64_words(rng, words[0])
    assert rng.bit_generator.state["state"] == orng.bit_generator.state["state"]     # same RNG consumption


def test_gbopd_agent_plugin_surface():
    from rl_agents_b200.agents.tree_search.graph_based import GraphBasedPlannerAgent
    from rl_agents_b200.envs import FiniteMDPEnv
    g = G["gbopd"]["large1_b500_g0.9_default"]
    agent = GraphBasedPlannerAgent(FiniteMDPEnv(M["large1_T"], M["large1_R"], M["large1_term"]),
                                   {"budget": g["budget"], "gamma": g["gamma"]})
    agent.seed(g["seed"])
    assert agent.plan(0) == g["plan"]
    assert agent.config["accuracy"] == 1e-2 and agent.config["sampling_timeout"] == 100
