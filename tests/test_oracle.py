"""Pin the oracle restatements (oracle/planners.py, oracle/envs.py) against
golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known answers
(tests/agents/test_utils.py:28-31 of the reference)."""
import numpy as np
import pytest

from oracle import envs, planners, ref_loader
from tests.util import assert_tree_matches, load_golden, load_mdps

G = load_golden("golden_finite.json")
H = load_golden("golden_highway.json")
M = load_mdps()


def np_random(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def finite(name="large1", terminal=None):
    term = M[name + "_term"] if terminal is None else terminal
    return envs.FiniteMDPLite(M[name + "_T"], M[name + "_R"], term)


def tree_dict(t, fields):
    d = {"parent": t.parent, "action": t.action, "count": t.count}
    for f in fields:
        d[f] = getattr(t, f)
    return d


@pytest.mark.parametrize("key,mdp", [("large1_b500_g0.9", "large1"), ("large1_b75_g0.7", "large1"),
                                     ("large1_b10000_g0.9", "large1"), ("large2_b2000_g0.8", "large2")])
def test_opd_finite(key, mdp):
    g = G["opd"][key]
    plan, t = planners.opd_plan(finite(mdp), g["budget"], g["gamma"], np_random=np_random(g["seed"]))
    assert plan == g["plan"]
    assert t.n_leaves == g["n_leaves"]
    assert_tree_matches(tree_dict(t, ["reward", "lower", "upper", "done"]), g["tree"],
                        ["reward", "lower", "upper"])


def test_opd_finite_terminal():
    g = G["opd"]["large1_terminal_b300_g0.85"]
    term = M["large1_term"].copy()
    term[[3, 17, 66, 91]] = True
    plan, t = planners.opd_plan(finite(terminal=term), 300, 0.85, np_random=np_random(0))
    assert plan == g["plan"]
    assert any(t.done)
    assert_tree_matches(tree_dict(t, ["reward", "lower", "upper", "done"]), g["tree"],
                        ["reward", "lower", "upper"])


@pytest.mark.parametrize("key", sorted(G["mcts"]))
def test_mcts_finite(key):
    g = G["mcts"][key]
    term = None
    if "terminal" in key:
        term = M["large1_term"].copy()
        term[[3, 17, 66, 91]] = True
    plan, t = planners.mcts_plan(finite(terminal=term), g["episodes"], g["horizon"], g["config"]["gamma"],
                                 g["temperature"], np_random(g["seed"]))
    assert plan == g["plan"]
    assert_tree_matches(tree_dict(t, ["value", "prior"]), g["tree"], ["value", "prior"])


def test_mcts_subtree_reuse_matches_reference():
    """step_strategy "subtree" (abstract.py:172-206): three consecutive decisions, the tree re-rooted at the
    executed action in between; compared in an id-independent breadth-first form."""
    from tests.util import canonical_tree
    g = G["mcts_subtree"]
    env = finite()
    rng = np_random(4)
    tree = None
    for k in range(3):
        assert env.mdp.state == g["states"][k]
        plan, tree = planners.mcts_plan(env, g["episodes"], g["horizon"], 0.85, g["temperature"], rng, tree=tree)
        assert plan == g["plans"][k]
        got = canonical_tree(tree.first_child, tree.n_children, [tree.count, tree.value, tree.prior])
        assert got == g["trees"][k], k
        env.step(plan[0])
        tree = planners.mcts_reroot(tree, plan[0])


def test_reroot_arrays_equals_oracle_reroot():
    from rl_agents_b200.engine.mcts import reroot_arrays
    from tests.util import canonical_tree
    _, t = planners.mcts_plan(finite(), 44, 9, 0.8, 10.0, np_random(1))
    meta = np.array([(a & 0xff) | (n << 8) for a, n in zip(t.action, t.n_children)], dtype=np.int32)
    arrays = {"parent": np.array(t.parent, dtype=np.int32), "first_child": np.array(t.first_child, dtype=np.int32),
              "count": np.array(t.count, dtype=np.int32), "meta": meta, "value": np.array(t.value),
              "prior": np.array(t.prior)}
    for action in range(5):
        ref = planners.mcts_reroot(t, action)
        out, kept = reroot_arrays(arrays, len(t), action)
        if ref is None:
            assert kept == 0
            continue
        assert kept == len(ref)
        assert out["parent"].tolist() == ref.parent and out["first_child"].tolist() == ref.first_child
        assert out["count"].tolist() == ref.count and np.array_equal(out["value"], np.array(ref.value))
        assert ((out["meta"] >> 8) & 0xff).tolist() == ref.n_children
        assert (out["meta"][1:] & 0xff).tolist() == ref.action[1:] and (out["meta"][0] & 0xff) == 0xff
        assert canonical_tree(ref.first_child, ref.n_children, [ref.count]) == \
            canonical_tree(out["first_child"], (out["meta"] >> 8) & 0xff, [out["count"].tolist()])


@pytest.mark.parametrize("key", sorted(G.get("gbopt", {})))
def test_state_aware_planner_gbopt(key):
    """Oracle groundwork for SURVEY 8f rank 3 (no device implementation yet): the flat restatement of
    StateAwarePlanner equals the reference (plan, state-value table, frontier after pruning)."""
    g = G["gbopt"][key]
    plan, t, state_values, leaves = planners.state_aware_plan(finite(), 0, g["budget"], g["gamma"], np_random(g["seed"]))
    assert plan == g["plan"]
    default = 1 / (1 - g["gamma"])      # the reference's defaultdict also materialises keys it merely reads
    assert all(state_values.get(int(k), default) == v for k, v in g["state_values"].items())
    assert all(str(k) in g["state_values"] for k in state_values)
    assert len(leaves) == g["n_leaves"] and len({o for o in t.obs}) == g["n_states"]
    assert sum(t.depth[l] for l in leaves) == g["leaf_depth_sum"]
    assert sum(t.lower[l] for l in leaves) == g["leaf_lower_sum"]


@pytest.mark.parametrize("key", sorted(G["olop"]))
def test_olop_finite(key):
    g = G["olop"][key]
    rng, _ = ref_loader.legacy_np_random(g["seed"])
    cfg = g["config"]
    plan, t = planners.olop_plan(envs.LegacyStepEnv(finite()), cfg["budget"], cfg["gamma"], rng,
                                 upper_bound=cfg["upper_bound"], continuation_type=cfg["continuation_type"])
    assert (t.episodes, t.horizon) == (g["episodes"], g["horizon"])
    assert plan == g["plan"]
    assert_tree_matches(tree_dict(t, ["cumulative_reward", "mu_ucb", "upper", "done"]), g["tree"],
                        ["cumulative_reward", "mu_ucb", "upper"])


def test_allocation_and_kl_known_answers():
    for key, (ep, hz) in G["allocation"].items():
        b, gm = key.split("_")
        assert planners.olop_allocation(int(b), float(gm)) == (ep, hz)
    # reference tests/agents/test_utils.py:29-31 (abs 1e-3)
    assert planners.kl_upper_bound(0.5, 1, np.log(10), eps=1e-3) == pytest.approx(0.997, abs=1e-3)
    assert planners.kl_upper_bound(5, 10, np.log(20), eps=1e-3) == pytest.approx(0.835, abs=1e-3)
    assert planners.kl_upper_bound(10, 20, np.log(40), eps=1e-3) == pytest.approx(0.777, abs=1e-3)
    for s, c, th, ref in G["kl_upper_bound"]:
        assert planners.kl_upper_bound(s, c, th, eps=1e-3) == ref
    for s, c, th, ref in G["kl_upper_bound_eps1e-2"]:
        assert planners.kl_upper_bound(s, c, th) == ref


def test_value_iteration():
    def run(name, gamma, it):
        q, _ = planners.value_iteration("deterministic", M[name + "_T"], M[name + "_R"], M[name + "_term"],
                                        gamma, it)
        return q
    for key, name in [("large1_g0.9_it100", "large1"), ("large1_g1.0_it2", "large1"),
                      ("trap_g0.9_it100", "trap"), ("loop_g0.9_it100", "loop")]:
        g = G["vi"][key]
        assert np.array_equal(run(name, g["gamma"], g["iterations"]), np.array(g["q"])), key
    # SURVEY appendix C anchors
    q = np.array(G["vi"]["large1_g0.9_it100"]["q"])
    assert q[0].tolist() == [8.586584113252275, 8.332319959947995, 8.33806448333072, 8.876425018786213,
                             8.505965200730994]
    assert q.sum() == 4193.544746115075
    rng = np.random.default_rng(0)
    P = rng.uniform(size=(100, 4, 100))
    P /= P.sum(-1, keepdims=True)
    R = rng.uniform(size=(100, 4))
    q, _ = planners.value_iteration("stochastic", P, R, np.zeros(100, bool), 0.95, 100)
    assert np.array_equal(q, np.array(G["vi"]["dense_c1_g0.95_it100"]["q"]))
    Ps, Ns, Rs = envs.garnet(500, 4, 3, seed=1)
    term = np.zeros(500, bool)
    term[::37] = True
    q, _ = planners.value_iteration("sparse", Ps, Rs, term, 0.95, 100, nxt=Ns)
    assert np.array_equal(q, np.array(G["vi"]["sparse_garnet500_g0.95_it100"]["q"]))


def robust_models(kind):
    if kind == "det":
        ms = [envs.garnet(300, 4, 1, seed=20 + m, deterministic=True) for m in range(3)]
        return "deterministic", np.array([m[0] for m in ms]), np.array([m[1] for m in ms])
    Ps, Rs = [], []
    for m in range(2):
        rng = np.random.default_rng(30 + m)
        P = rng.uniform(size=(40, 3, 40))
        P /= P.sum(-1, keepdims=True)
        Ps.append(P)
        Rs.append(rng.uniform(size=(40, 3)))
    return "stochastic", np.array(Ps), np.array(Rs)


def test_robust_value_iteration():
    for key, kind, gamma, it in [("det_3x300x4_g0.9_it60", "det", 0.9, 60), ("dense_2x40x3_g0.95_it100", "dense", 0.95, 100)]:
        mode, T, R = robust_models(kind)
        q, _ = planners.robust_value_iteration(mode, T, R, gamma, it)
        assert np.array_equal(q, np.array(G["robust_vi"][key]["q"])), key
        assert int(np.argmax(q[7])) == G["robust_vi"][key]["act7"]


# ------------------------- HighwayLite -------------------------
def test_highway_scene_and_traces_are_reproducible():
    for seed, words in H["states"].items():
        assert envs.make_highway_state(int(seed)).pack().tolist() == words
    for seed, steps in H["traces"].items():
        env = envs.HighwayLite(seed=int(seed))
        for st in steps[:12]:
            assert env.get_available_actions() == st["avail"]
            _, r, term, trunc, _ = env.step(st["a"])
            assert (r, term, trunc) == (st["r"], st["term"], st["trunc"])
            assert env.state.pack().tolist() == st["state"]


def test_highway_pack_roundtrip():
    s = envs.make_highway_state(3)
    s2 = envs.HighwayLiteState.unpack(s.pack())
    assert s2.pack().tolist() == s.pack().tolist()


@pytest.mark.parametrize("key", ["s0_b75_g0.7", "s1_b300_g0.8"])
def test_opd_highway(key):
    g = H["opd"][key]
    seed = int(key[1])
    plan, t = planners.opd_plan(envs.HighwayLite(seed=seed), g["budget"], g["gamma"], np_random=np_random(0))
    assert plan == g["plan"]
    assert_tree_matches(tree_dict(t, ["reward", "lower", "upper", "done"]), g["tree"],
                        ["reward", "lower", "upper"])


def test_mcts_highway():
    g = H["mcts"]["s0_ep60_h6_g0.8"]
    plan, t = planners.mcts_plan(envs.HighwayLite(seed=0), g["episodes"], g["horizon"], 0.8,
                                 g["temperature"], np_random(0))
    assert plan == g["plan"]
    assert_tree_matches(tree_dict(t, ["value", "prior"]), g["tree"], ["value", "prior"])


def test_vi_and_opd_agree_on_terminal_semantics():
    """done = terminal[state the action is taken in] (finite_mdp's MDP.step), the convention
    value_iteration.py:62 assumes: on a small deterministic MDP whose every path ends in a terminal
    state, the exhaustive OPD tree's root value_lower equals max_a Q*(s0, a) of value iteration."""
    T = np.array([[1, 2], [3, 3], [3, 4], [3, 3], [4, 4]], dtype=np.int32)
    R = np.array([[0.2, 0.5], [0.9, 0.1], [0.3, 0.6], [0.7, 0.4], [1.0, 0.8]])
    term = np.array([False, False, False, True, True])
    gamma = 0.9
    q, _ = planners.value_iteration("deterministic", T, R, term, gamma, 50)
    _, tree = planners.opd_plan(envs.FiniteMDPLite(T, R, term), 2000, gamma,
                                np_random=np.random.Generator(np.random.PCG64(0)))
    # terminal leaves keep being expanded by the reference (:111-112) but add nothing below value_lower's max
    # until the discount is exhausted; compare the part of the value collected up to the terminal step
    best = max(q[0])
    path_values = []
    for a0 in range(2):
        s1 = T[0, a0]
        for a1 in range(2):
            s2 = T[s1, a1]
            path_values.append(R[0, a0] + gamma * R[s1, a1] + gamma ** 2 * max(R[s2]))
    assert abs(best - max(path_values)) < 1e-12
    done_nodes = [i for i in range(len(tree)) if tree.done[i]]
    assert done_nodes and all(tree.depth[i] >= 3 for i in done_nodes)

    def path_return(i):
        g = 0.0
        while i > 0:
            g += gamma ** (tree.depth[i] - 1) * tree.reward[i]
            i = tree.parent[i]
        return g
    assert abs(max(path_return(i) for i in done_nodes if tree.depth[i] == 3) - best) < 1e-12


def test_mcts_policies_and_closed_loop_goldens():
    """The oracle's policies (mcts.py:34-97) against the reference, and: on a deterministic env the
    reference's closed_loop search has the statistics of the open-loop one."""
    for key, g in G["mcts_policies"].items():
        plan, t = planners.mcts_plan(finite(), g["episodes"], g["horizon"], g["config"]["gamma"], g["temperature"],
                                     np_random(g["seed"]), prior_policy=g["config"]["prior_policy"],
                                     rollout_policy=g["config"]["rollout_policy"])
        assert plan == g["plan"], key
        assert_tree_matches(tree_dict(t, ["value", "prior"]), g["tree"], ["value", "prior"])
    g = G["mcts_closed_loop"]
    episodes, horizon = planners.olop_allocation(g["config"]["budget"], g["config"]["gamma"])
    plan, t = planners.mcts_plan(finite(), episodes, horizon, g["config"]["gamma"], 10.0, np_random(g["seed"]))
    assert plan == g["plan_actions"] and g["plan_len"] == 2 * len(plan) - 1
    assert [[t.action[c], t.count[c], t.value[c]] for c in t.children(0)] == g["root"]
    assert t.count[0] == g["root_count"] and t.value[0] == g["root_value"]


def test_preference_tables_follow_numpy():
    from rl_agents_b200.engine.tables import preference_tables
    prior, cdf = preference_tables(5, 3)
    p = np.ones(4) / (4 - 1 + 3)
    p[2] *= 3
    assert np.array_equal(prior[4, 3, :4], p)
    c = p.cumsum()
    c /= c[-1]
    assert np.array_equal(cdf[4, 3, :4], c)
    assert np.array_equal(prior[3, 0, :3], np.ones(3) / 3)


def test_drop_oracle_matches_the_reference_goldens():
    """oracle.planners.robust_plan against DiscreteRobustPlanner of the unmodified reference (joint env of 2 / 3
    finite-MDP models, one with terminal states and a terminal reward)."""
    termx = M["large1_term"].copy()
    termx[[3, 17, 66, 91]] = True
    table = {"large1": ("large1", M["large1_term"]), "large1t": ("large1", termx), "large2": ("large2", M["large2_term"])}
    for key, g in G["drop"].items():
        models = [envs.FiniteMDPLite(M[table[n][0] + "_T"], M[table[n][0] + "_R"], table[n][1]) for n in g["models"]]
        plan, t = planners.robust_plan(models, g["budget"], g["gamma"], g["terminal_reward"], np_random(0))
        assert plan == g["plan"], key
        assert t.parent == g["tree"]["parent"] and t.action == g["tree"]["action"] and t.count == g["tree"]["count"]
        assert t.lower == g["tree"]["lower"] and t.upper == g["tree"]["upper"]


def test_graph_based_planner_gbopd():
    """oracle.planners.graph_based_plan against the unmodified GraphBasedPlanner (legacy 4-tuple env shim): exact
    for accuracy = 0, same plan / bounds within the accuracy for the default (the reference iterates parent sets)."""
    for key, g in G["gbopd"].items():
        plan, nodes = planners.graph_based_plan(envs.LegacyStepEnv(finite()), 0, g["budget"], g["gamma"], np_random(g["seed"]),
                                                g["accuracy"], g["sampling_timeout"])
        ref = {int(k): v for k, v in g["nodes"].items()}
        assert plan == g["plan"] and set(nodes) == set(ref), key
        tol = 0.0 if g["accuracy"] == 0 else 10 * g["accuracy"]
        assert all(abs(nodes[s]["lower"] - ref[s][0]) <= tol and abs(nodes[s]["upper"] - ref[s][1]) <= tol and
                   nodes[s]["expanded"] == ref[s][2] for s in ref), key
