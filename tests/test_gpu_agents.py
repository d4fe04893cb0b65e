"""GPU tests of the plugin surface: the drop-in agents return the reference's
plans on identical seeds (golden vectors) through plan()/act()."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_golden, load_mdps

pytestmark = pytest.mark.gpu
G = load_golden("golden_finite.json")
H = load_golden("golden_highway.json")
M = load_mdps()


def finite_env(state=0):
    from rl_agents_b200.envs import FiniteMDPEnv
    return FiniteMDPEnv(M["large1_T"], M["large1_R"], M["large1_term"], state=state)


def test_opd_agent_finite_matches_reference():
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    g = G["opd"]["large1_b500_g0.9"]
    agent = DeterministicPlannerAgent(finite_env(), {"budget": 500, "gamma": 0.9})
    agent.seed(0)
    assert agent.plan(None) == g["plan"]
    assert agent.act(None) == g["plan"][0]
    g = G["opd"]["large1_b75_g0.7"]
    agent = DeterministicPlannerAgent(finite_env(), {"budget": 75, "gamma": 0.7})
    agent.seed(0)
    assert agent.plan(None) == g["plan"]


def test_opd_agent_highway_episode_with_preprocessor_and_receding_horizon():
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.envs import HighwayLiteEnv
    env = HighwayLiteEnv(seed=0)
    cfg = {"budget": 75, "gamma": 0.7, "env_preprocessors": [{"method": "simplify"}], "receding_horizon": 2}
    agent = DeterministicPlannerAgent(env, cfg)
    agent.seed(0)
    g = H["opd"]["s0_b75_g0.7"]
    first = agent.plan(env.observation())
    assert first == g["plan"]
    # drive a short episode: env.step on the CUDA transition, oracle env in lock-step
    oenv = oenvs.HighwayLite(seed=0)
    actions = first
    for k in range(4):
        obs, r, term, trunc, _ = env.step(actions[0])
        _, r2, term2, trunc2, _ = oenv.step(actions[0])
        assert (r, term, trunc) == (r2, term2, trunc2)
        assert env.words.tolist() == oenv.state.pack().tolist()
        if term:
            break
        prev = actions
        actions = agent.plan(obs)
        if k % 2 == 0:                       # receding_horizon = 2: every other call reuses the plan
            assert actions == prev[1:]
        else:                                # replanned on the device from the new scene
            plan, _ = planners.opd_plan(oenvs.HighwayLite(oenv.state.copy()), 75, 0.7,
                                        np_random=np.random.default_rng(0))
            assert actions[:1] == plan[:1] or len(plan) == 0


def test_mcts_agent_matches_reference_and_keeps_rng_in_sync():
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    g = G["mcts"]["large1_b400_g0.8"]
    agent = MCTSAgent(finite_env(), {"budget": 400, "gamma": 0.8})
    agent.seed(g["seed"])
    assert agent.plan(None) == g["plan"]
    # a second decision continues the same stream the reference would continue
    ref_rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(g["seed"])))
    env = oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"])
    planners.mcts_plan(env, g["episodes"], g["horizon"], 0.8, g["temperature"], ref_rng)
    plan2, _ = planners.mcts_plan(env, g["episodes"], g["horizon"], 0.8, g["temperature"], ref_rng)
    assert agent.plan(None) == plan2


def test_vi_agent_matches_reference():
    from rl_agents_b200.agents.dynamic_programming.value_iteration import ValueIterationAgent
    g = G["vi"]["large1_g0.9_it100"]
    agent = ValueIterationAgent(finite_env(), {"gamma": 0.9, "iterations": 100})
    assert np.array_equal(agent.state_action_value, np.array(g["q"]))
    assert agent.act(0) == g["act0"] == 3
    assert agent.plan(0) == [3]
    # get_state_value: the reference's separate fixed point on V (value_iteration.py:37-40), restated with numpy
    v = np.zeros(100)
    for _ in range(100):
        nv = planners.bellman_expectation("deterministic", M["large1_T"], M["large1_R"], M["large1_term"].astype(bool), v,
                                          0.9).max(axis=-1)
        if np.allclose(v, nv):
            break
        v = nv
    assert np.array_equal(agent.get_state_value(), v)
    # non-finite env path: to_finite_mdp() re-solved on every act (value_iteration.py:31-34)
    agent2 = ValueIterationAgent(oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"]),
                                 {"gamma": 1.0, "iterations": 2})
    assert np.array_equal(agent2.state_action_value, np.array(G["vi"]["large1_g1.0_it2"]["q"]))
    assert agent2.act(None) == 3


def test_olop_agent_matches_reference():
    from oracle import ref_loader
    from rl_agents_b200.agents.tree_search.olop import OLOPAgent
    g = G["olop"]["large1_b200_g0.9_uniform"]
    agent = OLOPAgent(finite_env(), dict(g["config"]))
    agent.seed(g["seed"])
    assert (agent.planner.config["episodes"], agent.planner.config["horizon"]) == (g["episodes"], g["horizon"])
    assert agent.plan(None) == g["plan"]
    d = agent.planner.last_tree.tree_dict(0)
    assert d["count"].tolist() == g["tree"]["count"] and d["parent"].tolist() == g["tree"]["parent"]
    np.testing.assert_allclose(d["upper"][0], g["tree"]["upper"][0], rtol=1e-9)     # 7.7121380...
    # the planner's numpy stream continues exactly where the reference's would
    rng, _ = ref_loader.legacy_np_random(g["seed"])
    planners.olop_plan(oenvs.LegacyStepEnv(oenvs.FiniteMDPLite(M["large1_T"], M["large1_R"], M["large1_term"])),
                       g["config"]["budget"], g["config"]["gamma"], rng, upper_bound=g["config"]["upper_bound"],
                       continuation_type=g["config"]["continuation_type"])
    assert agent.planner.np_random.bit_generator.state["state"] == rng.bit_generator.state["state"]


def test_agents_reject_unsupported_envs_and_options():
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    from rl_agents_b200.envs import FiniteMDPEnv
    stochastic = FiniteMDPEnv(np.full((3, 2, 3), 1 / 3), np.zeros((3, 2)), mode="stochastic")
    with pytest.raises(ValueError):
        DeterministicPlannerAgent(stochastic, {"budget": 10}).plan(None)
    with pytest.raises(ValueError):
        MCTSAgent(finite_env(), {"rollout_policy": {"type": "nope"}})
    with pytest.raises(KeyError):
        MCTSAgent(finite_env(), {"horizon": 5}).plan(None)        # mcts.py:116-118,180: episodes missing


def test_mcts_root_parallel_extension():
    """"root_parallel": R (not in the reference): R trees of episodes/R from the same root, merged
    root statistics.  Deterministic for a given planner seed; every episode is accounted for."""
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    from rl_agents_b200.envs import HighwayLiteEnv
    cfg = {"episodes": 512, "horizon": 8, "gamma": 0.8, "root_parallel": 16}
    plans = []
    for _ in range(2):
        agent = MCTSAgent(HighwayLiteEnv(seed=2), dict(cfg))
        agent.seed(7)
        plans.append(agent.plan(None))
        stats = agent.planner.root_statistics
        assert stats["counts"].sum() == 16 * (512 // 16 - 1)       # the expanding episode of each tree selects no child
        assert plans[-1][0] == int(np.argmax(stats["counts"])) or (stats["counts"] == stats["counts"].max()).sum() > 1
    assert plans[0] == plans[1]


def test_robust_value_iteration_agent_matches_reference():
    from rl_agents_b200.agents.dynamic_programming.robust_value_iteration import RobustValueIterationAgent
    from tests.test_oracle import robust_models
    for key, kind, gamma, it in [("det_3x300x4_g0.9_it60", "det", 0.9, 60), ("dense_2x40x3_g0.95_it100", "dense", 0.95, 100)]:
        mode, T, R = robust_models(kind)
        models = [{"mode": mode, "transition": T[m].tolist(), "reward": R[m].tolist()} for m in range(len(T))]
        agent = RobustValueIterationAgent(None, {"gamma": gamma, "iterations": it, "models": models})
        q = agent.get_state_action_value()
        assert np.array_equal(q, np.array(G["robust_vi"][key]["q"])), key
        assert agent.act(7) == G["robust_vi"][key]["act7"]
        q_ref, sweeps = planners.robust_value_iteration(mode, T, R, gamma, it)
        assert agent.sweeps == sweeps
    with pytest.raises(ValueError):
        RobustValueIterationAgent(None, {})


def test_batched_evaluation_equals_per_episode_agents():
    """The lock-step batched episode runner takes the same actions, rewards and lengths as driving
    one DeterministicPlannerAgent per episode (which itself equals the reference's plans)."""
    from rl_agents_b200.agents.tree_search.deterministic import DeterministicPlannerAgent
    from rl_agents_b200.envs import HighwayLiteEnv
    from rl_agents_b200.evaluation import run_batched_episodes
    seeds = [0, 1, 2, 3, 4]
    out = run_batched_episodes("opd", seeds, 150, 0.8, max_steps=12, planner_seed=100)
    for i, s in enumerate(seeds):
        env = HighwayLiteEnv(seed=s)
        agent = DeterministicPlannerAgent(env, {"budget": 150, "gamma": 0.8})
        agent.seed(100 + i)
        total, steps = 0.0, 0
        for k in range(12):
            a = agent.act(None)
            assert a == out["actions"][i, k], (s, k)
            _, r, term, trunc, _ = env.step(a)
            total += float(np.float32(r))
            steps += 1
            if term or trunc:
                break
        assert steps == out["lengths"][i] and abs(total - out["returns"][i]) < 1e-9
    mc = run_batched_episodes("mcts", seeds, 100, 0.8, max_steps=5)
    ol = run_batched_episodes("olop", seeds, 100, 0.8, max_steps=5)
    assert mc["lengths"].min() >= 1 and ol["lengths"].min() >= 1


def test_mcts_agent_subtree_strategy_matches_reference():
    """step_strategy "subtree": the tree is re-rooted at the executed action between decisions
    (host compaction) and the kernel resumes from it; three decisions against the reference."""
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    from tests.util import canonical_tree
    g = G["mcts_subtree"]
    env = finite_env()
    agent = MCTSAgent(env, {"budget": 300, "gamma": 0.85, "step_strategy": "subtree"})
    agent.seed(4)
    for k in range(3):
        assert env.mdp.state == g["states"][k]
        plan = agent.plan(None)
        assert plan == g["plans"][k], k
        d = agent.planner.last_tree.tree_dict(0)
        got = canonical_tree(d["first_child"], d["n_children"], [d["count"].tolist(), d["value"].tolist(),
                                                                 d["prior"].tolist()])
        assert got == g["trees"][k], k
        env.step(plan[0])


def test_mcts_agent_preference_and_random_policies_match_reference():
    """prior / rollout policies other than random_available (mcts.py:34-97): goldens from the reference."""
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    for key, g in G["mcts_policies"].items():
        agent = MCTSAgent(finite_env(), dict(g["config"]))
        agent.seed(g["seed"])
        assert agent.plan(None) == g["plan"], key
        d = agent.planner.last_tree.tree_dict(0)
        full = "n_nodes" not in g["tree"]
        k = len(g["tree"]["parent"]) if full else 64
        assert d["parent"][:k].tolist() == g["tree"]["parent"][:k] and d["count"][:k].tolist() == g["tree"]["count"][:k]
        assert np.array_equal(d["value"][:k], np.array(g["tree"]["value"][:k]))
        assert np.array_equal(d["prior"][:k], np.array(g["tree"]["prior"][:k]))


def test_mcts_agent_closed_loop_on_a_deterministic_env_matches_reference():
    """closed_loop=True in the reference adds one observation-keyed node per action node on a deterministic
    env; statistics and the recommended actions are those of the open-loop search."""
    from rl_agents_b200.agents.tree_search.mcts import MCTSAgent
    g = G["mcts_closed_loop"]
    agent = MCTSAgent(finite_env(), dict(g["config"]))
    agent.seed(g["seed"])
    assert agent.plan(0) == g["plan_actions"]
    d = agent.planner.last_tree.tree_dict(0)
    fc, n = int(d["first_child"][0]), int(d["n_children"][0])
    assert [[int(d["action"][c]), int(d["count"][c]), float(d["value"][c])] for c in range(fc, fc + n)] == g["root"]
    assert int(d["count"][0]) == g["root_count"] and float(d["value"][0]) == g["root_value"]
