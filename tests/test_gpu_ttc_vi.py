"""GPU: ValueIterationAgent on HighwayLite scenes -- the batched conversion + fixed-point kernel (b2_highway_ttc_vi)
against the goldens of the unmodified reference agent and the oracle, the agent drop-in (device and host conversion),
and the batched evaluation harness."""
import numpy as np
import pytest

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_golden, ttc_edge_scenes

pytestmark = pytest.mark.gpu
V = load_golden("golden_highway_vi.json")
H = load_golden("golden_highway.json")


def solve(words, gamma, iterations):
    from rl_agents_b200.engine.ttc_vi import HighwayTTCVI
    out = HighwayTTCVI(gamma, iterations).solve(np.stack(words))
    return {k: v.cpu().numpy() for k, v in out.items()}


def test_ttc_vi_kernel_matches_the_reference_agent_goldens():
    for cfg in ({"iterations": 10}, {"gamma": 0.9, "iterations": 100}):
        cases = [c for c in V["cases"] if c["config"] == cfg]
        assert len(cases) >= 15
        out = solve([np.array(c["words"], dtype=np.int32) for c in cases], cfg.get("gamma", 1.0), cfg["iterations"])
        for i, c in enumerate(cases):
            assert out["state"][i] == c["state"] and out["action"][i] == c["act"], (c["seed"], c["step"])
            assert np.array_equal(out["q"][i], np.array(c["q"])), (c["seed"], c["step"])      # bit for bit


def test_ttc_vi_kernel_matches_the_oracle_on_many_scenes():
    words = [oenvs.make_highway_state(s).pack() for s in range(100, 164)]
    for steps in H["traces"].values():
        words += [np.array(s["state"], dtype=np.int32) for s in steps]
    words += ttc_edge_scenes()
    for gamma, iterations in ((1.0, 10), (0.95, 3), (0.8, 100), (1.0, 0)):
        out = solve(words, gamma, iterations)
        for i, w in enumerate(words):
            mdp = oenvs.highway_finite_mdp(oenvs.HighwayLiteState.unpack(w))
            q, sweeps = planners.value_iteration("deterministic", mdp.transition, mdp.reward, mdp.terminal, gamma, iterations)
            assert np.array_equal(out["q"][i], q), (i, gamma, iterations)
            assert out["sweeps"][i] == sweeps and out["state"][i] == mdp.state
            assert out["action"][i] == int(np.argmax(q[mdp.state]))


@pytest.mark.parametrize("conversion", ["device", "host"])
def test_vi_agent_on_highway_scenes_matches_reference(conversion):
    from rl_agents_b200.agents.dynamic_programming.value_iteration import ValueIterationAgent
    from rl_agents_b200.envs.highway_lite import HighwayLiteEnv
    for c in V["cases"][::3]:
        env = HighwayLiteEnv(words=np.array(c["words"], dtype=np.int32))
        cfg = dict(c["config"], conversion=conversion)
        agent = ValueIterationAgent(env, cfg)
        assert agent.config["gamma"] == c["config"].get("gamma", 1.0)
        assert np.array_equal(agent.state_action_value, np.array(c["q"]))
        assert agent.mdp.state == c["state"] and agent.act(None) == c["act"]
        states, actions = agent.plan_trajectory(agent.mdp.state)
        assert actions[0] == c["act"] and len(states) >= 2


def test_batched_vi_evaluation_on_highway_scenes():
    from rl_agents_b200.evaluation import run_batched_episodes
    res = run_batched_episodes("vi", list(range(6)), 10, 1.0, max_steps=6)
    first = {c["seed"]: c["act"] for c in V["cases"] if c["step"] == 0 and c["config"] == {"iterations": 10}}
    # golden scenes of step 0 are oracle make_highway_state(seed) = the harness's make_scene(seed)
    assert [int(a) for a in res["actions"][:, 0]] == [first[s] for s in range(6)]
    assert res["lengths"].min() >= 1 and res["returns"].shape == (6,)
