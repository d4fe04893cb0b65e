"""The time-to-collision grid MDP of a HighwayLite scene (`env.unwrapped.to_finite_mdp()`, value_iteration.py:17,32;
docs/HIGHWAY_LITE_SPEC.md section 9) -- CPU side: the oracle statement against the goldens made by the unmodified
reference ValueIterationAgent, and the product's host statement against the oracle."""
import numpy as np

from oracle import envs as oenvs
from oracle import planners
from tests.util import load_golden, ttc_edge_scenes

V = load_golden("golden_highway_vi.json")
H = load_golden("golden_highway.json")


def test_oracle_ttc_mdp_and_value_iteration_match_the_reference_agent():
    assert len(V["cases"]) >= 30
    for c in V["cases"]:
        st = oenvs.HighwayLiteState.unpack(np.array(c["words"], dtype=np.int32))
        assert np.array_equal(oenvs.highway_ttc_grid(st), np.array(c["grid"]))
        mdp = oenvs.highway_finite_mdp(st)
        assert mdp.state == c["state"] and list(mdp.original_shape) == c["shape"] == [3, 4, 10]
        if "transition" in c:
            assert np.array_equal(mdp.transition, np.array(c["transition"]))
            assert np.array_equal(mdp.reward, np.array(c["reward"]))
            assert np.array_equal(mdp.terminal, np.array(c["terminal"]))
        cfg = {"gamma": 1.0, "iterations": 100}
        cfg.update(c["config"])
        q, _ = planners.value_iteration("deterministic", mdp.transition, mdp.reward, mdp.terminal, cfg["gamma"],
                                        cfg["iterations"])
        assert np.array_equal(q, np.array(c["q"]))                 # the reference agent's Q, bit for bit
        assert int(np.argmax(q[mdp.state])) == c["act"]


def test_ttc_grid_is_a_cost_map_with_the_documented_structure():
    for c in V["cases"]:
        g = np.array(c["grid"])
        assert set(np.unique(g)) <= {0.0, 0.5, 1.0}
        mdp = oenvs.highway_finite_mdp(oenvs.HighwayLiteState.unpack(np.array(c["words"], dtype=np.int32)))
        t = mdp.transition.reshape(3, 4, 10, 5)
        hh, ii, jj = np.unravel_index(t, (3, 4, 10))
        assert (jj[:, :, :9] == np.arange(1, 10)[None, None, :, None]).all() and (jj[:, :, 9] == 9).all()   # time advances
        assert (hh[:, :, 1:, 3] == np.arange(3)[:, None, None]).all()       # FASTER / SLOWER only act at time 0
        assert (ii[:, :, :, 0] == np.maximum(np.arange(4) - 1, 0)[None, :, None]).all()                     # LEFT clips
        assert mdp.terminal.reshape(3, 4, 10)[:, :, 9].all()


def test_product_host_ttc_mdp_equals_the_oracle():
    from rl_agents_b200.envs.highway_lite import HighwayLiteEnv, ttc_grid
    words = [np.array(c["words"], dtype=np.int32) for c in V["cases"]]
    for steps in H["traces"].values():                                 # scenes in the middle of lane changes, crashes
        words += [np.array(s["state"], dtype=np.int32) for s in steps[::4]]
    words += ttc_edge_scenes()
    assert len(words) > 140
    for w in words:
        st = oenvs.HighwayLiteState.unpack(w)
        assert np.array_equal(ttc_grid(w), oenvs.highway_ttc_grid(st))
        a, b = HighwayLiteEnv(words=w).to_finite_mdp(), oenvs.highway_finite_mdp(st)
        assert a.mode == "deterministic" and a.state == b.state and tuple(a.original_shape) == (3, 4, 10)
        assert np.array_equal(a.transition, b.transition) and np.array_equal(a.reward, b.reward)
        assert np.array_equal(a.terminal, b.terminal)
