"""The C oracle (oracle/c) against the numpy oracle and the reference's golden vectors: three
independent statements of the HighwayLite spec (numpy, C, CUDA) must agree bit for bit."""
import time

import numpy as np

from oracle import c_oracle
from oracle import envs as oenvs
from oracle import planners
from tests.util import assert_tree_matches, load_golden

H = load_golden("golden_highway.json")


def test_c_env_matches_golden_traces_and_numpy():
    for seed, steps in H["traces"].items():
        w = np.array(H["states"][seed] if seed in H["states"] else oenvs.make_highway_state(int(seed)).pack(),
                     dtype=np.int32).reshape(1, -1)
        for st in steps:
            w, r, f = c_oracle.step_batch(w, [st["a"]])
            assert w[0].tolist() == st["state"]
            assert r[0] == np.float32(st["r"]) and f[0] == (1 if st["term"] else 0) | (2 if st["trunc"] else 0)
    # random sweep incl. exact x ties, absent and crashed slots
    rng = np.random.default_rng(3)
    scenes = []
    for seed in range(300, 340):
        s = oenvs.make_highway_state(seed)
        if seed % 5 == 0:
            s.x[6] = s.x[2]
            s.flags[13] = 0
            s.flags[14] = 3
        scenes.append(oenvs.HighwayLite(s))
    w = np.stack([e.state.pack() for e in scenes])
    for step in range(6):
        acts = [int(rng.choice(e.get_available_actions())) for e in scenes]
        w, r, f = c_oracle.step_batch(w, acts)
        for i, e in enumerate(scenes):
            _, r_np, term, trunc, _ = e.step(acts[i])
            assert np.array_equal(w[i], e.state.pack()), (step, i)
            assert r[i] == np.float32(r_np) and f[i] == (1 if term else 0) | (2 if trunc else 0)


def test_c_opd_matches_reference_golden_trees():
    for key in ("s0_b75_g0.7", "s1_b300_g0.8", "s2_b1000_g0.8", "s0_b10000_g0.8"):
        g = H["opd"][key]
        t0 = time.perf_counter()
        t = c_oracle.opd_plan(np.array(H["states"][key[1]], dtype=np.int32), g["budget"], g["gamma"])
        dt = time.perf_counter() - t0
        assert t["n_leaves"] == g["n_leaves"]
        assert_tree_matches(t, g["tree"], ["reward", "lower", "upper"])
        assert dt < 60
    # and against the Python restatement on a fresh scene
    _, tp = planners.opd_plan(oenvs.HighwayLite(seed=9), 200, 0.85, np_random=np.random.default_rng(0))
    tc = c_oracle.opd_plan(oenvs.make_highway_state(9).pack(), 200, 0.85)
    assert tc["parent"].tolist() == tp.parent and tc["count"].tolist() == tp.count
    assert np.array_equal(tc["upper"], np.array(tp.upper)) and np.array_equal(tc["lower"], np.array(tp.lower))


def test_c_mcts_matches_reference_golden_and_python_restatement():
    from oracle.pcg64 import PCG64

    def words(seed):
        return PCG64.from_numpy(np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))).words()
    for key in sorted(H["mcts"]):
        g = H["mcts"][key]
        t, _ = c_oracle.mcts_plan(np.array(H["states"][key[1]], dtype=np.int32), g["episodes"], g["horizon"],
                                  g["config"]["gamma"], g["temperature"], words(g["seed"]))
        assert_tree_matches(t, g["tree"], ["value", "prior"])
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(11)))
    plan, tp = planners.mcts_plan(oenvs.HighwayLite(seed=8), 150, 7, 0.85, 10.0, rng)
    tc, w = c_oracle.mcts_plan(oenvs.make_highway_state(8).pack(), 150, 7, 0.85, 10.0, words(11))
    assert tc["parent"].tolist() == tp.parent and tc["count"].tolist() == tp.count and tc["action"].tolist() == tp.action
    assert np.array_equal(tc["value"], np.array(tp.value))
    assert w.tolist() == PCG64.from_numpy(rng).words().tolist()          # same stream position afterwards


def test_c_ttc_value_iteration_matches_reference_goldens_and_numpy_oracle():
    """Third statement of docs/HIGHWAY_LITE_SPEC.md section 9 (TTC-grid MDP + the VI agent's fixed point): against the
    goldens of the unmodified reference agent and against the numpy oracle on the edge scenes."""
    from tests.util import load_golden, ttc_edge_scenes
    V = load_golden("golden_highway_vi.json")
    for c in V["cases"]:
        w = np.array(c["words"], dtype=np.int32)
        assert np.array_equal(c_oracle.ttc_grid(w), np.array(c["grid"]))
        q, st, act, _ = c_oracle.ttc_value_iteration(w, c["config"].get("gamma", 1.0), c["config"]["iterations"])
        assert np.array_equal(q, np.array(c["q"])) and st == c["state"] and act == c["act"]
    for w in ttc_edge_scenes():
        mdp = oenvs.highway_finite_mdp(oenvs.HighwayLiteState.unpack(w))
        for gamma, iterations in ((1.0, 10), (0.9, 100), (1.0, 0)):
            q_ref, sweeps_ref = planners.value_iteration("deterministic", mdp.transition, mdp.reward, mdp.terminal, gamma,
                                                         iterations)
            q, st, act, sweeps = c_oracle.ttc_value_iteration(w, gamma, iterations)
            assert np.array_equal(q, q_ref) and sweeps == sweeps_ref and st == mdp.state
            assert act == int(np.argmax(q_ref[mdp.state]))
