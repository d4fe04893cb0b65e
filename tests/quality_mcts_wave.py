#!/usr/bin/env python
"""Quality of the wavefront MCTS against the reference's sequential MCTS (CPU, C statements of both
specifications) -- not a pytest test; writes profiles/r02_mcts_quality.json.

For every scene: the reference-order MCTS (oracle/c mcts_highway_plan, numpy PCG64 stream), the wavefront
(widths 256 / 512 / 1024) and the root-parallel split (64 trees x 64 episodes, merged counts) at the C3 size
(4096 episodes x horizon 20), each over several seeds; reported: how often the recommended action equals the
reference-order MCTS's modal recommendation for that scene, and the total-variation distance between the root
visit distributions."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle, envs as oenvs      # noqa: E402
from oracle.pcg64 import PCG64   # noqa: E402


def words_from_generator(g):
    return PCG64.from_numpy(g).words()


def root_stats(t):
    fc, n = int(t["first_child"][0]), int(t["n_children"][0])
    counts, values = np.zeros(5), np.zeros(5)
    for c in range(fc, fc + n):
        counts[t["action"][c]], values[t["action"][c]] = t["count"][c], t["value"][c]
    return counts, values


def recommend(counts, values):
    ties = np.nonzero(counts == counts.max())[0]
    return int(max(ties, key=lambda i: values[i]))


def scene_runs(args):
    """All planners on one scene over the seeds: dict planner -> list of (root counts, root values)."""
    sc, seeds, E, H = args
    words = oenvs.make_highway_state(sc).pack()
    runs = {k: [] for k in KINDS}
    for sd in seeds:
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(sd)))
        t, _ = c_oracle.mcts_plan(words, E, H, 0.8, 10.0, words_from_generator(g))
        runs["strict"].append(root_stats(t))
        for w in (256, 512, 1024):
            runs["wave%d" % w].append(root_stats(c_oracle.mcts_plan_wave(words, E, H, 0.8, 10.0, w, sd)))
        counts, sums = np.zeros(5), np.zeros(5)
        for r, child in enumerate(np.random.SeedSequence(sd).spawn(64)):
            g = np.random.Generator(np.random.PCG64(child))
            t, _ = c_oracle.mcts_plan(words, 64, H, 0.8, 10.0, words_from_generator(g))
            c, v = root_stats(t)
            counts += c
            sums += c * v
        runs["root_parallel64"].append((counts, np.where(counts > 0, sums / np.maximum(counts, 1), 0.0)))
    return sc, runs


KINDS = ("strict", "wave256", "wave512", "wave1024", "root_parallel64")


def main():
    import multiprocessing as mp
    E, H = 4096, 20
    scenes, seeds = list(range(int(os.environ.get("SCENES", 6)))), list(range(int(os.environ.get("SEEDS", 4))))
    out = {"episodes": E, "horizon": H, "scenes": len(scenes), "seeds": len(seeds), "rows": {}, "per_scene": {}}
    agree = {k: [] for k in KINDS}
    tv = {k: [] for k in agree}
    regret = {k: [] for k in agree}
    c_oracle.build()
    with mp.Pool(min(len(scenes), len(os.sched_getaffinity(0)))) as pool:
        results = pool.map(scene_runs, [(sc, seeds, E, H) for sc in scenes])
    for sc, runs in results:
        acts = [recommend(*r) for r in runs["strict"]]
        modal = max(set(acts), key=acts.count)
        ref_dist = np.mean([c / c.sum() for c, _ in runs["strict"]], axis=0)
        # the reference-order MCTS's own estimate of the root action values (count-weighted over its seeds): the regret
        # of a recommendation is measured on it, so that disagreements between near-equal actions count as what they are
        cw = np.sum([c for c, _ in runs["strict"]], axis=0)
        q_ref = np.sum([c * v for c, v in runs["strict"]], axis=0) / np.maximum(cw, 1)
        q_ref = np.where(cw > 0, q_ref, -np.inf)
        for k in agree:
            agree[k] += [int(recommend(*r) == modal) for r in runs[k]]
            tv[k] += [0.5 * float(np.abs(c / c.sum() - ref_dist).sum()) for c, _ in runs[k]]
            regret[k] += [float(q_ref.max() - q_ref[recommend(*r)]) for r in runs[k]]
        out["per_scene"][str(sc)] = {"reference_modal_action": modal, "reference_root_values": [float(x) for x in q_ref],
                                     "agreement": {k: float(np.mean(agree[k][-len(seeds):])) for k in agree}}
        print("scene", sc, "modal", modal, out["per_scene"][str(sc)]["agreement"], flush=True)
    n = len(scenes) * len(seeds)
    for k in agree:
        p = float(np.mean(agree[k]))
        out["rows"][k] = {"agreement_with_reference_modal_action": p,
                          "agreement_standard_error": float(np.sqrt(max(p * (1 - p), 1e-12) / n)),
                          "root_visit_tv_distance_to_reference_mean": float(np.mean(tv[k])),
                          "value_regret_mean": float(np.mean(regret[k])), "value_regret_max": float(np.max(regret[k]))}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", os.environ.get("OUT", "r02_mcts_quality.json")), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["rows"], indent=1))


if __name__ == "__main__":
    main()
