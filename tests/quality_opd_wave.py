#!/usr/bin/env python
"""Quality of the wavefront OPD against the reference's strict best-first order (CPU, C statements of both
specifications, oracle/c) -- not a pytest test; writes profiles/r02d_opd_wave_quality.json.

For every scene (C2: budget 10 000, gamma 0.8): the strict tree and the wavefront trees of several widths; reported per
width: how often the root action (arg-max value_lower, first maximum) equals the strict one, the gap of the root's
value_lower and value_upper to the strict ones, the maximal depth reached, the number of waves."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle, envs as oenvs      # noqa: E402

WIDTHS = (16, 64, 128, 256)


def root_view(t):
    fc, n = int(t["first_child"][0]), int(t["n_children"][0])
    lo = t["lower"][fc:fc + n]
    return int(t["action"][fc + int(np.argmax(lo))]), float(t["lower"][0]), float(t["upper"][0]), int(t["depth"].max())


def scene_words(sc):
    """Generated scene `sc`, driven `sc % 13` decision steps with random available actions (C statement of the env):
    mid-episode scenes with vehicles changing lanes around the ego, not only the generator's tidy initial ones."""
    words = oenvs.make_highway_state(sc).pack()
    rng = np.random.default_rng(sc)
    for _ in range(sc % 13):
        st = oenvs.HighwayLiteState.unpack(words)
        avail = oenvs.highway_available_actions(st)
        w2, _, flags = c_oracle.step_batch(words.reshape(1, -1), np.array([avail[rng.integers(len(avail))]], dtype=np.int32))
        if flags[0] & 1:
            break           # keep the last state before a crash
        words = w2[0]
    return words


def scene_rows(sc):
    words = scene_words(sc)
    budget, gamma = int(os.environ.get("BUDGET", 10000)), 0.8
    strict = root_view(c_oracle.opd_plan(words, budget, gamma))
    rows = {}
    for w in WIDTHS:
        t = c_oracle.opd_plan_wave(words, budget, gamma, w)
        a, lo, up, d = root_view(t)
        rows[w] = {"same_action": int(a == strict[0]), "lower_gap": strict[1] - lo, "upper_gap": up - strict[2],
                   "depth": d, "waves": int(t["n_waves"])}
    return sc, strict, rows


def main():
    import multiprocessing as mp
    scenes = list(range(1000, 1000 + int(os.environ.get("SCENES", 64))))
    c_oracle.build()
    with mp.Pool(len(os.sched_getaffinity(0))) as pool:
        results = pool.map(scene_rows, scenes)
    out = {"workload": "C2: OPD on HighwayLite, budget 10000 (2000 expansions), gamma 0.8; scenes make_highway_state(1000..) "
                       "driven 0..12 random decision steps",
           "scenes": len(scenes), "strict_root_value_lower_mean": float(np.mean([s[1] for _, s, _ in results])),
           "strict_max_depth_mean": float(np.mean([s[3] for _, s, _ in results])), "widths": {}}
    for w in WIDTHS:
        r = [rows[w] for _, _, rows in results]
        out["widths"][str(w)] = {
            "root_action_agreement_vs_strict": float(np.mean([x["same_action"] for x in r])),
            "root_value_lower_gap_mean": float(np.mean([x["lower_gap"] for x in r])),
            "root_value_lower_gap_max": float(np.max([x["lower_gap"] for x in r])),
            "root_value_upper_gap_mean": float(np.mean([x["upper_gap"] for x in r])),
            "max_depth_mean": float(np.mean([x["depth"] for x in r])), "waves_mean": float(np.mean([x["waves"] for x in r]))}
    with open(os.path.join(ROOT, "profiles", "r02d_opd_wave_quality.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
