#!/usr/bin/env python
"""Throughput of b2_highway_ttc_vi: ValueIterationAgent decisions (TTC-grid conversion + fixed point, shipped config:
iterations 10, gamma 1) on a batch of HighwayLite scenes resident in HBM; CUDA events, one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from rl_agents_b200.engine.ttc_vi import HighwayTTCVI
    from rl_agents_b200.envs.highway_lite import make_scene
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=1 << 18)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    base = np.stack([make_scene(s) for s in range(256)])
    scenes = torch.from_numpy(np.tile(base, (a.scenes // 256, 1))).cuda()
    eng = HighwayTTCVI(1.0, a.iterations)
    for _ in range(3):
        out = eng.solve(scenes, want_q=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = eng.solve(scenes, want_q=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    n = scenes.shape[0]
    print(json.dumps({"kernel": "highway_ttc_vi_kernel", "scenes": n, "iterations": a.iterations, "ms_per_launch": ms,
                      "decisions_per_s": n / (ms * 1e-3), "sweeps_mean": float(out["sweeps"].float().mean().item()),
                      "q_entry_updates_per_s": n * 600 * float(out["sweeps"].float().mean().item()) / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
