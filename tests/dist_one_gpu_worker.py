"""Worker of tests/test_gpu_two_ranks_one_gpu.py: two ranks SHARING cuda:0 over the gloo backend (NCCL refuses two
ranks on one device), so the multi-rank host logic around the CUDA kernels runs on a single-GPU box too."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo")
    from oracle import envs as oenvs
    from oracle import planners
    from rl_agents_b200 import _lib
    from rl_agents_b200.distributed import (DistributedVI, ShardedOPD, merge_olop_root_statistics, merge_root_statistics,
                                            recommend, shard_range)
    from rl_agents_b200.engine.mcts import MCTSEngine, pcg64_words
    out = {}
    S, A, B = 2001, 4, 3
    P, N, R = oenvs.garnet(S, A, B, seed=8)
    term = np.zeros(S, bool)
    term[::97] = True
    q_ref, sweeps_ref = planners.value_iteration("sparse", P, R, term, 0.6, 80, nxt=N)
    b, e = shard_range(S, rank, world)
    slab = dict(transition=torch.as_tensor(P[b:e]).to(dev), reward=torch.as_tensor(R[b:e]).to(dev),
                terminal=torch.as_tensor(term[b:e].astype(np.uint8)).to(dev), nxt=torch.as_tensor(N[b:e].astype(np.int32)).to(dev))
    q, sweeps = DistributedVI("sparse", gamma=0.6, device=dev, tables_are_local=True, n_states=S, **slab).solve(80)
    out["vi_ok"] = bool(np.array_equal(q.cpu().numpy(), q_ref[b:e])) and sweeps == sweeps_ref
    q3, sweeps3 = DistributedVI("sparse", gamma=0.6, device=dev, tables_are_local=True, n_states=S, check_every=4, **slab).solve(80)
    out["vi_check_every_ok"] = bool(np.allclose(q3.cpu().numpy(), q_ref[b:e], rtol=1e-4, atol=1e-7)) and sweeps_ref <= sweeps3 <= sweeps_ref + 4
    # root-parallel MCTS merge
    words = oenvs.make_highway_state(3).pack()
    gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(11).spawn(world)[rank]))
    eng = MCTSEngine(_lib.ENV_HIGHWAY, 1, 5, 32, 6, 0.8, 10.0, device=dev)
    eng.plan(torch.tensor(words, dtype=torch.int32, device=dev).reshape(1, -1), pcg64_words(gen).reshape(1, -1))
    eng.finish()
    d = eng.tree_dict(0)
    counts = torch.zeros(5, dtype=torch.int32, device=dev)
    values = torch.zeros(5, dtype=torch.float64, device=dev)
    for c in range(d["first_child"][0], d["first_child"][0] + int(d["n_children"][0])):
        counts[d["action"][c]] = int(d["count"][c])
        values[d["action"][c]] = float(d["value"][c])
    mc, mv = merge_root_statistics(counts.cpu(), values.cpu())
    out["mcts_total"] = float(mc.sum().item())
    out["mcts_action"] = recommend(mc.numpy(), mv.numpy())
    oc, ou = merge_olop_root_statistics(counts.cpu(), values.cpu())
    out["olop_total"] = float(oc.sum().item())
    sharded = ShardedOPD(2000, 0.85, device=dev, wave_width=8).decide(words)
    out["sharded"] = [int(sharded["action"]), sharded["root_lower"], sharded["root_upper"], sharded["n_subtrees"]]
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
