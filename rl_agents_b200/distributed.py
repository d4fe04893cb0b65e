"""Multi-GPU partitioning of the planning path: one process per GPU
(torch.distributed, NCCL on the GPU box / gloo in CPU tests).

  OPD / batched decisions   independent trees are sharded over ranks; no
                            data-path collective (results gathered at the end).
  MCTS root parallelism     every rank grows its own tree on episodes/world
                            episodes from the same root with an independent RNG
                            stream; ONE all-reduce of the root's per-action
                            (count, count*value) decides the action.
  Value iteration           state slabs: rank g owns rows [g*S/G, (g+1)*S/G) of
                            P/N/R/Q; after every sweep the V slabs are
                            all-gathered and the allclose violation counter is
                            all-reduced (the only exchange step of the path).
The helpers below work on CPU tensors too, which is how the gloo tests cover them.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slab_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def allgather_slabs(full, n, group=None):
    """In place: every rank contributes full[begin:end] of its own slab and receives the rest."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = slab_sizes(n, world)
    if len(set(sizes)) == 1:
        b, e = shard_range(n, rank, world)
        dist.all_gather_into_tensor(full[:n], full[b:e].clone(), group=group)
        return full
    pad = max(sizes)
    b, e = shard_range(n, rank, world)
    mine = torch.zeros(pad, dtype=full.dtype, device=full.device)
    mine[:e - b] = full[b:e]
    out = torch.empty(pad * world, dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    for r in range(world):
        rb, re = shard_range(n, r, world)
        full[rb:re] = out[r * pad:r * pad + (re - rb)]
    return full


class DistributedVI(object):
    """Slab-sharded value iteration (value_iteration.py:42-73 over G GPUs)."""

    def __init__(self, mode, transition, reward, terminal, nxt=None, gamma=1.0, device="cuda", group=None,
                 tables_are_local=False):
        import torch.distributed as dist
        from rl_agents_b200.engine.vi import VIEngine
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if tables_are_local:
            raise NotImplementedError
        S = np.asarray(reward).shape[0]
        self.n_states = S
        b, e = shard_range(S, self.rank, self.world)
        self.engine = VIEngine(mode, np.asarray(transition)[b:e], np.asarray(reward)[b:e], np.asarray(terminal)[b:e],
                               nxt=None if nxt is None else np.asarray(nxt)[b:e], gamma=gamma, device=device,
                               row_begin=b, row_end=e, n_states=S)

    def solve(self, iterations):
        """Returns (this rank's Q slab on device, sweeps).  No host sync inside the loop:
        the sweep kernels read the all-reduced violation counters from device memory."""
        eng = self.engine
        eng.reset(iterations)
        for k in range(iterations):
            eng.sweep(k)
            allgather_slabs(eng.v[(k + 1) & 1], self.n_states, self.group)
            self.dist.all_reduce(eng.viol[k:k + 1], group=self.group)
        return eng.result(iterations)


def merge_root_statistics(counts, values, group=None):
    """MCTS root parallelisation: counts/values are [A] tensors of THIS rank's root
    children (count, mean value).  One all-reduce of [2, A]; returns the merged
    (count, mean value) on every rank."""
    import torch
    import torch.distributed as dist
    packed = torch.stack([counts.to(torch.float64), counts.to(torch.float64) * values.to(torch.float64)])
    dist.all_reduce(packed, group=group)
    merged_counts = packed[0]
    merged_values = torch.where(merged_counts > 0, packed[1] / merged_counts.clamp(min=1), torch.zeros_like(packed[1]))
    return merged_counts, merged_values


def recommend(counts, values):
    """MCTSNode.selection_rule (mcts.py:212-218) on merged statistics: most visited,
    ties -> highest value (first)."""
    counts = np.asarray(counts)
    values = np.asarray(values)
    ties = np.nonzero(counts == counts.max())[0]
    return int(max(ties, key=lambda i: values[i]))
