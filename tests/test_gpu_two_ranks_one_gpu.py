"""The multi-rank paths on a single-GPU box: two ranks share cuda:0 over gloo (the NCCL / peer-memory variants need
two devices: tests/test_gpu_multi.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_sharing_one_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dist_one_gpu_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert all(r["vi_ok"] and r["vi_check_every_ok"] for r in res)
    assert res[0]["mcts_total"] == res[1]["mcts_total"] == 64.0 - 2 and res[0]["mcts_action"] == res[1]["mcts_action"]
    assert res[0]["olop_total"] == res[1]["olop_total"]
    assert res[0]["sharded"] == res[1]["sharded"]
    import torch
    from oracle import envs as oenvs
    from rl_agents_b200.distributed import ShardedOPD
    single = ShardedOPD(2000, 0.85, device="cuda:0", wave_width=8).decide(oenvs.make_highway_state(3).pack())
    assert res[0]["sharded"] == [int(single["action"]), single["root_lower"], single["root_upper"], single["n_subtrees"]]
