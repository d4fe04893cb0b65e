"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_two_rank_vi_and_root_parallel_mcts():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert all(r["vi_ok"] for r in res)
    assert res[0]["vi_sweeps"] == res[1]["vi_sweeps"] < 80          # converged early, same sweep on both ranks
    # every episode of both trees counted once; the episode that expands a root descends into no child
    assert res[0]["mcts_total"] == res[1]["mcts_total"] == 64.0 - 2
    assert res[0]["mcts_action"] == res[1]["mcts_action"]
