"""The C ABI used from plain C (no torch, no CUDA calls in the caller): compile
examples/opd_host_example.c against libb2planner.so, run it, and check its output against the
Python engine and the oracle on the same tables."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tables():
    S, A = 64, 4
    x = 12345
    T = np.zeros(S * A, dtype=np.int64)
    R = np.zeros(S * A)
    for i in range(S * A):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        T[i] = (x >> 8) % S
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        R[i] = ((x >> 8) % 1000) / 1000.0
    term = np.array([s % 17 == 5 for s in range(S)])
    return T.reshape(S, A), R.reshape(S, A), term


def test_plain_c_caller_matches_engine_and_oracle(tmp_path):
    from oracle import envs as oenvs
    from oracle import planners
    exe = str(tmp_path / "opd_host_example")
    lib_dir = os.path.join(ROOT, "rl_agents_b200", "csrc")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "opd_host_example.c"),
                    "-L" + lib_dir, "-lb2planner", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()
    T, R, term = tables()
    for line, root in zip(out, [0, 7, 21, 63]):
        m = re.match(r"tree \d+ nodes (\d+) leaves (\d+) depth (\d+) tie (-?\d+) root_count (\d+) lower (\S+) upper (\S+) plan(.*)", line)
        nodes, leaves, depth, tie, count = [int(m.group(i)) for i in range(1, 6)]
        lower, upper = float(m.group(6)), float(m.group(7))
        plan_c = [int(a) for a in m.group(8).split()]
        plan, t = planners.opd_plan(oenvs.FiniteMDPLite(T, R, term, state=root), 500, 0.9,
                                    np_random=np.random.default_rng(0))
        assert nodes == len(t) and leaves == t.n_leaves and count == t.count[0]
        assert lower == t.lower[0] and upper == t.upper[0]          # C pow() tables == Python float ** tables
        assert depth == max(t.depth)
        assert plan_c == plan[:len(plan_c)] and (tie >= 0 or plan_c == plan)


def test_host_api_from_python_matches_device_api():
    import ctypes
    import torch
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDEngine
    from rl_agents_b200.envs.highway_lite import make_scene
    lib = _lib.load()
    n, budget, gamma = 24, 300, 0.8
    scenes = np.stack([make_scene(200 + i) for i in range(n)])
    hc = _lib.OPDHostConfig(_lib.ENV_HIGHWAY, n, 5, budget, 0, 0, gamma, 0.0, _lib.FiniteMDP())
    handle = ctypes.c_void_p()
    _lib.check(lib.b2_opd_create(ctypes.byref(hc), ctypes.byref(handle)))
    cap = lib.b2_opd_plan_capacity(handle)
    plan = np.zeros((n, cap), dtype=np.int8)
    result = np.zeros((n, _lib.OPD_RESULT_WORDS), dtype=np.int32)
    _lib.check(lib.b2_opd_plan_host(handle, scenes.ctypes.data_as(ctypes.c_void_p), plan.ctypes.data_as(ctypes.c_void_p),
                                    result.ctypes.data_as(ctypes.c_void_p)))
    eng = OPDEngine(_lib.ENV_HIGHWAY, n, 5, budget, gamma)
    eng.plan(torch.from_numpy(scenes).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(result[:, :7], eng.result.cpu().numpy()[:, :7])
    for i in range(n):
        assert np.array_equal(plan[i, :result[i, 5]], eng.plan_buf[i, :result[i, 5]].cpu().numpy())
    lower = np.zeros(int(result[3, 0]))
    _lib.check(lib.b2_opd_copy_tree(handle, 3, int(result[3, 0]), None, None, None, None, None,
                                    lower.ctypes.data_as(ctypes.c_void_p), None))
    assert np.array_equal(lower, eng.lower[3, :len(lower)].cpu().numpy())
    lib.b2_opd_destroy(handle)
