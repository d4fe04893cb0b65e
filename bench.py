#!/usr/bin/env python
"""Headline benchmark: OPD leaf-expansions/sec on HighwayLite (highway-v0 stand-in).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port of the reference)

Workload (BASELINE.json configs[1], C2): DeterministicPlannerAgent / OPD, budget
10 000 (=> 2 000 expand() calls per decision), gamma 0.8, on a batch of
`--trees` independent decisions (seeded scenes) per GPU -- eight search trees per
CTA, strict best-first order inside every tree (bit-exact with the reference).
A "step" is one plan() of the whole batch.  `value` = expand() calls per second
over all GPUs with the root scenes resident in HBM; `e2e` = the same through
the host-buffer path (pinned host scenes -> H2D -> search -> D2H of plans and
per-tree results) every step.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BUDGET = 10000
GAMMA = 0.8
N_ACTIONS = 5
STATE_BYTES = 136 * 4
NODE_BYTES = 5 * 4 + 3 * 8          # parent, first_child, depth, count, meta + reward, lower, upper


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trees", type=int, default=0, help="decisions per GPU per step (default 128 per SM, capped by free HBM)")
    ap.add_argument("--budget", type=int, default=BUDGET)
    ap.add_argument("--gamma", type=float, default=GAMMA)
    ap.add_argument("--keys-in-smem", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="OPD batch kernel variant (b2_opd_config.reserved)")
    ap.add_argument("--cpu-budget", type=int, default=2500, help="budget of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ----------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's plan() (reference/oracle is pure Python)
# ----------------------------------------------------------------------------
def _cpu_plan(args):
    seed, budget, gamma = args
    import numpy as np
    from oracle import envs as oenvs
    from oracle import planners
    t0 = time.perf_counter()
    _, t = planners.opd_plan(oenvs.HighwayLite(seed=seed), budget, gamma,
                             np_random=np.random.Generator(np.random.PCG64(np.random.SeedSequence(0))))
    return budget // N_ACTIONS, time.perf_counter() - t0


def cpu_baseline(budget, gamma):
    """One plan() of the oracle port (single core, as the reference planner is
    single-threaded Python) on the bounded sample `budget`."""
    n_exp, dt = _cpu_plan((0, budget, gamma))
    return {"value": n_exp / dt, "unit": "expansions/s", "cores": 1, "kind": "port",
            "sample": "1 plan() of oracle.planners.opd_plan on HighwayLite seed 0, budget %d (%d expansions, %.1f s)"
                      % (budget, n_exp, dt)}


def cpu_port_c(budget, gamma):
    """The C restatement (oracle/c: same spec, literal O(V^2) scans, heap frontier) on one core and on all
    host threads -- what an optimised CPU implementation of the same path does, next to the Python port."""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    from oracle import c_oracle
    from oracle import envs as oenvs
    scenes = [oenvs.make_highway_state(s).pack() for s in range(2 * (os.cpu_count() or 1))]
    c_oracle.opd_plan(scenes[0], 500, gamma)
    t0 = time.perf_counter()
    c_oracle.opd_plan(scenes[0], budget, gamma)
    single = (budget // N_ACTIONS) / (time.perf_counter() - t0)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda w: c_oracle.opd_plan(w, budget, gamma), scenes))
    multi = len(scenes) * (budget // N_ACTIONS) / (time.perf_counter() - t0)
    return {"value": multi, "unit": "expansions/s", "cores": cores, "single_core_value": single, "kind": "port",
            "sample": "oracle/c OPD, HighwayLite, budget %d (full C2 budget): 1 plan() on one core; %d plan()s on %d threads"
                      % (budget, len(scenes), cores)}


def run_reference(a):
    """--impl reference: oracle port on all host cores, one plan() per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    budget = a.cpu_budget
    pool = mp.Pool(cores)
    try:
        pool.map(_cpu_plan, [(s, 200, a.gamma) for s in range(cores)])           # spin-up (imports), untimed
        for w in range(min(a.warmup, 1)):
            pool.map(_cpu_plan, [(1000 + w * cores + s, budget, a.gamma) for s in range(cores)])
        t0 = time.perf_counter()
        total = 0
        for k in range(a.steps):
            out = pool.map(_cpu_plan, [(k * cores + s, budget, a.gamma) for s in range(cores)])
            total += sum(n for n, _ in out)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    value = total / dt
    sample = ("%d processes x 1 plan() per step of the oracle port (oracle.planners.opd_plan, HighwayLite, "
              "budget %d = %d expansions each; the reference is O(budget^2), full budget %d is slower per expansion)"
              % (cores, budget, budget // N_ACTIONS, a.budget))
    print(json.dumps({
        "impl": "reference", "metric": "OPD leaf-expansions/sec on highway-v0 (HighwayLite)", "value": value,
        "unit": "expansions/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "OPD plan() on HighwayLite (highway-v0 stand-in), gamma %g, CPU sample budget %d"
                               % (a.gamma, budget), "budget": budget, "gamma": a.gamma},
        "cpu_baseline": {"value": value, "unit": "expansions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "expansions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rl_agents_b200 import _lib
    from rl_agents_b200.engine.opd import OPDEngine
    from rl_agents_b200.envs.highway_lite import make_scene

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    n_exp = a.budget // N_ACTIONS
    if a.trees:
        trees = a.trees
    else:
        # default batch: 128 decisions per SM (more work in flight = better overlap of the trees' phases),
        # capped so that the tree arena (scene + node record + frontier key per node) takes <= 65 % of the free HBM
        per_tree = (1 + n_exp * N_ACTIONS) * (STATE_BYTES + NODE_BYTES + 8) + 16 * 1024
        free, _ = torch.cuda.mem_get_info(dev)
        trees = min(128 * sms, int(0.65 * free / per_tree))
        trees = max(8 * sms, trees // (8 * sms) * (8 * sms))

    eng = OPDEngine(_lib.ENV_HIGHWAY, trees, N_ACTIONS, a.budget, a.gamma, keys_in_smem=bool(a.keys_in_smem),
                    device=dev, kernel=a.kernel)
    # independent decisions: every (rank, tree) its own seeded scene; two alternating input sets
    host_scenes = [torch.from_numpy(np.stack([make_scene(1_000_000 * s + rank * trees + i) for i in range(trees)]))
                   .pin_memory() for s in range(2)]
    dev_scenes = [h.to(dev) for h in host_scenes]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            fn(k)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def device_step(k):
        eng.plan(dev_scenes[k & 1])

    for k in range(max(a.warmup, 3)):
        device_step(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(device_step, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    # sanity: the timed work really is the full search
    res = eng.result.cpu().numpy()
    assert (res[:, 0] > n_exp).all() and (res[:, 4] == 0).all()
    mean_children = float((res[:, 0] - 1).mean() / n_exp)
    # ---- e2e: the host-buffer C ABI (b2_opd_create / b2_opd_plan_host): pinned host scenes -> H2D -> search
    #      -> D2H of plans and per-tree results, synchronous, every step ----
    import ctypes
    capacity, plan_capacity = eng.capacity, eng.plan_capacity
    del eng
    torch.cuda.empty_cache()
    hc = _lib.OPDHostConfig(_lib.ENV_HIGHWAY, trees, N_ACTIONS, a.budget, int(a.keys_in_smem), a.kernel, a.gamma, 0.0,
                            _lib.FiniteMDP())
    handle = ctypes.c_void_p()
    _lib.check(lib.b2_opd_create(ctypes.byref(hc), ctypes.byref(handle)))
    plan_host = torch.empty((trees, plan_capacity), dtype=torch.int8).pin_memory()
    res_host = torch.empty((trees, _lib.OPD_RESULT_WORDS), dtype=torch.int32).pin_memory()

    def e2e_step(k):
        _lib.check(lib.b2_opd_plan_host(handle, ctypes.c_void_p(host_scenes[k & 1].data_ptr()),
                                        ctypes.c_void_p(plan_host.data_ptr()), ctypes.c_void_p(res_host.data_ptr())))

    for k in range(2):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        e2e_step(k)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item()) * 1e3
    assert (res_host.numpy()[:, 0] > n_exp).all() and (res_host.numpy()[:, 4] == 0).all()
    lib.b2_opd_destroy(handle)
    total_exp = float(world) * trees * n_exp * a.steps
    value = total_exp / (ms * 1e-3)
    e2e_value = total_exp / (ms_e2e * 1e-3)
    # roofline of the dominant (only) kernel, opd_highway_kernel: algorithmic HBM bytes per expand()
    # = parent scene read + children scenes written + node records + keys + bottom-up pass
    bytes_per_exp = STATE_BYTES * (1.0 + mean_children) + mean_children * (NODE_BYTES + 8 + 20) + 24
    launch_ms = ms / a.steps
    achieved = trees * n_exp * bytes_per_exp / (launch_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "opd_highway_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("trees") == trees and tr.get("budget") == a.budget:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass

    ncu = None      # static evidence from the committed ncu capture of this kernel (not measured in this run)
    try:
        with open(os.path.join(ROOT, "profiles", "r01_opd_highway_multi_ncu_summary.json")) as f:
            summ = json.load(f)
        ncu = {"source": "profiles/r01_opd_highway_multi_ncu_summary.json",
               "ipc_per_sm": float(summ["sm__inst_executed.avg.per_cycle_elapsed"][0]),
               "issue_slots_busy_pct": float(summ["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "alu_pipe_pct": float(summ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][0]),
               "fma_pipe_pct": float(summ["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][0]),
               "dram_pct": float(summ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0])}
    except Exception:
        pass
    out = {
        "metric": "OPD leaf-expansions/sec on highway-v0 (HighwayLite)", "value": value, "unit": "expansions/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: DeterministicPlannerAgent (OPD) on HighwayLite (highway-v0 stand-in, 16 vehicles, "
                               "15 sub-steps/step), budget %d, gamma %g, %d independent decisions per GPU per step, "
                               "strict best-first per tree" % (a.budget, a.gamma, trees),
                   "budget": a.budget, "gamma": a.gamma, "trees_per_gpu": trees, "expansions_per_tree": n_exp,
                   "mean_children_per_expansion": mean_children, "child_nodes_per_s": value * mean_children,
                   "l2": "working set %.1f GB per step >> 126 MB L2 (no flush needed)"
                         % (trees * capacity * (STATE_BYTES + NODE_BYTES) / 1e9),
                   "parallelism": "trees sharded over %d GPU(s), no data-path collective" % world,
                   "keys_in_smem": bool(a.keys_in_smem)},
        "e2e": {"value": e2e_value, "unit": "expansions/s", "h2d_bytes_per_step": int(trees * STATE_BYTES),
                "d2h_bytes_per_step": int(plan_host.numel() + res_host.numel() * 4), "ms_per_step": ms_e2e / a.steps,
                "path": "b2_opd_plan_host (C ABI, host buffers, synchronous); wall clock over the steps, max over ranks"},
        "gpu_launches": a.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "opd_highway_multi_kernel",
                     "bytes_per_expansion": bytes_per_exp, "limiter": "instruction issue (see ncu)", "ncu": ncu,
                     "note": "latency/FP32-issue bound by construction (15 dependent sub-steps per child); "
                             "HBM fraction reported as the contract asks, see DESIGN.md section 4"},
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a.cpu_budget, a.gamma)
        try:
            out["cpu_port_c"] = cpu_port_c(a.budget, a.gamma)
        except Exception as e:      # the C oracle is optional test infrastructure
            out["cpu_port_c"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
