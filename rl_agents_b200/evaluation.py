"""Batched closed-loop evaluation of the planners on HighwayLite: the working version of the
reference's disabled budget-sweep harness (scripts/planners_evaluation.py:287-289, SURVEY 8f
rank 4) in the shape the GPU wants -- all episodes advance in lock-step, every decision step is
ONE batched plan() launch over all live episodes and ONE batched env transition."""
import time

import numpy as np

from rl_agents_b200 import _lib
from rl_agents_b200.envs.highway_lite import make_scene


def _np_random(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def run_batched_episodes(planner, seeds, budget, gamma, max_steps=40, device="cuda", planner_seed=0, **kw):
    """planner: "opd" | "mcts" | "olop" | "vi" (ValueIterationAgent on the scenes' TTC-grid MDPs, `budget` = its
    `iterations`).  Every episode: scene make_scene(seed), replanning at every
    step (receding_horizon 1, step_strategy reset -- the reference defaults), until crash or `max_steps`.
    Returns dict(returns, lengths, crashed, decision_ms)."""
    import torch
    from rl_agents_b200.engine.mcts import MCTSEngine, pcg64_words, set_pcg64_words
    from rl_agents_b200.engine.olop import OLOPEngine
    from rl_agents_b200.engine.opd import OPDEngine
    from rl_agents_b200.agents.tree_search.mcts import allocation
    lib = _lib.load()
    dev = torch.device(device)
    n = len(seeds)
    scenes = torch.from_numpy(np.stack([make_scene(s) for s in seeds])).to(dev)
    rngs = [_np_random(planner_seed + i) for i in range(n)]
    if planner == "opd":
        eng = OPDEngine(_lib.ENV_HIGHWAY, n, 5, budget, gamma, kw.get("terminal_reward", 0.0), device=dev)
    elif planner == "mcts":
        episodes, horizon = allocation(budget, gamma)
        from rl_agents_b200.agents.tree_search.mcts import MCTS
        # the reference's default temperature comes from the CLASS default gamma (mcts.py:120-127), not the configured one
        eng = MCTSEngine(_lib.ENV_HIGHWAY, n, 5, episodes, horizon, gamma,
                         kw.get("temperature", MCTS.default_config()["temperature"]), device=dev)
    elif planner == "olop":
        episodes, horizon = allocation(max(5, budget), gamma)
        ub = kw.get("upper_bound", {"type": "kullback-leibler", "time": "global", "threshold": "2*np.log(time)"})
        eng = OLOPEngine(_lib.ENV_HIGHWAY, n, 5, episodes, horizon, gamma, ub, kw.get("continuation_type", "uniform"),
                         device=dev)
    elif planner == "vi":
        from rl_agents_b200.engine.ttc_vi import HighwayTTCVI
        eng = HighwayTTCVI(gamma, budget, device=dev)
    else:
        raise ValueError("unknown planner %r" % planner)
    returns = np.zeros(n)
    lengths = np.zeros(n, dtype=int)
    alive = np.ones(n, dtype=bool)
    crashed = np.zeros(n, dtype=bool)
    rew = torch.empty(n, dtype=torch.float32, device=dev)
    flg = torch.empty(n, dtype=torch.int32, device=dev)
    actions_log = []
    t_plan = 0.0
    for step in range(max_steps):
        if not alive.any():
            break
        t0 = time.perf_counter()
        if planner == "opd":
            eng.plan(scenes)
            plans, _ = eng.finish(rngs)
        elif planner == "vi":
            plans = [[int(a)] for a in eng.solve(scenes, want_q=False)["action"].cpu().numpy()]
        else:
            eng.plan(scenes, np.stack([pcg64_words(g) for g in rngs]))
            plans, _, words = eng.finish()
            for g, w in zip(rngs, words):
                set_pcg64_words(g, w)
        t_plan += time.perf_counter() - t0
        act = np.array([p[0] if p else 1 for p in plans], dtype=np.int32)      # empty plan: IDLE
        actions_log.append(act.copy())
        _lib.check(lib.b2_highway_step(_lib.ptr(scenes), _lib.ptr(torch.from_numpy(act).to(dev)), _lib.ptr(rew),
                                       _lib.ptr(flg), None, n, _lib.current_stream()))
        r, f = rew.cpu().numpy(), flg.cpu().numpy()
        returns += np.where(alive, r, 0.0)
        lengths += alive
        done = (f & 3) != 0
        crashed |= alive & ((f & 1) != 0)
        alive &= ~done
    return {"returns": returns, "lengths": lengths, "crashed": crashed, "actions": np.array(actions_log).T,
            "decision_ms": 1e3 * t_plan / max(len(actions_log), 1), "n": n}
