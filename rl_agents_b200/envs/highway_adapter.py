"""Hand-off of a LIVE `highway_env` highway-v0 environment to the device planners.

`preprocess_env` (rl_agents/agents/common/factory.py:97-116) gives the planner the env object itself (after
`simplify()`); the reference then deep-copies and steps it.  The device planners need the scene as the
136-word HighwayLite state instead (docs/HIGHWAY_LITE_SPEC.md): this module reads the upstream object model
by duck typing -- `env.unwrapped.vehicle` (controlled `MDPVehicle`), `env.unwrapped.road.vehicles`
(`IDMVehicle`s), their `position / heading / speed / crashed / lane_index / target_lane_index / target_speed /
timer / speed_index`, and `env.unwrapped.config` -- and packs it.

The search then runs on the HighwayLite MODEL of that scene (same structure as upstream: IDM + MOBIL traffic,
bicycle kinematics, 15 sub-steps per decision, highway-v0 reward), i.e. the device planner is a model-
predictive controller whose model is stated in the spec; the actions it returns are applied to the real env by
the caller as usual.  highway-env is not available in the build container, so this path is unit-tested with
objects carrying the upstream attribute names (tests/test_host.py), not against the package itself.
"""
import logging

import numpy as np

from rl_agents_b200 import _lib

logger = logging.getLogger(__name__)
V_SLOTS, N_LANES, LANE_WIDTH = 16, 4, 4.0
TARGET_SPEEDS = (20.0, 25.0, 30.0)


def looks_like_highway_env(u):
    road = getattr(u, "road", None)
    return road is not None and hasattr(road, "vehicles") and getattr(u, "vehicle", None) is not None


def _lane_id(index):
    """upstream lane indices are (from_node, to_node, lane_id) tuples"""
    if index is None:
        return None
    try:
        return int(index[2])
    except (TypeError, IndexError):
        return int(index)


def scene_from_highway_env(u, max_vehicles=V_SLOTS):
    """-> int32[136] HighwayLite words of the scene around `u.vehicle`."""
    cfg = getattr(u, "config", {}) or {}
    lanes = int(cfg.get("lanes_count", N_LANES))
    if lanes != N_LANES:
        raise TypeError("the HighwayLite kernel models %d lanes (env has lanes_count=%d)" % (N_LANES, lanes))
    act = cfg.get("action", {})
    if isinstance(act, dict) and act.get("type", "DiscreteMetaAction") != "DiscreteMetaAction":
        raise TypeError("the HighwayLite kernel models DiscreteMetaAction (env has %r)" % act.get("type"))
    ego = u.vehicle
    speeds = tuple(float(s) for s in getattr(ego, "target_speeds", TARGET_SPEEDS))
    if len(speeds) != 3 or any(abs(a - b) > 1e-6 for a, b in zip(speeds, TARGET_SPEEDS)):
        raise TypeError("the HighwayLite kernel models target speeds %r (env has %r)" % (TARGET_SPEEDS, speeds))
    ego_x = float(ego.position[0])
    others = [v for v in u.road.vehicles if v is not ego]
    others.sort(key=lambda v: abs(float(v.position[0]) - ego_x))
    if len(others) > max_vehicles - 1:
        logger.warning("HighwayLite keeps the %d vehicles nearest the ego (scene has %d); add "
                       "{'method': 'simplify'} to env_preprocessors" % (max_vehicles - 1, len(others)))
        others = others[:max_vehicles - 1]
    w = np.zeros(_lib.HW_STATE_WORDS, dtype=np.int32)
    f = w[:96].view(np.float32)
    for slot, v in enumerate([ego] + others):
        pos = v.position
        lane = _lane_id(getattr(v, "lane_index", None))
        if lane is None:
            lane = int(np.clip(np.rint(float(pos[1]) / LANE_WIDTH), 0, N_LANES - 1))
        tgt = _lane_id(getattr(v, "target_lane_index", None))
        speed = float(v.speed)
        f[0 * 16 + slot] = pos[0]
        f[1 * 16 + slot] = pos[1]
        f[2 * 16 + slot] = float(getattr(v, "heading", 0.0))
        f[3 * 16 + slot] = speed
        f[4 * 16 + slot] = float(getattr(v, "target_speed", speed))
        f[5 * 16 + slot] = float(getattr(v, "timer", 0.0))
        w[96 + slot] = int(np.clip(lane if tgt is None else tgt, 0, N_LANES - 1))
        w[112 + slot] = 1 | (2 if getattr(v, "crashed", False) else 0)
    si = getattr(ego, "speed_index", None)
    if si is None:
        si = int(np.argmin([abs(float(getattr(ego, "target_speed", ego.speed)) - s) for s in TARGET_SPEEDS]))
    w[129] = int(np.clip(si, 0, 2))
    f[4 * 16] = TARGET_SPEEDS[int(w[129])]
    # decisions elapsed: upstream counts `steps` (policy steps) or `time` in seconds at policy_frequency 1 Hz
    steps = getattr(u, "steps", None)
    if steps is None:
        steps = float(getattr(u, "time", 0.0)) * float(cfg.get("policy_frequency", 1))
    duration = float(cfg.get("duration", 40))
    w[128] = int(np.clip(round(float(steps) * 40.0 / duration) if duration > 0 else 0, 0, 40))
    return w
