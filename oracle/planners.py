"""Flat CPU restatements of the reference planners -- TEST INFRASTRUCTURE.

Each function restates one reference algorithm over struct-of-arrays lists
(node id = creation order) instead of the reference's Node object graph, and
is pinned against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> tests/golden/*.json, tests/test_oracle.py).

Envs are driven exactly like the reference drives them: deep copy + step.
"""
import copy
import math

import numpy as np


def _available_actions(state):
    # deterministic.py:32-35, olop.py:168-171, mcts.py:69-72
    try:
        return list(state.get_available_actions())
    except AttributeError:
        return list(range(state.action_space.n))


class Tree(object):
    """SoA tree dump shared by the oracle and the CUDA parity tests."""

    def __init__(self):
        self.parent = []
        self.action = []
        self.depth = []
        self.count = []
        self.first_child = []
        self.n_children = []

    def children(self, i):
        return range(self.first_child[i], self.first_child[i] + self.n_children[i])

    def __len__(self):
        return len(self.parent)


# --------------------------------------------------------------------------
# OPD -- rl_agents/agents/tree_search/deterministic.py
# --------------------------------------------------------------------------
def opd_plan(env, budget, gamma, terminal_reward=0.0, np_random=None, on_expansion=None):
    """OptimisticDeterministicPlanner.plan (deterministic.py:116-122) after a
    reset (deterministic.py:102-104).  Returns (plan, tree).  `on_expansion` (bench.py's
    time-boxed CPU arm) is called after every expand(); it does not touch the search."""
    t = Tree()
    t.reward, t.lower, t.upper, t.done = [], [], [], []
    states = []

    def new_node(parent, action, depth, state):
        # DeterministicNode.__init__ (deterministic.py:10-19): count = 1
        t.parent.append(parent)
        t.action.append(action)
        t.depth.append(depth)
        t.count.append(1)
        t.first_child.append(-1)
        t.n_children.append(0)
        t.reward.append(0.0)
        t.lower.append(0.0)
        t.upper.append(0.0)
        t.done.append(False)
        states.append(state)
        return len(t.parent) - 1

    root = new_node(-1, -1, 0, env)
    leaves = [root]
    terminal_expansions = 0
    for _ in range(int(budget) // env.action_space.n):          # :118
        # run(): first arg-max of value_upper over the leaves list (:110)
        best = leaves[0]
        for i in leaves[1:]:
            if t.upper[i] > t.upper[best]:
                best = i
        if t.done[best]:
            terminal_expansions += 1                            # :111-112 (warning)
        # expand() (:28-43)
        leaves.remove(best)
        actions = _available_actions(states[best])
        t.first_child[best] = len(t.parent)
        t.n_children[best] = len(actions)
        d = t.depth[best] + 1
        for a in actions:
            c = new_node(best, a, d, copy.deepcopy(states[best]))
            _, reward, done, _, _ = states[c].step(a)
            leaves.append(c)
            # update() (:45-65)
            if not (0 <= reward <= 1):
                raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
            t.reward[c] = reward
            t.done[c] = bool(done)
            t.lower[c] = t.lower[best] + (gamma ** (d - 1)) * reward
            t.upper[c] = t.lower[c] + (gamma ** d) / (1 - gamma)
            if done:
                t.lower[c] = t.upper[c] = t.lower[c] + terminal_reward * (gamma ** d) / (1 - gamma)
            n = c
            while n >= 0:                                       # sequence(): self..root
                t.count[n] += 1
                n = t.parent[n]
        states[best] = None
        # backup_to_root() (:74-79)
        n = best
        while n >= 0:
            t.lower[n] = max(t.lower[c] for c in t.children(n))
            t.upper[n] = max(t.upper[c] for c in t.children(n))
            n = t.parent[n]
        if on_expansion is not None:
            on_expansion()
    t.terminal_expansions = terminal_expansions
    t.n_leaves = len(leaves)
    return greedy_plan(t, t.lower, np_random), t


def opd_plan_wavefront(env, budget, gamma, width, terminal_reward=0.0, np_random=None):
    """SPECIFICATION of the device's wavefront OPD (b2_opd_plan_wave), not a reference algorithm: the
    reference expands ONE leaf per iteration (deterministic.py:106-122); the wavefront takes, per wave,
    the k = min(width, expansions left, frontier size) best leaves in the order (value_upper descending,
    node id ascending) -- the reference's own arg-max order, deterministic.py:110 -- and expands them in
    increasing node-id order (children in action order, ids in creation order).  width = 1 IS the
    reference's algorithm.  Node bounds, counts and the backup are the reference's (update :45-65,
    backup_to_root :74-79): they do not depend on the expansion order.  Returns (plan, tree); tree.waves
    lists the leaves of every wave."""
    t = Tree()
    t.reward, t.lower, t.upper, t.done = [0.0], [0.0], [0.0], [False]
    t.parent, t.action, t.depth, t.count, t.first_child, t.n_children = [-1], [-1], [0], [1], [-1], [0]
    states = [env]
    leaves = {0}                                     # the frontier; its order is given by the sort key below
    remaining = int(budget) // env.action_space.n
    t.waves, t.terminal_expansions = [], 0
    while remaining > 0:
        k = min(int(width), remaining, len(leaves))
        chosen = sorted(sorted(leaves, key=lambda i: (-t.upper[i], i))[:k])
        t.waves.append(chosen)
        for best in chosen:
            t.terminal_expansions += 1 if t.done[best] else 0
            leaves.remove(best)
            actions = _available_actions(states[best])
            t.first_child[best] = len(t.parent)
            t.n_children[best] = len(actions)
            d = t.depth[best] + 1
            for a in actions:
                st = copy.deepcopy(states[best])
                _, reward, done, _, _ = st.step(a)
                if not (0 <= reward <= 1):
                    raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
                lo = t.lower[best] + (gamma ** (d - 1)) * reward
                up = lo + (gamma ** d) / (1 - gamma)
                if done:
                    lo = up = lo + terminal_reward * (gamma ** d) / (1 - gamma)
                c = len(t.parent)
                t.parent.append(best); t.action.append(a); t.depth.append(d); t.count.append(1)
                t.first_child.append(-1); t.n_children.append(0)
                t.reward.append(reward); t.lower.append(lo); t.upper.append(up); t.done.append(bool(done))
                states.append(st)
                leaves.add(c)
            states[best] = None
        remaining -= k
    # counts (:64-65: +1 on self and every ancestor per created node) and backups (:74-79), bottom-up
    for c in range(len(t.parent) - 1, 0, -1):
        n = c
        while n >= 0:
            t.count[n] += 1
            n = t.parent[n]
    for n in range(len(t.parent) - 1, -1, -1):
        if t.n_children[n]:
            t.lower[n] = max(t.lower[c] for c in t.children(n))
            t.upper[n] = max(t.upper[c] for c in t.children(n))
    t.n_leaves = len(leaves)
    return greedy_plan(t, t.lower, np_random), t


def robust_plan(envs, budget, gamma, terminal_reward=0.0, np_random=None, width=1):
    """DiscreteRobustPlanner.plan (rl_agents/agents/robust/robust.py:28-47, DROP) on the joint env of M models
    (JointEnv, robust.py:9-26).  Every node carries one value_lower / value_upper PER MODEL along its own
    path (deterministic.py:52-59: vector rewards and terminals); a leaf's bounds are the minima over the
    models (RobustNode, robust.py:40-47), the frontier arg-max and the backups work on those minima
    (deterministic.py:74-79 via get_value_*_bound).  Children follow JointEnv.get_available_actions: the
    set union of the models' available actions, i.e. ascending action ids.  width > 1: the device's
    wavefront (see opd_plan_wavefront); width = 1 is the reference's algorithm.  Returns (plan, tree);
    tree.lower / tree.upper are the robust (min over models) bounds, tree.lowerv the per-model path sums."""
    M = len(envs)
    n_actions = envs[0].action_space.n
    t = Tree()
    t.parent, t.action, t.depth, t.count, t.first_child, t.n_children = [-1], [-1], [0], [1], [-1], [0]
    t.lower, t.upper, t.lowerv, t.done = [0.0], [0.0], [[0.0] * M], [[False] * M]
    states = [list(envs)]
    leaves = {0}
    remaining = int(budget) // n_actions
    t.waves = []
    while remaining > 0:
        k = min(int(width), remaining, len(leaves))
        chosen = sorted(sorted(leaves, key=lambda i: (-t.upper[i], i))[:k])
        t.waves.append(chosen)
        for best in chosen:
            leaves.remove(best)
            actions = sorted(set().union(*[_available_actions(s) for s in states[best]]))
            t.first_child[best], t.n_children[best] = len(t.parent), len(actions)
            d = t.depth[best] + 1
            for a in actions:
                models = [copy.deepcopy(s) for s in states[best]]
                lov, upv, dn = [], [], []
                for m, st in enumerate(models):
                    _, reward, done, _, _ = st.step(a)
                    if not (0 <= reward <= 1):
                        raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
                    lo = t.lowerv[best][m] + (gamma ** (d - 1)) * reward
                    up = lo + (gamma ** d) / (1 - gamma)
                    if done:
                        lo = up = lo + terminal_reward * (gamma ** d) / (1 - gamma)
                    lov.append(lo); upv.append(up); dn.append(bool(done))
                c = len(t.parent)
                t.parent.append(best); t.action.append(a); t.depth.append(d); t.count.append(1)
                t.first_child.append(-1); t.n_children.append(0)
                t.lowerv.append(lov); t.done.append(dn)
                t.lower.append(min(lov)); t.upper.append(min(upv))
                states.append(models)
                leaves.add(c)
            states[best] = None
        remaining -= k
    for c in range(len(t.parent) - 1, 0, -1):
        n = c
        while n >= 0:
            t.count[n] += 1
            n = t.parent[n]
    for n in range(len(t.parent) - 1, -1, -1):
        if t.n_children[n]:
            t.lower[n] = max(t.lower[c] for c in t.children(n))
            t.upper[n] = max(t.upper[c] for c in t.children(n))
    t.n_leaves = len(leaves)
    return greedy_plan(t, t.lower, np_random), t


def graph_based_plan(env, observation, budget, gamma, np_random, accuracy=1e-2, sampling_timeout=100):
    """GraphBasedPlanner.plan (GBOP-D, rl_agents/agents/tree_search/graph_based.py:84-138) on an env whose
    observations are hashable state ids.  One graph node per state with value_lower / value_upper (:12-20); an epoch
    walks from the root along the optimistic action (:22-30, random tie-break through the planner RNG) to a node
    without children, expands it (:38-52; `done` is ignored there) and runs partial_value_iteration (:63-75): a FIFO
    of nodes whose two bounds are re-backed-up, pushing a node's parents while the change exceeds `accuracy`.
    The reference pushes `list(node.parents)` -- a Python set of objects, i.e. an order that depends on memory
    addresses; this restatement (and the device planner) pushes the parents in ASCENDING state id.  With
    accuracy = 0 the iteration runs to its fixed point, which does not depend on that order.
    Returns (plan, nodes) with nodes[s] = dict(lower, upper, expanded)."""
    A = env.action_space.n
    nodes = {}

    def get_node(o, state=None):
        if o not in nodes:
            nodes[o] = dict(state=state, lower=0.0, upper=1 / (1 - gamma), rewards={}, children={}, parents=set())
        if state is not None and nodes[o]["state"] is None:
            nodes[o]["state"] = state
        return nodes[o]

    def backup(n, field):
        return {a: n["rewards"][a] + gamma * nodes[n["children"][a]][field] for a in n["children"]}

    root = get_node(observation, env)
    for _ in range(int(budget) // A):
        o = observation
        for _k in range(sampling_timeout):
            n = nodes[o]
            if not n["children"]:
                for a in _available_actions(n["state"]):                      # expand (:38-52)
                    st = copy.deepcopy(n["state"])
                    nxt, reward, _done = st.step(a)[:3]
                    child = get_node(nxt)
                    child["state"] = st
                    child["parents"].add(o)
                    n["rewards"][a] = reward
                    n["children"][a] = nxt
                queue = [o]                                                   # partial_value_iteration (:63-75)
                while queue:
                    m = nodes[queue.pop(0)]
                    delta = 0
                    for field in ("lower", "upper"):
                        bound = np.amax(list(backup(m, field).values()))
                        delta = max(delta, abs(m[field] - bound))
                        m[field] = bound
                    if delta > accuracy:
                        queue.extend(sorted(m["parents"]))
                break
            q = backup(n, "upper")                                            # sampling_rule (:22-30)
            actions = list(q.keys())
            x = np.array([q[a] for a in actions])
            indices = np.nonzero(x == np.amax(x))[0]
            o = n["children"][actions[np_random.choice(indices)]]
    plan, n = [], root
    for _ in range(sampling_timeout):                                         # get_plan (:126-135)
        if not n["children"]:
            break
        ql = backup(n, "lower")
        a = max(ql.items(), key=lambda kv: kv[1])[0]
        plan.append(a)
        n = nodes[n["children"][a]]
    return plan, {o: dict(lower=float(n["lower"]), upper=float(n["upper"]), expanded=bool(n["children"])) for o, n in nodes.items()}


def greedy_plan(t, values, np_random):
    """AbstractPlanner.get_plan (abstract.py:143-156) with
    DeterministicNode.selection_rule (deterministic.py:21-26): arg-max of the
    children values, uniform tie-break through the planner RNG
    (abstract.py:296-311)."""
    plan = []
    node = 0
    while t.n_children[node] > 0:
        kids = list(t.children(node))
        x = np.array([values[c] for c in kids])
        indices = np.nonzero(x == np.amax(x))[0]
        index = np_random.choice(indices)
        plan.append(t.action[kids[index]])
        node = kids[index]
    return plan


# --------------------------------------------------------------------------
# MCTS -- rl_agents/agents/tree_search/mcts.py (open loop, K = 1)
# --------------------------------------------------------------------------
def olop_horizon(episodes, gamma):
    # olop.py:42-44
    return max(int(np.ceil(np.log(episodes) / (2 * np.log(1 / gamma)))), 1)


def olop_allocation(budget, gamma):
    # olop.py:50-62
    for episodes in range(1, int(budget)):
        if episodes * olop_horizon(episodes, gamma) > budget:
            episodes = max(episodes - 1, 1)
            horizon = olop_horizon(episodes, gamma)
            break
    else:
        raise ValueError("Could not split budget {} with gamma {}".format(budget, gamma))
    return episodes, horizon


def _policy(policy_config, state):
    """MCTSAgent.policy_factory policies (mcts.py:34-97) -> (actions, probs)."""
    kind = policy_config["type"]
    if kind == "random":
        actions = np.arange(state.action_space.n)
        return actions, np.ones(len(actions)) / len(actions)
    if hasattr(state, "get_available_actions"):
        available = state.get_available_actions()
    else:
        available = np.arange(state.action_space.n)
    uniform = np.ones(len(available)) / len(available)
    if kind == "random_available":
        return available, uniform
    if kind == "preference":
        for i in range(len(available)):
            if available[i] == policy_config["action"]:
                ratio = policy_config.get("ratio", 2)
                p = np.ones(len(available)) / (len(available) - 1 + ratio)
                p[i] *= ratio
                return available, p
        return available, uniform
    raise ValueError("Unknown policy type")


def mcts_reroot(t, action):
    """AbstractPlanner.step_by_subtree (abstract.py:195-206): the sub-tree under the root's child `action`
    becomes the tree (re-indexed breadth first); None when that action was never expanded."""
    child = next((c for c in t.children(0) if t.action[c] == action), None)
    if child is None:
        return None
    order, new_parent = [child], [-1]
    head = 0
    while head < len(order):
        for c in t.children(order[head]):
            order.append(c)
            new_parent.append(head)
        head += 1
    pos = {old: new for new, old in enumerate(order)}
    s = Tree()
    s.parent = new_parent
    s.action = [-1] + [t.action[o] for o in order[1:]]
    s.depth = [t.depth[o] - t.depth[child] for o in order]
    s.count = [t.count[o] for o in order]
    s.n_children = [t.n_children[o] for o in order]
    s.first_child = [pos[t.first_child[o]] if t.n_children[o] > 0 else -1 for o in order]
    s.value = [t.value[o] for o in order]
    s.prior = [t.prior[o] for o in order]
    return s


def mcts_plan(env, episodes, horizon, gamma, temperature, np_random,
              prior_policy=None, rollout_policy=None, tree=None):
    """MCTS.plan (mcts.py:179-184) from a fresh root, or continuing `tree` (step_strategy "subtree").
    Returns (plan, tree)."""
    prior_policy = prior_policy or {"type": "random_available"}
    rollout_policy = rollout_policy or {"type": "random_available"}
    t = tree if tree is not None else Tree()
    if tree is None:
        t.value, t.prior = [], []

    def new_node(parent, action, depth, prior):
        t.parent.append(parent)
        t.action.append(action)
        t.depth.append(depth)
        t.count.append(0)                                       # abstract.py:231
        t.first_child.append(-1)
        t.n_children.append(0)
        t.value.append(0.0)
        t.prior.append(prior)
        return len(t.parent) - 1

    if tree is None:
        new_node(-1, -1, 0, 1)
    for _ in range(episodes):
        state = copy.deepcopy(env)                              # :183
        node, total, depth, terminal = 0, 0, 0, False
        while depth < horizon and t.n_children[node] > 0 and not terminal:   # :141
            kids = list(t.children(node))
            # selection_strategy (:275-286)
            x = np.array([t.value[c] + temperature * len(kids) * t.prior[c] / (t.count[c] + 1)
                          for c in kids])
            indices = np.nonzero(x == np.amax(x))[0]
            child = kids[np_random.choice(indices)]             # random_argmax
            _, reward, terminal, _, _ = state.step(t.action[child])
            total += gamma ** depth * reward
            node = child
            depth += 1
        if t.n_children[node] == 0 and depth < horizon and (not terminal or node == 0):   # :151-154
            actions, probs = _policy(prior_policy, state)
            t.first_child[node] = len(t.parent)
            t.n_children[node] = len(actions)
            for a, p in zip(actions, probs):
                new_node(node, int(a), depth + 1, p)
        if not terminal:                                        # :156-157, evaluate :160-177
            for h in range(depth, horizon):
                actions, probs = _policy(rollout_policy, state)
                a = np_random.choice(actions, 1, p=np.array(probs))[0]
                _, reward, term, trunc, _ = state.step(a)
                total += gamma ** h * reward
                if np.all(term) or np.all(trunc):
                    break
        n = node                                                # update_branch :257-265
        while n >= 0:
            t.count[n] += 1
            t.value[n] += 1.0 / t.count[n] * (total - t.value[n])
            n = t.parent[n]
    # get_plan with MCTSNode.selection_rule (:212-218)
    plan, node = [], 0
    while t.n_children[node] > 0:
        kids = list(t.children(node))
        counts = np.array([t.count[c] for c in kids])
        ties = np.nonzero(counts == np.amax(counts))[0]
        best = max(ties, key=lambda i: t.value[kids[i]])
        plan.append(t.action[kids[best]])
        node = kids[best]
    return plan, t


_M64 = (1 << 64) - 1


def splitmix64(x):
    """The counter-based generator of the wavefront planners (same 64-bit arithmetic in C and CUDA)."""
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def wave_random(seed, episode, step, stream, n):
    """Uniform integer in [0, n) for (episode, step, stream): stream 0 = selection tie-break, 1 = rollout."""
    key = (seed + episode * 0x9E3779B97F4A7C15 + (step + 1) * 0xD1B54A32D192ED03 + stream * 0x8CB92BA72F3D8DD7) & _M64
    return int((splitmix64(key) >> 33) % n)


FIX_SCALE = float(1 << 40)


def mcts_plan_wavefront(env, episodes, horizon, gamma, temperature, width, seed):
    """SPECIFICATION of the device's wavefront MCTS (b2_mcts_plan_wave), not a reference algorithm.

    The reference's MCTS.plan (mcts.py:179-184) runs its episodes one after the other; every episode reads
    the statistics the previous one wrote and draws from one sequential PCG64 stream, so 4096 rollouts
    cannot run in parallel.  The wavefront keeps the reference's episode (selection :141-149, expansion
    :151-154, rollout :160-177, update_branch :257-265, recommendation :212-218) and changes three things:
      * episodes run in waves of `width`; inside a wave the selections are made in episode order on the
        statistics of the wave start plus VIRTUAL counts (an episode that passes through a child adds one
        to that child's count for the later selections of the same wave), so a wave spreads over the tree
        instead of repeating one path; a childless node reached by several episodes of a wave is expanded
        once, by the first of them, and every one of them rolls out from it;
      * randomness is counter based -- wave_random(seed, episode, step, stream) -- instead of a sequential
        stream: tie-breaks of the selection (random_argmax, abstract.py:296-311) and rollout actions
        (uniform over the available actions) do not depend on the execution order;
      * a node's value is the mean of the returns backed up through it, accumulated as an exact 64-bit
        fixed-point sum (2^-40 units), so the backup is order independent (the reference's incremental
        mean :253 gives the same mean up to rounding).
    Node ids: root 0; episode e owns the ids 1 + e * n_actions ... for the children it may create (unused
    ids keep parent = -2).  Returns (plan, tree)."""
    A = env.action_space.n
    cap = 1 + episodes * A
    t = Tree()
    t.parent = [-2] * cap
    t.action = [-1] * cap
    t.count = [0] * cap
    t.first_child = [-1] * cap
    t.n_children = [0] * cap
    t.vsum = [0] * cap
    t.parent[0] = -1
    t.env_steps = 0

    def value(c):
        return (float(t.vsum[c]) / FIX_SCALE) / t.count[c] if t.count[c] > 0 else 0.0

    for w0 in range(0, episodes, width):
        wave = range(w0, min(w0 + width, episodes))
        virtual, expander, paths = {}, {}, {}
        for e in wave:                                              # selections, episode order, virtual counts
            node, depth, path = 0, 0, []
            while depth < horizon and t.n_children[node] > 0:
                kids = list(t.children(node))
                n = len(kids)
                prior = 1.0 / n
                # the reference's score (mcts.py:275-286) with the division by (count + 1) written as a
                # multiplication by the reciprocal 1.0 / (count + virtual + 1)
                x = [value(c) + temperature * n * prior * (1.0 / (t.count[c] + virtual.get(c, 0) + 1)) for c in kids]
                best = max(x)
                ties = [i for i in range(n) if x[i] == best]
                child = kids[ties[wave_random(seed, e, depth, 0, len(ties))]]
                virtual[child] = virtual.get(child, 0) + 1
                path.append(child)
                node = child
                depth += 1
            paths[e] = path
            if depth < horizon and node not in expander:
                expander[node] = e
        results = {}
        for e in wave:                                              # simulations: independent of each other
            state = copy.deepcopy(env)
            path, total, terminal, reached = paths[e], 0.0, False, 0
            for h, c in enumerate(path):
                _, reward, terminal, _, _ = state.step(t.action[c])
                t.env_steps += 1
                total += gamma ** h * reward
                reached = h + 1
                if terminal:
                    break
            if not terminal:
                depth = len(path)
                leaf = path[-1] if path else 0
                if depth < horizon and expander.get(leaf) == e:
                    actions = _available_actions(state)
                    base = 1 + e * A
                    t.first_child[leaf], t.n_children[leaf] = base, len(actions)
                    for i, a in enumerate(actions):
                        t.parent[base + i], t.action[base + i] = leaf, int(a)
                for h in range(depth, horizon):
                    actions = _available_actions(state)
                    a = actions[wave_random(seed, e, h, 1, len(actions))]
                    _, reward, term, trunc, _ = state.step(a)
                    t.env_steps += 1
                    total += gamma ** h * reward
                    if term or trunc:
                        break
            results[e] = (reached, total)
        for e in wave:                                              # backup: exact integer sums
            reached, total = results[e]
            fixed = int(np.rint(total * FIX_SCALE))
            for c in [0] + paths[e][:reached]:
                t.count[c] += 1
                t.vsum[c] += fixed
    t.value = [value(c) for c in range(cap)]
    plan, node = [], 0
    while t.n_children[node] > 0:
        kids = list(t.children(node))
        counts = [t.count[c] for c in kids]
        ties = [i for i in range(len(kids)) if counts[i] == max(counts)]
        best = max(ties, key=lambda i: t.value[kids[i]])
        plan.append(t.action[kids[best]])
        node = kids[best]
    return plan, t


# --------------------------------------------------------------------------
# KL-UCB -- rl_agents/utils.py:89-203
# --------------------------------------------------------------------------
def bernoulli_kl(p, q):
    # utils.py:89-106
    kl1, kl2 = 0, math.inf
    if p > 0:
        if q > 0:
            kl1 = p * np.log(p / q)
    if q < 1:
        if p < 1:
            kl2 = (1 - p) * np.log((1 - p) / (1 - q))
        else:
            kl2 = 0
    return kl1 + kl2


def kl_upper_bound(_sum, count, threshold=1, eps=1e-2):
    """utils.py:123-147 + newton_iteration (:150-203), upper bound only."""
    if count == 0:
        return 1
    mu = _sum / count
    max_div = threshold / count
    a, b = mu, 1
    x = math.inf
    if a == b:
        return a
    x_next = (a + b) / 2
    iterations = 0
    while abs(x - x_next) > eps and iterations < 100:
        iterations += 1
        x = x_next
        f_x = bernoulli_kl(mu, x) - max_div
        try:
            df_x = (1 - mu) / (1 - x) - mu / x
        except ZeroDivisionError:
            df_x = (f_x - (bernoulli_kl(mu, x - eps) - max_div)) / eps
        if df_x != 0:
            x_next = x - f_x / df_x
        if x_next < a:
            x_next = 0.9 * a + (1 - 0.9) * x
        elif x_next > b:
            x_next = 0.9 * b + (1 - 0.9) * x
    if x_next < a:
        x_next = a
    if x_next > b:
        x_next = b
    return x_next


# --------------------------------------------------------------------------
# OLOP / KL-OLOP -- rl_agents/agents/tree_search/olop.py
# --------------------------------------------------------------------------
def olop_plan(env, budget, gamma, np_random, upper_bound=None,
              continuation_type="zeros", episodes=None, horizon=None):
    """OLOP.plan (olop.py:94-100) from reset (:36-40).  `env.step` follows the
    legacy 4-tuple API olop.py:87 expects.  Returns (plan, tree)."""
    upper_bound = upper_bound or {"type": "hoeffding", "time": "global",
                                  "threshold": "4*np.log(time)"}
    if horizon is None:
        budget = max(env.action_space.n, budget)                # :46-48
        episodes, horizon = olop_allocation(budget, gamma)
    kl = upper_bound["type"] == "kullback-leibler"
    t = Tree()
    t.cumulative_reward, t.mu_ucb, t.upper, t.done = [], [], [], []

    def new_node(parent, action, depth):
        # OLOPNode.__init__ (:106-124)
        t.parent.append(parent)
        t.action.append(action)
        t.depth.append(depth)
        t.count.append(0)
        t.first_child.append(-1)
        t.n_children.append(0)
        t.cumulative_reward.append(0)
        t.mu_ucb.append(1 if kl else math.inf)
        t.upper.append((1 - gamma ** (horizon + 1 - depth)) / (1 - gamma))
        t.done.append(False)
        return len(t.parent) - 1

    new_node(-1, -1, 0)
    for episode in range(episodes):
        state = copy.deepcopy(env)
        state.seed(np_random.randint(2 ** 30))                  # :73
        node = 0
        for h in range(horizon):
            if t.n_children[node] == 0:                         # :78-82
                actions = _available_actions(state)
                t.first_child[node] = len(t.parent)
                t.n_children[node] = len(actions)
                for a in actions:
                    new_node(node, a, t.depth[node] + 1)
                if continuation_type == "uniform":
                    action = np_random.choice(list(actions))
                else:
                    action = 0
                child = next(c for c in t.children(node) if t.action[c] == action)
            else:                                               # :84 first max
                child = t.first_child[node]
                for c in t.children(node):
                    if t.upper[c] > t.upper[child]:
                        child = c
                action = t.action[child]
            _, reward, done, _ = state.step(action)             # :87
            node = child
            # update (:132-142)
            if not 0 <= reward <= 1:
                raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
            if done:
                t.done[node] = True
            if t.done[node]:
                reward = 0
            t.cumulative_reward[node] += reward
            t.count[node] += 1
            if kl:                                              # :144-163
                time = (episode + 1) if upper_bound["time"] == "local" else episodes  # noqa: F841
                threshold = eval(upper_bound["threshold"])
                t.mu_ucb[node] = kl_upper_bound(t.cumulative_reward[node], t.count[node], threshold)
        n = node                                                # backup_to_root :182-193
        while n >= 0:
            if t.n_children[n] > 0:
                t.upper[n] = t.mu_ucb[n] + gamma * max(t.upper[c] for c in t.children(n))
            else:
                assert t.depth[n] == horizon
                t.upper[n] = t.mu_ucb[n]
            n = t.parent[n]
    t.episodes, t.horizon = episodes, horizon
    # get_plan with OLOPNode.selection_rule (:126-130)
    plan, node = [], 0
    while t.n_children[node] > 0:
        kids = list(t.children(node))
        counts = np.array([t.count[c] for c in kids])
        ties = np.nonzero(counts == np.amax(counts))[0]
        best = max(ties, key=lambda i: t.upper[kids[i]])
        plan.append(t.action[kids[best]])
        node = kids[best]
    return plan, t


# --------------------------------------------------------------------------
# Value iteration -- rl_agents/agents/dynamic_programming/value_iteration.py
# --------------------------------------------------------------------------
def bellman_expectation(mode, transition, reward, terminal, value, gamma, nxt=None):
    # value_iteration.py:51-63
    if mode == "deterministic":
        next_v = value[transition]
    elif mode == "stochastic":
        next_v = (transition * value.reshape((1, 1, value.size))).sum(axis=-1)
    elif mode == "sparse":
        next_v = (transition * np.take(value, nxt)).sum(axis=-1)
    else:
        raise ValueError("Unknown mode")
    next_v[terminal] = 0
    return reward + gamma * next_v


def value_iteration(mode, transition, reward, terminal, gamma, iterations, nxt=None):
    """get_state_action_value + fixed_point_iteration (value_iteration.py:42-45,
    65-73): iterates on Q; on np.allclose returns the PREVIOUS iterate."""
    q = np.zeros(reward.shape)
    sweeps = 0
    for _ in range(iterations):
        sweeps += 1
        nq = bellman_expectation(mode, transition, reward, terminal, q.max(axis=-1), gamma, nxt)
        if np.allclose(q, nq):
            break
        q = nq
    return q, sweeps


def robust_value_iteration(mode, transitions, rewards, gamma, iterations):
    """RobustValueIterationAgent.get_state_action_value (robust_value_iteration.py:39-58):
    Q <- min over the M models of R_m + gamma * E_m[max_a Q]; no terminal handling; the
    fixed-point loop and its allclose early exit are ValueIterationAgent's (:65-73)."""
    q = np.zeros(transitions.shape[1:3])
    sweeps = 0
    for _ in range(iterations):
        sweeps += 1
        value = q.max(axis=-1)
        if mode == "deterministic":
            next_v = value[transitions]
        elif mode == "stochastic":
            next_v = (transitions * value.reshape((1, 1, 1, np.size(value)))).sum(axis=-1)
        else:
            raise ValueError("Unknown mode")
        nq = np.min(rewards + gamma * next_v, axis=0)
        if np.allclose(q, nq):
            break
        q = nq
    return q, sweeps


# --------------------------------------------------------------------------
# GBOP-T / state-aware OPD -- rl_agents/agents/tree_search/state_aware.py
# (oracle groundwork for SURVEY 8f rank 3; no device implementation yet)
# --------------------------------------------------------------------------
def state_aware_plan(env, observation, budget, gamma, np_random, terminal_reward=0.0, backup_aggregated_nodes=True,
                     prune_suboptimal_leaves=True, accuracy=0):
    """StateAwarePlanner.plan (state_aware.py:118-130).  Observations must be hashable state ids
    (the reference keys its tables by str(observation)).  Returns (plan, tree, state_values, leaves)."""
    t = Tree()
    t.reward, t.lower, t.done, t.obs = [], [], [], []
    states = []
    state_nodes = {}                      # observation -> node ids in creation order (:20-22)
    default_value = 1 / (1 - gamma)
    state_values = {}

    def sv(o):
        return state_values.get(o, default_value)

    def update_value(o, value):           # :104-116
        delta = sv(o) - value
        if delta > 0:
            state_values[o] = value
        elif o not in state_values:
            state_values[o] = default_value          # defaultdict access materialises the key
        return delta

    def upper(i):                         # get_value_upper_bound (:66-68)
        return t.lower[i] + (gamma ** t.depth[i]) * sv(t.obs[i])

    def new_node(parent, action, depth, state):
        t.parent.append(parent); t.action.append(action); t.depth.append(depth); t.count.append(1)
        t.first_child.append(-1); t.n_children.append(0)
        t.reward.append(0.0); t.lower.append(0.0); t.done.append(False); t.obs.append(None)
        states.append(state)
        return len(t.parent) - 1

    root = new_node(-1, -1, 0, env)
    t.obs[root] = observation
    state_nodes[observation] = [root]
    state_values[observation] = default_value
    leaves = [root]
    for _ in range(int(budget) // env.action_space.n):
        best = leaves[0]                                    # run(): first max (:92)
        for i in leaves[1:]:
            if upper(i) > upper(best):
                best = i
        leaves.remove(best)                                 # expand (deterministic.py:28-43)
        actions = _available_actions(states[best])
        t.first_child[best] = len(t.parent)
        t.n_children[best] = len(actions)
        d = t.depth[best] + 1
        for a in actions:
            c = new_node(best, a, d, copy.deepcopy(states[best]))
            obs, reward, done, _, _ = states[c].step(a)
            leaves.append(c)
            if not (0 <= reward <= 1):
                raise ValueError("This planner assumes that all rewards are normalized in [0, 1]")
            t.reward[c] = reward
            t.done[c] = bool(done)
            t.lower[c] = t.lower[best] + (gamma ** (d - 1)) * reward
            if done:
                t.lower[c] = t.lower[c] + terminal_reward * (gamma ** d) / (1 - gamma)
            n = c
            while n >= 0:
                t.count[n] += 1
                n = t.parent[n]
            t.obs[c] = obs                                  # StateAwareNode.update (:15-26)
            state_nodes.setdefault(obs, []).append(c)
            if done:
                update_value(obs, 0)
        states[best] = None
        queue = [best]                                      # backup_to_root (:42-64)
        while queue:
            node = queue.pop(0)
            delta = 0
            if t.n_children[node] > 0:
                kids = list(t.children(node))
                bc = kids[0]
                for c in kids[1:]:
                    if upper(c) > upper(bc):
                        bc = c
                backup = t.reward[bc] + gamma * sv(t.obs[bc])
                if t.obs[bc] not in state_values:
                    state_values[t.obs[bc]] = default_value
                delta = update_value(t.obs[node], backup)
            for nb in state_nodes[t.obs[node]]:
                if t.parent[nb] >= 0 and (nb == node or backup_aggregated_nodes) and \
                        delta > accuracy * (1 - gamma) * gamma ** (t.depth[nb] - 1):
                    queue.append(t.parent[nb])
        if prune_suboptimal_leaves:                         # :98-99, prune (:28-40)
            for leaf in reversed(list(leaves)):
                ub = upper(leaf)
                for node in state_nodes[t.obs[leaf]]:
                    if node != leaf and upper(node) >= ub and t.depth[node] >= t.depth[leaf] and \
                            (t.n_children[node] > 0 or node in leaves):
                        leaves.remove(leaf)
                        break
    # StateAwarePlanner.plan calls super().plan() -- which already runs get_plan() -- and then get_plan()
    # again (:124,:130): the tie-breaking RNG is consumed twice and the second walk is the one returned
    greedy_plan(t, t.lower, np_random)
    return greedy_plan(t, t.lower, np_random), t, state_values, leaves
