// GBOP-T (state-aware optimistic planning) -- the plan() loop of StateAwarePlanner
// (rl_agents/agents/tree_search/state_aware.py:71-130) for a BATCH of independent decisions on deterministic
// finite MDPs: one warp per tree, node order / leaves / state values exactly the reference's.
//
// What differs from OPD (deterministic.py), and why the OPD kernels cannot be reused:
//   * a leaf's upper bound is value_lower + gamma^depth * U(state) (:66-68) where U is a table shared by all
//     nodes that reached the same state and tightened by every backup, so the frontier keys are not final
//     at creation: the arg-max (:92) is a fresh scan of the leaves (lane-strided, first max = lowest node id);
//   * backup_to_root (:42-64) is a breadth-first propagation through the nodes AGGREGATED by state: per state a
//     doubly linked list of its nodes in creation order, a FIFO of parents to revisit;
//   * after every expansion the leaves are pruned (:28-40, :98-99) in reverse order: a leaf goes when another
//     node of the same state has a bound at least as good at no smaller depth.  Leaves of different states do
//     not interact, so the prune runs one lane per state, each walking its state's list backwards.
#include "common.cuh"
#include "pcg64.cuh"

namespace b2 {

constexpr int GBOP_MAX_BRANCH = 8;
constexpr int META_LEAF = 1 << 17;

struct GbopArgs {
    b2_gbop_config cfg;
    b2_gbop_tree tree;
    const int32_t* root_states;
    char* workspace;
    int64_t ws_per_tree;
    int8_t* plan;
    int32_t* result;
};

struct GbopView {          // one tree's slices
    int32_t *parent, *first_child, *depth, *count, *meta, *obs;
    double *reward, *lower;
    double* sv;            // [S] state value upper bounds
    int32_t *head, *tail;  // [S] per-state node list
    int32_t *next, *prev;  // [cap]
    int32_t* queue;        // [queue_capacity]
    int32_t* exp_order;    // [n_expansions]
};

__device__ __forceinline__ double gbop_upper(const GbopArgs& a, const GbopView& v, int i) {
    return v.lower[i] + a.cfg.gamma_pow[v.depth[i]] * v.sv[v.obs[i]];       // get_value_upper_bound (:66-68)
}

__device__ __forceinline__ double update_value(const GbopView& v, int o, double value) {   // :104-116
    const double delta = v.sv[o] - value;
    if (delta > 0) v.sv[o] = value;
    return delta;
}

__global__ void __launch_bounds__(128) gbop_finite_kernel(GbopArgs a) {
    const int lane = threadIdx.x & 31;
    const int tree = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tree >= a.cfg.n_trees) return;
    const int cap = a.cfg.node_capacity, S = a.cfg.mdp.n_states, A = a.cfg.n_actions;
    const int64_t nb = (int64_t)tree * cap;
    GbopView v;
    v.parent = a.tree.parent + nb; v.first_child = a.tree.first_child + nb; v.depth = a.tree.depth + nb;
    v.count = a.tree.count + nb; v.meta = a.tree.meta + nb; v.obs = a.tree.obs + nb;
    v.reward = a.tree.reward + nb; v.lower = a.tree.lower + nb;
    char* ws = a.workspace + (int64_t)tree * a.ws_per_tree;
    v.sv = (double*)ws; ws += (int64_t)S * 8;
    v.head = (int32_t*)ws; ws += (int64_t)S * 4;
    v.tail = (int32_t*)ws; ws += (int64_t)S * 4;
    v.next = (int32_t*)ws; ws += (int64_t)cap * 4;
    v.prev = (int32_t*)ws; ws += (int64_t)cap * 4;
    v.exp_order = (int32_t*)ws; ws += (int64_t)(a.cfg.n_expansions + 1) * 4;
    v.queue = (int32_t*)ws;
    const b2_finite_mdp& m = a.cfg.mdp;
    const double gamma = a.cfg.gamma;
    for (int s = lane; s < S; s += 32) { v.sv[s] = a.cfg.default_value; v.head[s] = -1; v.tail[s] = -1; }
    __syncwarp();
    int error = 0, overflow = 0;
    if (lane == 0) {            // root (state_aware.py:118-123)
        const int s0 = a.root_states[tree];
        v.parent[0] = -1; v.first_child[0] = -1; v.depth[0] = 0; v.count[0] = 1; v.meta[0] = 0xff | META_LEAF;
        v.reward[0] = 0.0; v.lower[0] = 0.0; v.obs[0] = s0;
        v.head[s0] = v.tail[s0] = 0; v.next[0] = -1; v.prev[0] = -1;
    }
    __syncwarp();
    int n_nodes = 1, n_exp = 0;
    for (int it = 0; it < a.cfg.n_expansions; ++it) {
        // ---- run(): first arg-max of the leaves' upper bounds (:92) ----
        double best = -INFINITY;
        int best_i = 0x7fffffff;
        for (int i = lane; i < n_nodes; i += 32) {
            if (v.meta[i] & META_LEAF) {
                const double u = gbop_upper(a, v, i);
                if (u > best) { best = u; best_i = i; }       // ascending ids per lane: strict > keeps the first
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double b2v = __shfl_xor_sync(0xffffffffu, best, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (b2v > best || (b2v == best && i2 < best_i)) { best = b2v; best_i = i2; }
        }
        if (best_i == 0x7fffffff) break;                      // every leaf was pruned
        const int leaf = best_i;
        if (lane == 0) {
            // ---- expand (deterministic.py:28-43) + StateAwareNode.update (:15-26) ----
            const int s = v.obs[leaf], d = v.depth[leaf] + 1;
            v.meta[leaf] = (v.meta[leaf] & ~META_LEAF & ~0xff00) | (A << 8);
            v.first_child[leaf] = n_nodes;
            v.exp_order[n_exp] = leaf;
            for (int act = 0; act < A; ++act) {
                const int c = n_nodes + act;
                const double r = m.reward[(int64_t)s * A + act];
                const bool done = m.terminal[s] != 0;             // done = terminal[state the action is taken in]
                const int s2 = m.transition[(int64_t)s * A + act];
                if (!(r >= 0.0 && r <= 1.0)) error = 1;
                double lo = v.lower[leaf] + a.cfg.gamma_pow[d - 1] * r;
                if (done) lo = lo + a.cfg.terminal_bonus[d];
                v.parent[c] = leaf; v.first_child[c] = -1; v.depth[c] = d; v.count[c] = 2;
                v.meta[c] = act | (done ? 1 << 16 : 0) | META_LEAF;
                v.reward[c] = r; v.lower[c] = lo; v.obs[c] = s2;
                v.next[c] = -1; v.prev[c] = v.tail[s2];
                if (v.tail[s2] >= 0) v.next[v.tail[s2]] = c; else v.head[s2] = c;
                v.tail[s2] = c;
                if (done) update_value(v, s2, 0.0);
            }
            // ---- backup_to_root (:42-64): FIFO of nodes whose state bound may tighten ----
            int qh = 0, qt = 0;
            const int qcap = a.cfg.queue_capacity;
            v.queue[qt++] = leaf;
            while (qh < qt) {
                const int node = v.queue[qh++];
                double delta = 0.0;
                const int k = (v.meta[node] >> 8) & 0xff;
                if (k > 0) {
                    const int fc = v.first_child[node];
                    int bc = fc;
                    double bu = gbop_upper(a, v, fc);
                    for (int q = 1; q < k; ++q) {
                        const double u = gbop_upper(a, v, fc + q);
                        if (u > bu) { bu = u; bc = fc + q; }
                    }
                    const double backup = v.reward[bc] + gamma * v.sv[v.obs[bc]];
                    delta = update_value(v, v.obs[node], backup);
                }
                for (int nbn = v.head[v.obs[node]]; nbn >= 0; nbn = v.next[nbn]) {
                    if (v.parent[nbn] >= 0 && (nbn == node || a.cfg.backup_aggregated_nodes) &&
                        delta > a.cfg.accuracy_scale * a.cfg.gamma_pow[v.depth[nbn] - 1]) {
                        if (qt < qcap) v.queue[qt++] = v.parent[nbn];
                        else overflow = 1;
                    }
                }
            }
        }
        n_nodes += A;
        n_exp += 1;
        error = __shfl_sync(0xffffffffu, error, 0);
        overflow = __shfl_sync(0xffffffffu, overflow, 0);
        __syncwarp();
        if (error || overflow) break;
        // ---- prune (:28-40, :98-99): reverse leaf order; leaves of different states are independent, so
        //      one lane per state walks that state's node list backwards (= descending node ids) ----
        if (a.cfg.prune_suboptimal_leaves) {
            for (int s = lane; s < S; s += 32) {
                for (int leaf2 = v.tail[s]; leaf2 >= 0; leaf2 = v.prev[leaf2]) {
                    if (!(v.meta[leaf2] & META_LEAF)) continue;
                    const double ub = gbop_upper(a, v, leaf2);
                    const int dl = v.depth[leaf2];
                    for (int node = v.head[s]; node >= 0; node = v.next[node]) {
                        if (node != leaf2 && gbop_upper(a, v, node) >= ub && v.depth[node] >= dl &&
                            ((((v.meta[node] >> 8) & 0xff) > 0) || (v.meta[node] & META_LEAF))) {
                            v.meta[leaf2] &= ~META_LEAF;
                            break;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane != 0) return;
    // counts (deterministic.py:64-65), bottom-up in reverse expansion order
    for (int k = n_exp - 1; k >= 0; --k) {
        const int p = v.exp_order[k], fc = v.first_child[p], n = (v.meta[p] >> 8) & 0xff;
        int desc = 0;
        for (int q = 0; q < n; ++q) desc += v.count[fc + q] - 1;
        v.count[p] = (p == 0 ? 1 : 2) + desc;
    }
    // get_plan on value_lower (abstract.py:143-156); ties go to the host RNG
    int8_t* plan = a.plan + (int64_t)tree * a.cfg.plan_capacity;
    int node = 0, len = 0, tie_node = -1;
    while (v.first_child[node] >= 0) {
        const int fc = v.first_child[node], n = (v.meta[node] >> 8) & 0xff;
        double mx = -INFINITY;
        int cnt = 0, arg = fc;
        for (int q = 0; q < n; ++q) {
            const double x = v.lower[fc + q];
            if (x > mx) { mx = x; cnt = 1; arg = fc + q; }
            else if (x == mx) ++cnt;
        }
        if (cnt > 1) { tie_node = node; break; }
        if (len < a.cfg.plan_capacity) plan[len] = (int8_t)(v.meta[arg] & 0xff);
        ++len;
        node = arg;
    }
    int n_leaves = 0;
    for (int i = 0; i < n_nodes; ++i) n_leaves += (v.meta[i] & META_LEAF) ? 1 : 0;
    int32_t* res = a.result + (int64_t)tree * B2_OPD_RESULT_WORDS;
    res[0] = n_nodes; res[1] = n_leaves; res[2] = 0; res[3] = 0; res[4] = error; res[5] = len; res[6] = tie_node;
    res[7] = overflow; res[8] = n_exp;
}

// ---------------------------------------------------------------------------------------------------------
// GBOP-D (graph-based optimistic planning) -- GraphBasedPlanner.plan (graph_based.py:84-138), deterministic
// finite MDPs: one graph node per STATE with a lower and an upper value bound.  An epoch walks from the root along
// the optimistic action (random tie-break on the planner's numpy PCG64 stream) to a state without children,
// expands it and re-backs-up both bounds breadth first through the expanded parents (reverse transition CSR
// built on the host; the reference's `list(node.parents)` is a Python set, here: ascending state id).
// One warp per decision, the epoch loop on its first lane.
// ---------------------------------------------------------------------------------------------------------
struct GbopdArgs {
    b2_gbopd_config cfg;
    const int32_t* root_states;
    double* lower;          // [n_trees, S]
    double* upper;          // [n_trees, S]
    uint8_t* flags;         // [n_trees, S] bit0 node exists, bit1 expanded
    int32_t* queue;         // [n_trees, queue_capacity]
    uint64_t* rng;          // [n_trees, 6]
    int8_t* plan;
    int32_t* result;
};

__global__ void __launch_bounds__(128) gbopd_finite_kernel(GbopdArgs a) {
    const int lane = threadIdx.x & 31;
    const int tree = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tree >= a.cfg.n_trees) return;
    const b2_finite_mdp& m = a.cfg.mdp;
    const int S = m.n_states, A = a.cfg.n_actions;
    double* lo = a.lower + (int64_t)tree * S;
    double* up = a.upper + (int64_t)tree * S;
    uint8_t* fl = a.flags + (int64_t)tree * S;
    int32_t* queue = a.queue + (int64_t)tree * a.cfg.queue_capacity;
    for (int s = lane; s < S; s += 32) { lo[s] = 0.0; up[s] = a.cfg.default_value; fl[s] = 0; }
    __syncwarp();
    if (lane != 0) return;
    Pcg64 rng;
    rng.load(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
    const int root = a.root_states[tree];
    const double gamma = a.cfg.gamma;
    fl[root] = 1;                                                    // get_node(observation) (:110-116)
    int overflow = 0, expansions = 0, loops = 0;
    for (int epoch = 0; epoch < a.cfg.n_epochs && !overflow; ++epoch) {
        int s = root;
        bool sink = false;
        for (int k = 0; k < a.cfg.sampling_timeout; ++k) {
            if (!(fl[s] & 2)) {
                // expand (:38-52): children = T[s, :], rewards = R[s, :], this state joins its children's parents
                fl[s] |= 2;
                for (int act = 0; act < A; ++act) fl[m.transition[(int64_t)s * A + act]] |= 1;
                ++expansions;
                // partial_value_iteration (:63-75)
                int qh = 0, qt = 0;
                queue[qt++] = s;
                while (qh < qt) {
                    const int node = queue[qh++];
                    double bl = -INFINITY, bu = -INFINITY;
                    for (int act = 0; act < A; ++act) {
                        const int c = m.transition[(int64_t)node * A + act];
                        const double r = m.reward[(int64_t)node * A + act];
                        const double vl = r + gamma * lo[c], vu = r + gamma * up[c];
                        bl = vl > bl ? vl : bl;
                        bu = vu > bu ? vu : bu;
                    }
                    double delta = fabs(lo[node] - bl);
                    lo[node] = bl;
                    const double du = fabs(up[node] - bu);
                    delta = du > delta ? du : delta;
                    up[node] = bu;
                    if (delta > a.cfg.accuracy) {
                        for (int e = a.cfg.rev_ptr[node]; e < a.cfg.rev_ptr[node + 1]; ++e) {
                            const int p = a.cfg.rev_idx[e];
                            if (fl[p] & 2) {
                                if (qt < a.cfg.queue_capacity) queue[qt++] = p;
                                else overflow = 1;
                            }
                        }
                    }
                }
                sink = true;
                break;
            }
            // sampling_rule (:22-30): optimistic action, random_argmax over the tied maxima
            double q[GBOP_MAX_BRANCH], best = -INFINITY;
            int ties = 0;
            for (int act = 0; act < A; ++act) {
                q[act] = m.reward[(int64_t)s * A + act] + gamma * up[m.transition[(int64_t)s * A + act]];
                if (q[act] > best) { best = q[act]; ties = 1; }
                else if (q[act] == best) ++ties;
            }
            int pick = (int)rng.integers((uint32_t)ties), sel = 0;
            for (int act = 0; act < A; ++act) {
                if (q[act] == best) {
                    if (pick == 0) sel = act;
                    --pick;
                }
            }
            s = m.transition[(int64_t)s * A + sel];
        }
        if (!sink) ++loops;        // "could not find a sink" (:104-106)
    }
    rng.store(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
    // get_plan (:126-135): conservative action (first max of the lower-bound backups)
    int8_t* plan = a.plan + (int64_t)tree * a.cfg.plan_capacity;
    int node = root, len = 0;
    for (int k = 0; k < a.cfg.sampling_timeout; ++k) {
        if (!(fl[node] & 2)) break;
        double best = -INFINITY;
        int sel = 0;
        for (int act = 0; act < A; ++act) {
            const double v = m.reward[(int64_t)node * A + act] + gamma * lo[m.transition[(int64_t)node * A + act]];
            if (v > best) { best = v; sel = act; }
        }
        if (len < a.cfg.plan_capacity) plan[len] = (int8_t)sel;
        ++len;
        node = m.transition[(int64_t)node * A + sel];
    }
    int n_nodes = 0;
    for (int s2 = 0; s2 < S; ++s2) n_nodes += fl[s2] & 1;
    int32_t* res = a.result + (int64_t)tree * B2_OPD_RESULT_WORDS;
    res[0] = n_nodes; res[1] = expansions; res[2] = loops; res[3] = 0; res[4] = 0; res[5] = len; res[6] = -1;
    res[7] = overflow;
}

static int64_t gbop_ws_per_tree(const b2_gbop_config* c) {
    int64_t b = (int64_t)c->mdp.n_states * 16 + (int64_t)c->node_capacity * 8 + ((int64_t)c->n_expansions + 1) * 4 +
                (int64_t)c->queue_capacity * 4;
    return (b + 255) & ~(int64_t)255;
}

}  // namespace b2

using namespace b2;

extern "C" int64_t b2_gbop_workspace_bytes(const b2_gbop_config* cfg) {
    if (!cfg || cfg->n_trees <= 0 || cfg->node_capacity <= 0 || cfg->queue_capacity <= 0) return -1;
    return gbop_ws_per_tree(cfg) * cfg->n_trees;
}

extern "C" int b2_gbop_plan(const b2_gbop_config* cfg, const int32_t* root_states, const b2_gbop_tree* tree,
                            void* workspace, int8_t* plan, int32_t* result, void* stream) {
    B2_REQUIRE(cfg && root_states && tree && workspace && plan && result, "null pointer");
    B2_REQUIRE(cfg->n_trees > 0 && cfg->n_expansions >= 0, "bad batch / budget");
    B2_REQUIRE(cfg->n_actions > 0 && cfg->n_actions <= GBOP_MAX_BRANCH, "n_actions must be in 1..8");
    B2_REQUIRE((int64_t)cfg->node_capacity >= 1 + (int64_t)cfg->n_expansions * cfg->n_actions, "node_capacity too small");
    B2_REQUIRE(cfg->mdp.transition && cfg->mdp.reward && cfg->mdp.terminal && cfg->mdp.n_states > 0, "finite MDP tables missing");
    B2_REQUIRE(cfg->mdp.n_actions == cfg->n_actions, "mdp.n_actions != n_actions");
    B2_REQUIRE(cfg->gamma_pow && cfg->terminal_bonus, "gamma tables missing");
    B2_REQUIRE(cfg->queue_capacity > 0 && cfg->plan_capacity > 0, "capacities must be positive");
    GbopArgs a;
    a.cfg = *cfg; a.tree = *tree; a.root_states = root_states; a.workspace = (char*)workspace;
    a.ws_per_tree = gbop_ws_per_tree(cfg); a.plan = plan; a.result = result;
    gbop_finite_kernel<<<(cfg->n_trees + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a);
    B2_CUDA_CHECK(cudaGetLastError());
    return B2_OK;
}

extern "C" int b2_gbopd_plan(const b2_gbopd_config* cfg, const int32_t* root_states, double* lower, double* upper,
                             uint8_t* flags, int32_t* queue, uint64_t* rng, int8_t* plan, int32_t* result, void* stream) {
    B2_REQUIRE(cfg && root_states && lower && upper && flags && queue && rng && plan && result, "null pointer");
    B2_REQUIRE(cfg->n_trees > 0 && cfg->n_epochs >= 0 && cfg->sampling_timeout > 0, "bad batch / budget");
    B2_REQUIRE(cfg->n_actions > 0 && cfg->n_actions <= GBOP_MAX_BRANCH, "n_actions must be in 1..8");
    B2_REQUIRE(cfg->mdp.transition && cfg->mdp.reward && cfg->mdp.n_states > 0 && cfg->mdp.n_actions == cfg->n_actions,
               "finite MDP tables missing");
    B2_REQUIRE(cfg->rev_ptr && cfg->rev_idx, "reverse transition CSR missing");
    B2_REQUIRE(cfg->queue_capacity > 0 && cfg->plan_capacity > 0, "capacities must be positive");
    GbopdArgs a;
    a.cfg = *cfg; a.root_states = root_states; a.lower = lower; a.upper = upper; a.flags = flags; a.queue = queue;
    a.rng = rng; a.plan = plan; a.result = result;
    gbopd_finite_kernel<<<(cfg->n_trees + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a);
    B2_CUDA_CHECK(cudaGetLastError());
    return B2_OK;
}
