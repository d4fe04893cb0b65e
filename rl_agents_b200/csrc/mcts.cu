// MCTS / UCT -- the plan() loop of rl_agents/agents/tree_search/mcts.py for a
// BATCH of independent decisions.  Episodes inside one tree stay strictly
// sequential (every episode reads the statistics the previous one wrote) and
// consume that tree's numpy PCG64 stream exactly as the reference does, so
// tree indices, visit counts and values are bit-identical; the batch dimension
// (thousands of trees / root-parallel replicas) is what fills the GPU.
//
// One tree per lane group: 16 lanes for HighwayLite (lane = vehicle slot, the
// scene lives in registers and is "deep-copied" from the root once per episode,
// mcts.py:183), 1 lane for finite MDPs.  Tree bookkeeping is group-uniform
// scalar code; lane 0 of the group performs the stores.
#include "common.cuh"
#include "highway_lite.cuh"
#include "pcg64.cuh"

namespace b2 {

constexpr int MAX_BRANCH_MCTS = 8;

struct MctsArgs {
    b2_mcts_config cfg;
    b2_mcts_tree tree;
    const int32_t* root_states;
    uint64_t* rng;
    int8_t* plan;
    int32_t* result;
};

// ---------------------------------------------------------------- envs ----
struct FiniteEnv {
    static constexpr int GROUP = 1;
    int s;
    __device__ __forceinline__ void load_root(const MctsArgs& a, int tree, int li) { s = a.root_states[tree]; }
    __device__ __forceinline__ int avail(const MctsArgs& a, unsigned gmask) const { return (1 << a.cfg.n_actions) - 1; }
    __device__ __forceinline__ static int nth(int mask, int n) { return n; }
    // position of `action` among the available actions in the env's order, or -1
    __device__ __forceinline__ static int rank_of(int mask, int action) { return (action >= 0 && (mask >> action) & 1) ? action : -1; }
    __device__ __forceinline__ double step(const MctsArgs& a, int action, int li, unsigned gmask, float* gs, bool& term, bool& trunc) {
        const b2_finite_mdp& m = a.cfg.mdp;
        const double r = m.reward[(int64_t)s * m.n_actions + action];
        term = m.terminal[s] != 0;        // finite_mdp's MDP.step: done = terminal[state BEFORE the transition]
        s = m.transition[(int64_t)s * m.n_actions + action];
        trunc = false;
        return r;
    }
};

struct HighwayEnv {
    static constexpr int GROUP = 16;
    hw::Lane L;
    int t, si;
    __device__ __forceinline__ void load_root(const MctsArgs& a, int tree, int li) {
        hw::load_state(a.root_states + (int64_t)tree * hw::WORDS, li, L, t, si);
    }
    __device__ __forceinline__ int avail(const MctsArgs& a, unsigned gmask) const {
        const float ego_y = __shfl_sync(gmask, L.y, 0, 16);
        return hw::avail_mask(ego_y, si);
    }
    __device__ __forceinline__ static int nth(int mask, int n) { return hw::nth_action(mask, n); }
    __device__ __forceinline__ static int rank_of(int mask, int action) {
        if (action < 0 || !((mask >> action) & 1)) return -1;
        const int order[5] = {hw::A_IDLE, hw::A_LEFT, hw::A_RIGHT, hw::A_FASTER, hw::A_SLOWER};
        int k = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (order[i] == action) return k;
            k += (mask >> order[i]) & 1;
        }
        return -1;
    }
    __device__ __forceinline__ double step(const MctsArgs& a, int action, int li, unsigned gmask, float* gs, bool& term, bool& trunc) {
        return (double)hw::step(L, li, t, si, action, term, trunc, gmask, gs);
    }
};

// --------------------------------------------------------------- kernel ---
#ifndef B2_MCTS_MIN_BLOCKS
#define B2_MCTS_MIN_BLOCKS 8   // 64 registers, 32 warps/SM: measured best (6.5M vs 5.8M episodes/s at 4)
#endif
template <class Env>
__global__ void __launch_bounds__(128, B2_MCTS_MIN_BLOCKS) mcts_kernel(MctsArgs a) {
    constexpr int G = Env::GROUP;
    __shared__ float scratch[G == 16 ? 128 / 16 : 1][hw::SCRATCH_FLOATS];
    const int gtid = blockIdx.x * 128 + threadIdx.x;
    const int tree_raw = gtid / G, li = gtid % G;
    const bool live = tree_raw < a.cfg.n_trees;
    const int tree = live ? tree_raw : a.cfg.n_trees - 1;   // idle groups shadow the last tree, never store
    const bool writer = live && li == 0;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = G == 1 ? (1u << lane) : (0xFFFFu << (lane & 16));
    const int A = a.cfg.n_actions, H = a.cfg.horizon;
    const int64_t nb = (int64_t)tree * a.cfg.node_capacity;
    const b2_mcts_tree& tr = a.tree;

    Pcg64 rng;
    rng.load(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
    const int resume = a.cfg.resume_nodes ? a.cfg.resume_nodes[tree] : 0;
    if (writer && resume <= 0) {   // MCTSNode(parent=None) (mcts.py:129-130, :207-210)
        tr.parent[nb] = -1; tr.first_child[nb] = -1; tr.count[nb] = 0; tr.meta[nb] = 0xff;
        tr.value[nb] = 0.0; tr.prior[nb] = 1.0;
    }
    __syncwarp(gmask);
    int n_nodes = resume > 0 ? resume : 1, env_steps = 0;   // resume: a re-rooted sub-tree is already in place

    for (int ep = 0; ep < a.cfg.episodes; ++ep) {
        Env env;
        env.load_root(a, tree, li);                 // safe_deepcopy_env(state), mcts.py:183
        int node = 0;
        bool in_sel = true, active = live;
        double total = 0.0;
        for (int h = 0; h < H; ++h) {
            // the 16 lanes of a group leave the episode together (terminal / truncated); the
            // other half of the warp keeps stepping its own scene under its half mask
            if (!active) break;
            int action = hw::A_IDLE < A ? hw::A_IDLE : 0, child = -1;
            const int amask = env.avail(a, gmask);
            if (active && in_sel && tr.first_child[nb + node] < 0) {
                // expansion (mcts.py:151-154, :237-246): children for the policy's actions
                const int pm = a.cfg.prior_policy != 1 ? amask : (1 << A) - 1;
                const int n = __popc(pm);
                if (writer) {
                    const double p = 1.0 / (double)n;
                    const double* pt = nullptr;       // preference_policy (:76-97): host-made probabilities
                    if (a.cfg.prior_policy == 2)
                        pt = a.cfg.pref_prior + ((int64_t)n * (A + 1) + Env::rank_of(pm, a.cfg.prior_pref_action) + 1) * A;
                    for (int i = 0; i < n; ++i) {
                        const int c = n_nodes + i;
                        tr.parent[nb + c] = node; tr.first_child[nb + c] = -1; tr.count[nb + c] = 0;
                        tr.meta[nb + c] = a.cfg.prior_policy != 1 ? Env::nth(pm, i) : i;
                        tr.value[nb + c] = 0.0; tr.prior[nb + c] = pt ? pt[i] : p;
                    }
                    tr.first_child[nb + node] = n_nodes;
                    tr.meta[nb + node] = (tr.meta[nb + node] & 0xff) | (n << 8);
                }
                n_nodes += n;
                in_sel = false;
                __syncwarp(gmask);
            }
            if (active) {
                if (in_sel) {
                    // sampling_rule / selection_strategy (mcts.py:220-235, :275-286)
                    const int fc = tr.first_child[nb + node];
                    const int n = (tr.meta[nb + node] >> 8) & 0xff;
                    double best = -INFINITY;
                    int ties = 0;
                    double sc[MAX_BRANCH_MCTS];
#pragma unroll
                    for (int i = 0; i < MAX_BRANCH_MCTS; ++i) {
                        if (i < n) {
                            const double v = tr.value[nb + fc + i];
                            const double p = tr.prior[nb + fc + i];
                            const int cnt = tr.count[nb + fc + i];
                            sc[i] = v + a.cfg.temperature * (double)n * p / (double)(cnt + 1);
                            if (sc[i] > best) { best = sc[i]; ties = 1; }
                            else if (sc[i] == best) ++ties;
                        }
                    }
                    int pick = (int)rng.integers((uint32_t)ties);   // random_argmax
                    int sel = 0;
#pragma unroll
                    for (int i = 0; i < MAX_BRANCH_MCTS; ++i) {
                        if (i < n && sc[i] == best) {
                            if (pick == 0) sel = i;
                            --pick;
                        }
                    }
                    child = fc + sel;
                    action = tr.meta[nb + child] & 0xff;
                } else {
                    // rollout policy (mcts.py:171-172): choice(actions, 1, p)
                    const int pm = a.cfg.rollout_policy != 1 ? amask : (1 << A) - 1;
                    const int n = __popc(pm);
                    const double u = rng.random();
                    const double* cdf = a.cfg.rollout_policy == 2
                        ? a.cfg.pref_cdf + ((int64_t)n * (A + 1) + Env::rank_of(pm, a.cfg.rollout_pref_action) + 1) * A
                        : a.cfg.uniform_cdf + (int64_t)n * A;
                    int idx = 0;
                    for (int i = 0; i < n; ++i) idx += cdf[i] <= u ? 1 : 0;   // searchsorted(side='right')
                    idx = min(idx, n - 1);
                    action = a.cfg.rollout_policy != 1 ? Env::nth(pm, idx) : idx;
                }
            }
            bool term, trunc;
            Env next = env;
            const double r = next.step(a, action, li, gmask, scratch[(threadIdx.x >> 4) % (128 / 16)], term, trunc);
            if (active) {
                env = next;
                ++env_steps;
                total += a.cfg.gamma_pow[h] * r;          // mcts.py:146,174
                if (in_sel) {
                    node = child;
                    if (term) active = false;              // :141 loop exit, no expansion, no rollout
                } else if (term || trunc) {
                    active = false;                        // :175
                }
            }
        }
        // update_branch (mcts.py:257-265)
        if (writer) {
            int n = node;
            while (n >= 0) {
                const int c = tr.count[nb + n] + 1;
                const double v = tr.value[nb + n];
                tr.count[nb + n] = c;
                tr.value[nb + n] = v + 1.0 / (double)c * (total - v);
                n = tr.parent[nb + n];
            }
        }
        __syncwarp(gmask);
    }

    if (writer) {
        rng.store(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
        // get_plan with MCTSNode.selection_rule (mcts.py:212-218)
        int8_t* plan = a.plan + (int64_t)tree * a.cfg.horizon;
        int node = 0, len = 0;
        while (tr.first_child[nb + node] >= 0) {
            const int fc = tr.first_child[nb + node];
            const int n = (tr.meta[nb + node] >> 8) & 0xff;
            int best = 0;
            for (int i = 1; i < n; ++i) {
                const int ci = tr.count[nb + fc + i], cb = tr.count[nb + fc + best];
                if (ci > cb || (ci == cb && tr.value[nb + fc + i] > tr.value[nb + fc + best])) best = i;
            }
            if (len < a.cfg.horizon) plan[len] = (int8_t)(tr.meta[nb + fc + best] & 0xff);
            ++len;
            node = fc + best;
        }
        int32_t* res = a.result + (int64_t)tree * B2_MCTS_RESULT_WORDS;
        res[0] = n_nodes;
        res[1] = len;
        res[2] = env_steps;
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_mcts_plan(const b2_mcts_config* cfg, const int32_t* root_states, const b2_mcts_tree* tree,
                            uint64_t* rng, int8_t* plan, int32_t* result, void* stream_) {
    B2_REQUIRE(cfg && root_states && tree && rng && plan && result, "null pointer");
    B2_REQUIRE(cfg->n_trees > 0 && cfg->episodes >= 0 && cfg->horizon >= 0, "bad batch / budget");
    B2_REQUIRE(cfg->n_actions > 0 && cfg->n_actions <= MAX_BRANCH_MCTS, "n_actions must be in 1..8");
    B2_REQUIRE((int64_t)cfg->node_capacity >= 1 + (int64_t)cfg->episodes * cfg->n_actions, "node_capacity too small");
    B2_REQUIRE(cfg->gamma_pow && cfg->uniform_cdf, "gamma / cdf tables missing");
    B2_REQUIRE(cfg->rollout_policy >= 0 && cfg->rollout_policy <= 2 && cfg->prior_policy >= 0 && cfg->prior_policy <= 2,
               "policy must be 0 (random_available), 1 (random) or 2 (preference)");
    B2_REQUIRE((cfg->prior_policy != 2 || cfg->pref_prior) && (cfg->rollout_policy != 2 || cfg->pref_cdf),
               "preference policy tables missing");
    cudaStream_t stream = (cudaStream_t)stream_;
    MctsArgs a;
    a.cfg = *cfg; a.tree = *tree; a.root_states = root_states; a.rng = rng; a.plan = plan; a.result = result;
    if (cfg->env_kind == B2_ENV_FINITE) {
        B2_REQUIRE(cfg->mdp.transition && cfg->mdp.reward && cfg->mdp.terminal, "finite MDP tables missing");
        B2_REQUIRE(cfg->mdp.n_actions == cfg->n_actions, "mdp.n_actions != n_actions");
        const int grid = (cfg->n_trees + 127) / 128;
        mcts_kernel<FiniteEnv><<<grid, 128, 0, stream>>>(a);
    } else if (cfg->env_kind == B2_ENV_HIGHWAY) {
        B2_REQUIRE(cfg->n_actions == B2_HW_ACTIONS, "HighwayLite has 5 actions");
        const int grid = (cfg->n_trees * 16 + 127) / 128;
        mcts_kernel<HighwayEnv><<<grid, 128, 0, stream>>>(a);
    } else {
        set_error("unknown env_kind %d", cfg->env_kind);
        return B2_ERR_INVALID;
    }
    B2_CUDA_CHECK(cudaGetLastError());
    return B2_OK;
}
