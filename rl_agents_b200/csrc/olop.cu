// OLOP / KL-OLOP -- the plan() loop of rl_agents/agents/tree_search/olop.py for
// a BATCH of independent decisions, strict episode order inside each tree, the
// planner's numpy PCG64 stream consumed exactly as the reference does
// (`np_random.randint(2**30)` per episode, olop.py:73; `choice(children)` for the
// "uniform" continuation, :80-81).  The KL upper confidence bound
// (rl_agents/utils.py:123-203: damped Newton iteration on the Bernoulli KL) runs
// in-kernel in fp64.
//
// Same lane-group mapping as mcts.cu: one tree per 16-lane group (HighwayLite,
// lane = vehicle slot) or per lane (finite MDP).
#include "common.cuh"
#include "highway_lite.cuh"
#include "pcg64.cuh"

namespace b2 {

struct OlopArgs {
    b2_olop_config cfg;
    b2_olop_tree tree;
    const int32_t* root_states;
    uint64_t* rng;
    int8_t* plan;
    int32_t* result;
};

// bernoulli_kullback_leibler (utils.py:89-106)
__device__ __forceinline__ double bernoulli_kl(double p, double q) {
    double kl1 = 0.0, kl2 = INFINITY;
    if (p > 0.0 && q > 0.0) kl1 = p * log(p / q);
    if (q < 1.0) kl2 = p < 1.0 ? (1.0 - p) * log((1.0 - p) / (1.0 - q)) : 0.0;
    return kl1 + kl2;
}

// kl_upper_bound(_sum, count, threshold) with eps = 1e-2 (utils.py:123-147) through
// newton_iteration (:150-203): start at the midpoint of [mu, 1], pull back with
// weight 0.9 when a step leaves the interval, stop on |dx| <= eps or 100 iterations.
__device__ double kl_upper_bound(double sum, int count, double threshold) {
    if (count == 0) return 1.0;
    const double eps = 1e-2, weight = 0.9;
    const double mu = sum / (double)count;
    const double max_div = threshold / (double)count;
    const double a = mu, b = 1.0;
    if (a == b) return a;
    double x = INFINITY, x_next = (a + b) / 2.0;
    int iterations = 0;
    while (fabs(x - x_next) > eps && iterations < 100) {
        ++iterations;
        x = x_next;
        const double f_x = bernoulli_kl(mu, x) - max_div;
        double df_x;
        if (x == 0.0 || x == 1.0)      // Python float division raises ZeroDivisionError (:183-186)
            df_x = (f_x - (bernoulli_kl(mu, x - eps) - max_div)) / eps;
        else
            df_x = (1.0 - mu) / (1.0 - x) - mu / x;
        if (df_x != 0.0) x_next = x - f_x / df_x;
        if (x_next < a) x_next = weight * a + (1.0 - weight) * x;
        else if (x_next > b) x_next = weight * b + (1.0 - weight) * x;
    }
    if (x_next < a) x_next = a;
    if (x_next > b) x_next = b;
    return x_next;
}

struct OFiniteEnv {
    static constexpr int GROUP = 1;
    int s;
    __device__ __forceinline__ void load_root(const OlopArgs& a, int tree, int li) { s = a.root_states[tree]; }
    __device__ __forceinline__ int avail(const OlopArgs& a, unsigned gmask) const { return (1 << a.cfg.n_actions) - 1; }
    __device__ __forceinline__ static int nth(int mask, int n) { return n; }
    __device__ __forceinline__ double step(const OlopArgs& a, int action, int li, unsigned gmask, float* gs, bool& term) {
        const b2_finite_mdp& m = a.cfg.mdp;
        const double r = m.reward[(int64_t)s * m.n_actions + action];
        term = m.terminal[s] != 0;        // done = terminal[state BEFORE the transition]
        s = m.transition[(int64_t)s * m.n_actions + action];
        return r;
    }
};

struct OHighwayEnv {
    static constexpr int GROUP = 16;
    hw::Lane L;
    int t, si;
    __device__ __forceinline__ void load_root(const OlopArgs& a, int tree, int li) {
        hw::load_state(a.root_states + (int64_t)tree * hw::WORDS, li, L, t, si);
    }
    __device__ __forceinline__ int avail(const OlopArgs& a, unsigned gmask) const {
        return hw::avail_mask(__shfl_sync(gmask, L.y, 0, 16), si);
    }
    __device__ __forceinline__ static int nth(int mask, int n) { return hw::nth_action(mask, n); }
    __device__ __forceinline__ double step(const OlopArgs& a, int action, int li, unsigned gmask, float* gs, bool& term) {
        bool trunc;
        return (double)hw::step(L, li, t, si, action, term, trunc, gmask, gs);
    }
};

template <class Env>
__global__ void __launch_bounds__(128, 8) olop_kernel(OlopArgs a) {
    constexpr int G = Env::GROUP;
    __shared__ float scratch[G == 16 ? 128 / 16 : 1][hw::SCRATCH_FLOATS];
    const int gtid = blockIdx.x * 128 + threadIdx.x;
    const int tree_raw = gtid / G, li = gtid % G;
    const bool live = tree_raw < a.cfg.n_trees;
    const int tree = live ? tree_raw : a.cfg.n_trees - 1;
    const bool writer = live && li == 0;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = G == 1 ? (1u << lane) : (0xFFFFu << (lane & 16));
    const int L = a.cfg.horizon;
    const int64_t nb = (int64_t)tree * a.cfg.node_capacity;
    const b2_olop_tree& tr = a.tree;
    const double gamma = a.cfg.gamma;

    Pcg64 rng;
    rng.load(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
    if (writer) {   // OLOPNode(parent=None) (olop.py:106-124)
        tr.parent[nb] = -1; tr.first_child[nb] = -1; tr.count[nb] = 0; tr.meta[nb] = 0xff;
        tr.cumulative[nb] = 0.0; tr.mu_ucb[nb] = a.cfg.kl ? 1.0 : INFINITY; tr.upper[nb] = a.cfg.init_upper[0];
    }
    __syncwarp(gmask);
    int n_nodes = 1, error = 0;

    for (int ep = 0; ep < a.cfg.episodes; ++ep) {
        Env env;
        env.load_root(a, tree, li);                 // safe_deepcopy_env(state), olop.py:98
        if (live) rng.integers(1u << 30);            // state.seed(np_random.randint(2**30)), :73
        int node = 0;
        const double threshold = a.cfg.thresholds[ep];
        for (int h = 0; h < L; ++h) {
            int action = 0, child = 0;
            const int amask = env.avail(a, gmask);
            if (live && !error) {
                int fc = tr.first_child[nb + node];
                if (fc < 0) {
                    // expand (olop.py:165-180): one child per available action
                    const int n = __popc(amask);
                    if (writer) {
                        for (int i = 0; i < n; ++i) {
                            const int c = n_nodes + i;
                            tr.parent[nb + c] = node; tr.first_child[nb + c] = -1; tr.count[nb + c] = 0;
                            tr.meta[nb + c] = Env::nth(amask, i);
                            tr.cumulative[nb + c] = 0.0; tr.mu_ucb[nb + c] = a.cfg.kl ? 1.0 : INFINITY;
                            tr.upper[nb + c] = a.cfg.init_upper[h + 1];
                        }
                        tr.first_child[nb + node] = n_nodes;
                        tr.meta[nb + node] = (tr.meta[nb + node] & ~0xff00) | (n << 8);
                    }
                    fc = n_nodes;
                    n_nodes += n;
                    __syncwarp(gmask);
                    if (a.cfg.continuation == 1) {           // "uniform": np_random.choice(children keys)
                        child = fc + (int)rng.integers((uint32_t)n);
                    } else {                                 // "zeros": children[0] -- KeyError when unavailable
                        child = -1;
                        for (int i = 0; i < n; ++i)
                            if (Env::nth(amask, i) == 0) child = fc + i;
                        if (child < 0) { error = 2; child = fc; }
                    }
                } else {
                    // UCB elsewhere: first arg-max of value_upper (olop.py:84)
                    const int n = (tr.meta[nb + node] >> 8) & 0xff;
                    child = fc;
                    double best = tr.upper[nb + fc];
                    for (int i = 1; i < n; ++i) {
                        const double u = tr.upper[nb + fc + i];
                        if (u > best) { best = u; child = fc + i; }
                    }
                }
                action = tr.meta[nb + child] & 0xff;
            }
            bool term;
            const double r = env.step(a, action, li, gmask, scratch[(threadIdx.x >> 4) % (128 / 16)], term);     // olop.py:87
            if (live && !error) {
                node = child;
                // update (olop.py:132-142)
                if (!(r >= 0.0 && r <= 1.0)) error = 1;
                if (writer) {
                    int meta = tr.meta[nb + node];
                    if (term) meta |= 1 << 16;
                    const double rr = (meta >> 16) & 1 ? 0.0 : r;
                    const double cum = tr.cumulative[nb + node] + rr;
                    const int cnt = tr.count[nb + node] + 1;
                    tr.meta[nb + node] = meta;
                    tr.cumulative[nb + node] = cum;
                    tr.count[nb + node] = cnt;
                    if (a.cfg.kl) tr.mu_ucb[nb + node] = kl_upper_bound(cum, cnt, threshold);   // :144-163
                }
                __syncwarp(gmask);
            }
        }
        // backup_to_root (olop.py:182-193)
        if (writer && !error) {
            int n = node;
            while (n >= 0) {
                const int fc = tr.first_child[nb + n];
                if (fc >= 0) {
                    const int k = (tr.meta[nb + n] >> 8) & 0xff;
                    double m = tr.upper[nb + fc];
                    for (int i = 1; i < k; ++i) {
                        const double u = tr.upper[nb + fc + i];
                        m = u > m ? u : m;
                    }
                    tr.upper[nb + n] = tr.mu_ucb[nb + n] + gamma * m;
                } else {
                    tr.upper[nb + n] = tr.mu_ucb[nb + n];
                }
                n = tr.parent[nb + n];
            }
        }
        __syncwarp(gmask);
    }

    if (writer) {
        rng.store(a.rng + (int64_t)tree * B2_PCG64_STATE_WORDS);
        // get_plan with OLOPNode.selection_rule (olop.py:126-130)
        int8_t* plan = a.plan + (int64_t)tree * max(L, 1);
        int node = 0, len = 0;
        while (tr.first_child[nb + node] >= 0) {
            const int fc = tr.first_child[nb + node];
            const int n = (tr.meta[nb + node] >> 8) & 0xff;
            int best = 0;
            for (int i = 1; i < n; ++i) {
                const int ci = tr.count[nb + fc + i], cb = tr.count[nb + fc + best];
                if (ci > cb || (ci == cb && tr.upper[nb + fc + i] > tr.upper[nb + fc + best])) best = i;
            }
            if (len < L) plan[len] = (int8_t)(tr.meta[nb + fc + best] & 0xff);
            ++len;
            node = fc + best;
        }
        int32_t* res = a.result + (int64_t)tree * B2_OLOP_RESULT_WORDS;
        res[0] = n_nodes;
        res[1] = len;
        res[2] = error;
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_olop_plan(const b2_olop_config* cfg, const int32_t* root_states, const b2_olop_tree* tree,
                            uint64_t* rng, int8_t* plan, int32_t* result, void* stream_) {
    B2_REQUIRE(cfg && root_states && tree && rng && plan && result, "null pointer");
    B2_REQUIRE(cfg->n_trees > 0 && cfg->episodes >= 0 && cfg->horizon >= 1, "bad batch / budget");
    B2_REQUIRE(cfg->n_actions > 0 && cfg->n_actions <= 8, "n_actions must be in 1..8");
    B2_REQUIRE((int64_t)cfg->node_capacity >= 1 + (int64_t)cfg->episodes * cfg->horizon * cfg->n_actions,
               "node_capacity too small");
    B2_REQUIRE(cfg->thresholds && cfg->init_upper, "threshold / initial bound tables missing");
    cudaStream_t stream = (cudaStream_t)stream_;
    OlopArgs a;
    a.cfg = *cfg; a.tree = *tree; a.root_states = root_states; a.rng = rng; a.plan = plan; a.result = result;
    if (cfg->env_kind == B2_ENV_FINITE) {
        B2_REQUIRE(cfg->mdp.transition && cfg->mdp.reward && cfg->mdp.terminal, "finite MDP tables missing");
        B2_REQUIRE(cfg->mdp.n_actions == cfg->n_actions, "mdp.n_actions != n_actions");
        olop_kernel<OFiniteEnv><<<(cfg->n_trees + 127) / 128, 128, 0, stream>>>(a);
    } else if (cfg->env_kind == B2_ENV_HIGHWAY) {
        B2_REQUIRE(cfg->n_actions == B2_HW_ACTIONS, "HighwayLite has 5 actions");
        olop_kernel<OHighwayEnv><<<(cfg->n_trees * 16 + 127) / 128, 128, 0, stream>>>(a);
    } else {
        set_error("unknown env_kind %d", cfg->env_kind);
        return B2_ERR_INVALID;
    }
    B2_CUDA_CHECK(cudaGetLastError());
    return B2_OK;
}
