// Handle-based host-buffer API over the device-pointer entry points: for callers
// that do not manage CUDA memory themselves (a cgo / JNI / plain-C host, or the
// `e2e` path of bench.py).  A handle owns the HBM arena of a batch of OPD trees;
// b2_opd_plan_host copies the root states in, searches, copies the plans and the
// per-tree result records out and synchronises.
#include <math.h>

#include <vector>

#include "common.cuh"

struct b2_opd_handle {
    b2_opd_config cfg;
    b2_opd_tree tree;
    void* workspace;
    int8_t* plan;
    int32_t* result;
    int32_t* roots;
    double *gamma_pow, *gamma_pow_div, *terminal_bonus;
    int32_t* mdp_T;
    double* mdp_R;
    uint8_t* mdp_term;
    cudaStream_t stream;
    size_t root_words;
};

#define B2_ALLOC(ptr, bytes) B2_CUDA_CHECK(cudaMalloc((void**)&(ptr), (bytes) ? (bytes) : 8))

extern "C" void b2_opd_destroy(b2_opd_handle* h) {
    if (!h) return;
    void* ptrs[] = {h->tree.parent, h->tree.first_child, h->tree.depth, h->tree.count, h->tree.meta, h->tree.reward,
                    h->tree.lower, h->tree.upper, h->tree.state, h->workspace, h->plan, h->result, h->roots,
                    h->gamma_pow, h->gamma_pow_div, h->terminal_bonus, h->mdp_T, h->mdp_R, h->mdp_term};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int b2_opd_create(const b2_opd_host_config* hc, b2_opd_handle** out) {
    B2_REQUIRE(hc && out, "null pointer");
    B2_REQUIRE(hc->n_trees > 0 && hc->n_actions > 0 && hc->budget >= 0, "bad batch / budget");
    B2_REQUIRE(1 + (int64_t)(hc->budget / hc->n_actions) * hc->n_actions < (int64_t)1 << 31, "budget too large for int32 node ids");
    B2_REQUIRE(hc->gamma >= 0.0 && hc->gamma < 1.0, "gamma must be in [0, 1)");
    b2_opd_handle* h = new b2_opd_handle();
    memset(h, 0, sizeof(*h));
    *out = nullptr;
    b2_opd_config& c = h->cfg;
    c.env_kind = hc->env_kind;
    c.n_trees = hc->n_trees;
    c.n_actions = hc->n_actions;
    c.n_expansions = hc->budget / hc->n_actions;                     // deterministic.py:118
    c.node_capacity = 1 + c.n_expansions * c.n_actions;
    c.plan_capacity = c.n_expansions + 1;
    c.keys_in_smem = hc->keys_in_smem;
    c.reserved = hc->kernel;
    c.terminal_reward = hc->terminal_reward;
    const size_t n = (size_t)c.n_trees * c.node_capacity;
    const bool hwy = c.env_kind == B2_ENV_HIGHWAY;
    h->root_words = hwy ? B2_HW_STATE_WORDS : 1;
    int rc = B2_OK;
    auto fail = [&](int code) { b2_opd_destroy(h); return code; };
#define TRY(expr) do { rc = [&]() -> int { expr; return B2_OK; }(); if (rc) return fail(rc); } while (0)
    TRY(B2_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)));
    TRY(B2_ALLOC(h->tree.parent, n * 4));
    TRY(B2_ALLOC(h->tree.first_child, n * 4));
    TRY(B2_ALLOC(h->tree.depth, n * 4));
    TRY(B2_ALLOC(h->tree.count, n * 4));
    TRY(B2_ALLOC(h->tree.meta, n * 4));
    TRY(B2_ALLOC(h->tree.reward, n * 8));
    TRY(B2_ALLOC(h->tree.lower, n * 8));
    TRY(B2_ALLOC(h->tree.upper, n * 8));
    TRY(B2_ALLOC(h->tree.state, n * 4 * h->root_words));
    TRY(B2_ALLOC(h->plan, (size_t)c.n_trees * c.plan_capacity));
    TRY(B2_ALLOC(h->result, (size_t)c.n_trees * B2_OPD_RESULT_WORDS * 4));
    TRY(B2_ALLOC(h->roots, (size_t)c.n_trees * h->root_words * 4));
    // gamma**d tables: C pow() is what CPython's float ** uses, so the values equal the reference's
    std::vector<double> gp(c.n_expansions + 2), gd(c.n_expansions + 2), tb(c.n_expansions + 2);
    for (int d = 0; d < c.n_expansions + 2; ++d) {
        gp[d] = pow(hc->gamma, (double)d);
        gd[d] = gp[d] / (1.0 - hc->gamma);
        tb[d] = (hc->terminal_reward * gp[d]) / (1.0 - hc->gamma);      // deterministic.py:60-63
    }
    TRY(B2_ALLOC(h->terminal_bonus, tb.size() * 8));
    TRY(B2_CUDA_CHECK(cudaMemcpy(h->terminal_bonus, tb.data(), tb.size() * 8, cudaMemcpyHostToDevice)));
    c.terminal_bonus = h->terminal_bonus;
    TRY(B2_ALLOC(h->gamma_pow, gp.size() * 8));
    TRY(B2_ALLOC(h->gamma_pow_div, gd.size() * 8));
    TRY(B2_CUDA_CHECK(cudaMemcpy(h->gamma_pow, gp.data(), gp.size() * 8, cudaMemcpyHostToDevice)));
    TRY(B2_CUDA_CHECK(cudaMemcpy(h->gamma_pow_div, gd.data(), gd.size() * 8, cudaMemcpyHostToDevice)));
    c.gamma_pow = h->gamma_pow;
    c.gamma_pow_div = h->gamma_pow_div;
    if (!hwy) {
        if (!(hc->mdp.transition && hc->mdp.reward && hc->mdp.terminal && hc->mdp.n_states > 0)) {
            b2::set_error("invalid argument: finite MDP host tables missing");
            return fail(B2_ERR_INVALID);
        }
        const size_t sa = (size_t)hc->mdp.n_states * hc->n_actions;
        TRY(B2_ALLOC(h->mdp_T, sa * 4));
        TRY(B2_ALLOC(h->mdp_R, sa * 8));
        TRY(B2_ALLOC(h->mdp_term, (size_t)hc->mdp.n_states));
        TRY(B2_CUDA_CHECK(cudaMemcpy(h->mdp_T, hc->mdp.transition, sa * 4, cudaMemcpyHostToDevice)));
        TRY(B2_CUDA_CHECK(cudaMemcpy(h->mdp_R, hc->mdp.reward, sa * 8, cudaMemcpyHostToDevice)));
        TRY(B2_CUDA_CHECK(cudaMemcpy(h->mdp_term, hc->mdp.terminal, hc->mdp.n_states, cudaMemcpyHostToDevice)));
        c.mdp.n_states = hc->mdp.n_states;
        c.mdp.n_actions = hc->n_actions;
        c.mdp.transition = h->mdp_T;
        c.mdp.reward = h->mdp_R;
        c.mdp.terminal = h->mdp_term;
    }
    const int64_t ws = b2_opd_workspace_bytes(&c);
    if (ws < 0) {
        b2::set_error("unsupported OPD configuration (node_capacity %d)", c.node_capacity);
        return fail(B2_ERR_UNSUPPORTED);
    }
    TRY(B2_ALLOC(h->workspace, (size_t)ws));
#undef TRY
    *out = h;
    return B2_OK;
}

extern "C" int b2_opd_plan_host(b2_opd_handle* h, const int32_t* root_states_host, int8_t* plan_host,
                                int32_t* result_host) {
    B2_REQUIRE(h && root_states_host && plan_host && result_host, "null pointer");
    const b2_opd_config& c = h->cfg;
    B2_CUDA_CHECK(cudaMemcpyAsync(h->roots, root_states_host, (size_t)c.n_trees * h->root_words * 4,
                                  cudaMemcpyHostToDevice, h->stream));
    int rc = b2_opd_plan(&h->cfg, h->roots, &h->tree, h->workspace, h->plan, h->result, h->stream);
    if (rc) return rc;
    B2_CUDA_CHECK(cudaMemcpyAsync(plan_host, h->plan, (size_t)c.n_trees * c.plan_capacity, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA_CHECK(cudaMemcpyAsync(result_host, h->result, (size_t)c.n_trees * B2_OPD_RESULT_WORDS * 4,
                                  cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

extern "C" int b2_opd_copy_tree(b2_opd_handle* h, int32_t tree, int32_t n_nodes, int32_t* parent, int32_t* first_child,
                                int32_t* count, int32_t* meta, double* reward, double* lower, double* upper) {
    B2_REQUIRE(h && tree >= 0 && tree < h->cfg.n_trees && n_nodes >= 0 && n_nodes <= h->cfg.node_capacity, "bad tree / size");
    const size_t off = (size_t)tree * h->cfg.node_capacity;
    struct { void* dst; const void* src; size_t elem; } jobs[] = {
        {parent, h->tree.parent + off, 4}, {first_child, h->tree.first_child + off, 4}, {count, h->tree.count + off, 4},
        {meta, h->tree.meta + off, 4},     {reward, h->tree.reward + off, 8},           {lower, h->tree.lower + off, 8},
        {upper, h->tree.upper + off, 8}};
    for (auto& j : jobs)
        if (j.dst) B2_CUDA_CHECK(cudaMemcpyAsync(j.dst, j.src, j.elem * n_nodes, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

extern "C" int32_t b2_opd_plan_capacity(const b2_opd_handle* h) { return h ? h->cfg.plan_capacity : -1; }
