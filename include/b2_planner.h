/*
 * b2_planner.h -- C ABI of libb2planner.so, the B200 (sm_100a) planning engine
 * behind the rl-agents plugin surface.
 *
 * Every entry point replaces the inner loop of one reference method (cited as
 * file:line under /root/reference).  Conventions:
 *   - plain C, no allocation, no ownership: every buffer is caller-owned device
 *     memory (e.g. torch.Tensor.data_ptr()) unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and return
 *     without synchronising; the caller synchronises before reading results;
 *   - return 0 on success, otherwise a non-zero code; b2_last_error() gives the
 *     message for the calling thread;
 *   - distinct handles / buffers are independent: safe from different host
 *     threads or processes (scripts/experiments.py:105 forks worker processes).
 */
#ifndef B2_PLANNER_H
#define B2_PLANNER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_INVALID 1   /* bad argument */
#define B2_ERR_CUDA 2      /* CUDA runtime error, see b2_last_error() */
#define B2_ERR_UNSUPPORTED 3

const char* b2_last_error(void);
int b2_version(void);
/* number of SMs / device name of the current device (diagnostics for bench.py) */
int b2_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ------------------------------------------------------------------------
 * Environment models (the batched transition the planners call)
 * ---------------------------------------------------------------------- */
#define B2_ENV_FINITE 0   /* deterministic finite MDP tables               */
#define B2_ENV_HIGHWAY 1  /* HighwayLite, docs/HIGHWAY_LITE_SPEC.md         */

#define B2_ENV_INTERSECTION 2 /* IntersectionLite, docs/INTERSECTION_LITE_SPEC.md (wavefront OPD + b2_intersection_step) */

#define B2_HW_STATE_WORDS 136 /* 32-bit words of one HighwayLite / IntersectionLite state */
#define B2_HW_ACTIONS 5
#define B2_IL_ACTIONS 3

typedef struct b2_finite_mdp {
    int32_t n_states;
    int32_t n_actions;
    const int32_t* transition; /* [S, A] next state (deterministic mode)     */
    const double* reward;      /* [S, A]                                     */
    const uint8_t* terminal;   /* [S]                                        */
} b2_finite_mdp;

/* One decision step of n_envs HighwayLite states (15 physics sub-steps each).
 * Replaces `env.step(a)` on a deep-copied env: deterministic.py:36-43,
 * mcts.py:145,173.  states: [n_envs, 136] words, updated in place.
 * actions: [n_envs] int32.  reward: [n_envs] float; flags: [n_envs] int32,
 * bit0 terminated, bit1 truncated.  avail_mask (nullable): [n_envs] int32
 * bitmask of get_available_actions() in the NEW state. */
int b2_highway_step(int32_t* states, const int32_t* actions, float* reward, int32_t* flags,
                    int32_t* avail_mask, int32_t n_envs, void* stream);

/* ValueIterationAgent on HighwayLite scenes, batched: per scene, the time-to-collision grid MDP that
 * `env.unwrapped.to_finite_mdp()` hands the reference's agent (rl_agents/agents/dynamic_programming/
 * value_iteration.py:17,32; docs/HIGHWAY_LITE_SPEC.md section 9: 3 speeds x 4 lanes x 10 s = 120 states, 5 actions,
 * deterministic) is built and the agent's fixed-point iteration (value_iteration.py:42-73, incl. the np.allclose early
 * exit that returns the previous iterate) is run by one warp in shared memory.
 * states [n_envs,136] i32; q_out [n_envs,120,5] f64 or NULL; action_out [n_envs] i32 = argmax_a Q[state] (:35);
 * mdp_state_out [n_envs] i32 (the scene's cell, `mdp.state`) or NULL; sweeps_out [n_envs] i32 or NULL. */
#define B2_TTC_STATES 120
int b2_highway_ttc_vi(const int32_t* states, int32_t n_envs, double gamma, int32_t iterations, double rtol,
                      double atol, double* q_out, int32_t* action_out, int32_t* mdp_state_out, int32_t* sweeps_out,
                      void* stream);

/* The same for IntersectionLite (BASELINE config C5's env model): 3 actions (0 SLOWER, 1 IDLE, 2 FASTER),
 * states [n_envs, 136] words, avail_mask in the NEW state. */
int b2_intersection_step(int32_t* states, const int32_t* actions, float* reward, int32_t* flags,
                         int32_t* avail_mask, int32_t n_envs, void* stream);

/* Self-test (tests/test_gpu_engines.py): the HighwayLite kernel divides by two constants of the spec with a
 * 3-instruction sequence instead of the full IEEE division; this compares both, exhaustively over every fp32
 * mantissa, both signs and exponents -60..60, and writes the number of differing results (must be 0). */
int b2_selftest_const_division(unsigned long long* mismatches_dev, void* stream);

/* ------------------------------------------------------------------------
 * Value iteration -- rl_agents/agents/dynamic_programming/value_iteration.py
 * ---------------------------------------------------------------------- */
#define B2_VI_DETERMINISTIC 0 /* next_v = V[T]                (:52-53)       */
#define B2_VI_STOCHASTIC 1    /* next_v = sum_s' P[s,a,s']V[s'] (:54-55)     */
#define B2_VI_SPARSE 2        /* next_v = sum_b P[s,a,b]V[N[s,a,b]] (:56-59) */

typedef struct b2_vi_problem {
    int32_t mode;
    int32_t n_actions;   /* A                                                */
    int32_t n_next;      /* B (sparse), S (stochastic), ignored otherwise    */
    int32_t reserved;    /* kernel choice: 0 auto (register kernel when the shape
                            allows, else tiled), 1 tiled, 2 TMA-staged tiled    */
    int64_t n_states;    /* S of the whole MDP (length of V)                 */
    int64_t row_begin;   /* state slab [row_begin, row_end) owned by the call */
    int64_t row_end;
    double gamma;
    double rtol, atol;   /* np.allclose tolerances (1e-5, 1e-8)              */
    /* slab-local tables: row 0 is state row_begin */
    const void* transition; /* int32 [rows,A] | double [rows,A,S] | double [rows,A,B] */
    const int32_t* next;    /* int32 [rows,A,B] (sparse only)                */
    const double* reward;   /* [rows, A]                                     */
    const uint8_t* terminal;/* [rows]                                        */
} b2_vi_problem;

/* One Bellman sweep  Q' = R + gamma * E[V(s')]  fused with V' = max_a Q' and the
 * np.allclose(Q, Q') test of fixed_point_iteration (value_iteration.py:51-63,
 * 65-73).  v_in/v_out: [S] (v_out rows of the slab are written); q_old/q_new:
 * slab-local [rows, A].  viol: int32 [>= sweep_index+1] device counters,
 * zero-initialised by the caller: the sweep adds the number of elements that
 * break allclose to viol[sweep_index]; if sweep_index > 0 and
 * viol[sweep_index-1] == 0 (previous sweep converged) the launch does nothing,
 * so a whole fixed-point loop can be enqueued without a host round trip. */
int b2_vi_sweep(const b2_vi_problem* p, const double* v_in, const double* q_old, double* q_new,
                double* v_out, int32_t* viol, int32_t sweep_index, void* stream);

/* get_state_action_value (value_iteration.py:42-45): `iterations` sweeps
 * ping-ponging q[0]/q[1] and v[0]/v[1] (q[0], v[0] must hold the initial
 * zeros).  After synchronising, the first k with viol[k]==0 marks convergence:
 * the result (the OLD iterate, :70-72) is q[k%2]; without one it is
 * q[iterations%2].  Single device (row slab = all states). */
int b2_vi_solve(const b2_vi_problem* p, double* q0, double* q1, double* v0, double* v1,
                int32_t* viol, int32_t iterations, void* stream);

/* ---- slab-sharded value iteration with the exchange fused into the sweep over NVLink peer memory ----
 * Replaces "sweep; ncclAllGather(V); ncclAllReduce(violations)" of the sharded fixed-point loop
 * (value_iteration.py:42-73 over G GPUs, one process per GPU): every rank keeps the V ping-pong buffers, an
 * arrival-flag array [world] and a violation table [iterations, world] in CUDA-IPC shared device memory
 * (b2_p2p_alloc / export / import); the sweep kernel stores V' into every rank's copy and publishes its count and
 * flag when its last CTA retires; the next sweep acquires the flags.  No NCCL call inside the loop. */
#define B2_MAX_PEERS 8
int b2_p2p_alloc(int64_t bytes, void** ptr);                          /* zeroed, IPC-exportable device memory */
int b2_p2p_free(void* ptr);
int b2_p2p_export(void* ptr, unsigned char* handle64);                /* cudaIpcGetMemHandle (64 bytes)        */
int b2_p2p_import(const unsigned char* handle64, void** peer_ptr);    /* cudaIpcOpenMemHandle + peer access    */
int b2_p2p_close(void* peer_ptr);
int b2_p2p_memset(void* ptr, int32_t value, int64_t bytes, void* stream);
int b2_p2p_read(void* dst_host, const void* src_dev, int64_t bytes, void* stream);   /* synchronous D2H */

typedef struct b2_vi_p2p {
    int32_t world, rank;
    double* v[2][B2_MAX_PEERS];      /* v[i][r]: V ping-pong buffer i ([S] doubles) in rank r's memory          */
    int32_t* flags[B2_MAX_PEERS];    /* flags[r]: rank r's [world] arrival flags; this rank writes flags[r][rank] */
    int32_t* parts[B2_MAX_PEERS];    /* parts[r]: rank r's [iterations, world] violation table                   */
    int32_t* viol_local;             /* [iterations] this rank's own counters (local scratch, zeroed)            */
    uint32_t* done;                  /* [iterations] retired-CTA counters (local scratch, zeroed)                */
    int32_t* status;                 /* [1] local scratch, zeroed: set to 1 when a peer's flag did not arrive within
                                        ~1 s (the sweep then proceeds on stale data instead of hanging the GPU)    */
} b2_vi_p2p;

/* Sweep `sweep_index` of the slab [row_begin, row_end): reads v[sweep&1][rank] and q_old, writes q_new and
 * V' into v[(sweep+1)&1][r] of every rank r.  Sparse / deterministic mode, A a power of two <= 32,
 * B in {1,2,4,8}.  After synchronising, sum_r parts[rank][k*world + r] is sweep k's allclose violation count
 * (same convergence protocol as b2_vi_sweep). */
int b2_vi_sweep_p2p(const b2_vi_problem* p, const b2_vi_p2p* x, const double* q_old, double* q_new,
                    int32_t sweep_index, void* stream);

/* Robust value iteration (rl_agents/agents/dynamic_programming/robust_value_iteration.py:39-58):
 * Q' = min over n_models models of R_m + gamma * E_m[V(s')], no terminal handling.
 * p->transition: int32 [M,S,A] (deterministic) or double [M,S,A,S] (stochastic);
 * p->reward: double [M,S,A]; p->terminal / p->next unused; rows = all states.
 * Same viol / early-exit protocol as b2_vi_sweep. */
int b2_vi_robust_sweep(const b2_vi_problem* p, int32_t n_models, const double* v_in, const double* q_old,
                       double* q_new, double* v_out, int32_t* viol, int32_t sweep_index, void* stream);

/* ------------------------------------------------------------------------
 * OPD -- rl_agents/agents/tree_search/deterministic.py
 * A batch of n_trees independent decisions, one tree per CTA, strict
 * best-first order inside each tree (bit-exact node order).
 * ---------------------------------------------------------------------- */
typedef struct b2_opd_config {
    int32_t env_kind;       /* B2_ENV_*                                      */
    int32_t n_trees;
    int32_t n_actions;      /* action_space.n: budget divisor (:118) and max branching */
    int32_t n_expansions;   /* budget // n_actions (:118)                    */
    int32_t node_capacity;  /* per tree, >= 1 + n_expansions * n_actions     */
    int32_t plan_capacity;  /* per tree, >= n_expansions + 1                 */
    int32_t keys_in_smem;   /* 1: frontier keys in shared memory when they fit */
    int32_t reserved;       /* HighwayLite batch kernel: 0 default (8 trees per CTA, packed
                               slots, block barriers between the phases), 1 one tree per warp,
                               2 8 trees per CTA as a dataflow over a work ring in shared
                               memory (no block barriers); identical trees, 0 is fastest */
    double terminal_reward; /* config["terminal_reward"] (:60-63)            */
    const double* gamma_pow;     /* [n_expansions+2] gamma**d   (host floats) */
    const double* gamma_pow_div; /* [n_expansions+2] gamma**d / (1 - gamma)   */
    b2_finite_mdp mdp;      /* env_kind == FINITE                            */
    const double* terminal_bonus;/* [n_expansions+2] (terminal_reward * gamma**d) / (1 - gamma), the
                                    reference's association (:60-63), host floats */
} b2_opd_config;

/* Node arrays, each [n_trees, node_capacity] (struct-of-arrays in HBM) */
typedef struct b2_opd_tree {
    int32_t* parent;       /* -1 for the root                                */
    int32_t* first_child;  /* -1 for a leaf                                  */
    int32_t* depth;
    int32_t* count;        /* DeterministicNode.count (:18,:64-65)           */
    int32_t* meta;         /* action | n_children << 8 | done << 16          */
    double* reward;
    double* lower;         /* value_lower                                    */
    double* upper;         /* value_upper                                    */
    int32_t* state;        /* FINITE: [n_trees, cap] state ids;
                              HIGHWAY: [n_trees, cap, 136] words             */
} b2_opd_tree;

#define B2_OPD_RESULT_WORDS 16
/* per tree int32 result record:
 * [0] n_nodes [1] n_leaves [2] max_depth [3] terminal_expansions
 * [4] error (1: reward outside [0,1], deterministic.py:46-47)
 * [5] plan_len [6] tie_node (-1, or the node where get_plan met a tie that the
 *     host must break with the planner RNG, abstract.py:304-311) */

/* bytes of scratch the call needs (frontier keys + tournament + expansion order) */
int64_t b2_opd_workspace_bytes(const b2_opd_config* cfg);

/* OptimisticDeterministicPlanner.plan (:116-122): root_states [n_trees] int32
 * state ids or [n_trees,136] words.  plan: int8 [n_trees, plan_capacity]
 * greedy value_lower path (abstract.py:143-156); result: int32
 * [n_trees, B2_OPD_RESULT_WORDS]. */
int b2_opd_plan(const b2_opd_config* cfg, const int32_t* root_states, const b2_opd_tree* tree,
                void* workspace, int8_t* plan, int32_t* result, void* stream);

/* ------------------------------------------------------------------------
 * Wavefront OPD: ONE decision searched by the whole GPU.  Per wave the
 * k = min(width, expansions left, frontier size) best leaves -- in the
 * reference's arg-max order (value_upper descending, node id ascending,
 * deterministic.py:110) -- are expanded in increasing node-id order and all their
 * children simulated at once.  width = 1 is the reference's algorithm
 * (deterministic.py:106-122); for any width the result is bit-identical with the
 * specification oracle/planners.py::opd_plan_wavefront.
 * ---------------------------------------------------------------------- */
#define B2_MAX_MODELS 8
typedef struct b2_opd_wave_config {
    int32_t env_kind;       /* B2_ENV_*                                      */
    int32_t n_actions;      /* action_space.n (:118)                         */
    int32_t n_expansions;   /* budget // n_actions (:118)                    */
    int32_t node_capacity;  /* >= 1 + n_expansions * n_actions, < 2^28       */
    int32_t plan_capacity;
    int32_t width;          /* leaves expanded per wave (>= 1)               */
    int32_t max_ctas;       /* 0: one CTA per SM                             */
    int32_t n_models;       /* 0: plain OPD.  M >= 1: DROP, the joint env of M models (rl_agents/agents/robust/
                               robust.py:9-47): root_state / tree.state hold M states per node ([M] ids or
                               [M,136] words), children follow the union of the models' available actions in
                               ascending order, a node's bounds are the minima over the models of the per-model
                               path bounds (deterministic.py:52-59 with vector rewards)            */
    const double* gamma_pow;      /* [n_expansions+2] gamma**d               */
    const double* gamma_pow_div;  /* [n_expansions+2] gamma**d / (1 - gamma) */
    const double* terminal_bonus; /* [n_expansions+2] terminal_reward * gamma**d / (1 - gamma) (:60-63) */
    b2_finite_mdp mdp;
    b2_finite_mdp model_mdps[8];  /* env_kind FINITE and n_models > 0: one deterministic MDP per model */
} b2_opd_wave_config;

int64_t b2_opd_wave_workspace_bytes(const b2_opd_wave_config* cfg);
/* tree: node arrays of ONE tree ([node_capacity] each, state [node_capacity(,136)]); root_state: [1] state id
 * or [136] words; plan: int8 [plan_capacity]; result: int32 [B2_OPD_RESULT_WORDS] as for b2_opd_plan, plus
 * [7] number of waves.  Cooperative launch on `stream` (one CTA per SM). */
int b2_opd_plan_wave(const b2_opd_wave_config* cfg, const int32_t* root_state, const b2_opd_tree* tree,
                     void* workspace, int8_t* plan, int32_t* result, void* stream);

/* Speculative strict search: the reference's own one-leaf-per-iteration order (deterministic.py:106-114) -- the
 * tree is bit-identical with b2_opd_plan / width 1 -- searched by the whole GPU.  Per wave the `width` best
 * frontier leaves (value_upper descending, node id ascending) are simulated on every SM unless already cached,
 * and the longest prefix the strict order would have expanded is committed.  Same config struct (width in
 * 1..256, n_models = 0, node_capacity <= 24576); tree->state is an ARENA of b2_opd_spec_arena_slots(cfg)
 * states (a node's state is not at its own index); result as for b2_opd_plan_wave. */
int64_t b2_opd_spec_workspace_bytes(const b2_opd_wave_config* cfg);
int64_t b2_opd_spec_arena_slots(const b2_opd_wave_config* cfg);
int b2_opd_plan_spec(const b2_opd_wave_config* cfg, const int32_t* root_state, const b2_opd_tree* tree,
                     void* workspace, int8_t* plan, int32_t* result, void* stream);

/* ------------------------------------------------------------------------
 * GBOP-T -- rl_agents/agents/tree_search/state_aware.py (StateAwarePlanner), deterministic finite MDPs:
 * OPD whose leaf bounds share one value table per STATE (:66-68), tightened by a breadth-first backup through
 * the nodes aggregated by state (:42-64), with pruning of dominated leaves after every expansion (:28-40).
 * A batch of n_trees independent decisions, one warp per tree, node order / leaves / state values exactly
 * the reference's.
 * ---------------------------------------------------------------------- */
typedef struct b2_gbop_config {
    int32_t n_trees;
    int32_t n_actions;
    int32_t n_expansions;        /* budget // n_actions (deterministic.py:118)      */
    int32_t node_capacity;       /* per tree, >= 1 + n_expansions * n_actions       */
    int32_t plan_capacity;
    int32_t queue_capacity;      /* entries of the backup FIFO (result[7] = 1 on overflow) */
    int32_t backup_aggregated_nodes;   /* config key of the same name (:80-85)      */
    int32_t prune_suboptimal_leaves;
    double gamma;
    double default_value;        /* 1 / (1 - gamma), host float (:76-77)            */
    double accuracy_scale;       /* accuracy * (1 - gamma), host float (:61)        */
    const double* gamma_pow;     /* [n_expansions+2] gamma**d                       */
    const double* terminal_bonus;/* [n_expansions+2] terminal_reward * gamma**d / (1 - gamma) */
    b2_finite_mdp mdp;
} b2_gbop_config;

typedef struct b2_gbop_tree {    /* [n_trees, node_capacity] each */
    int32_t* parent;
    int32_t* first_child;
    int32_t* depth;
    int32_t* count;
    int32_t* meta;               /* action | n_children << 8 | done << 16 | still-a-leaf << 17 */
    double* reward;
    double* lower;               /* value_lower (the path sum: GBOP-T never backs it up) */
    int32_t* obs;                /* the state the node reached                      */
} b2_gbop_tree;

int64_t b2_gbop_workspace_bytes(const b2_gbop_config* cfg);
/* The first 8 * n_states bytes of every tree's workspace slice hold the state value table afterwards.
 * result: int32 [n_trees, B2_OPD_RESULT_WORDS]: [0] n_nodes [1] n_leaves [4] error [5] plan_len [6] tie_node
 * [7] queue overflow [8] expansions done. */
int b2_gbop_plan(const b2_gbop_config* cfg, const int32_t* root_states, const b2_gbop_tree* tree, void* workspace,
                 int8_t* plan, int32_t* result, void* stream);

/* GBOP-D -- rl_agents/agents/tree_search/graph_based.py (GraphBasedPlanner), deterministic finite MDPs: one graph
 * node per state with value_lower / value_upper, optimistic descent to an unexpanded state, breadth-first partial
 * value iteration through the expanded parents (pushed in ascending state id; the reference iterates a Python set).
 * n_trees independent decisions, one warp each; rng: numpy PCG64 states [n_trees, 6] consumed by the tie-breaks of
 * sampling_rule (:22-30) exactly as Generator.choice does, advanced in place. */
typedef struct b2_gbopd_config {
    int32_t n_trees;
    int32_t n_actions;
    int32_t n_epochs;            /* budget // n_actions (:119)                       */
    int32_t sampling_timeout;    /* config["sampling_timeout"] (default 100)         */
    int32_t plan_capacity;       /* >= sampling_timeout                              */
    int32_t queue_capacity;      /* per tree; result[7] = 1 on overflow              */
    double gamma;
    double default_value;        /* 1 / (1 - gamma) (:17)                            */
    double accuracy;             /* config["accuracy"] (default 1e-2, :74)           */
    b2_finite_mdp mdp;           /* terminal unused: GBOP-D ignores `done` (:46)     */
    const int32_t* rev_ptr;      /* [S+1] reverse transitions: states p with T[p, a] == s for some a, */
    const int32_t* rev_idx;      /* ascending, unique                                */
} b2_gbopd_config;
/* lower / upper: double [n_trees, S]; flags: uint8 [n_trees, S] (bit0 node exists, bit1 expanded); queue: int32
 * [n_trees, queue_capacity] scratch; plan: int8 [n_trees, plan_capacity]; result: [0] nodes [1] expansions
 * [2] epochs that found no sink [5] plan_len [7] queue overflow. */
int b2_gbopd_plan(const b2_gbopd_config* cfg, const int32_t* root_states, double* lower, double* upper, uint8_t* flags,
                  int32_t* queue, uint64_t* rng, int8_t* plan, int32_t* result, void* stream);

/* Host-buffer convenience API (callers that do not manage CUDA memory: plain C, cgo, JNI ...).
 * A handle owns the device arena of a batch of trees; *_host pointers are ordinary host memory
 * (pinned memory makes the copies asynchronous); b2_opd_plan_host is synchronous. */
typedef struct b2_opd_handle b2_opd_handle;
typedef struct b2_opd_host_config {
    int32_t env_kind, n_trees, n_actions;
    int32_t budget;          /* config["budget"]; n_expansions = budget / n_actions (:118) */
    int32_t keys_in_smem, kernel;
    double gamma;            /* config["gamma"], 0 <= gamma < 1                     */
    double terminal_reward;
    b2_finite_mdp mdp;       /* HOST tables when env_kind == B2_ENV_FINITE          */
} b2_opd_host_config;
int b2_opd_create(const b2_opd_host_config* cfg, b2_opd_handle** out);
void b2_opd_destroy(b2_opd_handle* h);
int32_t b2_opd_plan_capacity(const b2_opd_handle* h);   /* bytes per tree in plan_host */
/* root_states_host: [n_trees] state ids or [n_trees,136] words; plan_host: int8
 * [n_trees, plan_capacity]; result_host: int32 [n_trees, B2_OPD_RESULT_WORDS]. */
int b2_opd_plan_host(b2_opd_handle* h, const int32_t* root_states_host, int8_t* plan_host, int32_t* result_host);
/* Node arrays of one tree (first n_nodes entries) to host buffers; any pointer may be NULL. */
int b2_opd_copy_tree(b2_opd_handle* h, int32_t tree, int32_t n_nodes, int32_t* parent, int32_t* first_child,
                     int32_t* count, int32_t* meta, double* reward, double* lower, double* upper);

/* ------------------------------------------------------------------------
 * MCTS -- rl_agents/agents/tree_search/mcts.py (open loop)
 * ---------------------------------------------------------------------- */
typedef struct b2_mcts_config {
    int32_t env_kind;
    int32_t n_trees;
    int32_t n_actions;
    int32_t episodes;        /* config["episodes"] (:180)                    */
    int32_t horizon;         /* config["horizon"]                            */
    int32_t node_capacity;   /* per tree, >= 1 + episodes * n_actions        */
    int32_t rollout_policy;  /* 0 random_available, 1 random, 2 preference (:46-97) */
    int32_t prior_policy;    /* idem, for expansion priors                   */
    double temperature;      /* config["temperature"] (:127)                 */
    const double* gamma_pow; /* [horizon+1] gamma**d                         */
    const double* uniform_cdf; /* [(n_actions+1), n_actions]: row n = cumsum(ones(n)/n)/last,
                                  the cdf Generator.choice(a, 1, p) searches (host numpy)  */
    b2_finite_mdp mdp;
    /* "preference" policies (mcts.py:76-97): the preferred action label and, for n available actions with the
     * preferred one at position k-1 of the env's action order (k = 0: not available -> uniform), row
     * [(n * (n_actions+1) + k) * n_actions ..] of a host-made table: probabilities (prior policy) / the cdf
     * Generator.choice searches (rollout policy).  Only read when the policy is 2. */
    int32_t prior_pref_action, rollout_pref_action;
    const double* pref_prior;    /* [(n_actions+1), (n_actions+1), n_actions] */
    const double* pref_cdf;      /* [(n_actions+1), (n_actions+1), n_actions] */
    const int32_t* resume_nodes; /* nullable [n_trees]: > 0 -> the tree arrays already hold that many nodes
                                    (a re-rooted sub-tree, step_strategy "subtree", abstract.py:195-206,
                                    mcts.py:129-130) and the search continues from them; the caller sizes
                                    node_capacity >= resume + episodes * n_actions                      */
} b2_mcts_config;

typedef struct b2_mcts_tree {
    int32_t* parent;
    int32_t* first_child;
    int32_t* count;        /* MCTSNode.count (:248-255)                      */
    int32_t* meta;         /* action | n_children << 8                       */
    double* value;         /* MCTSNode.value                                 */
    double* prior;
} b2_mcts_tree;

#define B2_PCG64_STATE_WORDS 6 /* uint64: state hi,lo, inc hi,lo, has_uint32, uinteger */
#define B2_MCTS_RESULT_WORDS 8
/* per tree int32 result: [0] n_nodes [1] plan_len [2] env steps taken */

/* MCTS.plan (:179-184) for n_trees independent decisions, strict episode order
 * inside each tree, consuming each tree's numpy PCG64 stream exactly as
 * Generator.choice does (abstract.py:304-311, mcts.py:172).  rng: uint64
 * [n_trees, 6] numpy bit-generator states, advanced in place.  plan: int8 [n_trees, horizon]. */
int b2_mcts_plan(const b2_mcts_config* cfg, const int32_t* root_states, const b2_mcts_tree* tree,
                 uint64_t* rng, int8_t* plan, int32_t* result, void* stream);

/* ------------------------------------------------------------------------
 * Wavefront MCTS: ONE decision searched by the whole GPU.  The reference's episode (selection mcts.py:141-149,
 * expansion :151-154, rollout :160-177, backup :257-265, recommendation :212-218) run in waves of `width`
 * episodes: selections in episode order with virtual counts, counter-based randomness, exact fixed-point
 * value sums.  Bit-identical with the specification oracle/planners.py::mcts_plan_wavefront.
 * ---------------------------------------------------------------------- */
typedef struct b2_mcts_wave_config {
    int32_t env_kind;
    int32_t n_actions;
    int32_t episodes;        /* config["episodes"]                              */
    int32_t horizon;         /* config["horizon"] (<= 64)                       */
    int32_t node_capacity;   /* >= 1 + episodes * n_actions (episode e owns ids 1 + e*n_actions ..) */
    int32_t width;           /* episodes per wave, 1..1024                      */
    int32_t rollout_policy;  /* 0 random_available (the only one implemented)   */
    int32_t prior_policy;
    double temperature;      /* config["temperature"] (:127)                    */
    uint64_t seed;           /* counter-based generator seed                    */
    const double* gamma_pow; /* [horizon+1] gamma**h                            */
    b2_finite_mdp mdp;
    int32_t max_ctas;        /* 0: one CTA per SM                               */
    int32_t reserved;
} b2_mcts_wave_config;

typedef struct b2_mcts_wave_tree {   /* [node_capacity] each; unused ids keep parent == -2 */
    int32_t* parent;
    int32_t* first_child;
    int32_t* count;
    int32_t* meta;           /* action | n_children << 8                        */
    int64_t* vsum;           /* sum of the returns backed up through the node, 2^-40 units */
    double* value;           /* vsum * 2^-40 / count (written at the end)       */
} b2_mcts_wave_tree;

int64_t b2_mcts_wave_workspace_bytes(const b2_mcts_wave_config* cfg);
/* root_state: [1] state id or [136] words; plan: int8 [horizon]; result: int32 [B2_MCTS_RESULT_WORDS]:
 * [0] node_capacity [1] plan_len [2] env steps [3] waves [4..7] phase clocks.  Cooperative launch. */
int b2_mcts_plan_wave(const b2_mcts_wave_config* cfg, const int32_t* root_state, const b2_mcts_wave_tree* tree,
                      void* workspace, int8_t* plan, int32_t* result, void* stream);

/* ------------------------------------------------------------------------
 * OLOP / KL-OLOP -- rl_agents/agents/tree_search/olop.py
 * ---------------------------------------------------------------------- */
typedef struct b2_olop_config {
    int32_t env_kind;
    int32_t n_trees;
    int32_t n_actions;
    int32_t episodes;        /* config["episodes"] (OLOP.allocation, :50-62)  */
    int32_t horizon;         /* config["horizon"]                            */
    int32_t node_capacity;   /* per tree, >= 1 + episodes*horizon*n_actions   */
    int32_t kl;              /* 1: upper_bound.type == "kullback-leibler"; 0: the
                                reference leaves mu_ucb = inf (:153-163)      */
    int32_t continuation;    /* 0 "zeros", 1 "uniform" (:79-82)               */
    double gamma;
    const double* thresholds;/* [episodes] eval(upper_bound.threshold) per episode (:160) */
    const double* init_upper;/* [horizon+2] (1 - gamma**(L+1-d)) / (1 - gamma) (:118-119) */
    b2_finite_mdp mdp;
} b2_olop_config;

typedef struct b2_olop_tree {
    int32_t* parent;
    int32_t* first_child;
    int32_t* count;
    int32_t* meta;           /* action | n_children << 8 | done << 16          */
    double* cumulative;      /* OLOPNode.cumulative_reward                     */
    double* mu_ucb;
    double* upper;           /* value_upper                                    */
} b2_olop_tree;

#define B2_OLOP_RESULT_WORDS 8
/* per tree int32 result: [0] n_nodes [1] plan_len
 * [2] error (1: reward outside [0,1], olop.py:133-134; 2: "zeros" continuation
 *     with action 0 unavailable -- a KeyError in the reference, :82,:88) */

/* OLOP.plan (:94-100); rng as in b2_mcts_plan; plan: int8 [n_trees, horizon]. */
int b2_olop_plan(const b2_olop_config* cfg, const int32_t* root_states, const b2_olop_tree* tree,
                 uint64_t* rng, int8_t* plan, int32_t* result, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2_PLANNER_H */
