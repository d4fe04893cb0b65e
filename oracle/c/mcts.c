/* MCTS (rl_agents/agents/tree_search/mcts.py, open loop, random_available policies) in plain C
 * on HighwayLite, consuming a numpy PCG64 stream exactly like Generator.choice does -- TEST
 * INFRASTRUCTURE (oracle/).  Literal restatement of oracle/planners.py::mcts_plan; lets the C3
 * configuration (4096 episodes x horizon 20) be checked bit for bit at full size. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "highway_lite.h"

typedef unsigned __int128 u128;
typedef struct { u128 state, inc; uint32_t has32, uinteger; } pcg64;

static uint64_t pcg_next64(pcg64* g) {
    const u128 mult = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    g->state = g->state * mult + g->inc;
    const uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}
static uint32_t pcg_next32(pcg64* g) {
    if (g->has32) { g->has32 = 0; return g->uinteger; }
    const uint64_t n = pcg_next64(g);
    g->has32 = 1;
    g->uinteger = (uint32_t)(n >> 32);
    return (uint32_t)n;
}
static double pcg_random(pcg64* g) { return (double)(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0); }
static uint32_t pcg_integers(pcg64* g, uint32_t n) {   /* Generator.integers(0, n) */
    const uint32_t rng = n - 1;
    if (rng == 0) return 0;
    const uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)pcg_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) { m = (uint64_t)pcg_next32(g) * rng_excl; leftover = (uint32_t)m; }
    }
    return (uint32_t)(m >> 32);
}

/* rng_words: 6 x uint64 (state hi, lo, inc hi, lo, has_uint32, uinteger), advanced in place.
 * cdf: [(5+1), 5] uniform cdfs from numpy.  Arrays have capacity 1 + episodes*5.  Returns #nodes. */
int mcts_highway_plan(const int32_t* root_words, int episodes, int horizon, double gamma, double temperature,
                      uint64_t* rng_words, const double* cdf, int32_t* parent, int32_t* action, int32_t* count,
                      int32_t* first_child, int32_t* n_children, double* value, double* prior) {
    pcg64 g;
    g.state = ((u128)rng_words[0] << 64) | rng_words[1];
    g.inc = ((u128)rng_words[2] << 64) | rng_words[3];
    g.has32 = (uint32_t)rng_words[4];
    g.uinteger = (uint32_t)rng_words[5];
    int n = 1;
    parent[0] = -1; action[0] = -1; count[0] = 0; first_child[0] = -1; n_children[0] = 0; value[0] = 0.0; prior[0] = 1.0;
    hl_state root, st;
    memcpy(&root, root_words, sizeof(root));
    for (int ep = 0; ep < episodes; ++ep) {
        st = root;
        int node = 0, depth = 0, terminal = 0;
        double total = 0.0;
        while (depth < horizon && n_children[node] > 0 && !terminal) {          /* mcts.py:141-149 */
            const int fc = first_child[node], k = n_children[node];
            double sc[8], best = -INFINITY;
            int ties = 0;
            for (int i = 0; i < k; ++i) {
                sc[i] = value[fc + i] + temperature * (double)k * prior[fc + i] / (double)(count[fc + i] + 1);
                if (sc[i] > best) { best = sc[i]; ties = 1; } else if (sc[i] == best) ++ties;
            }
            int pick = (int)pcg_integers(&g, (uint32_t)ties), sel = 0;
            for (int i = 0; i < k; ++i)
                if (sc[i] == best) { if (pick == 0) sel = i; --pick; }
            int flags;
            const double r = (double)hl_step(&st, action[fc + sel], &flags);
            terminal = flags & 1;
            total += pow(gamma, depth) * r;
            node = fc + sel;
            depth += 1;
        }
        if (n_children[node] == 0 && depth < horizon && (!terminal || node == 0)) {   /* :151-154 */
            int acts[5];
            const int k = hl_available_actions(&st, acts);
            first_child[node] = n;
            n_children[node] = k;
            for (int i = 0; i < k; ++i) {
                parent[n] = node; action[n] = acts[i]; count[n] = 0; first_child[n] = -1; n_children[n] = 0;
                value[n] = 0.0; prior[n] = 1.0 / (double)k;
                ++n;
            }
        }
        if (!terminal) {                                                          /* evaluate :160-177 */
            for (int h = depth; h < horizon; ++h) {
                int acts[5];
                const int k = hl_available_actions(&st, acts);
                const double u = pcg_random(&g);
                const double* c = cdf + (size_t)k * 5;
                int idx = 0;
                for (int i = 0; i < k; ++i) idx += c[i] <= u ? 1 : 0;
                if (idx > k - 1) idx = k - 1;
                int flags;
                const double r = (double)hl_step(&st, acts[idx], &flags);
                total += pow(gamma, h) * r;
                if (flags & 3) break;
            }
        }
        for (int a = node; a >= 0; a = parent[a]) {                               /* update_branch :257-265 */
            count[a] += 1;
            value[a] += 1.0 / (double)count[a] * (total - value[a]);
        }
    }
    rng_words[0] = (uint64_t)(g.state >> 64); rng_words[1] = (uint64_t)g.state;
    rng_words[2] = (uint64_t)(g.inc >> 64); rng_words[3] = (uint64_t)g.inc;
    rng_words[4] = g.has32; rng_words[5] = g.uinteger;
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * SPECIFICATION of the wavefront MCTS (oracle/planners.py::mcts_plan_wavefront, same statement in C).
 * Arrays have capacity 1 + episodes*5 (episode e owns the ids 1 + 5e ..); unused ids keep parent -2.
 * gamma_pow: [horizon+1] gamma**h (host floats).  Returns the number of env steps taken.
 * ------------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static int wave_random(uint64_t seed, int episode, int step, int stream, int n) {
    const uint64_t key = seed + (uint64_t)episode * 0x9E3779B97F4A7C15ULL + (uint64_t)(step + 1) * 0xD1B54A32D192ED03ULL +
                         (uint64_t)stream * 0x8CB92BA72F3D8DD7ULL;
    return (int)((splitmix64(key) >> 33) % (uint64_t)n);
}

int mcts_highway_plan_wave(const int32_t* root_words, int episodes, int horizon, const double* gamma_pow,
                           double temperature, int width, uint64_t seed, int32_t* parent, int32_t* action,
                           int32_t* count, int32_t* first_child, int32_t* n_children, int64_t* vsum, double* value) {
    const int cap = 1 + episodes * 5;
    const double FIX = 1099511627776.0;   /* 2^40 */
    for (int i = 0; i < cap; ++i) { parent[i] = -2; action[i] = -1; count[i] = 0; first_child[i] = -1; n_children[i] = 0; vsum[i] = 0; }
    parent[0] = -1;
    hl_state root, st;
    memcpy(&root, root_words, sizeof(root));
    int32_t* virt = (int32_t*)calloc((size_t)cap, sizeof(int32_t));
    int32_t* expander = (int32_t*)malloc((size_t)cap * sizeof(int32_t));
    int32_t* paths = (int32_t*)malloc((size_t)width * (size_t)(horizon + 1) * sizeof(int32_t));
    int32_t* plen = (int32_t*)malloc((size_t)width * sizeof(int32_t));
    int32_t* reached = (int32_t*)malloc((size_t)width * sizeof(int32_t));
    double* totals = (double*)malloc((size_t)width * sizeof(double));
    long env_steps = 0;
    for (int w0 = 0; w0 < episodes; w0 += width) {
        const int nw = episodes - w0 < width ? episodes - w0 : width;
        for (int i = 0; i < cap; ++i) expander[i] = -1;
        for (int j = 0; j < nw; ++j) {                       /* selections, episode order, virtual counts */
            const int e = w0 + j;
            int node = 0, depth = 0;
            int32_t* path = paths + (size_t)j * (horizon + 1);
            while (depth < horizon && n_children[node] > 0) {
                const int fc = first_child[node], k = n_children[node];
                const double prior = 1.0 / (double)k;
                double sc[8], best = -INFINITY;
                int ties = 0;
                for (int i = 0; i < k; ++i) {
                    const int c = fc + i;
                    const double v = count[c] > 0 ? ((double)vsum[c] / FIX) / (double)count[c] : 0.0;
                    sc[i] = v + temperature * (double)k * prior * (1.0 / (double)(count[c] + virt[c] + 1));
                    if (sc[i] > best) { best = sc[i]; ties = 1; } else if (sc[i] == best) ++ties;
                }
                int pick = wave_random(seed, e, depth, 0, ties), sel = 0;
                for (int i = 0; i < k; ++i)
                    if (sc[i] == best) { if (pick == 0) sel = i; --pick; }
                const int child = fc + sel;
                virt[child] += 1;
                path[depth] = child;
                node = child;
                depth += 1;
            }
            plen[j] = depth;
            if (depth < horizon && expander[node] < 0) expander[node] = e;
        }
        for (int j = 0; j < nw; ++j) {                       /* simulations */
            const int e = w0 + j;
            const int32_t* path = paths + (size_t)j * (horizon + 1);
            st = root;
            double total = 0.0;
            int terminal = 0, rch = 0;
            for (int h = 0; h < plen[j]; ++h) {
                int flags;
                const double r = (double)hl_step(&st, action[path[h]], &flags);
                ++env_steps;
                total += gamma_pow[h] * r;
                rch = h + 1;
                terminal = flags & 1;
                if (terminal) break;
            }
            if (!terminal) {
                const int depth = plen[j];
                const int leaf = depth > 0 ? path[depth - 1] : 0;
                if (depth < horizon && expander[leaf] == e) {
                    int acts[5];
                    const int k = hl_available_actions(&st, acts);
                    const int base = 1 + e * 5;
                    first_child[leaf] = base;
                    n_children[leaf] = k;
                    for (int i = 0; i < k; ++i) { parent[base + i] = leaf; action[base + i] = acts[i]; }
                }
                for (int h = depth; h < horizon; ++h) {
                    int acts[5];
                    const int k = hl_available_actions(&st, acts);
                    int flags;
                    const double r = (double)hl_step(&st, acts[wave_random(seed, e, h, 1, k)], &flags);
                    ++env_steps;
                    total += gamma_pow[h] * r;
                    if (flags & 3) break;
                }
            }
            reached[j] = rch;
            totals[j] = total;
        }
        for (int j = 0; j < nw; ++j) {                       /* backup: exact integer sums; virtual counts cleared */
            const int32_t* path = paths + (size_t)j * (horizon + 1);
            const int64_t fixed = (int64_t)llrint(totals[j] * FIX);
            count[0] += 1;
            vsum[0] += fixed;
            for (int h = 0; h < plen[j]; ++h) {
                virt[path[h]] = 0;
                if (h < reached[j]) { count[path[h]] += 1; vsum[path[h]] += fixed; }
            }
        }
    }
    for (int i = 0; i < cap; ++i) value[i] = count[i] > 0 ? ((double)vsum[i] / FIX) / (double)count[i] : 0.0;
    free(virt); free(expander); free(paths); free(plen); free(reached); free(totals);
    return (int)env_steps;
}
