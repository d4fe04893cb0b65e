/* Plain-C caller of libb2planner.so through the host-buffer API: no CUDA calls, no Python.
 *
 *   gcc -O2 -Iinclude examples/opd_host_example.c -Lrl_agents_b200/csrc -lb2planner \
 *       -Wl,-rpath,$PWD/rl_agents_b200/csrc -o /tmp/opd_host_example && /tmp/opd_host_example
 *
 * Plans OPD (budget 500, gamma 0.9) on a small deterministic MDP from 4 root states and prints
 * the plans; tests/test_gpu_cabi.py checks the output against the Python engine. */
#include <stdio.h>
#include <stdlib.h>

#include "b2_planner.h"

int main(void) {
    enum { S = 64, A = 4, TREES = 4, BUDGET = 500 };
    static int32_t T[S * A];
    static double R[S * A];
    static uint8_t term[S];
    uint32_t x = 12345u;                       /* tiny LCG: the tables are part of the test vector */
    for (int i = 0; i < S * A; ++i) {
        x = x * 1664525u + 1013904223u;
        T[i] = (int32_t)((x >> 8) % S);
        x = x * 1664525u + 1013904223u;
        R[i] = (double)((x >> 8) % 1000) / 1000.0;
    }
    for (int s = 0; s < S; ++s) term[s] = (uint8_t)(s % 17 == 5);

    b2_opd_host_config cfg = {0};
    cfg.env_kind = B2_ENV_FINITE; cfg.n_trees = TREES; cfg.n_actions = A; cfg.budget = BUDGET;
    cfg.gamma = 0.9; cfg.terminal_reward = 0.0;
    cfg.mdp.n_states = S; cfg.mdp.n_actions = A; cfg.mdp.transition = T; cfg.mdp.reward = R; cfg.mdp.terminal = term;
    b2_opd_handle* h = NULL;
    if (b2_opd_create(&cfg, &h)) { fprintf(stderr, "create: %s\n", b2_last_error()); return 1; }
    const int cap = b2_opd_plan_capacity(h);
    int32_t roots[TREES] = {0, 7, 21, 63};
    int8_t* plan = (int8_t*)malloc((size_t)TREES * cap);
    int32_t result[TREES * B2_OPD_RESULT_WORDS];
    if (b2_opd_plan_host(h, roots, plan, result)) { fprintf(stderr, "plan: %s\n", b2_last_error()); return 1; }
    for (int t = 0; t < TREES; ++t) {
        const int32_t* r = result + t * B2_OPD_RESULT_WORDS;
        double lower = 0, upper = 0;
        int32_t count = 0;
        b2_opd_copy_tree(h, t, 1, NULL, NULL, &count, NULL, NULL, &lower, &upper);
        printf("tree %d nodes %d leaves %d depth %d tie %d root_count %d lower %.17g upper %.17g plan", t, r[0], r[1], r[2],
               r[6], count, lower, upper);
        for (int k = 0; k < r[5]; ++k) printf(" %d", plan[t * cap + k]);
        printf("\n");
    }
    b2_opd_destroy(h);
    free(plan);
    return 0;
}
