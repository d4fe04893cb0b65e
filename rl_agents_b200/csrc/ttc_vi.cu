// ValueIterationAgent on HighwayLite scenes, batched (b2_highway_ttc_vi): one warp per scene builds the scene's
// time-to-collision grid MDP (docs/HIGHWAY_LITE_SPEC.md section 9 -- what `env.unwrapped.to_finite_mdp()` hands the
// reference's agent, value_iteration.py:17,32) and runs the agent's fixed-point iteration on it
// (value_iteration.py:42-73: Q' = R + gamma * V[T], rows of terminal source states zeroed, np.allclose early exit that
// returns the PREVIOUS iterate), all in shared memory.  The reference converts and re-solves on the host at every
// act(); here thousands of scenes are one launch.  fp64 throughout, every operation a single IEEE operation in the
// reference's order (the file is compiled with -fmad=false), so Q is bit-identical with the reference agent's.
#include "common.cuh"
#include "highway_lite.cuh"

namespace b2 {

constexpr int TTC_H = 3, TTC_L = hw::N_LANES, TTC_T = 10;      // speeds {20, 25, 30} m/s x 4 lanes x 10 s
constexpr int TTC_S = TTC_H * TTC_L * TTC_T, TTC_A = B2_HW_ACTIONS;
constexpr int TTC_WARPS = 4;
static_assert(TTC_S == B2_TTC_STATES, "header constant");

struct TtcWarp {
    double q[2][TTC_S * TTC_A];
    double v[TTC_S];
    int grid[TTC_S];          // cost x 2: 0, 1 (= 0.5), 2 (= 1)
    int violations;
};

__device__ __forceinline__ int ttc_cell(int h, int i, int j) {
    h = min(max(h, 0), TTC_H - 1);
    i = min(max(i, 0), TTC_L - 1);
    j = min(max(j, 0), TTC_T - 1);
    return (h * TTC_L + i) * TTC_T + j;
}

__global__ void __launch_bounds__(TTC_WARPS * 32) highway_ttc_vi_kernel(const int32_t* __restrict__ states, int n_envs,
                                                                        double gamma, int iterations, double rtol,
                                                                        double atol, double* __restrict__ q_out,
                                                                        int32_t* __restrict__ action_out,
                                                                        int32_t* __restrict__ state_out,
                                                                        int32_t* __restrict__ sweeps_out) {
    __shared__ TtcWarp sm[TTC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env = blockIdx.x * TTC_WARPS + warp;
    if (env >= n_envs) return;                    // whole warp; nothing below synchronises across warps
    TtcWarp& w = sm[warp];
    const int32_t* st = states + (int64_t)env * hw::WORDS;
    for (int s = lane; s < TTC_S; s += 32) w.grid[s] = 0;
    __syncwarp();

    // ---- compute_ttc_grid: (speed h, vehicle k, collision point p) triples over the lanes ----
    const float x0 = __int_as_float(st[0 * hw::V]), h0 = __int_as_float(st[2 * hw::V]);
    for (int c = lane; c < TTC_H * (hw::V - 1) * 3; c += 32) {
        const int p = c % 3, k = 1 + (c / 3) % (hw::V - 1), h = c / (3 * (hw::V - 1));
        if (!(st[7 * hw::V + k] & 1)) continue;                                   // absent slot
        const double ego_speed = 20.0 + 5.0 * (double)h;
        const float vk = __int_as_float(st[3 * hw::V + k]);
        if (ego_speed == (double)vk) continue;
        const float dh = __int_as_float(st[2 * hw::V + k]) - h0;
        const double projected = (double)vk * (double)hw::cos_p(dh);
        const double diff = ego_speed - projected;
        const double nz = fabs(diff) > 0.01 ? diff : (diff >= 0.0 ? 0.01 : -0.01);
        const double off = p == 0 ? 0.0 : (p == 1 ? -5.0 : 5.0);
        const double distance = ((double)__int_as_float(st[0 * hw::V + k]) - (double)x0) + off;
        const double ttc = distance / nz;
        if (ttc < 0.0 || !(ttc < (double)TTC_T)) continue;       // int(ttc) >= 10: both quantized times fall off the grid
        const int i = hw::lane_of(__int_as_float(st[1 * hw::V + k]));
        const int cost = p == 0 ? 2 : 1;
        const int t_lo = (int)ttc, t_hi = (int)ceil(ttc);
        atomicMax(&w.grid[(h * TTC_L + i) * TTC_T + t_lo], cost);
        if (t_hi < TTC_T) atomicMax(&w.grid[(h * TTC_L + i) * TTC_T + t_hi], cost);
    }
    __syncwarp();

    // ---- fixed_point_iteration on Q (value_iteration.py:42-49,65-73) ----
    for (int e = lane; e < TTC_S * TTC_A; e += 32) w.q[0][e] = 0.0;
    __syncwarp();
    int cur = 0, sweeps = 0;
    for (int it = 0; it < iterations; ++it) {
        const double* q = w.q[cur];
        double* qn = w.q[cur ^ 1];
        for (int s = lane; s < TTC_S; s += 32) {                  // best_action_value: max over actions (numpy .max)
            double m = q[s * TTC_A];
#pragma unroll
            for (int a = 1; a < TTC_A; ++a) m = q[s * TTC_A + a] > m ? q[s * TTC_A + a] : m;
            w.v[s] = m;
        }
        if (lane == 0) w.violations = 0;
        __syncwarp();
        int bad = 0;
        for (int e = lane; e < TTC_S * TTC_A; e += 32) {
            const int s = e / TTC_A, a = e - s * TTC_A;
            const int j = s % TTC_T, i = (s / TTC_T) % TTC_L, h = s / (TTC_T * TTC_L);
            const int g = w.grid[s];
            int nh = h, ni = i;
            if (a == hw::A_LEFT) ni = i - 1;
            else if (a == hw::A_RIGHT) ni = i + 1;
            else if (a == hw::A_FASTER && j == 0) nh = h + 1;
            else if (a == hw::A_SLOWER && j == 0) nh = h - 1;
            const bool terminal = g == 2 || j == TTC_T - 1;
            const double next_v = terminal ? 0.0 : w.v[ttc_cell(nh, ni, j + 1)];
            // state reward: collision * grid + right_lane * lane / 3 + high_speed * speed / 2, then + action reward (0)
            const double sr = ((-1.0 * (0.5 * (double)g)) + (0.1 * ((double)i / 3.0))) + (0.4 * ((double)h / 2.0));
            const double r = sr + 0.0;
            const double nq = r + gamma * next_v;
            qn[e] = nq;
            bad += !(fabs(q[e] - nq) <= atol + rtol * fabs(nq));     // np.allclose(value, next_value), element test
        }
        bad = __reduce_add_sync(0xffffffffu, bad);
        __syncwarp();
        ++sweeps;
        if (bad == 0) break;                   // converged: the PREVIOUS iterate is returned (:71-72)
        cur ^= 1;
    }
    __syncwarp();
    const double* q = w.q[cur];
    const int ego_lane = hw::lane_of(__int_as_float(st[1 * hw::V]));
    const int s0 = ttc_cell(st[8 * hw::V + 1], ego_lane, 0);
    if (q_out)
        for (int e = lane; e < TTC_S * TTC_A; e += 32) q_out[(int64_t)env * TTC_S * TTC_A + e] = q[e];
    if (lane == 0) {
        int best = 0;                          // np.argmax: first maximum
        for (int a = 1; a < TTC_A; ++a)
            if (q[s0 * TTC_A + a] > q[s0 * TTC_A + best]) best = a;
        action_out[env] = best;
        if (state_out) state_out[env] = s0;
        if (sweeps_out) sweeps_out[env] = sweeps;
    }
}

}  // namespace b2

extern "C" int b2_highway_ttc_vi(const int32_t* states, int32_t n_envs, double gamma, int32_t iterations, double rtol,
                                 double atol, double* q_out, int32_t* action_out, int32_t* mdp_state_out,
                                 int32_t* sweeps_out, void* stream) {
    using namespace b2;
    B2_REQUIRE(states && action_out && n_envs > 0, "null pointer / empty batch");
    B2_REQUIRE(iterations >= 0, "iterations < 0");
    const int grid = (n_envs + TTC_WARPS - 1) / TTC_WARPS;
    highway_ttc_vi_kernel<<<grid, TTC_WARPS * 32, 0, (cudaStream_t)stream>>>(states, n_envs, gamma, iterations, rtol, atol,
                                                                           q_out, action_out, mdp_state_out, sweeps_out);
    B2_CUDA_CHECK(cudaGetLastError());
    return B2_OK;
}
