// IntersectionLite: one decision step (15 physics sub-steps) of a 16-slot unsignalised four-way intersection,
// executed by a 16-lane group of a warp -- lane = vehicle slot.  Spec: docs/INTERSECTION_LITE_SPEC.md; the CPU
// statement of the same spec is oracle/intersection.py::intersection_step; every arithmetic operation below is
// a single IEEE fp32 operation in the same order (compiled with -fmad=false, IEEE division), so results are
// bit-identical.  Stands for `intersection-v0` of BASELINE config C5 (curved turn lanes, yielding traffic,
// in-step spawning), which is third-party, unpinned and absent from the reference tree.
#pragma once
#include "common.cuh"
#include "highway_lite.cuh"   // sin_p / cos_p / not_zero and the IDM constants shared by both env models

namespace b2 {
namespace il {

constexpr int V = 16;
constexpr int WORDS = B2_HW_STATE_WORDS;       // same 136-word arena slot as HighwayLite
constexpr int N_ACTIONS = 3;
constexpr int A_SLOWER = 0, A_IDLE = 1, A_FASTER = 2;
constexpr int SUBSTEPS = 15;
constexpr int DURATION = 13;
constexpr int N_ROUTES = 12;
constexpr int SPAWN_PERIOD = 23;

#define IL_CONST(name, lit) constexpr float name = lit
IL_CONST(DT, 0x1.111112p-4f);
IL_CONST(KP_A, 0x1.aaaaaap+0f);
IL_CONST(APPROACH, 0x1.4p+5f);                 // 40
IL_CONST(ARC_LEFT, 0x1.921fb6p+3f);            // f32(4) * PI
IL_CONST(ARC_RIGHT, 0x1.921fb6p+2f);           // f32(2) * PI
IL_CONST(LEN_LEFT, 0x1.7243f8p+6f);            // (40 + ARC_LEFT) + 40
IL_CONST(LEN_STRAIGHT, 0x1.7p+6f);             // 92
IL_CONST(LEN_RIGHT, 0x1.5921fcp+6f);
IL_CONST(PRIO_END_LEFT, 0x1.c487eep+5f);       // (40 + ARC_LEFT) + 4
IL_CONST(PRIO_END_STRAIGHT, 0x1.cp+5f);        // 56
IL_CONST(PRIO_END_RIGHT, 0x1.9243f6p+5f);
IL_CONST(LENGTH, 0x1.4p+2f);                   // 5
IL_CONST(HIT_D2, 0x1.88p+3f);                  // 12.25
IL_CONST(ACC_MAX, 0x1.8p+2f);
IL_CONST(OTHER_TS, 0x1.2p+3f);                 // 9
IL_CONST(STOP_LINE, 0x1.3p+5f);                // 38
IL_CONST(YIELD_FROM, 0x1.ep+3f);               // 15
IL_CONST(PRIO_FROM, 0x1.9p+4f);                // 25
IL_CONST(ENTRY_CLEAR, 0x1.cp+3f);              // 14
IL_CONST(SPAWN_SPEED, 0x1.0p+3f);              // 8
IL_CONST(SPEED_STEP, 0x1.2p+2f);               // 4.5
#undef IL_CONST

struct Lane {
    float s, v;
    int route, flags;
};
struct Globals {
    int t, si, spawn_clock, spawn_seq, arrived;
};

template <bool CG>
__device__ __forceinline__ int ldw(const int32_t* p) { return CG ? __ldcg(p) : *p; }

template <bool CG>
__device__ __forceinline__ void load_state(const int32_t* w, int li, Lane& L, Globals& g) {
    L.s = __int_as_float(ldw<CG>(w + 0 * V + li));
    L.v = __int_as_float(ldw<CG>(w + 1 * V + li));
    L.route = ldw<CG>(w + 2 * V + li);
    L.flags = ldw<CG>(w + 3 * V + li);
    g.t = ldw<CG>(w + 128);
    g.si = ldw<CG>(w + 129);
    g.spawn_clock = ldw<CG>(w + 130);
    g.spawn_seq = ldw<CG>(w + 131);
    g.arrived = ldw<CG>(w + 132);
}

__device__ __forceinline__ void store_state(int32_t* w, int li, const Lane& L, const Globals& g) {
    w[0 * V + li] = __float_as_int(L.s);
    w[1 * V + li] = __float_as_int(L.v);
    w[2 * V + li] = L.route;
    w[3 * V + li] = L.flags;
    w[4 * V + li] = 0; w[5 * V + li] = 0; w[6 * V + li] = 0; w[7 * V + li] = 0;
    if (li < 8) {
        const int vals[8] = {g.t, g.si, g.spawn_clock, g.spawn_seq, g.arrived, 0, 0, 0};
        int v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v = li == k ? vals[k] : v;
        w[128 + li] = v;
    }
}

// get_available_actions(): IDLE, FASTER (speed index below the top), SLOWER (above the bottom), in this order
__device__ __forceinline__ int avail_mask(int si) { return (1 << A_IDLE) | (si < 2 ? 1 << A_FASTER : 0) | (si > 0 ? 1 << A_SLOWER : 0); }
__device__ __forceinline__ int nth_action(int mask, int n) {
    const int order[3] = {A_IDLE, A_FASTER, A_SLOWER};
    int k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (mask & (1 << order[i])) {
            if (k == n) return order[i];
            ++k;
        }
    }
    return -1;
}

__device__ __forceinline__ void position(int route, float s, float& wx, float& wy) {
    const int entry = route / 3, turn = route - 3 * entry;
    float x = 2.0f, y = -46.0f + s;
    const float u = s - APPROACH;
    if (turn == 0 && s >= APPROACH) {                   // left: radius 8 about (-6, -6)
        const float th = u * 0.125f;
        const float lx = -6.0f + 8.0f * hw::cos_p(th), ly = -6.0f + 8.0f * hw::sin_p(th);
        const float ul = u - ARC_LEFT;
        const bool past = u >= ARC_LEFT;
        x = past ? -6.0f - ul : lx;
        y = past ? 2.0f : ly;
    } else if (turn == 2 && s >= APPROACH) {            // right: radius 4 about (6, -6)
        const float th = u * 0.25f;
        const float rx = 6.0f - 4.0f * hw::cos_p(th), ry = -6.0f + 4.0f * hw::sin_p(th);
        const float ur = u - ARC_RIGHT;
        const bool past = u >= ARC_RIGHT;
        x = past ? 6.0f + ur : rx;
        y = past ? -2.0f : ry;
    }
    wx = entry == 0 ? x : (entry == 1 ? y : (entry == 2 ? -x : -y));
    wy = entry == 0 ? y : (entry == 1 ? -x : (entry == 2 ? -y : x));
}

__device__ __forceinline__ float idm(float v, bool has_front, float d, float v_f) {
    const float ts = fminf(fmaxf(OTHER_TS, 0.0f), hw::SPEED_LIMIT);
    const float ratio = fmaxf(v, 0.0f) / fabsf(hw::not_zero(ts));
    const float r2 = ratio * ratio;
    const float r4 = r2 * r2;
    const float acc = hw::COMFORT_ACC_MAX * (1.0f - r4);
    const float gap = (hw::D0 + v * hw::TAU) + (v * (v - v_f)) / hw::TWO_SQRT_AB;
    const float q = gap / hw::not_zero(d);
    return has_front ? acc - hw::COMFORT_ACC_MAX * (q * q) : acc;
}

#define IL_SHFL(val, src) __shfl_sync(gmask, (val), (src), V)

// One decision step.  Returns the reward (group-uniform).
__device__ __forceinline__ float step(Lane& L, int li, Globals& g, int action, bool& term, bool& trunc, unsigned gmask) {
    if (action == A_FASTER) g.si = min(g.si + 1, 2);
    else if (action == A_SLOWER) g.si = max(g.si - 1, 0);
    const float ts0 = SPEED_STEP * (float)g.si;
    const float INF = __int_as_float(0x7f800000);
    for (int sub = 0; sub < SUBSTEPS; ++sub) {
        const bool present = (L.flags & 1) != 0, crashed = (L.flags & 2) != 0;
        const int entry = L.route / 3;
        // ---- front vehicle and priority traffic: one pass over the slots ----
        float best_s = INF, best_v = 0.0f;
        bool has_f = false, prio = false;
        for (int j = 0; j < V; ++j) {
            const float sj = IL_SHFL(L.s, j), vj = IL_SHFL(L.v, j);
            const int rj = IL_SHFL(L.route, j), fj = IL_SHFL(L.flags, j);
            const bool pj = (fj & 1) != 0;
            const int ej = rj / 3, tj = rj - 3 * ej;
            const bool same_lane = rj == L.route || (ej == entry && sj < APPROACH);
            if (pj && j != li && same_lane && sj > L.s && sj < best_s) { best_s = sj; best_v = vj; has_f = true; }
            const float pend = tj == 0 ? PRIO_END_LEFT : (tj == 1 ? PRIO_END_STRAIGHT : PRIO_END_RIGHT);
            prio = prio || (pj && ej != entry && rj > L.route && sj > PRIO_FROM && sj < pend);
        }
        float acc = idm(L.v, has_f, best_s - L.s, best_v);
        const bool yields = L.s > YIELD_FROM && L.s < STOP_LINE && prio;
        const float acc_y = idm(L.v, true, STOP_LINE - L.s, 0.0f);
        if (yields) acc = fminf(acc, acc_y);
        acc = fminf(fmaxf(acc, -ACC_MAX), ACC_MAX);
        if (li == 0) acc = KP_A * (ts0 - L.v);
        if (crashed) acc = -L.v;
        // ---- integrate ----
        if (present) {
            const float ns = L.s + L.v * DT;
            const float nv = fmaxf(L.v + acc * DT, 0.0f);
            L.s = ns; L.v = nv;
        }
        // ---- collisions on the new positions ----
        float x, y;
        position(L.route, L.s, x, y);
        bool hit = false;
        for (int j = 0; j < V; ++j) {
            const float sj = IL_SHFL(L.s, j), xj = IL_SHFL(x, j), yj = IL_SHFL(y, j);
            const int rj = IL_SHFL(L.route, j), fj = IL_SHFL(L.flags, j);
            const bool lane = rj == L.route || (rj / 3 == entry && sj < APPROACH && L.s < APPROACH);
            const float dx = xj - x, dy = yj - y;
            const float d2 = dx * dx + dy * dy;
            const bool close = lane ? fabsf(sj - L.s) < LENGTH : d2 < HIT_D2;
            hit = hit || ((fj & 1) != 0 && present && j != li && close);
        }
        if (hit) L.flags |= 2;
        // ---- end of route: others leave, the ego has arrived ----
        const int turn = L.route - 3 * entry;
        const float length = turn == 0 ? LEN_LEFT : (turn == 1 ? LEN_STRAIGHT : LEN_RIGHT);
        const bool gone = present && L.s >= length;
        if (IL_SHFL(gone ? 1 : 0, 0)) g.arrived = 1;
        if (gone && li != 0) L.flags = 0;
        // ---- deterministic spawning ----
        g.spawn_clock += 1;
        if (g.spawn_clock >= SPAWN_PERIOD) {
            g.spawn_clock = 0;
            const unsigned k = (unsigned)g.spawn_seq;
            g.spawn_seq = (int)((k + 1u) & 0x7fffffffu);
            const unsigned h = k * 2654435761u + 40503u;
            const int route = (int)((h >> 16) % (unsigned)N_ROUTES);
            const bool pres = (L.flags & 1) != 0;
            // the spawn clock differs between the two scenes of a warp (leaves of different depths): this
            // branch is uniform per 16-lane group only, so the votes synchronise the group, not the warp
            const unsigned half_shift = threadIdx.x & 16;
            const unsigned hmask = 0xffffu << half_shift;
            const unsigned blocked = (__ballot_sync(hmask, pres && L.route / 3 == route / 3 && L.s < ENTRY_CLEAR) >> half_shift) & 0xffffu;
            const unsigned freem = (__ballot_sync(hmask, !pres && li > 0) >> half_shift) & 0xffffu;
            if (blocked == 0 && freem != 0 && li == __ffs(freem) - 1) {
                L.s = 0.0f; L.v = SPAWN_SPEED; L.route = route; L.flags = 1;
            }
        }
    }
    g.t += 1;
    const bool crashed0 = IL_SHFL(L.flags, 0) & 2;
    const float v0 = IL_SHFL(L.v, 0);
    float rew;
    if (crashed0) rew = 0.0f;
    else if (g.arrived) rew = 1.0f;
    else {
        const float sc = fminf(fmaxf((v0 - 7.0f) * 0.5f, 0.0f), 1.0f);
        rew = (5.0f + sc) / 6.0f;
    }
    term = crashed0 || g.arrived != 0;
    trunc = g.t >= DURATION;
    return rew;
}

}  // namespace il
}  // namespace b2
