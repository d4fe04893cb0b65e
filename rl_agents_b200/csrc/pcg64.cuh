// numpy's PCG64 bit generator and the two Generator draws the reference
// planners consume, bit for bit (CPU twin + derivation: oracle/pcg64.py):
//   random_argmax -> np_random.choice(indices)   = integers(0, n)   (abstract.py:304-311)
//   rollout       -> np_random.choice(a, 1, p=p) = searchsorted(cdf, random(), 'right') (mcts.py:172)
#pragma once
#include <stdint.h>

namespace b2 {

struct Pcg64 {
    unsigned __int128 state, inc;
    uint32_t has_uint32, uinteger;

    __device__ __forceinline__ void load(const uint64_t* w) {
        state = ((unsigned __int128)w[0] << 64) | w[1];
        inc = ((unsigned __int128)w[2] << 64) | w[3];
        has_uint32 = (uint32_t)w[4];
        uinteger = (uint32_t)w[5];
    }
    __device__ __forceinline__ void store(uint64_t* w) const {
        w[0] = (uint64_t)(state >> 64);
        w[1] = (uint64_t)state;
        w[2] = (uint64_t)(inc >> 64);
        w[3] = (uint64_t)inc;
        w[4] = has_uint32;
        w[5] = uinteger;
    }
    __device__ __forceinline__ uint64_t next64() {
        const unsigned __int128 mult =
            ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    __device__ __forceinline__ uint32_t next32() {
        if (has_uint32) {
            has_uint32 = 0;
            return uinteger;
        }
        const uint64_t n = next64();
        has_uint32 = 1;
        uinteger = (uint32_t)(n >> 32);
        return (uint32_t)n;
    }
    __device__ __forceinline__ double random() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
    // Generator.integers(0, n), 1 <= n < 2^32 (Lemire, buffered 32-bit halves)
    __device__ __forceinline__ uint32_t integers(uint32_t n) {
        const uint32_t rng = n - 1;
        if (rng == 0) return 0;
        const uint32_t rng_excl = rng + 1;
        uint64_t m = (uint64_t)next32() * rng_excl;
        uint32_t leftover = (uint32_t)m;
        if (leftover < rng_excl) {
            const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
            while (leftover < threshold) {
                m = (uint64_t)next32() * rng_excl;
                leftover = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

}  // namespace b2
