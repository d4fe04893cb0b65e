"""Pure-Python restatement of the numpy PCG64 bit generator and of the two
Generator draws the reference planners consume -- TEST INFRASTRUCTURE.

  random_argmax -> np_random.choice(indices)        (abstract.py:304-311)
      = integers(0, n): Lemire bounded 32-bit draw, nothing for n == 1
  rollout      -> np_random.choice(actions, 1, p)   (mcts.py:172)
      = searchsorted(cumsum(p)/sum, random(), 'right')

Algorithm source: numpy/random/src/pcg64/pcg64.h (PCG XSL-RR 128/64, cheap
multiplier NOT used by PCG64) and numpy/random/src/distributions/
distributions.c (buffered_bounded_lemire_uint32).  numpy is a dependency of
the reference (setup.py:20), not vendored; pinned here against numpy itself
(tests/test_pcg64.py).  The CUDA twin is rl_agents_b200/csrc/pcg64.cuh.
"""
import numpy as np

MASK128 = (1 << 128) - 1
MASK64 = (1 << 64) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


class PCG64(object):
    def __init__(self, state, inc, has_uint32=0, uinteger=0):
        self.state, self.inc = int(state), int(inc)
        self.has_uint32, self.uinteger = int(has_uint32), int(uinteger)

    @classmethod
    def from_numpy(cls, gen):
        st = gen.bit_generator.state
        assert st["bit_generator"] == "PCG64"
        return cls(st["state"]["state"], st["state"]["inc"], st["has_uint32"], st["uinteger"])

    def to_numpy(self, gen):
        st = gen.bit_generator.state
        st["state"]["state"], st["state"]["inc"] = self.state, self.inc
        st["has_uint32"], st["uinteger"] = self.has_uint32, self.uinteger
        gen.bit_generator.state = st

    def words(self):
        """6 x uint64 layout shared with the CUDA engine."""
        return np.array([self.state >> 64, self.state & MASK64, self.inc >> 64, self.inc & MASK64,
                         self.has_uint32, self.uinteger], dtype=np.uint64)

    @classmethod
    def from_words(cls, w):
        w = [int(x) for x in w]
        return cls((w[0] << 64) | w[1], (w[2] << 64) | w[3], w[4], w[5])

    def next64(self):
        self.state = (self.state * PCG_MULT + self.inc) & MASK128
        hi, lo = self.state >> 64, self.state & MASK64
        x = hi ^ lo
        rot = hi >> 58
        return ((x >> rot) | (x << ((-rot) & 63))) & MASK64

    def next32(self):
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        n = self.next64()
        self.has_uint32 = 1
        self.uinteger = n >> 32
        return n & 0xFFFFFFFF

    def random(self):
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    def integers(self, n):
        """Generator.integers(0, n) for 1 <= n <= 2**32."""
        rng = n - 1
        if rng == 0:
            return 0
        if rng == 0xFFFFFFFF:
            return self.next32()
        rng_excl = rng + 1
        m = self.next32() * rng_excl
        leftover = m & 0xFFFFFFFF
        if leftover < rng_excl:
            threshold = (0xFFFFFFFF - rng) % rng_excl
            while leftover < threshold:
                m = self.next32() * rng_excl
                leftover = m & 0xFFFFFFFF
        return m >> 32

    def choice_p(self, cdf):
        """Generator.choice(a, 1, p=p)[0] index, cdf = cumsum(p)/cumsum(p)[-1]."""
        u = self.random()
        return int(np.searchsorted(cdf, u, side="right"))


def uniform_cdf(n):
    p = np.ones(n) / n
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf
