"""Pin oracle/pcg64.py (and through it the CUDA twin) against numpy's own
Generator(PCG64) on interleaved draws of the kinds the planners make."""
import numpy as np

from oracle.pcg64 import PCG64, uniform_cdf


def test_interleaved_draws_match_numpy():
    for seed in (0, 1, 12345):
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        mine = PCG64.from_numpy(gen)
        rng = np.random.default_rng(seed + 7)
        for _ in range(3000):
            kind = rng.integers(3)
            if kind == 0:
                n = int(rng.integers(1, 7))
                idx = np.arange(n) + 3
                assert gen.choice(idx) == idx[mine.integers(n)]
            elif kind == 1:
                n = int(rng.integers(1, 6))
                acts = np.arange(n)
                p = np.ones(n) / n
                assert gen.choice(acts, 1, p=np.array(p))[0] == mine.choice_p(uniform_cdf(n))
            else:
                assert gen.random() == mine.random()
        st = gen.bit_generator.state
        assert (st["state"]["state"], st["has_uint32"], st["uinteger"]) == (mine.state, mine.has_uint32,
                                                                             mine.uinteger)


def test_words_roundtrip_and_large_bounds():
    gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(5)))
    mine = PCG64.from_words(PCG64.from_numpy(gen).words())
    for n in (2 ** 30, 2 ** 32, 3, 2 ** 31 + 11):
        for _ in range(50):
            assert int(gen.integers(0, n)) == mine.integers(n)
    gen2 = np.random.Generator(np.random.PCG64(0))
    mine.to_numpy(gen2)
    assert gen2.random() == gen.random()
