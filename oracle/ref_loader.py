"""Import the UNMODIFIED reference (eleurent/rl-agents) from /root/reference.

Test infrastructure (build container only: /root/reference does not exist on
the GPU box).  Uses the stand-in `gymnasium` / `matplotlib` packages under
oracle/shims when the real ones are absent, and restores `np.infty`
(removed in numpy 2; used at rl_agents/agents/tree_search/olop.py:112 and
rl_agents/utils.py:97).  Recipe: SURVEY.md Appendix E.
"""
import importlib.util
import os
import sys

import numpy as np

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rl_agents"))


def load_reference():
    """Make `import rl_agents` resolve to the reference; returns the package."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "infty"):
        np.infty = np.inf
    for name in ("gymnasium", "matplotlib"):
        if importlib.util.find_spec(name) is None and _SHIMS not in sys.path:
            sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import rl_agents  # noqa: F401
    return rl_agents


class LegacyGenerator(np.random.Generator):
    """`np_random.randint` for olop.py:73 (legacy gym RandomState API)."""

    def randint(self, low, high=None, size=None):
        return self.integers(low, high, size=size)


def legacy_np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    return LegacyGenerator(np.random.PCG64(seed_seq)), seed_seq.entropy
