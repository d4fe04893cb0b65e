"""Minimal stand-in for the `gymnasium` names the reference planners import.

TEST INFRASTRUCTURE ONLY (oracle/): lets the unmodified reference under
/root/reference be imported in the build container, where gymnasium is absent
(SURVEY.md Appendix E).  Never imported by the product package.
"""


class Env(object):
    metadata = {}

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    @property
    def unwrapped(self):
        return self.env.unwrapped


from . import core, error, utils  # noqa: E402,F401
