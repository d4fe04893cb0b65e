from . import Env, Wrapper  # noqa: F401
