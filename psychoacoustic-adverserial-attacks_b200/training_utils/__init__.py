from . import build, parser, sharding, train  # noqa: F401
