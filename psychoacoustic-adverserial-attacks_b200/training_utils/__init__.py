from . import build, parser, save, sharding, train  # noqa: F401
