from . import build, parser, train  # noqa: F401
