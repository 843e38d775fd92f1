from . import la  # noqa: F401
