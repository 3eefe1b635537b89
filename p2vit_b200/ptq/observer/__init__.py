from .build import build_observer, str2observer  # noqa: F401
