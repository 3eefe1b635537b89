from .build import build_quantizer, str2quantizer  # noqa: F401
