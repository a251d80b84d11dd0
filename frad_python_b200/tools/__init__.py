from . import asfh, ecc  # noqa: F401
