from . import fft, linalg, special  # noqa: F401
