import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_library():
    """The product path needs libjfx.so.  _build.py is loaded by path (importing the package needs the
    library) and runs before collection, because the test modules import jaxfun_b200 at import time.
    A missing library is always built; a stale one only where there is no GPU (the dev container), so
    the GPU box never spends its time in nvcc on the prebuilt file that travelled with the snapshot."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_jfx_build", os.path.join(ROOT, "jaxfun_b200", "_build.py"))
    _build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_build)
    if not os.path.exists(_build.LIB):
        _build.build_library(force=True)
    elif _build.needs_build():
        import torch
        if not torch.cuda.is_available():
            _build.build_library()


_ensure_library()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
