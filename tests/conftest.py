import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", "logmel_golden.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libwlm.so is git-ignored: build it in-tree when missing/stale and nvcc is present (the
    build container); on the GPU box the prebuilt file travels with the snapshot."""
    import importlib.util
    import shutil

    spec = importlib.util.spec_from_file_location("_wlm_build", os.path.join(ROOT, "whisper_context_biasing_b200", "build.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)          # by path: importing the package would dlopen a stale library first
    if B.needs_build() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        B.build_library()
    yield
