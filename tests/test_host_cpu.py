"""CPU-side checks: the C-ABI library loads and exports every symbol include/wlm.h declares, the
host-side constants are bit-identical to the reference's, and the library refuses to run without
a B200 (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "wlm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wlm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from whisper_context_biasing_b200 import _native as N

    declared = _header_symbols()
    assert declared, "no symbols parsed from include/wlm.h"
    assert sorted(N.SYMBOLS) == declared
    raw = C.CDLL(N.lib_path())
    for name in declared:
        assert getattr(raw, name) is not None
    assert N.LIB.wlm_version() == 10100


def test_mel_table_bit_identical_to_reference(golden):
    from whisper_context_biasing_b200 import slaney_mel_filters

    z, _ = golden
    for m in (80, 128):
        assert np.array_equal(slaney_mel_filters(m), z[f"mel_filters_m{m}"])


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from whisper_context_biasing_b200 import B200WhisperFeatureExtractor, _native as N, slaney_mel_filters

    with pytest.raises(RuntimeError):
        B200WhisperFeatureExtractor()
    table = np.ascontiguousarray(slaney_mel_filters(80).astype(np.float32))
    h = C.c_void_p()
    rc = N.LIB.wlm_plan_create(0, 80, table.ctypes.data, C.byref(h))
    assert rc == N.WLM_ERR_NO_DEVICE and h.value is None
    assert "no CPU fallback" in N.last_error()


def test_abi_argument_validation_without_gpu():
    from whisper_context_biasing_b200 import _native as N

    h = C.c_void_p()
    assert N.LIB.wlm_plan_create(0, 80, None, C.byref(h)) == N.WLM_ERR_BAD_ARG
    assert N.LIB.wlm_plan_create(0, 80, None, None) == N.WLM_ERR_BAD_ARG
    table = np.zeros((201, 300), np.float32)
    assert N.LIB.wlm_plan_create(0, 300, table.ctypes.data, C.byref(h)) == N.WLM_ERR_UNSUPPORTED
    assert N.LIB.wlm_logmel(None, None, 0, None, None, 0, 1, None, None, None, 0, None) == N.WLM_ERR_BAD_ARG
    assert N.LIB.wlm_plan_destroy(None) == N.WLM_OK
    assert N.LIB.wlm_plan_n_mels(None) == N.WLM_ERR_BAD_ARG
    assert N.LIB.wlm_workspace_bytes(None, 4) == 0
    with pytest.raises(ValueError):
        N.check(N.WLM_ERR_BAD_ARG)
    with pytest.raises(N.WlmError):
        N.check(N.WLM_ERR_CUDA)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "whisper_context_biasing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_committed_profiles_belong_to_the_shipped_kernel_sources():
    """bench.py prints `roofline.traffic` only when profiles/<tag>_summary.json was captured from a build of exactly the
    kernel sources in the tree (tools/ncu_summary.py records their sha).  The newest C2 and C3 summaries must match, or
    the bench line would silently lose its ncu evidence."""
    import importlib.util
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_wlm_build_t", os.path.join(root, "whisper_context_biasing_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sha = mod.kernel_sources_sha()
    pdir = os.path.join(root, "profiles")
    for suffix in ("_c2_80mel_summary.json", "_c3_128mel_summary.json"):
        cands = sorted(f for f in os.listdir(pdir) if f.endswith(suffix))
        assert cands, suffix
        doc = json.load(open(os.path.join(pdir, cands[-1])))
        assert doc.get("kernel_sources_sha") == sha, (cands[-1], doc.get("kernel_sources_sha"), sha)
        assert doc["kernels"] if "kernels" in doc else True
