"""The PFA 16x25 real-FFT templates (csrc/fft_pfa.cuh) instantiated on the host with V=double,
against numpy.  Proves the index maps, the 13-slot scheme and every butterfly before the same
code is trusted on the device."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("fft") / "libfft_host.so")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           "-I", os.path.join(ROOT, "whisper_context_biasing_b200", "csrc"),
                           os.path.join(ROOT, "tests", "fft_host_harness.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.pfa_rfft400.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pfa_output_bin.restype = C.c_int
    return lib


def test_pfa_rfft400_matches_numpy(harness):
    rng = np.random.default_rng(0)
    for trial in range(8):
        x = rng.standard_normal(400)
        if trial == 0:
            x = np.zeros(400); x[1] = 1.0
        re = np.full(201, np.nan); im = np.full(201, np.nan)
        harness.pfa_rfft400(x.ctypes.data, re.ctypes.data, im.ctypes.data)
        ref = np.fft.rfft(x)
        assert not np.isnan(re).any() and not np.isnan(im).any()      # every bin 0..200 produced
        # constants are float32 literals -> ~1e-7 relative
        assert np.abs(re - ref.real).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        assert np.abs(im - ref.imag).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_every_bin_exactly_once(harness):
    slots = [0, 5, 10, 1, 6, 11, 16, 21, 2, 7, 12, 17, 22]
    seen = {}
    for k2 in slots:
        for k1 in range(16 if k2 else 9):
            seen.setdefault(harness.pfa_output_bin(k1, k2), []).append((k1, k2))
    assert sorted(seen) == list(range(201))
    assert all(len(v) == 1 for v in seen.values())


def test_power_row_map_is_the_inverse_of_the_stage2_layout(harness):
    """Stage 2 writes |X|^2 of (slot, FFT16 output position i) to row 16*slot + i; the mel stage reads FFT bin k from
    row_of_bin(k).  Every bin must map to a row that holds exactly that bin, and position <-> k1 must be a bijection."""
    for f in (harness.pfa_row_of_bin, harness.pfa_bin_of_row, harness.pfa_fft16_pos_of_k1):
        f.restype = C.c_int
    rows = [harness.pfa_row_of_bin(k) for k in range(201)]
    assert all(0 <= r < 13 * 16 for r in rows) and len(set(rows)) == 201
    assert all(harness.pfa_bin_of_row(r) == k for k, r in enumerate(rows))
    assert sorted(harness.pfa_fft16_pos_of_k1(k1) for k1 in range(16)) == list(range(16))
    # every row of every slot holds a bin in range (slot 0 holds seven of its bins twice: k1 and 16 - k1)
    bins = [harness.pfa_bin_of_row(r) for r in range(13 * 16)]
    assert min(bins) == 0 and max(bins) == 200 and len(set(bins)) == 201
