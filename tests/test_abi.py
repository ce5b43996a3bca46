"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a
GPU, exports every symbol include/ctcasr.h declares, and rejects bad arguments with the documented
error convention.  No compute call is made here (there is no GPU in this container)."""
import ctypes
import os
import re
import subprocess

import pytest

from ctc_asr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctcasr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctcasr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libctcasr.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared      # the ctypes table binds exactly the header


def test_abi_version_and_workspace_queries(lib):
    assert lib.ctcasr_abi_version() == 1
    # cfg5: B=512, T=1700, L=84, V=29 -> checkpoints only, never the [T,S] alpha table
    ws = lib.ctcasr_ctc_workspace_bytes(1700, 512, 29, 84)
    assert 0 < ws < 512 * 1700 * 169 * 8 / 4        # (hi, lo) checkpoint rows every 8 frames, not the table
    assert lib.ctcasr_ctc_workspace_bytes(10, 1, 500, 4) == 0          # V > 128 is unsupported
    rb = lib.ctcasr_birnn_reserve_bytes(1000, 32, 2048, 2048, _lib.CELL_LSTM)
    assert rb >= 1000 * 32 * (2 * 4 * 2048 + 2 * 2048) * 4


def test_argument_validation_needs_no_gpu(lib):
    rc = lib.ctcasr_ctc_loss(None, 5, 1, 6, 5, None, 1, None, None, None, None, 1.0, None, 1, None, 0, None)
    assert rc == -1 and b"null" in lib.ctcasr_last_error()
    rc = lib.ctcasr_adam(None, None, None, None, 4, 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "adam")


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
