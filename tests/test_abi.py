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
    # cfg5: B=512, T=1700, L=84, V=29 -> the warp kernel's spilled rows: [B][T][32 lanes x 6 states] (m, e) pairs, one table
    # shared by the alpha half and the beta half of every utterance
    ws = lib.ctcasr_ctc_workspace_bytes(1700, 512, 29, 84)
    assert ws == 512 * 1700 * 192 * 8
    # long labels (2L+1 > 384) take the block kernel: checkpoint rows every 8 frames, never the [T,S] table
    ws = lib.ctcasr_ctc_workspace_bytes(1700, 64, 29, 422)
    assert 0 < ws < 64 * 1700 * 845 * 8 / 4
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


def test_host_side_queries_of_the_entry_points_added_for_the_next_rows(lib):
    """conv / beam-search / feature entry points: everything that does not need a device."""
    to, fo = ctypes.c_int(), ctypes.c_int()
    # the reference's three conv layers on a 10 s clip (asr/util/tf_contrib.py:66-67): 999x80 -> 500x40 -> 500x20 -> 500x10
    for (T, F, kt, kf, st, sf), want in [((999, 80, 11, 41, 2, 2), (500, 40)), ((500, 40, 11, 21, 1, 2), (500, 20)),
                                         ((500, 20, 11, 21, 1, 2), (500, 10))]:
        assert lib.ctcasr_conv2d_out_dims(T, F, kt, kf, st, sf, ctypes.byref(to), ctypes.byref(fo)) == 0
        assert (to.value, fo.value) == want
    # patch matrix of layer 2 at B=32: 500*32*20 positions x roundup8(11*21*32) floats
    assert lib.ctcasr_conv2d_workspace_bytes(500, 32, 40, 32, 11, 21, 1, 2) == 500 * 32 * 20 * 7392 * 4
    assert lib.ctcasr_feature_frames(160000, 16000) == 999 and lib.ctcasr_feature_frames(401, 16000) == 2
    bins = (ctypes.c_int32 * 82)()
    assert lib.ctcasr_feature_filterbank_bins(16000, 80, bins) == 0
    assert bins[0] == 4 and bins[81] == 512 and all(bins[i] <= bins[i + 1] for i in range(81))
    assert lib.ctcasr_beam_search_workspace_bytes(1000, 32, 29, 1024) > 32 * 1000 * 1024 * 8
    assert lib.ctcasr_beam_search_workspace_bytes(1000, 32, 64, 1024) == 0         # V > 32 unsupported
    assert lib.ctcasr_beam_search_workspace_bytes(1000, 32, 29, 2048) == 0         # beam_width > 1024 unsupported


def test_argument_validation_of_the_new_entry_points_needs_no_gpu(lib):
    one = ctypes.c_void_p(16)            # any non-null pointer: validation must fail before it is touched
    rc = lib.ctcasr_beam_search(one, 10, 1, 29, 5, one, 16, 0, one, one, None, one, 1 << 20, None)
    assert rc == -1 and b"blank" in lib.ctcasr_last_error()                         # TF: blank = num_classes - 1
    rc = lib.ctcasr_beam_search(one, 10, 1, 29, 28, one, 4096, 0, one, one, None, one, 1 << 20, None)
    assert rc == -1 and b"beam_width" in lib.ctcasr_last_error()
    n = (ctypes.c_int32 * 1)(400)
    rc = lib.ctcasr_featurize(one, 1, 400, n, 1, 1, 0, 16000, 80, one, 1, one, one, 1 << 20, None)
    assert rc == -1 and b"to short" in lib.ctcasr_last_error()                      # asr/input_functions.py:213-214
    n[0] = 16000
    rc = lib.ctcasr_featurize(one, 1, 16000, n, 2, 1, 0, 16000, 80, one, 99, one, one, 1 << 20, None)
    assert rc == -1 and b"isn't supported" in lib.ctcasr_last_error()               # :196
    rc = lib.ctcasr_featurize(one, 1, 16000, n, 1, 1, 0, 16000, 80, one, 50, one, one, 1 << 20, None)
    assert rc == -1 and b"99" in lib.ctcasr_last_error()                            # output too short for 99 frames
    rc = lib.ctcasr_conv2d_fwd(one, 1, one, one, one, 10, 1, 8, 2, 3, 3, 1, 1, 64, 1, 20.0, 0.0, 0, 0, None, 0, None)
    assert rc == -1                                                                 # pitch 1 < 2 channels
    rc = lib.ctcasr_conv2d_fwd(one, 2, one, one, one, 10, 1, 8, 2, 3, 3, 1, 1, 64, 1, 20.0, 0.0, 0, 0, None, 0, None)
    assert rc == -3 and b"workspace" in lib.ctcasr_last_error()
    rc = lib.ctcasr_conv2d_fwd(one, 2, one, one, one, 10, 1, 8, 2, 3, 3, 1, 1, 64, 1, 20.0, 1.0, 0, 0, None, 0, None)
    assert rc == -1 and b"drop_rate" in lib.ctcasr_last_error()                     # keep probability 0
    assert lib.ctcasr_dropout(one, one, 4, 1.5, 0, None) == -1


def test_contexts_keep_their_own_scratch_arena(lib):
    """ctcasr_handle_t: the scratch arena belongs to a context; a thread sees the context it bound with
    ctcasr_use(), other threads the process-default one."""
    import threading
    h1, h2 = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.ctcasr_create(ctypes.byref(h1)) == 0 and lib.ctcasr_create(ctypes.byref(h2)) == 0
    fake = ctypes.c_void_p(1 << 20)                         # 1024-B aligned "device" address; never dereferenced here
    assert lib.ctcasr_set_scratch(None, 0) == 0             # default context: no arena
    assert lib.ctcasr_use(h1) == 0 and lib.ctcasr_set_scratch(fake, 4096) == 0
    assert lib.ctcasr_use(h2) == 0 and lib.ctcasr_set_scratch(fake, 8192) == 0
    assert lib.ctcasr_scratch_bytes() == 8192
    assert lib.ctcasr_use(h1) == 0 and lib.ctcasr_scratch_bytes() == 4096
    seen = []
    t = threading.Thread(target=lambda: seen.append(lib.ctcasr_scratch_bytes()))    # another thread: default context
    t.start(); t.join()
    assert seen == [0]
    assert lib.ctcasr_use(None) == 0 and lib.ctcasr_scratch_bytes() == 0
    assert lib.ctcasr_set_scratch(ctypes.c_void_p(12), 64) == -1                    # misaligned
    assert lib.ctcasr_destroy(h1) == 0 and lib.ctcasr_destroy(h2) == 0 and lib.ctcasr_destroy(None) == 0
