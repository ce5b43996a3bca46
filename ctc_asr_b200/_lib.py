"""ctypes binding of libctcasr.so (include/ctcasr.h).  No fallback: if the CUDA library is missing or
does not load, importing callers get a loud error — the product path never computes on the CPU."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctcasr.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
CELL_RNN_TANH, CELL_RNN_RELU, CELL_LSTM, CELL_GRU = 0, 1, 2, 3
COMPUTE_FP32, COMPUTE_TF32, COMPUTE_BF16X3, COMPUTE_BF16 = 0, 1, 2, 3
COMPUTE_ID = {"fp32": COMPUTE_FP32, "tf32": COMPUTE_TF32, "bf16x3": COMPUTE_BF16X3, "bf16": COMPUTE_BF16}

_vp, _i, _f, _u32, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint32, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ctcasr.h one to one (tests check the export list)
SIGNATURES = {
    "ctcasr_abi_version": (_i, []),
    "ctcasr_last_error": (ctypes.c_char_p, []),
    "ctcasr_launch_count": (ctypes.c_uint64, []),
    "ctcasr_ctc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ctcasr_ctc_loss": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _i, _vp, _sz, _vp]),
    "ctcasr_ctc_loss_host": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp]),
    "ctcasr_greedy_decode": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ctcasr_edit_distance": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "ctcasr_dense_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _u32, _i, _vp]),
    "ctcasr_dense_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _u32, _i, _vp]),
    "ctcasr_birnn_reserve_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ctcasr_birnn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ctcasr_birnn_stream_bytes": (ctypes.c_double, [_i, _i, _i, _i, _i, _i]),
    "ctcasr_birnn_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _sz, _vp]),
    "ctcasr_birnn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ctcasr_beam_search_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ctcasr_beam_search": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ctcasr_feature_frames": (_i, [_i, _i]),
    "ctcasr_feature_filterbank_bins": (_i, [_i, _i, _vp]),
    "ctcasr_featurize_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ctcasr_featurize": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "ctcasr_conv2d_out_dims": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ctcasr_conv2d_workspace_bytes": (_sz, [_i] * 8),
    "ctcasr_conv2d_fwd": (_i, [_vp, _i, _vp, _vp, _vp] + [_i] * 10 + [_f, _f, _u32, _i, _vp, _sz, _vp]),
    "ctcasr_conv2d_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp] + [_i] * 10 + [_f, _f, _u32, _i, _vp, _sz, _vp]),
    "ctcasr_dropout": (_i, [_vp, _vp, _sz, _f, _u32, _vp]),
    "ctcasr_profile_enable": (_i, [_i]),
    "ctcasr_profile_collect": (_i, [_vp, _vp, _i]),
    "ctcasr_set_scratch": (_i, [_vp, _sz]),
    "ctcasr_scratch_needed": (_sz, []),
    "ctcasr_scratch_bytes": (_sz, []),
    "ctcasr_create": (_i, [_vp]),
    "ctcasr_use": (_i, [_vp]),
    "ctcasr_destroy": (_i, [_vp]),
    "ctcasr_transpose01": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ctcasr_adam": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _f, _f, _f, _f, _f, _vp]),
    "ctcasr_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


class CtcAsrError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every .cu for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CtcAsrError("libctcasr.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                              "There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != OK:
        msg = load().ctcasr_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise CtcAsrError("%s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
