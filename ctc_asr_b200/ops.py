"""torch-tensor front-ends of the C-ABI (include/ctcasr.h).  torch is only the owner of device
memory and streams here; every computation happens in libctcasr.so's kernels."""
import torch

from . import _lib
from ._lib import check, ptr


def _stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)
    return t


def _i32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous int32 CUDA tensor" % name)
    return t


_ws_cache = {}
_scratch = {}


def _call(fn, what, *args):
    """Run a C-ABI call; when it asks for a bigger split-operand scratch arena, grow it and retry."""
    lib = _lib.load()
    rc = fn(*args)
    for _ in range(3):
        if not (rc == -3 and lib.ctcasr_scratch_needed() > 0 and b"scratch" in lib.ctcasr_last_error()):
            break
        need = int(lib.ctcasr_scratch_needed() * 1.25) + (1 << 20)
        dev = torch.cuda.current_device()
        _scratch[dev] = None
        buf = torch.empty(need + 1024, dtype=torch.uint8, device="cuda")
        _scratch[dev] = buf
        torch.cuda.current_stream().synchronize()
        import ctypes
        base = (buf.data_ptr() + 1023) // 1024 * 1024          # the arena must be 1024-B aligned (swizzle atoms)
        check(lib.ctcasr_set_scratch(ctypes.c_void_p(base), need), "set_scratch")
        rc = fn(*args)
    check(rc, what)


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, tag)."""
    key = (str(device), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def ctc_loss(logits, labels, label_len, seq_len, blank=None, grad=True, grad_scale=1.0, out_grad=None):
    """logits [T,B,V]; labels [B,Lmax] int32 (0-padded); -> (loss[B], grad[T,B,V] or None, status[B]).
    Replaces tf.nn.ctc_loss at asr/model.py:259-264."""
    lib = _lib.load()
    _f32(logits, "logits"); _i32(labels, "labels"); _i32(label_len, "label_len"); _i32(seq_len, "seq_len")
    if logits.dim() != 3 or labels.dim() != 2:
        raise ValueError("logits must be [T,B,V] and labels [B,Lmax]")
    T, B, V = logits.shape
    if labels.shape[0] != B or label_len.numel() != B or seq_len.numel() != B:
        raise ValueError("batch size mismatch between logits, labels and lengths")
    blank = V - 1 if blank is None else blank
    lmax = labels.shape[1]
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    status = torch.empty(B, dtype=torch.int32, device=logits.device)
    g = None
    if grad:
        g = out_grad if out_grad is not None else torch.empty_like(logits)
    wsb = lib.ctcasr_ctc_workspace_bytes(T, B, V, lmax)
    if wsb == 0:
        check(-2, "ctc_workspace_bytes")
    ws = workspace(wsb, logits.device, "ctc")
    check(lib.ctcasr_ctc_loss(ptr(logits), T, B, V, blank, ptr(labels), labels.stride(0), ptr(label_len),
                              ptr(seq_len), ptr(loss), ptr(g), float(grad_scale), ptr(status), lmax,
                              ptr(ws), ws.numel(), _stream()), "ctc_loss")
    return loss, g, status


def greedy_decode(logits, seq_len, blank=None):
    lib = _lib.load()
    _f32(logits, "logits"); _i32(seq_len, "seq_len")
    T, B, V = logits.shape
    blank = V - 1 if blank is None else blank
    ids = torch.empty((B, max(T, 1)), dtype=torch.int32, device=logits.device)
    n = torch.empty(B, dtype=torch.int32, device=logits.device)
    check(lib.ctcasr_greedy_decode(ptr(logits), T, B, V, blank, ptr(seq_len), ptr(ids), ptr(n), _stream()),
          "greedy_decode")
    return ids[:, :T], n


def beam_search(logits, seq_len, beam_width=1024, merge_repeated=False, blank=None):
    """tf.nn.ctc_beam_search_decoder(top_paths=1) (asr/model.py:292-296) -> (ids [B,T] -1 padded, lengths [B],
    log-score of the best path [B])."""
    lib = _lib.load()
    _f32(logits, "logits"); _i32(seq_len, "seq_len")
    T, B, V = logits.shape
    blank = V - 1 if blank is None else blank
    ids = torch.empty((B, max(T, 1)), dtype=torch.int32, device=logits.device)
    n = torch.empty(B, dtype=torch.int32, device=logits.device)
    lp = torch.empty(B, dtype=torch.float32, device=logits.device)
    wsb = lib.ctcasr_beam_search_workspace_bytes(T, B, V, int(beam_width))
    if wsb == 0:
        raise ValueError("beam_search: unsupported shape (num_classes %d <= 32, beam_width %d <= 1024)" % (V, beam_width))
    ws = workspace(wsb, logits.device, "beam")
    check(lib.ctcasr_beam_search(ptr(logits), T, B, V, blank, ptr(seq_len), int(beam_width), int(merge_repeated),
                                 ptr(ids), ptr(n), ptr(lp), ptr(ws), ws.numel(), _stream()), "beam_search")
    return ids[:, :T], n, lp


def edit_distance(hyp, hyp_len, truth, truth_len, normalize=True):
    lib = _lib.load()
    _i32(hyp, "hyp"); _i32(hyp_len, "hyp_len"); _i32(truth, "truth"); _i32(truth_len, "truth_len")
    B = hyp.shape[0]
    out = torch.empty(B, dtype=torch.float32, device=hyp.device)
    check(lib.ctcasr_edit_distance(ptr(hyp), hyp.stride(0), ptr(hyp_len), ptr(truth), truth.stride(0), ptr(truth_len),
                                   B, hyp.shape[1], int(normalize), ptr(out), _stream()), "edit_distance")
    return out


def transpose01(x, out=None):
    lib = _lib.load()
    _f32(x, "x")
    A, B, C = x.shape
    out = torch.empty((B, A, C), dtype=torch.float32, device=x.device) if out is None else out
    check(lib.ctcasr_transpose01(ptr(x), ptr(out), A, B, C, _stream()), "transpose01")
    return out


def dense_fwd(x, w, b, act=1, cutoff=20.0, drop_rate=0.0, seed=0, compute=_lib.COMPUTE_FP32, out=None):
    lib = _lib.load()
    _f32(x, "x"); _f32(w, "w")
    M, K = x.shape
    N = w.shape[1]
    if w.shape[0] != K:
        raise ValueError("dense: x [%d,%d] and w %s do not match" % (M, K, tuple(w.shape)))
    y = torch.empty((M, N), dtype=torch.float32, device=x.device) if out is None else out
    _call(lib.ctcasr_dense_fwd, "dense_fwd", ptr(x), ptr(w), ptr(b), ptr(y), M, K, N, act, cutoff, drop_rate, seed,
          compute, _stream())
    return y


def dense_bwd(x, w, y, dy, dw, db, dx=None, act=1, cutoff=20.0, drop_rate=0.0, seed=0,
              compute=_lib.COMPUTE_FP32):
    """dy is clobbered (becomes dz).  dw/db/dx are written in place."""
    lib = _lib.load()
    M, K = x.shape
    N = w.shape[1]
    _call(lib.ctcasr_dense_bwd, "dense_bwd", ptr(x), ptr(w), ptr(y), ptr(dy), ptr(dx), ptr(dw), ptr(db), M, K, N, act,
          cutoff, drop_rate, seed, compute, _stream())


def conv2d_fwd(x, x_pitch, w, b, y, T, B, F, C, kt, kf, st, sf, act=1, cutoff=20.0, drop_rate=0.0, seed=0,
               compute=_lib.COMPUTE_FP32):
    """x [T,B,F,x_pitch] (C real channels), w [Kp,N] padded HWIO kernel, b [N] -> y [To,B,Fo,N] (pre-allocated).
    One conv layer of asr/util/tf_contrib.py:123-135 (conv2d + relu + minimum + dropout)."""
    lib = _lib.load()
    _f32(x, "x"); _f32(w, "w"); _f32(y, "y")
    N = w.shape[1]
    wsb = lib.ctcasr_conv2d_workspace_bytes(T, B, F, C, kt, kf, st, sf)
    ws = workspace(wsb, x.device, "conv")
    _call(lib.ctcasr_conv2d_fwd, "conv2d_fwd", ptr(x), x_pitch, ptr(w), ptr(b), ptr(y), T, B, F, C, kt, kf, st, sf, N,
          act, cutoff, drop_rate, seed & 0xffffffff, compute, ptr(ws), ws.numel(), _stream())
    return y


def dropout(x, rate, seed, out=None):
    """out = dropout(x) with the library's counter-hash keep-mask over the flat index (out may be x).  The same call on
    a gradient is the backward pass.  RNN input / output / inter-layer dropout (asr/util/tf_contrib.py:190-194,
    asr/model.py:201-206)."""
    lib = _lib.load()
    _f32(x, "x")
    out = torch.empty_like(x) if out is None else out
    _call(lib.ctcasr_dropout, "dropout", ptr(x), ptr(out), x.numel(), rate, seed & 0xffffffff, _stream())
    return out


def conv2d_bwd(x, x_pitch, w, y, dy, dx, dw, db, T, B, F, C, kt, kf, st, sf, act=1, cutoff=20.0, drop_rate=0.0, seed=0,
               compute=_lib.COMPUTE_FP32):
    """dy [To,B,Fo,N] is clobbered (becomes dz); dx [T,B,F,x_pitch] may be None; dw [Kp,N], db [N] overwritten."""
    lib = _lib.load()
    N = w.shape[1]
    wsb = lib.ctcasr_conv2d_workspace_bytes(T, B, F, C, kt, kf, st, sf)
    ws = workspace(wsb, x.device, "conv")
    _call(lib.ctcasr_conv2d_bwd, "conv2d_bwd", ptr(x), x_pitch, ptr(w), ptr(y), ptr(dy), ptr(dx), ptr(dw), ptr(db),
          T, B, F, C, kt, kf, st, sf, N, act, cutoff, drop_rate, seed & 0xffffffff, compute, ptr(ws), ws.numel(), _stream())


def birnn_sizes(T, B, nin, H, cell):
    lib = _lib.load()
    return lib.ctcasr_birnn_reserve_bytes(T, B, nin, H, cell), lib.ctcasr_birnn_workspace_bytes(T, B, nin, H, cell)


def birnn_fwd(x, seq_len, wx, wh, bias, y, reserve, cell, use_len, forget_bias=1.0, compute=_lib.COMPUTE_FP32):
    """x [T,B,in] -> y [T,B,2H] (pre-allocated); reserve: uint8 buffer of birnn_sizes()[0] bytes."""
    lib = _lib.load()
    _f32(x, "x"); _f32(wx, "wx"); _f32(wh, "wh"); _f32(bias, "bias"); _f32(y, "y")
    T, B, nin = x.shape
    H = wh.shape[1]
    _, wsb = birnn_sizes(T, B, nin, H, cell)
    ws = workspace(wsb, x.device, "rnn")
    _call(lib.ctcasr_birnn_fwd, "birnn_fwd", ptr(x), ptr(seq_len), ptr(wx), ptr(wh), ptr(bias), ptr(y), ptr(reserve),
          T, B, nin, H, cell, int(use_len), forget_bias, compute, ptr(ws), ws.numel(), _stream())
    return y


def birnn_bwd(x, seq_len, wx, wh, y, reserve, dy, dx, dwx, dwh, dbias, cell, use_len,
              compute=_lib.COMPUTE_FP32):
    lib = _lib.load()
    T, B, nin = x.shape
    H = wh.shape[1]
    _, wsb = birnn_sizes(T, B, nin, H, cell)
    ws = workspace(wsb, x.device, "rnn")
    _call(lib.ctcasr_birnn_bwd, "birnn_bwd", ptr(x), ptr(seq_len), ptr(wx), ptr(wh), ptr(y), ptr(reserve), ptr(dy),
          ptr(dx), ptr(dwx), ptr(dwh), ptr(dbias), T, B, nin, H, cell, int(use_len), compute, ptr(ws), ws.numel(),
          _stream())


def adam(p, m, v, g, step, lr, beta1, beta2, eps, grad_scale=1.0):
    lib = _lib.load()
    check(lib.ctcasr_adam(ptr(p), ptr(m), ptr(v), ptr(g), p.numel(), int(step), lr, beta1, beta2, eps,
                          grad_scale, _stream()), "adam")


def gemm(a, b, ta=False, tb=False, out=None, accumulate=False, compute=_lib.COMPUTE_FP32):
    lib = _lib.load()
    _f32(a, "a"); _f32(b, "b")
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    c = torch.empty((M, N), dtype=torch.float32, device=a.device) if out is None else out
    _call(lib.ctcasr_gemm, "gemm", ptr(a), ptr(b), ptr(c), M, N, K, int(ta), int(tb), a.stride(0), b.stride(0),
          c.stride(0), int(accumulate), compute, _stream())
    return c
