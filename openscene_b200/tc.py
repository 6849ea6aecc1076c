"""Thin wrappers over the tensor-core entry points of libosb200 (csrc/conv_tc.cu, conv_stem.cu, split.cu).

"split rows": uint8 tensor [N, 4*C]; every 32-channel block is a 128-byte line [bf16 hi x32 | bf16 lo x32]."""
import ctypes

import torch

from . import _cabi as C


def to_split(x):
    x = x.contiguous()
    n, c = x.shape
    out = torch.empty((n, 4 * c), dtype=torch.uint8, device=x.device)
    C.call('osb_f32_to_split', C.ptr(x), n, c, C.ptr(out), C.stream_ptr())
    return out


def from_split(xs, c):
    n = xs.shape[0]
    out = torch.empty((n, c), dtype=torch.float32, device=xs.device)
    C.call('osb_split_to_f32', C.ptr(xs), n, c, C.ptr(out), C.stream_ptr())
    return out


def pack_weights(w3, transpose_w=False):
    """w3 fp32 [K, cin, cout] (or [K, cout, cin] with transpose_w) -> packed K-major split-bf16 B operand."""
    w3 = w3.detach().contiguous().float()
    K = w3.shape[0]
    cin, cout = (w3.shape[2], w3.shape[1]) if transpose_w else (w3.shape[1], w3.shape[2])
    nbytes = C.lib().osb_conv_packed_weight_bytes(K, cin, cout)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w3.device)
    C.call('osb_conv_pack_weights', C.ptr(w3), K, cin, cout, int(transpose_w), C.ptr(out), C.stream_ptr())
    return out


_WS = {}


def _workspace(dev, nbytes):
    """Grow-only per-device scratch for split-mode convolutions (stream-ordered reuse)."""
    w = _WS.get(dev)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(nbytes, 32 << 20), dtype=torch.uint8, device=dev)
        _WS[dev] = w
    return w


def conv_tc(src0, c0, src1, c1, nbr, n_out, K, wpack, cout, scale=None, shift=None, res=None, relu=False,
            out_split=True, out_f32=False, out_row_map=None):
    """One fused sparse convolution.  Returns (split rows or None, fp32 rows or None)."""
    dev = src0.device
    os_ = torch.empty((n_out, 4 * cout), dtype=torch.uint8, device=dev) if out_split else None
    of_ = torch.empty((n_out, cout), dtype=torch.float32, device=dev) if out_f32 else None
    ws_bytes = C.lib().osb_conv_tc_workspace_bytes(n_out, K, c0 + c1, cout)
    ws = _workspace(dev, ws_bytes) if ws_bytes else None
    C.call('osb_conv_fwd_tc', C.ptr(src0), c0, src0.shape[0], C.ptr(src1), c1, 0 if src1 is None else src1.shape[0],
           C.ptr(nbr), n_out, K, C.ptr(wpack), cout, C.ptr(scale), C.ptr(shift), C.ptr(res), int(relu),
           C.ptr(os_), C.ptr(of_), C.ptr(out_row_map), C.ptr(ws), ws_bytes, 0, C.stream_ptr())
    return os_, of_


def conv_stem(x, coords, slots, cap, ks, step, w3, scale=None, shift=None, relu=False, out_split=True, out_f32=False):
    n, cin = x.shape
    cout = w3.shape[2]
    os_ = torch.empty((n, 4 * cout), dtype=torch.uint8, device=x.device) if out_split else None
    of_ = torch.empty((n, cout), dtype=torch.float32, device=x.device) if out_f32 else None
    C.call('osb_conv_stem_fused', C.ptr(x.contiguous()), cin, C.ptr(coords), n, C.ptr(slots), cap, ks, step,
           C.ptr(w3.contiguous()), cout, C.ptr(scale), C.ptr(shift), int(relu), C.ptr(os_), C.ptr(of_), C.stream_ptr())
    return os_, of_


def debug_set_tc(use_gather4=-1, smem_budget=0, dbg_skip=-1, force_split=-1, target_ctas=0, pf_dist=None, small_nt=None, min_stages=None, lazy=None):
    fn = C.lib().osb_debug_set_tc
    fn.restype, fn.argtypes = None, [ctypes.c_int, ctypes.c_int]
    fn(use_gather4, smem_budget)
    fn2 = C.lib().osb_debug_set_tc2
    fn2.restype, fn2.argtypes = None, [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    fn2(dbg_skip, force_split, target_ctas)
    if pf_dist is not None:
        fn3 = C.lib().osb_debug_set_tc3
        fn3.restype, fn3.argtypes = None, [ctypes.c_int]
        fn3(pf_dist)
    if small_nt is not None:
        fn4 = C.lib().osb_debug_set_tc4
        fn4.restype, fn4.argtypes = None, [ctypes.c_int, ctypes.c_int]
        fn4(small_nt, 0)
    if min_stages is not None:
        fn5 = C.lib().osb_debug_set_tc5
        fn5.restype, fn5.argtypes = None, [ctypes.c_int]
        fn5(min_stages)
    if lazy is not None:
        fn6 = C.lib().osb_debug_set_tc6
        fn6.restype, fn6.argtypes = None, [ctypes.c_int]
        fn6(lazy)


def debug_set_clock(buf):
    """tuning: int64 CUDA tensor [n_tiles, 8] receiving per-CTA clock64 stamps (None disables)."""
    fn = C.lib().osb_debug_set_clock
    fn.restype, fn.argtypes = None, [ctypes.c_void_p]
    fn(C.ptr(buf))
