"""Thin wrappers over the tensor-core entry points of libosb200 (csrc/conv_tc.cu, conv_stem.cu, split.cu).

"split rows": uint8 tensor [N, 4*C]; every 32-channel block is a 128-byte line [bf16 hi x32 | bf16 lo x32]."""
import ctypes
import weakref

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


def conv_wgrad_tc(x_split, cin, n_in, nbr, n_out, K, gout_split, cout):
    """gw fp32 [K, cin, cout] = sum_o x[nbr[k][o]]^T gout[o] on tensor cores (csrc/conv_wgrad_tc.cu)."""
    dev = x_split.device
    gw = torch.empty((K, cin, cout), dtype=torch.float32, device=dev)
    ws_bytes = C.lib().osb_conv_wgrad_tc_workspace_bytes(n_out, K, cin, cout)
    ws = _workspace(dev, ws_bytes)
    C.call('osb_conv_wgrad_tc', C.ptr(x_split), cin, n_in, C.ptr(nbr), n_out, K, C.ptr(gout_split), cout, C.ptr(gw), C.ptr(ws),
           ws_bytes, C.stream_ptr())
    return gw


_PACK_CACHE = {}          # id(parameter) -> [weakref(parameter), version, address, {(transposed, shape): packed operand}]


def packed_weights_cached(w3, transpose_w=False):
    """``pack_weights`` memoised per PARAMETER OBJECT and version counter: the packed operand is rebuilt only when the
    parameter was written (optimizer step, load_state_dict), not on every forward / backward.  The entry holds a weak
    reference to the parameter: an address + version key alone is not enough -- a freed model's storage is handed to the next
    model with the same address and the same version 0, and a stale operand would be multiplied."""
    base = w3._base if w3._base is not None else w3          # kernel.unsqueeze(0) of a 1x1x1 layer -> the parameter itself
    ent = _PACK_CACHE.get(id(base))
    if ent is None or ent[0]() is not base or ent[1] != base._version or ent[2] != base.data_ptr():
        if len(_PACK_CACHE) > 2048:                             # drop the entries of parameters that no longer exist
            for k in [k for k, e in _PACK_CACHE.items() if e[0]() is None]:
                del _PACK_CACHE[k]
        ent = _PACK_CACHE[id(base)] = [weakref.ref(base), base._version, base.data_ptr(), {}]
    k = (bool(transpose_w), tuple(w3.shape))
    hit = ent[3].get(k)
    if hit is None:
        hit = ent[3][k] = pack_weights(w3, transpose_w)
    return hit


def conv_stem(x, coords, slots, cap, ks, step, w3, scale=None, shift=None, relu=False, out_split=True, out_f32=False):
    n, cin = x.shape
    cout = w3.shape[2]
    os_ = torch.empty((n, 4 * cout), dtype=torch.uint8, device=x.device) if out_split else None
    of_ = torch.empty((n, cout), dtype=torch.float32, device=x.device) if out_f32 else None
    C.call('osb_conv_stem_fused', C.ptr(x.contiguous()), cin, C.ptr(coords), n, C.ptr(slots), cap, ks, step,
           C.ptr(w3.contiguous()), cout, C.ptr(scale), C.ptr(shift), int(relu), C.ptr(os_), C.ptr(of_), C.stream_ptr())
    return os_, of_


def pack_weight_tiles(w3, transpose_w=False):
    """w3 fp32 [K, cin, cout] (or [K, cout, cin] with transpose_w) -> tile-major pre-swizzled B operands of the
    persistent kernel (csrc/conv_chain.cu)."""
    w3 = w3.detach().contiguous().float()
    K = w3.shape[0]
    cin, cout = (w3.shape[2], w3.shape[1]) if transpose_w else (w3.shape[1], w3.shape[2])
    nbytes = C.lib().osb_conv_weight_tiles_bytes(K, cin, cout)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w3.device)
    C.call('osb_conv_pack_weight_tiles', C.ptr(w3), K, cin, cout, int(transpose_w), C.ptr(out), C.stream_ptr())
    return out


class ConvChain:
    """A list of convolution layers executed by osb_conv_chain_launch (one persistent launch per group of layers).

    Descriptors are filled into host memory with raw device addresses (ints) and handed to the kernel as launch
    parameters, group by group: ``begin()``, ``add(...)`` per layer, ``cut()`` between launches, ``run()``."""

    def __init__(self, device, max_layers=160):
        self.device = torch.device(device)
        self.dbytes = C.lib().osb_conv_desc_bytes()
        self.max_layers = max_layers
        self.host = torch.zeros(max_layers * self.dbytes, dtype=torch.uint8)          # copied into the launch parameters
        with torch.cuda.device(self.device):
            self.gbar = torch.zeros(4, dtype=torch.int32, device=self.device)       # {count, generation}: zeroed once
        self.host_a, self.gbar_a = self.host.data_ptr(), self.gbar.data_ptr()
        self._fill = C.lib().osb_conv_desc_fill
        self.begin()

    def begin(self):
        self.n = 0
        self.groups = []
        self._g0 = 0

    def add(self, src0, c0, src1, c1, nbr, n_out, K, wtiles, cout, scale=0, shift=0, res=0, relu=0, out_split=0, out_f32=0,
            row_map=0, cmap=0, cmap_cout=0, ws=0, ws_bytes=0, barrier_before=0):
        """All pointer arguments are raw device addresses (0 = NULL)."""
        if self.n >= self.max_layers:
            raise RuntimeError("ConvChain: too many layers")
        rc = self._fill(self.host_a + self.n * self.dbytes, src0, c0, src1 or None, c1, nbr or None, n_out, K, wtiles, cout,
                        scale or None, shift or None, res or None, int(relu), out_split or None, out_f32 or None,
                        row_map or None, cmap or None, cmap_cout, ws or None, ws_bytes,
                        int(bool(barrier_before) and self.n > self._g0))
        if rc:
            C.check(rc, 'osb_conv_desc_fill')
        self.n += 1

    def cut(self):
        """End the current launch group (the next layer starts a new launch)."""
        if self.n > self._g0:
            self.groups.append((self._g0, self.n - self._g0))
            self._g0 = self.n

    def run(self, flags=0, stream=None):
        self.cut()
        stream = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        fn = C.lib().osb_conv_chain_launch
        for (g0, cnt) in self.groups:
            rc = fn(self.host_a + g0 * self.dbytes, cnt, self.gbar_a, flags, stream)
            if rc:
                C.check(rc, 'osb_conv_chain_launch')


_CHAINS = {}


def conv_chain_single(src0, c0, src1, c1, nbr, n_out, K, wtiles, cout, scale=None, shift=None, res=None, relu=False,
                      out_split=True, out_f32=False, out_row_map=None, cmap=None, cmap_cout=0, n_rows_out=None):
    """One convolution through the persistent kernel (tests / module surface).  Same meaning as ``conv_tc``;
    ``cmap`` selects the dense transposed form (outputs then have ``n_rows_out`` rows of ``cmap_cout`` channels)."""
    dev = src0.device
    rows = n_rows_out if n_rows_out is not None else n_out
    oc = cmap_cout if cmap is not None else cout
    a = lambda t: 0 if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        os_ = torch.empty((rows, 4 * oc), dtype=torch.uint8, device=dev) if out_split else None
        of_ = torch.empty((rows, oc), dtype=torch.float32, device=dev) if out_f32 else None
        ws_bytes = 0 if cmap is not None else C.lib().osb_conv_chain_workspace_bytes(n_out, K, c0 + c1, cout)
        ws = _workspace(dev, ws_bytes) if ws_bytes else None
        ch = _CHAINS.get(dev)
        if ch is None:
            ch = _CHAINS[dev] = ConvChain(dev, 8)
        ch.begin()
        ch.add(a(src0), c0, a(src1), c1, a(nbr), n_out, K, a(wtiles), cout, a(scale), a(shift), a(res), int(relu), a(os_), a(of_),
               a(out_row_map), a(cmap), cmap_cout, a(ws), ws_bytes, 0)
        ch.run()
    return os_, of_


def tuning_set(name, value):
    """Process-wide tuning knob of libosb200 (osb_tuning_set; never changes results)."""
    C.call('osb_tuning_set', name.encode(), int(value))


def debug_set_tc(use_gather4=-1, smem_budget=0, dbg_skip=-1, force_split=-1, target_ctas=0, pf_dist=None, small_nt=None, min_stages=None, lazy=None):
    """Knobs of the first-generation kernel (csrc/conv_tc.cu)."""
    if use_gather4 >= 0:
        tuning_set('tc_a_path', use_gather4)
    if smem_budget > 0:
        tuning_set('tc_smem_budget', smem_budget)
    if dbg_skip >= 0:
        tuning_set('tc_dbg_skip', dbg_skip)
    if force_split >= 0:
        tuning_set('tc_force_split', force_split)
    if target_ctas > 0:
        tuning_set('tc_target_ctas', target_ctas)
    for name, v in (('tc_pf_dist', pf_dist), ('tc_small_nt', small_nt), ('tc_min_stages', min_stages), ('tc_lazy', lazy)):
        if v is not None:
            tuning_set(name, v)


def debug_set_clock(buf):
    """tuning: int64 CUDA tensor [n_tiles, 8] receiving per-CTA clock64 stamps (None disables)."""
    tuning_set('tc_dbg_clock', buf.data_ptr() if buf is not None else 0)
