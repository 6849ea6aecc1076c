"""Fused inference engine for the MinkUNet family (eval mode) on libosb200.

Takes a network built from the MinkowskiEngine surface (``openscene_b200.minkunet.MinkUNet`` or the
reference's own ``models/mink_unet.py`` classes running on this repository's ``MinkowskiEngine``
package) and executes ``MinkUNetBase.forward`` (models/mink_unet.py:116-174) as one C-ABI call per
convolution:

* BatchNorm (eval) is folded into the producing convolution's epilogue (scale/shift), ReLU and the
  BasicBlock residual add likewise; ``ME.cat`` is never materialised (the next convolution reads two
  sources); activations stay in the split-bf16 layout between layers;
* the 5x5x5 stem fuses its 125 neighbour lookups (occupancy grid, or the hash where that does not fit) with the 3->32 FMA
  (no 5^3 kernel map in HBM);
* the final 1x1x1 convolution writes fp32 rows straight into the caller's row order.

Results equal the module-by-module path within the bf16x3 tolerance (tests/test_gpu_engine.py).
"""
import os

import torch

from . import _cabi as C
from . import tc
from .coords import CoordinateManager


def _al(x):
    return (x + 255) & ~255


def _fold_bn(bn_module):
    bn = bn_module.bn
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).float().contiguous()
    return scale, shift


class _Conv:
    """Packed weights + folded BN of one convolution."""
    __slots__ = ('K', 'cin', 'cout', 'wpack', 'w3', 'scale', 'shift', 'ks', 'stride', 'transpose', 'wpack_a', 'scale_a', 'shift_a',
                 'wtiles', 'wtiles_a', 'n_ntiles')

    def __init__(self, conv, bn=None, keep_f32=False):
        if getattr(conv, 'bias', None) is not None:
            raise NotImplementedError("FusedMinkUNet: convolutions with bias are not folded (no MinkUNet layer has one: "
                                      "models/mink_unet.py builds every convolution with bias=False)")
        w3 = conv.kernel.detach()
        w3 = w3.unsqueeze(0) if w3.dim() == 2 else w3
        self.K, self.cin, self.cout = w3.shape
        self.ks, self.stride = conv.kernel_size, conv.stride
        self.transpose = conv.TRANSPOSE
        self.w3 = w3.float().contiguous() if keep_f32 else None
        self.wpack = tc.pack_weights(w3) if (self.cin % 32 == 0 and self.cout % 32 == 0) else None
        self.scale, self.shift = _fold_bn(bn) if bn is not None else (None, None)
        # raw device addresses for the low-overhead launch path (tensors above keep the memory alive)
        self.wpack_a = self.wpack.data_ptr() if self.wpack is not None else 0
        self.wtiles = tc.pack_weight_tiles(w3) if self.wpack is not None else None        # persistent-kernel packing
        self.wtiles_a = self.wtiles.data_ptr() if self.wtiles is not None else 0
        self.n_ntiles = max(1, -(-self.cout // 256))
        self.scale_a = self.scale.data_ptr() if self.scale is not None else 0
        self.shift_a = self.shift.data_ptr() if self.shift is not None else 0


class FusedMinkUNet:
    def __init__(self, model):
        """model: eval-mode MinkUNet (BasicBlock variants) whose parameters live on a CUDA device."""
        net = model.net3d if hasattr(model, 'net3d') else model
        p = next(net.parameters())
        C.require_cuda(p, 'model parameters')
        if net.training:
            raise RuntimeError("FusedMinkUNet folds BatchNorm running statistics: call model.eval() first")
        self.device = p.device
        self.dense_up = os.environ.get('OSB_DENSE_UP', '1') != '0'
        # The packed weights and folded BatchNorm constants are COPIES: every tensor they were made from is tracked, and a
        # forward that finds one changed (load_state_dict, an optimiser step, .to()) re-packs before it runs.
        self._net = net
        self._tracked = list(net.parameters()) + list(net.buffers())
        self._build()
        self.out_channels = self.final.cout
        self.last_cm = None
        self._ws = None
        self._arena = None
        self.use_pdl = os.environ.get('OSB_PDL', '1') != '0'
        # persistent chain kernel (csrc/conv_chain.cu); OSB_CHAIN=0 restores one launch of the first-generation kernel per layer
        self.use_chain = os.environ.get('OSB_CHAIN', '1') != '0'
        self.chain_max_tiles = int(os.environ.get('OSB_CHAIN_MAX_TILES', '-1'))   # layers up to this many (row x N) tiles share a launch
        self._chain = None
        self.layer_log = None                     # profiling: set to [] to record (rows, K, cin, cout, tag) per convolution
        self.use_pyramid = os.environ.get('OSB_PYRAMID', '1') != '0'
        if 'OSB_TC_LAZY' in os.environ:                      # tuning: 0 = smem index prologue, 1 = lazy on >= 2-wave launches, 2 = always
            tc.debug_set_tc(lazy=int(os.environ['OSB_TC_LAZY']))

    def _signature(self):
        # ~20 us for the 373 tensors of MinkUNet34C: in-place updates (optimiser steps, load_state_dict) bump `_version`; a
        # re-allocation (.to(), assign=True) moves the first / last tensor along with all others
        t = self._tracked
        return (sum(x._version for x in t), t[0].data_ptr(), t[-1].data_ptr(), len(t))

    def refresh(self):
        """Re-pack the weights and re-fold BatchNorm from the source module (called automatically when a tracked tensor changed)."""
        if self._net.training:
            raise RuntimeError("FusedMinkUNet folds BatchNorm running statistics: call model.eval() first")
        self._tracked = list(self._net.parameters()) + list(self._net.buffers())
        self._build()

    def _build(self):
        net = self._net
        with torch.cuda.device(self.device), torch.no_grad():
            self.stem = _Conv(net.conv0p1s1, net.bn0, keep_f32=True)
            if self.stem.cin > 3 or self.stem.cout != 32:
                raise NotImplementedError("fused stem supports cin <= 3, cout == 32 (every MinkUNet: INIT_DIM = 32, 3 input features)")
            self.enc, self.dec = [], []
            for i in range(1, 5):
                down = _Conv(getattr(net, f'conv{i}p{2 ** (i - 1)}s2'), getattr(net, f'bn{i}'))
                self.enc.append((down, self._blocks(getattr(net, f'block{i}'))))
            for j in range(4, 8):
                m = getattr(net, f'convtr{j}p{2 ** (8 - j)}s2')
                up = _Conv(m, getattr(net, f'bntr{j}'))
                if self.dense_up and up.wpack is not None:
                    # dense transposed conv: one [cin, 8*cout] matrix, column block k = W[k]
                    wide = m.kernel.detach().permute(1, 0, 2).reshape(1, up.cin, up.K * up.cout).contiguous()
                    up.wpack = tc.pack_weights(wide)
                    up.wpack_a = up.wpack.data_ptr()
                    up.wtiles = tc.pack_weight_tiles(wide)
                    up.wtiles_a = up.wtiles.data_ptr()
                    up.n_ntiles = max(1, -(-(up.K * up.cout) // 256))
                self.dec.append((up, self._blocks(getattr(net, f'block{j + 1}'))))
            self.final = _Conv(net.final, None, keep_f32=True)
        self._sig = self._signature()

    @staticmethod
    def _blocks(seq):
        out = []
        for b in seq:
            if not hasattr(b, 'conv2') or hasattr(b, 'conv3'):
                raise NotImplementedError("FusedMinkUNet supports BasicBlock networks (all MinkUNet14/18/34 variants)")
            ds = _Conv(b.downsample[0], b.downsample[1]) if b.downsample is not None else None
            out.append((_Conv(b.conv1, b.norm1), _Conv(b.conv2, b.norm2), ds))
        return out

    # ---------------------------------------------------------------------------------------
    # Launch path: activations of one forward live in ONE arena tensor; layers are addressed by raw
    # device pointers and every convolution is a single ctypes call with prebuilt integer arguments
    # (no per-layer torch allocation, pointer boxing or workspace query: the Python dispatch cost per
    # layer has to stay below the ~20 us the coarse-level kernels take).
    def _plan_bytes(self, n):
        """Upper bound of split-row bytes for all activations of one forward (256-byte aligned slices)."""
        total = _al(n[0] * 4 * self.stem.cout)
        for l, (dconv, blocks) in enumerate(self.enc):
            total += _al(n[l + 1] * 4 * dconv.cout)
            for (c1, c2, ds) in blocks:
                total += _al(n[l + 1] * 4 * c1.cout) + _al(n[l + 1] * 4 * c2.cout) + (_al(n[l + 1] * 4 * ds.cout) if ds else 0)
        for j, (uconv, blocks) in enumerate(self.dec):
            l = 3 - j
            total += _al(n[l] * 4 * uconv.cout)
            for (c1, c2, ds) in blocks:
                total += _al(n[l] * 4 * c1.cout) + _al(n[l] * 4 * c2.cout) + (_al(n[l] * 4 * ds.cout) if ds else 0)
        return total

    def _chain_add(self, cv, s0, c0, s1, c1, nbr_a, n_rows, K, cout, res_a, relu, out_a, out_f32_a, row_map_a, cmap_a, cmap_cout,
                   independent=False):
        """Record one layer of the persistent chain.  Layers whose (row tile x N tile) count is at most `chain_max_tiles`
        share a launch with their neighbours (grid barrier between dependent layers); larger layers get their own launch.
        `independent`: the layer reads nothing the previous layer of the chain wrote (BasicBlock downsample next to conv1)."""
        ch = self._chain
        if self.layer_log is not None:
            self.layer_log.append((n_rows, K, c0 + c1, cout, 'dense-up' if cmap_a else ('res' if res_a else '')))
        tiles = -(-n_rows // 128) * max(1, -(-cout // 256))
        small = tiles <= self._chain_small
        if not (small and self._chain_prev_small):
            ch.cut()
        self._chain_prev_small = small
        ws_bytes = 0 if cmap_a else self._ws_query(n_rows, K, c0 + c1, cout)
        ws_a = 0
        if ws_bytes:
            self._ws_flip ^= 1                                   # consecutive split layers never share scratch
            ws_a = self._ws_a + self._ws_flip * (self._ws_bytes // 2)
            if ws_bytes > self._ws_bytes // 2:
                raise RuntimeError(f"FusedMinkUNet: split workspace of {ws_bytes} bytes exceeds the {self._ws_bytes // 2} provided")
        ch.add(s0, c0, s1, c1, nbr_a, n_rows, K, cv.wtiles_a, cout, cv.scale_a, cv.shift_a, res_a, relu, out_a, out_f32_a, row_map_a,
               cmap_a, cmap_cout, ws_a, ws_bytes, 0 if independent else 1)

    def _conv(self, cv, srcs, nbr_a, n_out, res_a=0, relu=1, out_f32_a=0, row_map_a=0, independent=False):
        """srcs: [(addr, channels, rows)] (one or two).  Returns the address of the split output (or 0)."""
        (s0, c0, r0) = srcs[0]
        (s1, c1, r1) = srcs[1] if len(srcs) > 1 else (0, 0, 0)
        out_a = 0
        if not out_f32_a:
            out_a = self._cursor
            self._cursor += _al(n_out * 4 * cv.cout)
        if self._chain_on:
            self._chain_add(cv, s0, c0, s1, c1, nbr_a, n_out, cv.K, cv.cout, res_a, relu, out_a, out_f32_a, row_map_a, 0, 0,
                            independent=independent)
            return out_a
        if self.layer_log is not None:
            self.layer_log.append((n_out, cv.K, c0 + c1, cv.cout, 'res' if res_a else ''))
        rc = self._fn(s0, c0, r0, s1, c1, r1, nbr_a, n_out, cv.K, cv.wpack_a, cv.cout, cv.scale_a, cv.shift_a, res_a, relu,
                      out_a, out_f32_a, row_map_a, self._ws_a, self._ws_bytes, self._flags, self._stream)
        if rc:
            C.check(rc, 'osb_conv_fwd_tc')
        return out_a

    def _chain_run(self):
        if self._chain_on:
            self._chain.run(self._flags, self._stream)

    def _stage(self, blocks, srcs, nbr3_a, n):
        x = srcs
        for (c1, c2, ds) in blocks:
            y = self._conv(c1, x, nbr3_a, n)
            if ds is not None:
                r = self._conv(ds, x, 0, n, relu=0, independent=True)
            else:
                r = x[0][0]
            x = [(self._conv(c2, [(y, c1.cout, n)], nbr3_a, n, res_a=r), c2.cout, n)]
        return x[0]

    @torch.no_grad()
    def forward(self, coords, feats, coordinate_manager=None, head=None):
        """coords int32 [N,4] (batch,x,y,z), feats fp32 [N,cin], both CUDA, caller order.
        Returns fp32 [N, out_channels] in the caller's row order (== ``model(SparseTensor(feats, coords))``)."""
        C.require_cuda(feats, 'features')
        if self._sig != self._signature():                     # the source module changed since the weights were packed
            self.refresh()
        with torch.cuda.device(self.device):
            cm = coordinate_manager or CoordinateManager(coords, pyramid_levels=4 if self.use_pyramid else 0)
            self.last_cm = cm
            ts_list = [1]
            for _ in range(4):
                ts_list.append(cm.stride(ts_list[-1], 2))
            n = [cm.sets[t].n for t in ts_list]
            nbr3 = [cm.kernel_map(t, t, 3).nbr for t in ts_list]
            down = [cm.kernel_map(ts_list[l], ts_list[l + 1], 2) for l in range(4)]
            up_nbr = [d.transposed().nbr for d in down] if not self.dense_up else None
            nbr3_a = [t.data_ptr() for t in nbr3]

            # Grow-only activation arena reused by every forward (activations never outlive one; the result is a separate
            # tensor).  Allocating ~1 GB per call made the caching allocator fragment against the 600 MB outputs and fall
            # back to cudaMalloc inside steps (10-120 ms stalls, 78 of them in 200 steps).
            need = self._plan_bytes(n) + 256
            if self._arena is None or self._arena.numel() < need:
                self._arena = None
                self._arena = torch.empty(int(need * 1.25), dtype=torch.uint8, device=self.device)
            self._cursor = _al(self._arena.data_ptr())
            if self._ws is None:
                self._ws = torch.empty(192 << 20, dtype=torch.uint8, device=self.device)      # two halves: consecutive split layers alternate
            self._ws_a, self._ws_bytes = self._ws.data_ptr(), self._ws.numel()
            self._stream = torch.cuda.current_stream().cuda_stream
            self._fn = C.lib().osb_conv_fwd_tc
            self._chain_on = self.use_chain
            if self._chain_on:
                if self._chain is None:
                    self._chain = tc.ConvChain(self.device, 160)
                    self._ws_query = C.lib().osb_conv_chain_workspace_bytes
                self._chain.begin()
                grid = C.lib().osb_conv_chain_grid()
                self._chain_small = self.chain_max_tiles if self.chain_max_tiles >= 0 else 2 * grid
                self._chain_prev_small = False
                self._ws_flip = 0
            # PDL: every kernel map / packed weight / BN constant is complete before the chain starts (maps are built
            # above, the stem kernel sits between them and the first convolution)
            self._flags = 1 if self.use_pdl else 0

            cs0 = cm.sets[1].ensure_lookup()
            f32 = feats.float().contiguous()
            x_int = torch.empty_like(f32)
            C.call('osb_gather_rows_f32', C.ptr(f32), C.ptr(cm.perm), n[0], f32.shape[1], C.ptr(x_int), C.stream_ptr())
            st = self.stem
            x_a = self._cursor
            self._cursor += _al(n[0] * 4 * st.cout)
            if cs0.grid is not None:
                C.call('osb_conv_stem_fused_grid', C.ptr(x_int), st.cin, C.ptr(cs0.coords), n[0], C.ptr(cs0.grid), *cs0.grid_args,
                       st.ks, 1, C.ptr(st.w3), st.cout, st.scale_a, st.shift_a, 1, x_a, None, self._stream)
            else:
                C.call('osb_conv_stem_fused', C.ptr(x_int), st.cin, C.ptr(cs0.coords), n[0], C.ptr(cs0.slots), cs0.cap, st.ks, 1,
                       C.ptr(st.w3), st.cout, st.scale_a, st.shift_a, 1, x_a, None, self._stream)
            skips = [(x_a, st.cout, n[0])]
            cur = skips[0]
            for l, (dconv, blocks) in enumerate(self.enc):
                y = self._conv(dconv, [cur], down[l].nbr.data_ptr(), n[l + 1])
                cur = self._stage(blocks, [(y, dconv.cout, n[l + 1])], nbr3_a[l + 1], n[l + 1])
                skips.append(cur)
            for j, (uconv, blocks) in enumerate(self.dec):
                l = 3 - j                                   # output level of this transposed conv
                if self.dense_up and self._chain_on:
                    y = self._cursor
                    self._cursor += _al(n[l] * 4 * uconv.cout)
                    self._chain_add(uconv, cur[0], cur[1], 0, 0, 0, n[l + 1], 1, uconv.K * uconv.cout, 0, 1, y, 0, 0,
                                    down[l].nbr.data_ptr(), uconv.cout)
                elif self.dense_up:
                    y = self._cursor
                    self._cursor += _al(n[l] * 4 * uconv.cout)
                    if self.layer_log is not None:
                        self.layer_log.append((n[l + 1], 1, cur[1], uconv.K * uconv.cout, 'dense-up'))
                    rc = C.lib().osb_convtr_fwd_tc(cur[0], cur[1], n[l + 1], down[l].nbr.data_ptr(), uconv.K, uconv.wpack_a,
                                                   uconv.cout, uconv.scale_a, uconv.shift_a, 1, y, 0, self._flags, self._stream)
                    if rc:
                        C.check(rc, 'osb_convtr_fwd_tc')
                else:
                    y = self._conv(uconv, [cur], up_nbr[l].data_ptr(), n[l])
                cur = self._stage(blocks, [(y, uconv.cout, n[l]), skips[l]], nbr3_a[l], n[l])
            if head is not None:                           # folded head: 96 -> (96 + K) conv, rows straight in caller order
                z = torch.empty((n[0], head.cout), dtype=torch.float32, device=self.device)
                self._conv(head, [cur], 0, n[0], relu=0, out_f32_a=z.data_ptr(), row_map_a=cm.perm.data_ptr())
                self._chain_run()
                return z
            fin = self.final
            out = torch.empty((n[0], fin.cout), dtype=torch.float32, device=self.device)
            if fin.wpack is not None:
                self._conv(fin, [cur], 0, n[0], relu=0, out_f32_a=out.data_ptr(), row_map_a=cm.perm.data_ptr())
                self._chain_run()
                return out
            self._chain_run()
            # odd head width (e.g. 20 classes): generic fp32 kernel, then restore the caller's order
            xf = torch.empty((n[0], cur[1]), dtype=torch.float32, device=self.device)
            C.call('osb_split_to_f32', cur[0], n[0], cur[1], C.ptr(xf), C.stream_ptr())
            C.call('osb_conv_fwd_f32', C.ptr(xf), fin.cin, None, n[0], 1, C.ptr(fin.w3), fin.cin, fin.cout, 0, C.ptr(out),
                   C.stream_ptr())
            ext = torch.empty_like(out)
            C.call('osb_gather_rows_f32', C.ptr(out), C.ptr(cm.inv_perm), n[0], fin.cout, C.ptr(ext), C.stream_ptr())
            return ext

    __call__ = forward

    # ---------------------------------------------------------------------------------------
    def fold_head(self, text_features):
        """Pre-compute the folded head for a set of unit-norm text embeddings [K, C_out]:
        W W^T = L L^T (Cholesky, fp64) and U = W T^T, packed as one 1x1x1 convolution 96 -> (96 + K)."""
        W = self.final.w3[0].double()                                    # [cin, cout]
        T = text_features.to(W.device).double()                          # [K, cout]
        G = W @ W.t()
        G = G + 1e-12 * torch.eye(G.shape[0], device=G.device, dtype=G.dtype) * G.diagonal().mean()
        L = torch.linalg.cholesky(G)                                     # x G x^T = |x L|^2
        U = W @ T.t()
        cin, k = W.shape[0], T.shape[0]
        cout = ((cin + k + 31) // 32) * 32
        w = torch.zeros((1, cin, cout), dtype=torch.float32, device=W.device)
        w[0, :, :cin] = L.float()
        w[0, :, cin:cin + k] = U.float()
        cv = _Conv.__new__(_Conv)
        cv.K, cv.cin, cv.cout, cv.ks, cv.stride, cv.transpose = 1, cin, cout, 1, 1, False
        cv.w3, cv.wpack = w, tc.pack_weights(w)
        cv.wtiles = tc.pack_weight_tiles(w)
        cv.scale = cv.shift = None
        cv.wpack_a, cv.wtiles_a, cv.scale_a, cv.shift_a = cv.wpack.data_ptr(), cv.wtiles.data_ptr(), 0, 0
        cv.n_ntiles = max(1, -(-cout // 256))
        return (cv, cin, k, self._signature())               # the signature lets forward_scores refuse a head folded from older weights

    @torch.no_grad()
    def forward_scores(self, coords, feats, folded, want_scores=True):
        """Cosine scores / labels of every voxel against the folded text set (``fold_head``), equal to
        ``match(normalize(forward(coords, feats)), text)`` up to rounding, without the 768-d features.
        Returns (scores fp16 [N,K] or None, label int64 [N], smax fp32 [N]) in the caller's row order."""
        cv, cin, k = folded[:3]
        if len(folded) > 3 and folded[3] != self._signature():
            raise RuntimeError("forward_scores: the folded head was built from weights that have changed since; call fold_head again")
        z = self.forward(coords, feats, head=cv)
        n = z.shape[0]
        scores = torch.empty((n, k), dtype=torch.float16, device=self.device) if want_scores else None
        label = torch.empty(n, dtype=torch.int64, device=self.device)
        smax = torch.empty(n, dtype=torch.float32, device=self.device)
        C.call('osb_folded_head_finish', C.ptr(z), n, cv.cout, cin, k, C.ptr(scores), C.ptr(label), C.ptr(smax), C.stream_ptr())
        return scores, label, smax

    def conv_census(self, cm):
        """Per-convolution (name, pairs, cin, cout, n_in, n_out) of the last forward: algorithmic flops / bytes
        (SURVEY.md 8d definitions) for bench.py's roofline accounting."""
        ts = [1, 2, 4, 8, 16]
        n = [cm.sets[t].n for t in ts]
        p3 = [cm.kernel_map(t, t, 3).num_pairs() for t in ts]
        rows = []
        k5 = cm.kmaps.get((1, 1, 5, 1))
        rows.append(('stem', k5.num_pairs() if k5 is not None else None, self.stem.cin, self.stem.cout, n[0], n[0], 125))

        def stage(tag, blocks, l, cin_first):
            for bi, (c1, c2, ds) in enumerate(blocks):
                rows.append((f'{tag}.{bi}.conv1', p3[l], c1.cin, c1.cout, n[l], n[l], 27))
                rows.append((f'{tag}.{bi}.conv2', p3[l], c2.cin, c2.cout, n[l], n[l], 27))
                if ds is not None:
                    rows.append((f'{tag}.{bi}.downsample', n[l], ds.cin, ds.cout, n[l], n[l], 1))
        for l, (dconv, blocks) in enumerate(self.enc):
            rows.append((f'down{l + 1}', n[l], dconv.cin, dconv.cout, n[l], n[l + 1], 8))
            stage(f'block{l + 1}', blocks, l + 1, dconv.cout)
        for j, (uconv, blocks) in enumerate(self.dec):
            l = 3 - j
            rows.append((f'up{j + 4}', n[l], uconv.cin, uconv.cout, n[l + 1], n[l], 8))
            stage(f'block{j + 5}', blocks, l, None)
        rows.append(('final', n[0], self.final.cin, self.final.cout, n[0], n[0], 1))
        return rows
