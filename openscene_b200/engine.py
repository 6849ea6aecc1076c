"""Fused inference engine for the MinkUNet family (eval mode) on libosb200.

Takes a network built from the MinkowskiEngine surface (``openscene_b200.minkunet.MinkUNet`` or the
reference's own ``models/mink_unet.py`` classes running on this repository's ``MinkowskiEngine``
package) and executes ``MinkUNetBase.forward`` (models/mink_unet.py:116-174) as one C-ABI call per
convolution:

* BatchNorm (eval) is folded into the producing convolution's epilogue (scale/shift), ReLU and the
  BasicBlock residual add likewise; ``ME.cat`` is never materialised (the next convolution reads two
  sources); activations stay in the split-bf16 layout between layers;
* the 5x5x5 stem fuses its 125-offset hash probe with the 3->32 FMA (no 5^3 kernel map in HBM);
* the final 1x1x1 convolution writes fp32 rows straight into the caller's row order.

Results equal the module-by-module path within the bf16x3 tolerance (tests/test_gpu_engine.py).
"""
import torch

from . import _cabi as C
from . import tc
from .coords import CoordinateManager


def _fold_bn(bn_module):
    bn = bn_module.bn
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).float().contiguous()
    return scale, shift


class _Conv:
    """Packed weights + folded BN of one convolution."""
    __slots__ = ('K', 'cin', 'cout', 'wpack', 'w3', 'scale', 'shift', 'ks', 'stride', 'transpose')

    def __init__(self, conv, bn=None, keep_f32=False):
        w3 = conv.kernel.detach()
        w3 = w3.unsqueeze(0) if w3.dim() == 2 else w3
        self.K, self.cin, self.cout = w3.shape
        self.ks, self.stride = conv.kernel_size, conv.stride
        self.transpose = conv.TRANSPOSE
        self.w3 = w3.float().contiguous() if keep_f32 else None
        self.wpack = tc.pack_weights(w3) if (self.cin % 32 == 0 and self.cout % 32 == 0) else None
        self.scale, self.shift = _fold_bn(bn) if bn is not None else (None, None)


class FusedMinkUNet:
    def __init__(self, model):
        """model: eval-mode MinkUNet (BasicBlock variants) whose parameters live on a CUDA device."""
        net = model.net3d if hasattr(model, 'net3d') else model
        p = next(net.parameters())
        C.require_cuda(p, 'model parameters')
        if net.training:
            raise RuntimeError("FusedMinkUNet folds BatchNorm running statistics: call model.eval() first")
        self.device = p.device
        with torch.cuda.device(self.device), torch.no_grad():
            self.stem = _Conv(net.conv0p1s1, net.bn0, keep_f32=True)
            if self.stem.cin > 3 or self.stem.cout != 32:
                raise NotImplementedError("fused stem supports cin <= 3, cout == 32 (every MinkUNet: INIT_DIM = 32, 3 input features)")
            self.enc, self.dec = [], []
            for i in range(1, 5):
                down = _Conv(getattr(net, f'conv{i}p{2 ** (i - 1)}s2'), getattr(net, f'bn{i}'))
                self.enc.append((down, self._blocks(getattr(net, f'block{i}'))))
            for j in range(4, 8):
                up = _Conv(getattr(net, f'convtr{j}p{2 ** (8 - j)}s2'), getattr(net, f'bntr{j}'))
                self.dec.append((up, self._blocks(getattr(net, f'block{j + 1}'))))
            self.final = _Conv(net.final, None, keep_f32=True)
        self.out_channels = self.final.cout
        self.last_cm = None

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
    @staticmethod
    def _run(cv, srcs, nbr, n_out, res=None, relu=True, out_f32=False, row_map=None):
        (s0, c0), (s1, c1) = srcs[0], (srcs[1] if len(srcs) > 1 else (None, 0))
        assert c0 + c1 == cv.cin
        o_split, o_f32 = tc.conv_tc(s0, c0, s1, c1, nbr, n_out, cv.K, cv.wpack, cv.cout, cv.scale, cv.shift, res, relu,
                                    out_split=not out_f32, out_f32=out_f32, out_row_map=row_map)
        return o_f32 if out_f32 else o_split

    def _stage(self, blocks, srcs, nbr3, n):
        x = srcs
        for (c1, c2, ds) in blocks:
            y = self._run(c1, x, nbr3, n, relu=True)
            if ds is not None:
                r = self._run(ds, x, None, n, relu=False)
            else:
                assert len(x) == 1
                r = x[0][0]
            x = [(self._run(c2, [(y, c1.cout)], nbr3, n, res=r, relu=True), c2.cout)]
        return x[0]

    @torch.no_grad()
    def forward(self, coords, feats, coordinate_manager=None):
        """coords int32 [N,4] (batch,x,y,z), feats fp32 [N,cin], both CUDA, caller order.
        Returns fp32 [N, out_channels] in the caller's row order (== ``model(SparseTensor(feats, coords))``)."""
        C.require_cuda(feats, 'features')
        with torch.cuda.device(self.device):
            cm = coordinate_manager or CoordinateManager(coords)
            self.last_cm = cm
            ts_list = [1]
            for _ in range(4):
                ts_list.append(cm.stride(ts_list[-1], 2))
            n = [cm.sets[t].n for t in ts_list]
            nbr3 = [cm.kernel_map(t, t, 3).nbr for t in ts_list]
            down = [cm.kernel_map(ts_list[l], ts_list[l + 1], 2) for l in range(4)]

            cs0 = cm.sets[1].ensure_hash()
            x_int = torch.empty_like(feats, dtype=torch.float32)
            f32 = feats.float().contiguous()
            C.call('osb_gather_rows_f32', C.ptr(f32), C.ptr(cm.perm), n[0], f32.shape[1], C.ptr(x_int), C.stream_ptr())
            st = self.stem
            x, _ = tc.conv_stem(x_int, cs0.coords, cs0.slots, cs0.cap, st.ks, 1, st.w3, st.scale, st.shift, True, True, False)
            skips = [(x, st.cout)]
            cur = skips[0]
            for l, (dconv, blocks) in enumerate(self.enc):
                y = self._run(dconv, [cur], down[l].nbr, n[l + 1], relu=True)
                cur = self._stage(blocks, [(y, dconv.cout)], nbr3[l + 1], n[l + 1])
                skips.append(cur)
            for j, (uconv, blocks) in enumerate(self.dec):
                l = 3 - j                                   # output level of this transposed conv
                y = self._run(uconv, [cur], down[l].transposed().nbr, n[l], relu=True)
                cur = self._stage(blocks, [(y, uconv.cout), skips[l]], nbr3[l], n[l])
            fin = self.final
            if fin.wpack is not None:
                return self._run(fin, [cur], None, n[0], relu=False, out_f32=True, row_map=cm.perm)
            # odd head width (e.g. 20 classes): generic fp32 kernel, then restore the caller's order
            xf = tc.from_split(cur[0], cur[1])
            out = torch.empty((n[0], fin.cout), dtype=torch.float32, device=self.device)
            C.call('osb_conv_fwd_f32', C.ptr(xf), fin.cin, None, n[0], 1, C.ptr(fin.w3), fin.cin, fin.cout, 0, C.ptr(out),
                   C.stream_ptr())
            ext = torch.empty_like(out)
            C.call('osb_gather_rows_f32', C.ptr(out), C.ptr(cm.inv_perm), n[0], fin.cout, C.ptr(ext), C.stream_ptr())
            return ext

    __call__ = forward

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
