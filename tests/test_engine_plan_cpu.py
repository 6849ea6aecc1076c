"""Static check of the launch plan ``FusedMinkUNet.forward`` (openscene_b200/engine.py) builds, without a GPU.

The fused engine turns ``MinkUNetBase.forward`` (models/mink_unet.py:116-174) into 18-20 launches of the persistent kernel:
layers are appended to a chain, cut into launch groups, and inside a group a layer only sees what an earlier layer wrote if a
grid barrier lies between them (``barrier_before``).  A missing barrier is a data race the GPU tests would catch only by luck,
and only for the two architectures they run.  Here the engine's own Python runs unmodified on CPU tensors with the DEVICE
entry points of the library replaced by recorders -- the host-only planner functions (``osb_conv_desc_fill``,
``osb_conv_chain_workspace_bytes``, ``osb_conv_chain_grid``) are the real ones -- and the recorded descriptor lists are checked:

  * every buffer a layer reads (both sources, the residual) was written by an earlier layer or the stem, and when producer and
    consumer share a launch a barrier separates them;
  * no two layers that may run concurrently (same launch, no barrier between them) write the same buffer or share split-K
    scratch; every activation lies inside the arena, no two outputs of a forward overlap;
  * the layer list is the network: one descriptor per convolution of the architecture, channel widths chained correctly,
    the final layer writes fp32 rows through the caller-order row map;
for all ten architectures ``mink_unet()`` accepts (models/mink_unet.py:241-263) and scenes on both sides of the "small layer"
threshold."""
import contextlib
import ctypes
import struct
import types

import pytest
import torch

from openscene_b200 import _cabi as C
from openscene_b200 import engine, minkunet, synth, tc

_FMT = '<12Q q 14i 8i'
_PTRS = ('src0', 'src1', 'nbr', 'wtiles', 'scale', 'shift', 'res', 'out_split', 'out_f32', 'out_row_map', 'cmap', 'partial')
_INTS = ('K', 'nb0', 'nb1', 'cout', 'cout_pad', 'nt', 'n_ntiles', 'relu', 'cmap_cout', 'nsplit', 'm_tiles', 'nsub_max',
         'barrier_before', 'stages_per_split')
HOST_ONLY = {'osb_conv_desc_fill', 'osb_conv_chain_workspace_bytes', 'osb_conv_chain_grid', 'osb_conv_desc_bytes', 'osb_last_error'}


def _decode(raw):
    v = struct.unpack(_FMT, raw)
    d = dict(zip(_PTRS, v[:12]))
    d['n_out'] = v[12]
    d.update(zip(_INTS, v[13:27]))
    return d


class _Set:
    def __init__(self, n):
        self.n = n
        self.coords = torch.zeros(4, dtype=torch.int32)
        self.grid, self.grid_args = torch.zeros(8, dtype=torch.uint8), (0, 6, 1)

    def ensure_lookup(self):
        return self


class _Map:
    def __init__(self, K, n_in, n_out):
        self.nbr, self.K, self.n_in, self.n_out = torch.zeros(4, dtype=torch.int32), K, n_in, n_out
        self._t = None

    def transposed(self):
        if self._t is None:
            self._t = _Map(self.K, self.n_out, self.n_in)
        return self._t


class _FakeCM:
    """what engine.forward asks of a CoordinateManager, with given level sizes"""

    def __init__(self, n_levels):
        self.sets = {1 << l: _Set(n) for l, n in enumerate(n_levels)}
        self.perm = torch.zeros(4, dtype=torch.int32)
        self.inv_perm = torch.zeros(4, dtype=torch.int32)
        self.kmaps = {}

    def stride(self, ts, s):
        return ts * s

    def kernel_map(self, ts_in, ts_out, ks, dilation=1):
        key = (ts_in, ts_out, ks, dilation)
        if key not in self.kmaps:
            self.kmaps[key] = _Map(ks ** 3, self.sets[ts_in].n, self.sets[ts_out].n)
        return self.kmaps[key]


@pytest.fixture
def recorded(monkeypatch):
    """engine + tc running on CPU tensors; device entry points record instead of launching"""
    real = C.lib()
    rec = types.SimpleNamespace(launches=[], calls=[])

    def chain_launch(descs_host, n_layers, gbar, flags, stream):
        raw = ctypes.string_at(descs_host, 192 * n_layers)
        rec.launches.append([_decode(raw[192 * i:192 * (i + 1)]) for i in range(n_layers)])
        return 0

    class Lib:
        def __getattr__(self, name):
            if name in HOST_ONLY:
                return getattr(real, name)
            if name == 'osb_conv_chain_launch':
                return chain_launch
            return lambda *a: (rec.calls.append((name, a)), 0)[1]
    lib = Lib()
    monkeypatch.setattr(C, 'lib', lambda: lib)
    monkeypatch.setattr(C, 'call', lambda name, *a: rec.calls.append((name, a)))
    monkeypatch.setattr(C, 'require_cuda', lambda t, what: None)
    monkeypatch.setattr(C, 'stream_ptr', lambda: None)
    monkeypatch.setattr(tc, 'pack_weights', lambda w3, transpose_w=False: torch.zeros(64, dtype=torch.uint8))
    monkeypatch.setattr(tc, 'pack_weight_tiles', lambda w3, transpose_w=False: torch.zeros(64, dtype=torch.uint8))
    monkeypatch.setattr(torch.cuda, 'device', lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: types.SimpleNamespace(cuda_stream=0))
    return rec


def _expected_convs(arch, out_channels):
    """(cin, cout, K, level) of every tensor-core convolution of the architecture in forward order (stem excluded)."""
    kind, layers, planes = minkunet.ARCHS[arch]
    convs, width, skips = [], minkunet.INIT_DIM, [minkunet.INIT_DIM]

    def stage(plane, n_blocks, lvl, width_in):
        w = width_in
        for b in range(n_blocks):
            convs.append((w, plane, 27, lvl))
            if w != plane:
                convs.append((w, plane, 1, lvl))
            convs.append((plane, plane, 27, lvl))
            w = plane
        return plane
    for i in range(1, 5):
        convs.append((width, width, 8, i))
        width = stage(planes[i - 1], layers[i - 1], i, width)
        skips.append(width)
    for j in range(4, 8):
        lvl = 7 - j
        convs.append((width, planes[j], 8, lvl))          # transposed, dense form: one [cin, 8*cout] product over the coarse rows
        width = stage(planes[j], layers[j], lvl, planes[j] + skips[lvl])
    convs.append((width, out_channels, 1, 0))
    return convs


SCENES = {'bench': [197382, 40640, 9674, 2136, 473],       # levels 0-1 one launch per layer, levels 2-4 chained
          'mid': [30011, 6400, 1500, 350, 90],             # level 1 below the small-layer threshold as well
          'tiny': [1500, 380, 97, 31, 9]}                  # everything small: the whole network in a few launches


@pytest.mark.parametrize('scene', list(SCENES))
@pytest.mark.parametrize('arch', sorted(minkunet.ARCHS))
def test_launch_plan_has_no_unordered_dependency(recorded, arch, scene):
    n = SCENES[scene]
    model = synth.build_model(arch, 768, seed=0).eval()
    eng = engine.FusedMinkUNet(model)
    cm = _FakeCM(n)
    out = eng(torch.zeros(n[0], 4, dtype=torch.int32), torch.ones(n[0], 3), coordinate_manager=cm)
    assert out.shape == (n[0], 768)
    launches = recorded.launches
    flat = [d for g in launches for d in g]
    arena_lo = eng._arena.data_ptr()
    arena_hi = arena_lo + eng._arena.numel()

    # ---- the layer list is the network ----------------------------------------------------------
    want = _expected_convs(arch, 768)
    assert len(flat) == len(want), (len(flat), len(want))
    for d, (cin, cout, K, lvl) in zip(flat, want):
        assert 32 * (d['nb0'] + d['nb1']) == cin, (d, cin)
        if d['cmap']:                                          # dense transposed conv over the COARSE rows
            assert d['K'] == 1 and d['cout'] == 8 * cout and d['cmap_cout'] == cout and d['n_out'] == n[lvl + 1]
        else:
            assert d['K'] == K and d['cout'] == cout and d['n_out'] == n[lvl]
    last = flat[-1]
    assert last['out_f32'] == out.data_ptr() and last['out_row_map'] == cm.perm.data_ptr() and last['out_split'] == 0
    assert last['relu'] == 0 and last['scale'] == 0 and last['res'] == 0
    assert all(d['out_split'] and not d['out_f32'] for d in flat[:-1])

    # ---- producers: the stem output + every split output, each written once, inside the arena, disjoint ----
    stem_calls = [a for (nm, a) in recorded.calls if nm.startswith('osb_conv_stem_fused')]
    assert len(stem_calls) == 1
    stem_out = stem_calls[0][-3]                              # (..., relu, out_split, out_f32, stream)
    stem_out = stem_out if isinstance(stem_out, int) else stem_out.value
    produced = {stem_out: (-1, -1, n[0] * 4 * eng.stem.cout)}                 # address -> (group, index in group, bytes)
    for gi, g in enumerate(launches):
        for li, d in enumerate(g):
            if d['out_split']:
                rows = n[[l for l in range(5) if n[l] == d['n_out']][0] - 1] if d['cmap'] else d['n_out']
                width = d['cmap_cout'] if d['cmap'] else d['cout']
                assert d['out_split'] not in produced, "two layers write the same activation"
                produced[d['out_split']] = (gi, li, rows * 4 * width)
    spans = sorted((a, a + b) for a, (_, _, b) in produced.items())
    assert spans[0][0] >= arena_lo and spans[-1][1] <= arena_hi, "an activation lies outside the arena"
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:])), "two activations of one forward overlap"

    # ---- consumers: ordered behind their producers ------------------------------------------------
    for gi, g in enumerate(launches):
        assert g[0]['barrier_before'] == 0
        for li, d in enumerate(g):
            for name in ('src0', 'src1', 'res'):
                a = d[name]
                if not a:
                    continue
                assert a in produced, f"{arch}/{scene}: group {gi} layer {li} reads {name} that nothing wrote"
                pg, pl, _ = produced[a]
                assert (pg, pl) < (gi, li), "a layer reads what a LATER layer writes"
                if pg == gi:
                    assert any(g[j]['barrier_before'] for j in range(pl + 1, li + 1)), \
                        f"{arch}/{scene}: group {gi}: layer {li} reads {name} written by layer {pl} of the same launch without a barrier"
            # layers that may overlap in time (no barrier since layer j) must not share split-K scratch
            j = li
            while j > 0 and not g[j]['barrier_before']:
                j -= 1
                if d['partial'] and g[j]['partial']:
                    assert d['partial'] != g[j]['partial'], f"{arch}/{scene}: group {gi}: layers {j} and {li} share split scratch"
            if d['nsplit'] > 1:
                assert d['partial'] and d['partial'] >= eng._ws.data_ptr()
                assert d['partial'] + d['nsplit'] * d['n_out'] * d['cout_pad'] * 4 <= eng._ws.data_ptr() + eng._ws.numel()

    # ---- launch grouping: big layers alone, consecutive small ones together -------------------------
    grid = C.lib().osb_conv_chain_grid()
    for g in launches:
        tiles = [d['m_tiles'] * d['n_ntiles'] for d in g]
        if len(g) > 1:
            assert all(t <= 2 * grid for t in tiles), "a layer above the small-layer threshold shares a launch"
    assert len(launches) <= {'bench': 24, 'mid': 16, 'tiny': 8}[scene]


def test_engine_refuses_what_it_cannot_fuse(recorded):
    model = synth.build_model('MinkUNet18A', 768, seed=0)          # train mode: BatchNorm cannot be folded
    with pytest.raises(RuntimeError, match='eval'):
        engine.FusedMinkUNet(model.train())
    model.eval()
    model.block1[0].conv1.bias = torch.nn.Parameter(torch.zeros(1, 32))
    with pytest.raises(NotImplementedError, match='bias'):
        engine.FusedMinkUNet(model)


def test_checker_catches_a_missing_barrier(recorded, monkeypatch):
    """Negative control: the same plan with the residual convolution of every BasicBlock marked 'independent' (no barrier in
    front of the layer that reads conv1's output) must be rejected by the check above."""
    orig = engine.FusedMinkUNet._conv

    def conv(self, cv, srcs, nbr_a, n_out, res_a=0, relu=1, out_f32_a=0, row_map_a=0, independent=False):
        return orig(self, cv, srcs, nbr_a, n_out, res_a=res_a, relu=relu, out_f32_a=out_f32_a, row_map_a=row_map_a,
                    independent=bool(res_a) or independent)
    monkeypatch.setattr(engine.FusedMinkUNet, '_conv', conv)
    with pytest.raises(AssertionError, match='without a barrier'):
        test_launch_plan_has_no_unordered_dependency(recorded, 'MinkUNet34C', 'bench')


def test_folded_head_is_the_same_cosine_score(recorded):
    """``fold_head`` re-associates the last layer with the text matrix (W W^T = L L^T, U = W T^T) so that the [N,768] features
    are never written.  The algebra, on the CPU in fp64: for rows x, (x U_k) / |x L| equals cos(x W, T_k), and the packed
    96 -> 96 + K convolution the engine launches carries exactly [L | U]."""
    model = synth.build_model('MinkUNet34C', 768, seed=0).eval()
    eng = engine.FusedMinkUNet(model)
    text = torch.from_numpy(synth.text_embeddings(20)).float()
    cv, cin, k, sig = eng.fold_head(text)
    assert (cin, k) == (96, 20) and cv.cout % 32 == 0 and cv.cout >= cin + k and sig == eng._signature()
    W = model.final.kernel.detach().double()                         # [96, 768]
    x = torch.randn(500, 96, dtype=torch.float64)
    z = x @ cv.w3[0].double()                                        # what the folded convolution produces per row
    got = z[:, cin:cin + k] / z[:, :cin].norm(dim=1, keepdim=True)
    f = x @ W
    want = (f / f.norm(dim=1, keepdim=True)) @ text.double().t()     # run/evaluate.py:305-310: text rows used as they are
    assert float((got - want).abs().max()) < 1e-6                    # fp32 storage of L and U (scores are rounded to fp16 later)
    assert torch.all(cv.w3[0, :, cin + k:] == 0)                     # padding columns
    # a head folded from older weights is refused (the engine re-packs, the signature moves on)
    with torch.no_grad():
        model.final.kernel.mul_(1.5)
    n = SCENES['tiny']
    with pytest.raises(RuntimeError, match='folded head was built from weights that have changed'):
        eng.forward_scores(torch.zeros(n[0], 4, dtype=torch.int32), torch.ones(n[0], 3), (cv, cin, k, sig))
