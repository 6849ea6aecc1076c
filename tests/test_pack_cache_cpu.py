"""Host logic of the packed-weight cache behind the training path (openscene_b200/tc.py: packed_weights_cached), without a GPU.

``SparseConvFunction`` (me.py) asks for the packed operand of a layer's kernel in forward and of its transpose in backward.
The cache is keyed on the PARAMETER OBJECT (weak reference) + version counter + address; this test pins the PyTorch behaviour
that key relies on -- the tensor an autograd Function receives in forward and gets back from ``ctx.saved_tensors`` in
backward is the parameter itself (or a view whose ``_base`` is the parameter, for the 2-D kernel of a 1x1x1 layer) -- and the
invalidation rules, with ``pack_weights`` (a CUDA call) replaced by a counter."""
import gc

import pytest
import torch

from openscene_b200 import tc


@pytest.fixture
def counted(monkeypatch):
    calls = []

    def fake_pack(w3, transpose_w=False):
        calls.append((tuple(w3.shape), bool(transpose_w)))
        return torch.tensor([len(calls)])
    monkeypatch.setattr(tc, 'pack_weights', fake_pack)
    monkeypatch.setattr(tc, '_PACK_CACHE', {})
    return calls


class _Probe(torch.autograd.Function):
    """same save / unpack pattern as SparseConvFunction"""
    seen = []

    @staticmethod
    def forward(ctx, x, w3):
        _Probe.seen.append(('fwd', w3, tc.packed_weights_cached(w3)))
        ctx.save_for_backward(x, w3)
        return x @ w3[0]

    @staticmethod
    def backward(ctx, g):
        x, w3 = ctx.saved_tensors
        _Probe.seen.append(('bwd', w3, tc.packed_weights_cached(w3, transpose_w=True)))
        return g @ w3[0].t(), (x.t() @ g).unsqueeze(0)


def _step(kernel, x):
    w3 = kernel.unsqueeze(0) if kernel.dim() == 2 else kernel
    out = _Probe.apply(x, w3)
    out.sum().backward()


@pytest.mark.parametrize('two_d', [False, True])
def test_one_pack_per_direction_until_the_parameter_is_written(counted, two_d):
    _Probe.seen.clear()
    kernel = torch.nn.Parameter(torch.randn(4, 3) if two_d else torch.randn(1, 4, 3))
    x = torch.randn(5, 4, requires_grad=True)
    for _ in range(3):
        _step(kernel, x)
    # what the Function saw is the parameter (or a view of it) in forward AND in backward
    for (_, w3, _) in _Probe.seen:
        assert (w3._base if w3._base is not None else w3) is kernel
    assert counted == [((1, 4, 3), False), ((1, 4, 3), True)]               # three steps, two packs
    with torch.no_grad():
        kernel.add_(1.0)                                                     # an optimiser step: version bump
    _step(kernel, x)
    assert counted[2:] == [((1, 4, 3), False), ((1, 4, 3), True)]
    kernel.data = torch.randn_like(kernel)                                   # .to() / assign: new storage, same object
    _step(kernel, x)
    assert len(counted) == 6
    with torch.no_grad():                                                    # load_state_dict copies in place
        kernel.copy_(torch.zeros_like(kernel))
    _step(kernel, x)
    assert len(counted) == 8


def test_views_created_without_grad_still_name_their_parameter(counted):
    kernel = torch.nn.Parameter(torch.randn(4, 3))
    with torch.no_grad():
        a = tc.packed_weights_cached(kernel.unsqueeze(0))
        b = tc.packed_weights_cached(kernel.unsqueeze(0))
    assert a is b and len(counted) == 1


def test_a_new_parameter_never_inherits_a_dead_ones_operand(counted):
    """The failure the weak reference exists for: model A is freed, model B's kernel lands on the same address (and the same
    Python id) with the same version 0 -- it must get its own packed operand."""
    hits = []
    for _ in range(50):
        k = torch.nn.Parameter(torch.randn(1, 4, 3))
        hits.append(int(tc.packed_weights_cached(k)))
        del k
        gc.collect()
    assert hits == list(range(1, 51))
    assert len(tc._PACK_CACHE) <= 50


def test_dead_entries_are_dropped(counted):
    keep = torch.nn.Parameter(torch.randn(1, 4, 3))
    tc.packed_weights_cached(keep)
    params = [torch.nn.Parameter(torch.randn(1, 2, 2)) for _ in range(2100)]
    for p in params:
        tc.packed_weights_cached(p)
    del params, p
    gc.collect()
    tc.packed_weights_cached(torch.nn.Parameter(torch.randn(1, 2, 2)))       # beyond 2048 entries: sweep the dead ones
    assert len(tc._PACK_CACHE) < 100
    n = len(counted)
    tc.packed_weights_cached(keep)
    assert len(counted) == n                                                  # the live entry survived the sweep
