"""The C-ABI library loads and exports every symbol include/osb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'osb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(osb_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree():
    from openscene_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == _declared()


def test_library_exports_every_declared_symbol():
    from openscene_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    assert _cabi.lib().osb_version() == 100


def test_no_cpu_fallback():
    import torch
    from openscene_b200 import me
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        me.SparseTensor(torch.ones(4, 3), torch.zeros(4, 4, dtype=torch.int32))


def test_product_never_imports_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b', re.M)
    for pkg in ('openscene_b200', 'MinkowskiEngine'):
        for dp, _, fns in os.walk(os.path.join(ROOT, pkg)):
            for fn in fns:
                if fn.endswith('.py'):
                    src = open(os.path.join(dp, fn)).read()
                    hits = [m for m in pat.finditer(src)]
                    assert not hits, f"{pkg}/{fn} imports the oracle"
