"""The table-driven mirror (openscene_b200/minkunet.py) against activations of the REFERENCE's unmodified
models/mink_unet.py (golden, produced in the build container).  Both run on the CPU oracle here, in fp64:
same seed -> same state dict -> same output, so this pins names, shapes, construction order and dataflow."""
import numpy as np
import pytest
import torch

from openscene_b200 import minkunet, synth
from oracle import me_cpu
from tests.util import golden, rel_row_err


@pytest.mark.parametrize('arch', ['MinkUNet18A', 'MinkUNet34C'])
def test_mirror_reproduces_reference_model(arch):
    g = golden(f'unet_{arch}.npz')
    model = synth.build_model(arch, 768, seed=0, ME=me_cpu.as_module()).double().eval()
    sd = model.state_dict()
    assert list(sd.keys()) == g['state_keys'].tolist()
    assert [str(tuple(v.shape)) for v in sd.values()] == g['state_shapes'].tolist()
    assert sum(p.numel() for p in model.parameters()) == int(g['n_params'])
    with torch.no_grad():
        x = me_cpu.SparseTensor(torch.from_numpy(g['feats']).double(), torch.from_numpy(g['coords']))
        out = model(x).numpy()
    assert rel_row_err(out[g['rows']], g['out_rows']) < 1e-6          # golden rows stored as fp32
    assert np.allclose(np.linalg.norm(out, axis=1), g['row_norm'], rtol=1e-5)
    assert np.allclose(out.sum(0), g['col_sum'], rtol=1e-8, atol=1e-8)


def test_expected_checkpoint_key_names():
    g = golden('unet_MinkUNet18A.npz')
    keys = set(g['state_keys'].tolist())
    for k in ['conv0p1s1.kernel', 'bn0.bn.weight', 'bn0.bn.running_mean', 'bn0.bn.num_batches_tracked',
              'block1.0.conv1.kernel', 'block1.0.norm1.bn.bias', 'block2.0.downsample.0.kernel',
              'block2.0.downsample.1.bn.running_var', 'convtr4p16s2.kernel', 'bntr7.bn.weight', 'final.kernel']:
        assert k in keys, k
    shapes = dict(zip(g['state_keys'].tolist(), g['state_shapes'].tolist()))
    assert shapes['conv0p1s1.kernel'] == '(125, 3, 32)'
    assert shapes['block2.0.downsample.0.kernel'] == '(32, 64)'
    assert shapes['convtr4p16s2.kernel'] == '(8, 256, 128)'
    assert shapes['final.kernel'] == '(96, 768)'


def test_disnet_prefix():
    import types
    cfg = types.SimpleNamespace(arch_3d='MinkUNet14A', feature_2d_extractor='lseg')
    net = minkunet.DisNet(cfg, ME=me_cpu.as_module())
    assert all(k.startswith('net3d.') for k in net.state_dict())
    assert net.net3d.final.kernel.shape == (96, 512)
