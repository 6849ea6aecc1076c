"""Generate tests/golden/*.npz in the build container (needs /root/reference; NOT run on the GPU box).

1. voxelizer_*.npz : outputs of the reference's own ``dataset/voxelizer.py`` (imported unmodified, with the
   ``collections.Sequence/Iterable`` aliases Python 3.12 needs) for seeded inputs + the exact 4x4 matrix it drew.
2. unet_*.npz      : activations of the reference's unmodified ``models/mink_unet.py`` run on the CPU oracle
   (oracle/me_cpu.py registered as ``MinkowskiEngine``) in fp64, seeded weights, eval-mode BN with randomised
   statistics, on a small synthetic room.  These pin the *topology*; the ME arithmetic itself is the oracle's
   (parity unpinned against real MinkowskiEngine -- see oracle/__init__.py).

Usage: python scripts/make_golden.py
"""
import collections
import collections.abc
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')


def golden_voxelizer():
    collections.Sequence = collections.abc.Sequence      # voxelization_utils.py:6
    collections.Iterable = collections.abc.Iterable      # voxelizer.py:55
    sys.path.insert(0, REF)
    from dataset.voxelizer import Voxelizer
    cases = {
        'aug_f64': dict(n=4000, extent=3.0, voxel=0.05, aug=True, dtype=np.float64, seed=1),
        'noaug_f32': dict(n=3000, extent=2.0, voxel=0.05, aug=False, dtype=np.float32, seed=2),
        'dups_f64': dict(n=5000, extent=0.6, voxel=0.05, aug=True, dtype=np.float64, seed=3),
        'neg_f64': dict(n=2000, extent=4.0, voxel=0.02, aug=True, dtype=np.float64, seed=4, shift=-2.0),
    }
    for name, c in cases.items():
        rng = np.random.RandomState(c['seed'])
        pts = (rng.rand(c['n'], 3) * c['extent'] + c.get('shift', 0.0)).astype(c['dtype'])
        vox = Voxelizer(voxel_size=c['voxel'], clip_bound=None, use_augmentation=c['aug'],
                        scale_augmentation_bound=(0.9, 1.1),
                        rotation_augmentation_bound=((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64), (-np.pi, np.pi)),
                        translation_augmentation_ratio_bound=((-0.2, 0.2), (-0.2, 0.2), (0, 0)))
        np.random.seed(c['seed'] + 100)
        M_v, M_r = vox.get_transformation_matrix()
        rigid = (M_r @ M_v) if c['aug'] else M_v
        np.random.seed(c['seed'] + 100)          # same draws inside voxelize()
        feats = np.zeros((c['n'], 3), dtype=np.float32)
        labels = np.zeros(c['n'], dtype=np.int64)
        coords_aug, _, _, inds_rec, inds = vox.voxelize(pts, feats, labels, return_ind=True)
        np.savez_compressed(os.path.join(OUT, f'voxelizer_{name}.npz'), points=pts, matrix=rigid,
                            coords_vox=coords_aug, inds=np.asarray(inds), inds_reverse=np.asarray(inds_rec))
        print(name, 'points', c['n'], '-> voxels', len(inds))


def golden_unet():
    from oracle import me_cpu
    me_cpu.install_as_minkowski_engine()
    sys.path.insert(0, REF)
    from models.mink_unet import mink_unet as ref_mink_unet
    from openscene_b200 import synth
    coords = synth.scene('tiny')
    print('tiny scene voxels', len(coords))
    for arch in ('MinkUNet18A', 'MinkUNet34C'):
        torch.manual_seed(0)
        model = ref_mink_unet(in_channels=3, out_channels=768, D=3, arch=arch)
        synth.randomize_bn_stats(model, 1)
        model = model.double().eval()
        rng = np.random.RandomState(7)
        feats = torch.from_numpy(rng.rand(len(coords), 3).astype(np.float32)).double()   # fp32-representable
        with torch.no_grad():
            x = me_cpu.SparseTensor(feats, torch.from_numpy(coords))
            out = model(x)
        out = out.numpy()
        rows = np.sort(np.random.RandomState(11).choice(len(coords), 128, replace=False))
        keys = list(model.state_dict().keys())
        shapes = [tuple(v.shape) for v in model.state_dict().values()]
        np.savez_compressed(os.path.join(OUT, f'unet_{arch}.npz'), coords=coords, feats=feats.numpy().astype(np.float32),
                            rows=rows, out_rows=out[rows].astype(np.float32),
                            row_norm=np.linalg.norm(out, axis=1).astype(np.float32),
                            col_sum=out.sum(0).astype(np.float64),
                            state_keys=np.array(keys), state_shapes=np.array([str(s) for s in shapes]),
                            n_params=np.int64(sum(p.numel() for p in model.parameters())))
        print(arch, 'params', sum(p.numel() for p in model.parameters()), 'out', out.shape, 'abs mean', np.abs(out).mean())


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    golden_voxelizer()
    golden_unet()
