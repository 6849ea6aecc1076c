"""Seeded synthetic inputs shaped like the reference's data.  NumPy only.

* ``room_points``  -- ScanNet / Matterport-shaped indoor scene: floor + four walls (no ceiling) +
  random cuboids, surfaces sampled at 3 points per voxel face, 4 mm jitter.
* ``lidar_points`` -- nuScenes-shaped outdoor sweep.
* ``voxel_coords`` -- floor(p / voxel), shift to origin, unique: what ``Voxelizer.voxelize``
  (dataset/voxelizer.py:97-140) hands to the network, without the random rotation.
* ``text_embeddings`` -- K unit-norm fp16 rows standing in for CLIP text features
  (util/util.py:24-46 contract: unit-norm [K,768], fp16 on CUDA).
"""
import numpy as np

# (room size metres, number of cuboids) per BASELINE.json config
ROOMS = {
    'tiny': ((0.45, 0.40, 0.35), 1),          # a few thousand voxels: unit tests
    'config1_50k': ((1.36, 1.13, 1.21), 3),
    'config2_200k': ((3.2, 2.67, 2.22), 8),
    'config4_matterport': ((6.5, 5.0, 2.8), 18),
}


def _rect(rng, origin, u, v, density):
    area = np.linalg.norm(np.cross(u, v))
    n = max(int(area * density), 1)
    a, b = rng.rand(n, 1), rng.rand(n, 1)
    return origin + a * u + b * v


def room_points(size, n_cuboids, voxel=0.02, seed=0):
    rng = np.random.RandomState(seed)
    X, Y, Z = size
    dens = 3.0 / (voxel * voxel)
    ex, ey, ez = np.array([X, 0, 0.]), np.array([0, Y, 0.]), np.array([0, 0, Z])
    o = np.zeros(3)
    parts = [_rect(rng, o, ex, ey, dens),                       # floor
             _rect(rng, o, ex, ez, dens), _rect(rng, o + ey, ex, ez, dens),
             _rect(rng, o, ey, ez, dens), _rect(rng, o + ex, ey, ez, dens)]
    for _ in range(n_cuboids):
        sx, sy, sz = rng.uniform(0.3, 1.6), rng.uniform(0.3, 1.2), rng.uniform(0.3, 1.1)
        sx, sy, sz = min(sx, 0.8 * X), min(sy, 0.8 * Y), min(sz, 0.8 * Z)
        p = np.array([rng.uniform(0, X - sx), rng.uniform(0, Y - sy), 0.])
        bx, by, bz = np.array([sx, 0, 0.]), np.array([0, sy, 0.]), np.array([0, 0, sz])
        parts += [_rect(rng, p + bz, bx, by, dens),
                  _rect(rng, p, bx, bz, dens), _rect(rng, p + by, bx, bz, dens),
                  _rect(rng, p, by, bz, dens), _rect(rng, p + bx, by, bz, dens)]
    pts = np.concatenate(parts, 0)
    pts += rng.normal(0, 0.004, pts.shape)
    return pts


def lidar_points(target=790_000, seed=0):
    rng = np.random.RandomState(seed)
    n = int(1.6 * target)
    r = np.abs(rng.normal(0, 22.0, n)) + 1.0
    az = rng.uniform(0, 2 * np.pi, n)
    x, y = r * np.cos(az), r * np.sin(az)
    ground = rng.rand(n) < 0.65
    z = np.where(ground, rng.normal(0, 0.03, n), rng.uniform(0, 4.0, n))
    snap = rng.rand(n) < 0.25
    x = np.where(snap, np.round(x / 8.0) * 8.0, x)
    return np.stack([x, y, z], 1)


def voxel_coords(points, voxel, batch_index=0):
    """int32 [N,4] (batch, x, y, z), unique rows, in first-occurrence-of-sorted-key order."""
    c = np.floor(points / voxel).astype(np.int64)
    c -= c.min(0)
    c = np.unique(c, axis=0)
    rng = np.random.RandomState(12345)
    c = c[rng.permutation(len(c))]           # the voxeliser's FNV order is effectively random
    b = np.full((len(c), 1), batch_index, dtype=np.int64)
    return np.concatenate([b, c], 1).astype(np.int32)


def scene(name, voxel=None, seed=0, batch_index=0):
    """Named scene -> int32 coords [N,4]."""
    if name == 'config5_lidar':
        return voxel_coords(lidar_points(790_000, seed), voxel or 0.05, batch_index)
    if name.startswith('lidar_'):
        return voxel_coords(lidar_points(int(name.split('_')[1]), seed), voxel or 0.05, batch_index)
    size, nf = ROOMS[name]
    return voxel_coords(room_points(size, nf, voxel or 0.02, seed), voxel or 0.02, batch_index)


def scene_points(name, seed=0):
    """Raw float64 points [N_pts,3] (metres) and the voxel size of a named scene -- what ``Voxelizer.voxelize`` receives."""
    if name == 'config5_lidar':
        return lidar_points(790_000, seed), 0.05
    if name.startswith('lidar_'):
        return lidar_points(int(name.split('_')[1]), seed), 0.05
    size, nf = ROOMS[name]
    return room_points(size, nf, 0.02, seed), 0.02


def text_embeddings(k, c=768, seed=0):
    rng = np.random.RandomState(1000 + seed)
    t = rng.normal(size=(k, c))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return t.astype(np.float16)


def random_cloud(n, extent, seed=0, batch=1):
    """Small random unique coordinates for unit tests: ~n voxels inside [0,extent)^3 per batch."""
    rng = np.random.RandomState(seed)
    out = []
    for b in range(batch):
        c = np.unique(rng.randint(0, extent, size=(n, 3)), axis=0)
        c = c[rng.permutation(len(c))]
        out.append(np.concatenate([np.full((len(c), 1), b), c], 1))
    return np.concatenate(out, 0).astype(np.int32)


def randomize_bn_stats(model, seed=0):
    """Move every BatchNorm's running statistics / affine away from (0,1) with a seeded CPU generator so
    that BN folding is exercised (SURVEY.md 8d 'Weights').  Deterministic for a given torch build."""
    import torch
    g = torch.Generator(device='cpu').manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            c = m.num_features
            with torch.no_grad():
                m.running_mean.copy_((torch.rand(c, generator=g) - 0.5) * 0.2)
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(c, generator=g))
                m.bias.copy_((torch.rand(c, generator=g) - 0.5) * 0.2)
    return model


def build_model(arch='MinkUNet18A', out_channels=768, seed=0, ME=None, in_channels=3):
    """Seeded random-init network of the reference architecture (CPU tensors)."""
    import torch
    from .minkunet import mink_unet
    torch.manual_seed(seed)
    model = mink_unet(in_channels=in_channels, out_channels=out_channels, D=3, arch=arch, ME=ME)
    randomize_bn_stats(model, seed + 1)
    return model


def fusion_case(seed, n, with_depth, image_dim=(320, 240), n_frames=3):
    """Seeded points on the faces of a room, cameras inside it, depth = z-buffer of the points (mm, uint16) / 1000.
    Used by scripts/make_golden.py and the tests (which regenerate the inputs and compare with the stored reference outputs)."""
    rng = np.random.RandomState(seed)
    ext = np.array([4.0, 3.0, 2.5])
    pts = rng.rand(n, 3) * ext
    face = rng.randint(0, 5, n)                           # floor + 4 walls
    pts[face == 0, 2] = 0.0
    pts[face == 1, 0] = 0.0
    pts[face == 2, 0] = ext[0]
    pts[face == 3, 1] = 0.0
    pts[face == 4, 1] = ext[1]
    box = rng.rand(n) < 0.25                              # a box in the middle that occludes the walls behind it
    pts[box] = np.array([1.5, 1.0, 0.0]) + rng.rand(int(box.sum()), 3) * np.array([1.0, 0.8, 1.2])
    pts += rng.randn(n, 3) * 0.004
    W, H = image_dim
    intr = np.eye(4)
    intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2] = 577.870605 * 0.5, 577.870605 * 0.5, (W - 1) * 0.5, (H - 1) * 0.5
    poses, depths = [], []
    for f in range(n_frames):
        eye = np.array([0.4, 0.4, 1.2]) + rng.rand(3) * np.array([3.2, 2.2, 0.6])
        tgt = rng.rand(3) * ext * np.array([1, 1, 0.6])
        fwd = (tgt - eye) / np.linalg.norm(tgt - eye)
        right = np.cross(fwd, np.array([0, 0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
        poses.append(c2w)
        if with_depth:
            pc = (np.linalg.inv(c2w) @ np.concatenate([pts, np.ones((n, 1))], 1).T)
            z = pc[2]
            ok = z > 0.05
            u = np.round(pc[0, ok] * intr[0, 0] / z[ok] + intr[0, 2]).astype(int)
            v = np.round(pc[1, ok] * intr[1, 1] / z[ok] + intr[1, 2]).astype(int)
            inb = (u >= 0) & (u < W) & (v >= 0) & (v < H)
            zb = np.full((H, W), np.inf)
            np.minimum.at(zb, (v[inb], u[inb]), z[ok][inb])
            zb[~np.isfinite(zb)] = 0.0                    # holes: invalid depth, as in a real sensor image
            depths.append(np.round(zb * 1000.0).astype(np.uint16) / 1000.0)      # imread(uint16 png) / depth_scale
        else:
            depths.append(None)
    return pts, poses, depths, intr
