"""Multi-view feature fusion on the GPU behind the reference's interface:
``PointCloudToImageMapper`` (``scripts/feature_fusion/fusion_util.py:93-139``) and the accumulate / average loop of
``process_one_scene`` (``scripts/feature_fusion/scannet_openseg.py:74-108``).  Projection, depth test, pixel-feature
gather and the running fp32 sum run in csrc/fusion.cu; a batch of up to 32 frames is one native call."""
import math

import numpy as np
import torch

from . import _cabi as C

MAX_FRAMES_PER_CALL = 32


def make_intrinsic(fx, fy, mx, my):
    """fusion_util.py:17-25."""
    intrinsic = np.eye(4)
    intrinsic[0][0], intrinsic[1][1], intrinsic[0][2], intrinsic[1][2] = fx, fy, mx, my
    return intrinsic


def adjust_intrinsic(intrinsic, intrinsic_image_dim, image_dim):
    """fusion_util.py:27-39 (modifies and returns ``intrinsic`` like the reference)."""
    if intrinsic_image_dim == image_dim:
        return intrinsic
    resize_width = int(math.floor(image_dim[1] * float(intrinsic_image_dim[0]) / float(intrinsic_image_dim[1])))
    intrinsic[0, 0] *= float(resize_width) / float(intrinsic_image_dim[0])
    intrinsic[1, 1] *= float(image_dim[1]) / float(intrinsic_image_dim[1])
    intrinsic[0, 2] *= float(image_dim[0] - 1) / float(intrinsic_image_dim[0] - 1)
    intrinsic[1, 2] *= float(image_dim[1] - 1) / float(intrinsic_image_dim[1] - 1)
    return intrinsic


def _dev_points(coords, device):
    t = torch.as_tensor(coords)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    t = t.to(device).contiguous()
    assert t.dim() == 2 and t.shape[1] == 3, "points must be [N,3]"
    return t


def _cams(poses, intrinsics, device):
    w2c = np.stack([np.linalg.inv(np.asarray(p, dtype=np.float64)) for p in poses]).reshape(len(poses), 16)   # fusion_util.py:120
    k = np.array([[i[0][0], i[1][1], i[0][2], i[1][2]] for i in intrinsics], dtype=np.float64)
    return torch.from_numpy(w2c).to(device), torch.from_numpy(k).to(device)


def _launch(points, w2c, intr, depth, feat, F, H, W, Cw, cut, vis, sum_, counter, mapping):
    n = points.shape[0]
    ws_bytes = C.lib().osb_fusion_workspace_bytes(n, F)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device)
    C.call('osb_fusion_accumulate', C.ptr(points), int(points.dtype == torch.float64), n, C.ptr(w2c), C.ptr(intr),
           C.ptr(depth) if depth is not None else None, C.ptr(feat) if feat is not None else None, F, H, W, Cw, int(cut),
           float(vis), C.ptr(sum_) if sum_ is not None else None, C.ptr(counter) if counter is not None else None,
           C.ptr(mapping) if mapping is not None else None, C.ptr(ws), ws_bytes, C.stream_ptr())


class PointCloudToImageMapper:
    """Same constructor and ``compute_mapping`` signature as the reference; image_dim = (W, H)."""

    def __init__(self, image_dim, visibility_threshold=0.25, cut_bound=0, intrinsics=None, device='cuda'):
        self.image_dim = image_dim
        self.vis_thres = visibility_threshold
        self.cut_bound = cut_bound
        self.intrinsics = intrinsics
        self.device = torch.device(device)

    def compute_mapping(self, camera_to_world, coords, depth=None, intrinsic=None, as_tensor=False):
        """[N,3] = (row, col, visible).  Default: a NumPy int array, exactly what the reference returns (its caller writes
        it into a NumPy buffer, scannet_openseg.py:95); ``as_tensor=True`` keeps the CUDA int32 tensor on the device."""
        if self.intrinsics is not None:
            intrinsic = self.intrinsics
        assert intrinsic is not None, "PointCloudToImageMapper: no intrinsics given (constructor or compute_mapping argument)"
        pts = _dev_points(coords, self.device)
        W, H = self.image_dim
        with torch.cuda.device(self.device):
            w2c, k = _cams([camera_to_world], [intrinsic], self.device)
            d = None
            if depth is not None:
                d = torch.as_tensor(depth).to(self.device, torch.float64).contiguous().view(1, H, W)
            out = torch.empty((1, pts.shape[0], 3), dtype=torch.int32, device=self.device)
            _launch(pts, w2c, k, d, None, 1, H, W, 8, self.cut_bound, self.vis_thres, None, None, out)
        return out[0] if as_tensor else out[0].cpu().numpy().astype(int)


class FeatureFusion:
    """Running mean of per-pixel features over the frames that see each point.

        fuser = FeatureFusion(locs_in, feat_dim=768, mapper=point2img_mapper)
        for batch of frames:  fuser.add_frames(poses, depths, feats)       # feats: fp16 [F,H,W,C] (or a list of [H,W,C])
        feat_bank, point_ids = fuser.finalize()                             # scannet_openseg.py:104-106
    """

    def __init__(self, points, feat_dim, mapper):
        self.mapper = mapper
        self.points = _dev_points(points, mapper.device)
        n = self.points.shape[0]
        self.feat_dim = feat_dim
        self.sum_features = torch.zeros((n, feat_dim), dtype=torch.float32, device=mapper.device)
        self.counter = torch.zeros(n, dtype=torch.float32, device=mapper.device)

    def add_frames(self, poses, depths, feats, intrinsics=None, channels_first=None):
        """feats: fp16 [F,H,W,C], or the reference's permuted [F,C,H,W] view (``channels_first=True``; None = detect from
        the shape, which is ambiguous only when the image width equals the feature width)."""
        dev = self.mapper.device
        W, H = self.mapper.image_dim
        if isinstance(feats, (list, tuple)):
            feats = torch.stack([torch.as_tensor(f) for f in feats])
        feats = feats.to(dev)
        if feats.dtype != torch.float16:
            feats = feats.half()
        F = feats.shape[0]
        if channels_first is None:
            channels_first = feats.shape[1] == self.feat_dim and feats.shape[-1] != self.feat_dim
        if channels_first:
            feats = feats.permute(0, 2, 3, 1)                    # the reference's [C,H,W] view of HWC memory
        feats = feats.contiguous()
        assert feats.shape == (F, H, W, self.feat_dim), f"features must be [F,{H},{W},{self.feat_dim}]"
        assert len(poses) == F
        if intrinsics is None:
            assert self.mapper.intrinsics is not None, "FeatureFusion.add_frames: no intrinsics (mapper has none, none passed)"
            intrinsics = [self.mapper.intrinsics] * F
        if depths is not None and any(d is None for d in depths):
            assert all(d is None for d in depths), "either every frame of a batch has a depth image or none"
            depths = None
        with torch.cuda.device(dev):
            d_all = None
            if depths is not None:
                d_all = torch.stack([torch.as_tensor(d) for d in depths]).to(dev, torch.float64).contiguous()
                assert d_all.shape == (F, H, W)
            for f0 in range(0, F, MAX_FRAMES_PER_CALL):
                f1 = min(F, f0 + MAX_FRAMES_PER_CALL)
                w2c, k = _cams(poses[f0:f1], intrinsics[f0:f1], dev)
                _launch(self.points, w2c, k, d_all[f0:f1] if d_all is not None else None, feats[f0:f1], f1 - f0, H, W,
                        self.feat_dim, self.mapper.cut_bound, self.mapper.vis_thres, self.sum_features, self.counter, None)

    def finalize(self):
        """Returns (feat_bank fp32 [N,C], point_ids int64): the mean feature per point and the points seen at least once."""
        with torch.cuda.device(self.mapper.device):
            bank = torch.empty_like(self.sum_features)
            C.call('osb_fusion_finalize', C.ptr(self.sum_features), C.ptr(self.counter), self.sum_features.shape[0], self.feat_dim,
                   C.ptr(bank), C.stream_ptr())
        return bank, torch.nonzero(self.counter > 0)[:, 0]
