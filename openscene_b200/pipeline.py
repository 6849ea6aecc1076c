"""One entry point for the whole hot path: points -> voxels -> MinkUNet -> open-vocabulary labels, all on the device.

Mirrors the inner loop of ``run/evaluate.py:283-323`` (feature types 'distill' and 'ensemble') for callers that hold raw
points instead of a pre-voxelised batch:

    seg = OpenVocabSegmenter(model, text_features, voxel_size=0.02)
    labels = seg.segment_points(points_xyz)                      # int64 [N_pts]; nothing but the labels leaves the GPU

* ``dataset/voxelizer.py:97-140`` (``Voxelizer.voxelize``: affine, floor, FNV key, first-occurrence unique) runs in
  csrc/voxelize.cu and hands ``inds_reverse`` straight to the matcher -- the voxel->point expansion
  ``predictions[inds_reverse]`` (evaluate.py:290) is fused into the matching kernel's operand load;
* when the caller does not ask for the 768-d features, the final 1x1x1 convolution is re-associated with the text matrix
  (``engine.fold_head``: W W^T = L L^T, U = W T^T) so the [N, 768] feature matrix is never written or read back.
"""
import numpy as np
import torch

from . import _cabi as C
from . import engine as _engine
from . import matching
from .voxelize import voxelize_points


class OpenVocabSegmenter:
    def __init__(self, model, text_features, voxel_size=0.02, normalize=True):
        """model: eval-mode MinkUNet / DisNet on a CUDA device (or a ready ``FusedMinkUNet``); text_features: unit-norm
        [K, C] CLIP text embeddings (util/util.py:24-46).  normalize=True gives cosine scores (the 'ensemble' branch's
        ``x / (|x| + 1e-5)``, evaluate.py:305-310); False the plain dot product of the 'distill' branch (:291)."""
        self.engine = model if isinstance(model, _engine.FusedMinkUNet) else _engine.FusedMinkUNet(model)
        self.device = self.engine.device
        self.text = text_features.to(self.device, torch.float16).contiguous()
        self.voxel_size = voxel_size
        self.normalize = normalize
        self._folded = None
        self._matrix = np.eye(4)
        np.fill_diagonal(self._matrix[:3, :3], 1.0 / voxel_size)

    def _fold(self):
        if self._folded is None or self._folded[3] != self.engine._signature():      # first use, or the model's weights changed
            self._folded = self.engine.fold_head(self.text.float())
        return self._folded

    @torch.no_grad()
    def segment_voxels(self, coords, feats, inds_reverse=None, want_features=False, want_scores=False):
        """coords int32 [Nv,4] (batch,x,y,z), feats fp32 [Nv,3].  Returns (labels int64 [Np], scores fp16 [Np,K] | None,
        features fp32 [Nv,C] | None) with Np = len(inds_reverse) (or Nv)."""
        if not want_features and self.normalize:
            scores, label, _ = self.engine.forward_scores(coords, feats, self._fold(), want_scores=want_scores)
            if inds_reverse is not None:                      # per-voxel results -> per-point (a [Np] / [Np,K] gather)
                inds_reverse = inds_reverse.to(label.device)
                label = label[inds_reverse]
                scores = scores[inds_reverse] if scores is not None else None
            return label, scores, None
        out = self.engine(coords, feats)
        scores, label, _ = matching._scores(out, inds_reverse, self.text, normalize=self.normalize, want_scores=want_scores)
        return label, scores, (out if want_features else None)

    @torch.no_grad()
    def segment_points(self, points, feats=None, matrix=None, want_scores=False):
        """points: CUDA float [N,3] (metres).  feats: per-POINT input features fp32 [N,3] or None (ones, the reference's
        default when ``input_color`` is off, dataset/feature_loader.py:180-184).  Returns int64 labels [N] (and fp16 scores
        [N,K] when asked): every point takes the label of its voxel, exactly ``pred[inds_reverse]`` of the reference."""
        C.require_cuda(points, 'points')
        with torch.cuda.device(self.device):
            cv, inds, inv, _ = voxelize_points(points, self._matrix if matrix is None else matrix)
            n_vox = cv.shape[0]
            coords = torch.zeros((n_vox, 4), dtype=torch.int32, device=self.device)       # batch index 0
            coords[:, 1:] = cv
            f = torch.ones((n_vox, 3), dtype=torch.float32, device=self.device) if feats is None else feats[inds].float()
            label, scores, _ = self.segment_voxels(coords, f, inv, want_features=False, want_scores=want_scores)
        return (label, scores) if want_scores else label
