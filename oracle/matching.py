"""CPU restatement of the open-vocabulary matching and distillation loss.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ``run/evaluate.py:283-326`` (feature_type 'distill' / 'fusion' / 'ensemble') and
``run/distill.py:322-328`` (cosine / L1 loss).  The reference multiplies fp16 tensors on CUDA
(``.half() @ text_features.t()``): fp16 operands, fp32 accumulation, one rounding of the result
to fp16.  That is restated here as ``(a.half().float() @ t.half().float().T).half()``.
"""
import torch


def _hmm(a, text):
    """fp16 x fp16 -> fp16 product with fp32 accumulation (cuBLAS HGEMM semantics)."""
    return (a.half().float() @ text.half().float().t()).half()


def match_distill(pred_vox, inds_reverse, text):
    """evaluate.py:288-292.  pred_vox fp32 [Nv,C]; returns (scores fp16 [Np,K], label int64 [Np])."""
    p = pred_vox[inds_reverse, :]
    pred = _hmm(p, text)
    return pred, torch.max(pred, 1)[1]


def match_fusion(feat_vox, inds_reverse, text):
    """evaluate.py:293-296."""
    return match_distill(feat_vox, inds_reverse, text)


def _l2n(x):
    return x / (x.norm(dim=-1, keepdim=True) + 1e-5)


def match_ensemble(pred_vox, feat_vox_fp16, inds_reverse, text):
    """evaluate.py:302-323.  feat_vox_fp16 is the fused 2-D feature, stored fp16
    (scripts/feature_fusion/fusion_util.py:87), so its norm is taken in fp16 like the reference.

    Returns (scores fp16 [Np,K], label int64 [Np], feat_ensemble fp16 [Np,C], mask bool [Np])."""
    feat_fuse = feat_vox_fp16[inds_reverse, :]
    pred_fusion = _hmm(_l2n(feat_fuse), text)
    predictions = pred_vox[inds_reverse, :]
    pred_distill = _hmm(_l2n(predictions), text)
    feat_ensemble = predictions.clone().half()
    mask_ = pred_distill.max(dim=-1)[0] < pred_fusion.max(dim=-1)[0]
    feat_ensemble[mask_] = feat_fuse[mask_]
    pred = _hmm(feat_ensemble, text)
    return pred, torch.max(pred, 1)[1], feat_ensemble, mask_


def distill_loss(output_3d, feat_3d, mask=None, loss_type='cosine'):
    """distill.py:322-328: ``(1 - CosineSimilarity()(o, f)).mean()`` or L1."""
    o = output_3d[mask] if mask is not None else output_3d
    f = feat_3d.to(o.dtype)
    if loss_type == 'cosine':
        return (1 - torch.nn.CosineSimilarity()(o, f)).mean()
    if loss_type == 'l1':
        return torch.nn.L1Loss()(o, f)
    raise NotImplementedError
