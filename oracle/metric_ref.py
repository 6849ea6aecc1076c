"""CPU restatement of the segmentation metrics.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PINNED: ``tests/golden/metric_*.npz`` hold outputs of the reference's own ``util/metric.py`` (``confusion_matrix``,
``evaluate``) and ``util/util.py`` (``intersectionAndUnion`` and ``intersectionAndUnionGPU``; the module imported with
stub ``open3d`` / ``clip`` / ``matplotlib`` modules those functions never touch, and ``Tensor.cuda`` patched to the
identity because the build container has no GPU) -- see ``scripts/make_golden.py``.
"""
import numpy as np

UNKNOWN_ID = 255         # util/metric.py:5
NO_FEATURE_ID = 256      # util/metric.py:6


def confusion_matrix(pred_ids, gt_ids, num_classes):
    """util/metric.py:9-25: rows = prediction, columns = ground truth; points with gt == 255 are skipped; predictions
    equal to 256 ("no feature") go to an extra row that is cut off again."""
    pred = np.asarray(pred_ids).astype(np.int64).copy()
    gt = np.asarray(gt_ids).astype(np.int64)
    keep = gt != UNKNOWN_ID
    conf = np.zeros((num_classes + 1, num_classes + 1), dtype=np.uint64)
    pred[pred == NO_FEATURE_ID] = num_classes
    np.add.at(conf, (pred[keep], gt[keep]), 1)
    return conf[:num_classes, :num_classes]


def mean_iou(pred_ids, gt_ids, num_classes):
    """util/metric.py:44-78 (``evaluate`` without the printing): classes absent from gt contribute 0, and the mean is
    over ALL classes.  Returns (mean_iou, mean_acc)."""
    conf = confusion_matrix(pred_ids, gt_ids, num_classes).astype(np.int64)
    gt = np.asarray(gt_ids)
    miou = macc = 0.0
    for i in range(num_classes):
        n_gt = int((gt == i).sum())
        if n_gt == 0:
            continue
        tp = int(conf[i, i])
        denom = tp + (int(conf[i, :].sum()) - tp) + (int(conf[:, i].sum()) - tp)
        if denom == 0:           # get_iou returns a bare nan here and the reference would fail on [0]; unreachable when n_gt > 0 and gt < C
            continue
        miou += tp / denom
        macc += tp / n_gt
    return miou / num_classes, macc / num_classes


def intersection_and_union(output, target, K, ignore_index=255):
    """util/util.py:117-145: labels in 0..K-1; where target == ignore the prediction is ignored too; values outside
    0..K-1 fall out of the K-bin histograms.  Returns int64 (area_intersection, area_union, area_target)."""
    out = np.asarray(output).reshape(-1).astype(np.int64).copy()
    tgt = np.asarray(target).reshape(-1).astype(np.int64)
    out[tgt == ignore_index] = ignore_index
    def hist(v):
        v = v[(v >= 0) & (v < K)]
        return np.bincount(v, minlength=K).astype(np.int64)
    inter, a_out, a_tgt = hist(out[out == tgt]), hist(out), hist(tgt)
    return inter, a_out + a_tgt - inter, a_tgt
