"""CPU restatement (numpy) of the reference's per-video detection post-processing — TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import anything under oracle/).

Follows eval_detection_results.py:91-183 (default branch: top_k <= 0, no external class scores) and ops/utils.py:
  softmax             ops/utils.py:38-40
  temporal_nms        ops/utils.py:56-82
  perform_regression  eval_detection_results.py:147-160
Pinned by tests/golden/detect.npz, produced by oracle/gen_golden_detect.py from the REAL reference functions
(ops.utils.softmax / temporal_nms imported, perform_regression compiled from the reference script's own source text).
"""
import numpy as np


def softmax(scores):
    es = np.exp(scores - scores.max(axis=-1)[..., None])
    return es / es.sum(axis=-1)[..., None]


def temporal_nms(bboxes, thresh):
    """[[st, ed, score, ...], ...] -> rows kept, in descending score order (IoU may be negative: disjoint boxes are kept)"""
    t1, t2, scores = bboxes[:, 0], bboxes[:, 1], bboxes[:, 2]
    durations = t2 - t1
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        tt1 = np.maximum(t1[i], t1[order[1:]])
        tt2 = np.minimum(t2[i], t2[order[1:]])
        intersection = tt2 - tt1
        iou = intersection / (durations[i] + durations[order[1:]] - intersection).astype(float)
        order = order[np.where(iou <= thresh)[0] + 1]
    return bboxes[keep, :]


def perform_regression(detections):
    t0, t1 = detections[:, 0], detections[:, 1]
    center, duration = (t0 + t1) / 2, (t1 - t0)
    new_center = center + duration * detections[:, 3]
    new_duration = duration * np.exp(detections[:, 4])
    return np.concatenate((np.clip(new_center - new_duration / 2, 0, 1)[:, None], np.clip(new_center + new_duration / 2, 0, 1)[:, None],
                           detections[:, 2:]), axis=1)


def video_detections(rel_props, act_scores, comp_scores, reg_scores, nms_thresh, regress=True):
    """-> list over classes of [n_kept_c, 5] arrays (t0, t1, score, loc, dur): gen_detection_results (:104-114), NMS (:139-142),
    regression (:163-168) for one video"""
    num_class = comp_scores.shape[1]
    reg = reg_scores.reshape((-1, num_class, 2))
    combined = softmax(act_scores)[:, 1:] * np.exp(comp_scores)
    out = []
    for c in range(num_class):
        det = np.concatenate((rel_props, combined[:, c][:, None], reg[:, c, 0][:, None], reg[:, c, 1][:, None]), axis=1)
        det = temporal_nms(det, nms_thresh)
        out.append(perform_regression(det) if regress else det)
    return out
