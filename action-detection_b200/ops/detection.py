"""Detection post-processing on the GPU — the per-video work of the reference's eval_detection_results.py:91-183
(combined scores, class-wise temporal NMS, location regression) and ops/utils.py:56-82 (temporal_nms), executed by
libssn_b200.so (csrc/detect.cu) instead of numpy loops.  Inputs and outputs are CUDA tensors; nothing here has a CPU path.
"""
import ctypes as C

import torch

from ssn_b200._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def video_detections(rel_props, act_scores, comp_scores, reg_scores, nms_threshold, regress=True):
    """rel_props [N,2], act_scores [N,K+1], comp_scores [N,K], reg_scores [N,K,2] (or [N,2K]) of ONE video ->
    (detections [K,N,5], counts [K] int32): for class c the first counts[c] rows of detections[c] are the surviving
    (t0, t1, score, loc, dur) in descending score order — `dataset_detections[c][video]` after gen_detection_results, NMS and
    perform_regression (eval_detection_results.py:104-114,139-142,147-168; default branch, top_k <= 0)."""
    for t, nm in ((rel_props, "rel_props"), (act_scores, "act_scores"), (comp_scores, "comp_scores"), (reg_scores, "reg_scores")):
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor: libssn_b200 has no CPU path" % nm)
    dev = act_scores.device
    props = rel_props.reshape(-1, 2).contiguous().float()
    act, comp = act_scores.contiguous().float(), comp_scores.contiguous().float()
    n, K = comp.shape
    reg = reg_scores.reshape(n, K, 2).contiguous().float()
    det = torch.zeros(K, max(n, 1), 5, dtype=torch.float32, device=dev)
    cnt = torch.zeros(K, dtype=torch.int32, device=dev)
    ws = torch.empty(max(n * K, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ssnb_detect_postprocess(props.data_ptr(), act.data_ptr(), comp.data_ptr(), reg.data_ptr(), n, K, float(nms_threshold),
                                          int(bool(regress)), det.data_ptr(), cnt.data_ptr(), ws.data_ptr(), _stream()), None, "detect_postprocess")
    return det, cnt


def temporal_nms(bboxes, thresh):
    """ops/utils.py:56-82 on the GPU: bboxes [n, >=3] rows (st, ed, score, ...) -> the kept rows in descending score order."""
    if not bboxes.is_cuda:
        raise RuntimeError("bboxes must be a CUDA tensor: libssn_b200 has no CPU path")
    n = bboxes.shape[0]
    if n == 0:
        return bboxes
    dev = bboxes.device
    b = bboxes.contiguous().float()
    props = b[:, :2].contiguous()
    scores = b[:, 2].contiguous()                       # [n, K=1]: ranked as they are
    # carry the row index through the "regression" slots: (loc, dur) = (row index, 0)
    reg = torch.stack([torch.arange(n, device=dev, dtype=torch.float32), torch.zeros(n, device=dev)], dim=1).contiguous()
    det = torch.zeros(1, n, 5, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ssnb_detect_postprocess(props.data_ptr(), None, None, reg.data_ptr(), n, 1, float(thresh), 0, det.data_ptr(), cnt.data_ptr(),
                                          scores.data_ptr(), _stream()), None, "detect_postprocess")
    keep = det[0, :int(cnt.item()), 3].long()
    return bboxes[keep]
