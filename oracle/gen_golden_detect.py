"""tests/golden/detect.npz from the REAL reference (build container only: python -m oracle.gen_golden_detect).

ops/utils.py imports cleanly (softmax, temporal_nms).  eval_detection_results.py is a script (argparse + dataset loading at
import time), so perform_regression is compiled from the function's own source text taken from the reference file with ast
— the reference's code, unedited — and run on the NMS output."""
import ast
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def synth_video(n, K, seed):
    """proposals in [0,1] with heavy overlap, scores without ties"""
    g = np.random.RandomState(seed)
    c = g.rand(n).astype(np.float32)
    d = (0.02 + 0.3 * g.rand(n)).astype(np.float32)
    props = np.stack([np.clip(c - d / 2, 0, 1), np.clip(c + d / 2, 0, 1)], axis=1).astype(np.float32)
    act = g.randn(n, K + 1).astype(np.float32) * 2
    comp = g.randn(n, K).astype(np.float32)
    reg = (g.randn(n, K, 2) * 0.3).astype(np.float32)
    return props, act, comp, reg


def main():
    sys.path.insert(0, REF)
    import yaml
    _orig = yaml.load
    yaml.load = lambda s, Loader=yaml.SafeLoader: _orig(s, Loader=Loader)
    from ops.utils import softmax, temporal_nms
    src = open(os.path.join(REF, "eval_detection_results.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "perform_regression"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "eval_detection_results.py", "exec"), ns)
    perform_regression = ns["perform_regression"]
    out = {}
    for tag, n, K, thr, seed in (("a", 300, 4, 0.6, 1), ("b", 1000, 20, 0.6, 2), ("c", 37, 3, 0.3, 3), ("d", 1, 2, 0.6, 4)):
        props, act, comp, reg = synth_video(n, K, seed)
        combined = softmax(act)[:, 1:] * np.exp(comp)
        out[tag + "_props"], out[tag + "_act"], out[tag + "_comp"], out[tag + "_reg"] = props, act, comp, reg
        out[tag + "_thr"] = np.float64(thr)
        out[tag + "_combined"] = combined
        for c in range(K):
            det = np.concatenate((props, combined[:, c][:, None], reg[:, c, 0][:, None], reg[:, c, 1][:, None]), axis=1)
            kept = temporal_nms(det, thr)
            out["%s_nms_%d" % (tag, c)] = kept
            out["%s_det_%d" % (tag, c)] = perform_regression(kept)
    np.savez_compressed(os.path.join(GOLD, "detect.npz"), **out)
    print("wrote detect.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
