"""Per-launch CUDA-event timing of every tensor-core convolution of the FAST engine at the bench shape
(288 frames): forward launches, L2 flushed before each.  Prints us, algorithmic TFLOP/s and the share of peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    sys.path.insert(0, p)
import torch
from oracle import synth
from ssn_b200 import _lib
from ssn_b200.engine import BackboneEngine, conv_table


def main():
    Fn = int(sys.argv[1]) if len(sys.argv) > 1 else 288
    dev = torch.device("cuda:0")
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    table = conv_table(3)
    names = [t[0] for t in table]
    spec = {t[0]: t[1:] for t in table}
    e = BackboneEngine(3, Fn, _lib.FAST_FP16, False, 1024.0, dev)
    e.pack([bb[n + ".weight"].to(dev) for n in names], [bb[n + ".bias"].to(dev) for n in names],
           [bb[n + "_bn.weight"].to(dev) for n in names], [bb[n + "_bn.bias"].to(dev) for n in names],
           [bb[n + "_bn.running_mean"].to(dev) for n in names], [bb[n + "_bn.running_var"].to(dev) for n in names])
    x = synth.synth_frames(Fn, 3, seed=1).to(dev)
    e.forward(x)                       # fills every activation (and the space-to-depth input of conv1)
    torch.cuda.synchronize()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        peak = 1590.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_us, tot_fl = 0.0, 0.0
    only = os.environ.get("SSNB_LAYERS")
    only = set(only.split(",")) if only else None
    for i, (kind, iname, oname) in enumerate(e.ops()):
        if kind != "conv" or (only and oname[:-3] not in only):
            continue
        ci, co, k, s, p = spec[oname[:-3]]
        _c, hh, ww = e.value_shape(oname)
        best = 1e9
        for _ in range(3):
            flush.zero_()
            a.record(); e.run_op(i, False); b.record(); b.synchronize()
            best = min(best, a.elapsed_time(b) * 1e3)
        fl = 2.0 * Fn * hh * ww * co * ci * k * k
        tot_us += best; tot_fl += fl
        print("%-34s ci %4d co %4d k%d s%d hw %3d  %8.1f us  %7.1f TF/s  %.3f" % (oname[:-3], ci, co, k, s, hh, best, fl / best / 1e6, fl / best / 1e6 / peak))
    print("TOTAL (unfused per-op launches) %.1f us, %.1f TF/s, frac %.3f" % (tot_us, tot_fl / tot_us / 1e6, tot_fl / tot_us / 1e6 / peak))


if __name__ == "__main__":
    main()
