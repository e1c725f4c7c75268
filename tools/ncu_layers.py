"""A handful of single tensor-core convolution launches at the bench shape (288 frames) for `ncu --set full`:
inputs are written straight into the engine's buffers, so the only umma_conv* launches are the listed layers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    sys.path.insert(0, p)
import torch
from oracle import synth
from ssn_b200 import _lib
from ssn_b200.engine import BackboneEngine, conv_table

LAYERS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["conv2_3x3_reduce", "conv2_3x3", "inception_3a_3x3", "inception_4a_double_3x3_2"]


def main():
    Fn = 288
    dev = torch.device("cuda:0")
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    names = [t[0] for t in conv_table(3)]
    e = BackboneEngine(3, Fn, _lib.FAST_FP16, False, 1024.0, dev)
    e.pack([bb[n + ".weight"].to(dev) for n in names], [bb[n + ".bias"].to(dev) for n in names],
           [bb[n + "_bn.weight"].to(dev) for n in names], [bb[n + "_bn.bias"].to(dev) for n in names],
           [bb[n + "_bn.running_mean"].to(dev) for n in names], [bb[n + "_bn.running_var"].to(dev) for n in names])
    g = torch.Generator().manual_seed(1)
    for i, (kind, iname, oname) in enumerate(e.ops()):
        if kind != "conv" or oname[:-3] not in LAYERS:
            continue
        c, h, w = e.value_shape(iname)
        e.write(iname, torch.randn(Fn, c, h, w, generator=g).to(dev))
        torch.cuda.synchronize()
        e.run_op(i, False)
        torch.cuda.synchronize()
        print("ran", oname)


if __name__ == "__main__":
    main()
