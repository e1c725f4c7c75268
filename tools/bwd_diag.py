"""GPU diagnostic: run the engine's FULL backward and compare every intermediate gradient with the
oracle's autograd, in reverse graph order, to locate the first diverging op."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    sys.path.insert(0, p)
import torch
from oracle import ssn_oracle as O, synth
from ssn_b200 import _lib
from ssn_b200.engine import BackboneEngine


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))


def main():
    prec = _lib.FAST_FP16 if (len(sys.argv) > 1 and sys.argv[1] == "fast") else _lib.EXACT_FP32
    Fn = 4
    dev = torch.device("cuda:0")
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    names = [n for (n, *_r) in O.conv_layers(3)]
    for k in bb:
        if "_bn." not in k:
            bb[k].requires_grad_(True)
    x = synth.synth_frames(Fn, 3, seed=3)
    taps = {}
    feat = O.backbone_forward(bb, x, 3, taps=taps)
    for t in taps.values():
        if t.requires_grad:
            t.retain_grad()
    g = torch.Generator().manual_seed(9)
    dfeat = torch.randn(feat.shape, generator=g) * 0.01
    feat.backward(dfeat)
    eng = BackboneEngine(3, Fn, prec, True, 4096.0, dev)
    eng.pack([bb[n + ".weight"].detach().to(dev) for n in names], [bb[n + ".bias"].detach().to(dev) for n in names],
             [bb[n + "_bn.weight"].to(dev) for n in names], [bb[n + "_bn.bias"].to(dev) for n in names],
             [bb[n + "_bn.running_mean"].to(dev) for n in names], [bb[n + "_bn.running_var"].to(dev) for n in names])
    out = eng.forward(x.to(dev))
    print("forward feat rel", rel(out, feat.detach()))
    dw = [torch.zeros_like(bb[n + ".weight"].detach()).to(dev) for n in names]
    db = [torch.zeros_like(bb[n + ".bias"].detach()).to(dev) for n in names]
    eng.backward(dfeat.to(dev), dw, db)
    torch.cuda.synchronize()
    ops = eng.ops()
    ci = len(names)
    for (kind, iname, oname) in reversed(ops):
        if kind == "gpool":
            e = rel(eng.read(iname, grad=True) * 0 + eng.read(iname, grad=True), taps[iname].grad * (taps[iname] > 0))
            print("%-8s d(%s) [masked] rel %.3e" % (kind, iname, e))
            continue
        gref = taps[oname].grad
        if kind == "conv":
            ci -= 1
            gref = gref * (taps[oname] > 0)
            ew, eb = rel(dw[ci], bb[names[ci] + ".weight"].grad), rel(db[ci], bb[names[ci] + ".bias"].grad)
            e = rel(eng.read(oname, grad=True), gref)
            print("%-8s dz(%-36s) rel %.3e   dW %.3e  db %.3e" % (kind, oname, e, ew, eb))
        else:
            e = rel(eng.read(oname, grad=True), gref)
            print("%-8s d(%-37s) rel %.3e" % (kind, oname, e))


if __name__ == "__main__":
    main()
