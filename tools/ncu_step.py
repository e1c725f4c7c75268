"""Minimal driver for ncu captures: 2 warm-up + 1 fused training steps of the bench workload (no
e2e / roofline / CPU-baseline legs, so the launch list stays short)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    sys.path.insert(0, p)
import torch
import ssn_models
from ssn_b200 import _lib
from oracle import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
model = ssn_models.SSN(20, 2, 5, 2, "RGB", base_model="BNInception", dropout=0)
bb = synth.synth_backbone(3, seed=0, calib_frames=2)
sd = model.state_dict()
for k, v in bb.items():
    sd["base_model." + k].copy_(v)
model = model.to(dev).train()
prec = {"fast": _lib.FAST_FP16, "exact": _lib.EXACT_FP32, "exact_tc": _lib.EXACT_TC}[sys.argv[2] if len(sys.argv) > 2 else "fast"]
model.set_precision(prec, 4096.0)
batch = tuple(t.to(dev) for t in synth.synth_batch(4, 20, 3, seed=0))
params = [p for p in model.parameters() if p.requires_grad]
flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)
off = 0
for p in params:
    p.grad = flat_grad[off:off + p.numel()].view_as(p)
    off += p.numel()
for i in range(steps):
    flat_grad.zero_()
    losses = model.fused_step(*batch)
torch.cuda.synchronize()
print("losses", losses.tolist())
