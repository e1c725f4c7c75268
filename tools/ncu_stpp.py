"""STPP kernels for `ncu --set full -k regex:stpp`: the vectorised forward / backward at 16384 proposals (1 GB of algorithmic
traffic) and the fused global-pool + STPP kernel at the bench shape (288 frames, fp32 5b output)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    sys.path.insert(0, p)
import torch
import ssn_models
from ssn_b200 import _lib
from oracle import synth

dev = torch.device("cuda:0")
model = ssn_models.SSN(20, 2, 5, 2, "RGB", base_model="BNInception", dropout=0)
bb = synth.synth_backbone(3, seed=0, calib_frames=2)
sd = model.state_dict()
for k, v in bb.items():
    sd["base_model." + k].copy_(v)
model = model.to(dev).train()
model.set_precision({"fast": _lib.FAST_FP16, "exact_tc": _lib.EXACT_TC}[sys.argv[1] if len(sys.argv) > 1 else "exact_tc"], 4096.0)
n = 16384
ft = torch.randn(n * 9, 1024, device=dev, requires_grad=True)
sc = torch.rand(n, 2, device=dev)
for _ in range(2):
    a, c = model.stpp(ft, sc, [2, 7, 9])
    torch.autograd.backward([a, c], [torch.ones_like(a), torch.ones_like(c)])
    ft.grad = None
batch = tuple(t.to(dev) for t in synth.synth_batch(4, 20, 3, seed=0))
for _ in range(2):
    model.fused_step(*batch)          # contains one gpool_stpp_v2 launch per step
torch.cuda.synchronize()
print("done")
