"""torchrun worker of tests/test_dp_nccl.py (2 ranks, NCCL): data-parallel equivalence on the GPU.

(1) heads + multi-task loss kernel, 64 GLOBAL videos (the case where the completeness denominator of the global batch,
    int(65.28) = 65, is not world x the per-rank one, ops/ssn_ops.py:236-239): each rank runs the fused CUDA kernel on its
    32 videos with global_videos=64 / loss_scale=1/2, gradients are summed with ncclAllReduce -> equal to the single-GPU
    kernel on all 64 videos.
(2) whole training step (SSN.fused_step, EXACT_TC), 4 global videos: rank-sharded + all-reduced gradients == the single-GPU
    step on the global batch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    import ssn_models
    from ssn_b200 import _lib
    from ssn_b200.engine import heads_loss_fused
    from oracle import synth
    out = {}
    # ---- (1) heads + loss kernel, 64 global videos ----
    K, M, D, videos = 20, 5, 1024, 64
    n = videos * 8
    g = torch.Generator().manual_seed(1)
    course, stpp = torch.randn(n, D, generator=g), torch.randn(n, D * M, generator=g)
    ptype = torch.tensor([0, 1, 1, 1, 1, 1, 1, 2]).repeat(videos)
    target = torch.randint(1, K + 1, (n,), generator=g); target[ptype == 2] = 0
    rtarget = torch.randn(n, 2, generator=g)
    hd = synth.synth_heads(K, M, seed=0, std=0.02, bias_std=0.1)

    def fcs():
        act = ssn_models._HeadLinear(D, K + 1).to(dev); comp = ssn_models._HeadLinear(D * M, K).to(dev); reg = ssn_models._HeadLinear(D * M, 2 * K).to(dev)
        for fc, nm in ((act, "activity_fc"), (comp, "completeness_fc"), (reg, "regressor_fc")):
            fc.weight.data.copy_(hd[nm + ".weight"]); fc.bias.data.copy_(hd[nm + ".bias"])
        return act, comp, reg
    keys = ("d_act_w", "d_act_b", "d_comp_w", "d_comp_b", "d_reg_w", "d_reg_b", "d_course", "d_stpp")
    full = heads_loss_fused(course.to(dev), stpp.to(dev), *fcs(), ptype.to(dev), target.to(dev), rtarget.to(dev), K, M)
    per = n // world
    rows = slice(rank * per, (rank + 1) * per)
    part = heads_loss_fused(course[rows].to(dev), stpp[rows].to(dev), *fcs(), ptype[rows].to(dev), target[rows].to(dev), rtarget[rows].to(dev),
                            K, M, global_videos=videos, loss_scale=1.0 / world)
    errs = {}
    for k in keys[:6]:
        t = part[k].clone()
        dist.all_reduce(t)
        errs[k] = rel(t, full[k])
    for k in keys[6:]:                       # per-row gradients: this rank's rows of the global result
        errs[k] = rel(part[k], full[k][rows])
    losses = part["losses"].clone()           # reported losses are per-rank means (only the gradients carry loss_scale): average them
    dist.all_reduce(losses)
    errs["losses"] = rel(losses / world, full["losses"])
    out["heads_64_videos"] = errs
    # ---- (2) whole step, 4 global videos, EXACT_TC ----
    K2 = 4
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    hd2 = synth.synth_heads(K2, 5, seed=0, std=0.02, bias_std=0.1)

    def model():
        m = ssn_models.SSN(K2, 2, 5, 2, "RGB", base_model="BNInception", dropout=0, stpp_cfg=(1, (1, 2), 1))
        sd = m.state_dict()
        for k, v in bb.items():
            sd["base_model." + k].copy_(v)
        for k, v in hd2.items():
            sd[k].copy_(v)
        m = m.to(dev).train()
        m.set_precision(_lib.EXACT_TC, 1024.0)
        return m
    batch = synth.synth_batch(4, K2, 3, seed=3)
    m_full = model()
    l_full = m_full.fused_step(*[t.to(dev) for t in batch])
    m_part = model()
    vs = slice(rank * 2, rank * 2 + 2)
    l_part = m_part.fused_step(*[t[vs].to(dev) for t in batch], global_videos=4, loss_scale=1.0 / world)
    l_sum = l_part.clone(); dist.all_reduce(l_sum); l_sum /= world
    worst, worst_name = 0.0, ""
    num = den = 0.0
    for (n_, p), (_n2, q) in zip(m_part.named_parameters(), m_full.named_parameters()):
        if p.grad is None:
            continue
        gsum = p.grad.clone()
        dist.all_reduce(gsum)
        e = rel(gsum, q.grad)
        num += float((gsum.double() - q.grad.double()).pow(2).sum()); den += float(q.grad.double().pow(2).sum())
        if e > worst:
            worst, worst_name = e, n_
    out["step_4_videos"] = {"losses": rel(l_sum, l_full), "aggregate_grad": (num / den) ** 0.5, "worst_grad": worst, "worst_name": worst_name}
    if rank == 0:
        print("DP_NCCL_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
