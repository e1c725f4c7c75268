"""Data-parallel host logic: one process per GPU, videos sharded over ranks, one all-reduce of a
flat fp32 gradient buffer (replaces nn.DataParallel, ssn_train.py:67).

Loss normalisation stays exact under sharding: cross-entropy and regression are means
over per-video fixed row counts, so rank means average to the global mean; the completeness loss
divides by pos_cnt + int(neg_cnt * ohem_ratio) of the GLOBAL batch (ops/ssn_ops.py:236-239), which
is NOT world * the per-rank value in general (B=64: int(65.28)=65 vs 8*int(8.16)=64).
"""
import torch
import torch.distributed as dist


def completeness_denominator(global_videos, fg_per_video=1, comp_group=7, ohem_ratio=0.17):
    """pos_cnt + int(neg_cnt * ratio) of the global batch, Python int() semantics."""
    neg = comp_group - fg_per_video
    return global_videos * fg_per_video + int(global_videos * neg * ohem_ratio)


def shard_loss_config(global_videos, world, fg_per_video=1, comp_group=7, ohem_ratio=0.17):
    """(comp_denom, loss_scale) for one rank such that SUMMING the per-rank gradients over ranks
    reproduces the global-batch gradient."""
    assert global_videos % world == 0, "equal videos per rank"
    denom = completeness_denominator(global_videos, fg_per_video, comp_group, ohem_ratio)
    return float(denom) / world, 1.0 / world


class FlatGrads:
    """Views every parameter's .grad into one flat fp32 buffer so the gradient exchange is a single
    collective (42.3 MB for K=20 RGB)."""

    def __init__(self, params, device):
        self.params = [p for p in params if p.requires_grad]
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
