"""Deterministic synthetic weights and inputs for the SSN hot path.
TEST INFRASTRUCTURE ONLY (see oracle/ssn_oracle.py header for who may import this).

Pretrained weights are unobtainable offline, so every parity test, the golden generator and the
bench use the same seeded generator: kaiming-uniform convs, random BN affine, and BN running
statistics calibrated by one batch-statistics pass over 8 synthetic frames so that activations
stay O(1) through the 69 layers.  Everything is a function of (seed, in_channels, num_class,
stpp multiplier) only, so the GPU box regenerates identical tensors (same torch build).
"""
import torch

from . import ssn_oracle as O

MEAN_RGB = (104.0, 117.0, 128.0)        # ssn_models.py:126 (BGR order after transforms.Stack roll)


def synth_frames(n_frames, in_channels=3, size=224, seed=0):
    """uint8 - mean frames, NCHW fp32 (ssn_models.py:126-130, transforms.py:67-90)."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randint(0, 256, (n_frames, in_channels, size, size), generator=g).float()
    if in_channels == 3:
        x -= torch.tensor(MEAN_RGB).view(1, 3, 1, 1)
    else:
        x -= 128.0
    return x


def synth_backbone(in_channels=3, seed=0, calib_frames=8, size=224):
    """Seeded BNInception parameters keyed like the reference state_dict."""
    g = torch.Generator().manual_seed(2000 + seed)
    p = {}
    for (id_, cin, cout, k, st, pad) in O.conv_layers(in_channels):
        fan_in = cin * k * k
        bound = (6.0 / fan_in) ** 0.5
        p[id_ + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        p[id_ + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.1
        p[id_ + "_bn.weight"] = 0.75 + 0.5 * torch.rand(cout, generator=g)
        p[id_ + "_bn.bias"] = 0.3 * (torch.rand(cout, generator=g) * 2 - 1) + 0.1
        p[id_ + "_bn.running_mean"] = torch.zeros(cout)
        p[id_ + "_bn.running_var"] = torch.ones(cout)
    # calibration: one pass with batch statistics; copy them into the running buffers
    stats = {}
    with torch.no_grad():
        O.backbone_forward(p, synth_frames(calib_frames, in_channels, size, seed=seed + 77), in_channels,
                           bn_training=True, bn_stats_out=stats)
    for id_, (m, v) in stats.items():
        p[id_ + "_bn.running_mean"] = m
        p[id_ + "_bn.running_var"] = v.clamp_min(1e-3)
    return p


def synth_heads(num_class, feat_mult, feat_dim=1024, seed=0, std=0.001, bias_std=0.0):
    """activity_fc / completeness_fc / regressor_fc, N(0, std) like ssn_models.py:80-89.
    (bias_std > 0 and larger std give the parity tests non-trivial values.)"""
    g = torch.Generator().manual_seed(3000 + seed)
    K = num_class
    h = {}
    for name, o, i in (("activity_fc", K + 1, feat_dim), ("completeness_fc", K, feat_dim * feat_mult),
                       ("regressor_fc", 2 * K, feat_dim * feat_mult)):
        h[name + ".weight"] = torch.randn(o, i, generator=g) * std
        h[name + ".bias"] = torch.randn(o, generator=g) * bias_std
    return h


def synth_batch(n_videos, num_class, in_channels=3, seg=9, size=224, seed=0, props=8):
    """One training batch shaped like SSNDataSet.get_training_data (ssn_dataset.py:455-490):
    per video 1 fg + 6 incomplete + 1 bg (dataset_cfg.yaml:11-14)."""
    g = torch.Generator().manual_seed(4000 + seed)
    frames = synth_frames(n_videos * props * seg, in_channels, size, seed=seed)
    x = frames.view(n_videos, props * seg * in_channels, size, size)
    scaling = torch.rand(n_videos, props, 2, generator=g)
    prop_type = torch.tensor([0, 1, 1, 1, 1, 1, 1, 2]).repeat(n_videos, 1)
    target = torch.randint(1, num_class + 1, (n_videos, props), generator=g)
    target[prop_type == 2] = 0
    reg_target = torch.randn(n_videos, props, 2, generator=g)
    reg_target[prop_type != 0] = 0
    return x, scaling, target, reg_target, prop_type
