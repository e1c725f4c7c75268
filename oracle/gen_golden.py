"""Generate tests/golden/*.npz|json from the REAL reference (imports /root/reference).

Run in the build container only:  python -m oracle.gen_golden
The reference ships no golden vectors, so the oracle is pinned against outputs of
the reference itself.  Four monkey-patches are applied from outside, no reference file is edited
(SURVEY.md §8c): yaml.load default Loader; model_zoo.load_url -> None; BNInception.load_state_dict
no-op during construction; Tensor.cuda -> identity (ops/ssn_ops.py hard-codes .cuda()).
"""
import contextlib
import io
import json
import os
import sys
import types
import warnings

import numpy as np
import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    _orig = yaml.load
    yaml.load = lambda s, Loader=yaml.SafeLoader: _orig(s, Loader=Loader)
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    os.chdir(REF)                      # pytorch_load.py:9 uses a relative yaml path
    warnings.filterwarnings("ignore")
    import model_zoo.bninception.pytorch_load as pl
    pl.BNInception.load_state_dict = lambda self, sd, *a, **k: None
    import ssn_models
    import ops.ssn_ops as ssn_ops
    return ssn_models, ssn_ops, pl


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import synth, ssn_oracle as O
    ssn_models, R, pl = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)

    # ---- 1. graph ---------------------------------------------------------------------------
    with contextlib.redirect_stdout(io.StringIO()):
        net = pl.BNInception()
    graph = []
    for (id_, op, out, ins) in net._op_list:
        m = getattr(net, id_, None)
        e = {"id": id_, "op": op, "out": out, "in": ins if isinstance(ins, list) else [ins]}
        if isinstance(m, torch.nn.Conv2d):
            e.update(cin=m.in_channels, cout=m.out_channels, k=m.kernel_size[0], stride=m.stride[0],
                     pad=m.padding[0], bias=m.bias is not None)
        elif isinstance(m, (torch.nn.MaxPool2d, torch.nn.AvgPool2d)) and op == "Pooling":
            e.update(mode="max" if isinstance(m, torch.nn.MaxPool2d) else "ave", k=m.kernel_size,
                     stride=m.stride, pad=m.padding, ceil=m.ceil_mode)
        elif isinstance(m, torch.nn.BatchNorm2d):
            e.update(eps=m.eps, c=m.num_features)
        graph.append(e)
    json.dump({"ops": graph, "state_dict_keys": list(net.state_dict().keys())},
              open(os.path.join(GOLD, "bninception_graph.json"), "w"))

    # ---- 2. STPP ----------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, cfg, seg in (("pyr", (1, (1, 2), 1), (2, 5, 2)), ("flat", [1, 1, 1], (2, 5, 2)),
                          ("seg3", [1, 1, 1], (1, 1, 1)), ("seg3nan", (1, (1, 2), 1), (1, 1, 1)),
                          ("deep", ((1, 2), (1, 2, 4), 2), (4, 8, 4))):
        n, D = 6, 40
        S = sum(seg)
        ft = torch.randn(n * S, D, generator=g, requires_grad=True)
        sc = torch.rand(n, 2, generator=g)
        mod = R.StructuredTemporalPyramidPooling(D, True, configs=cfg)
        a, c = mod(ft, sc, [seg[0], seg[0] + seg[1], S])
        wa, wc = torch.randn(a.shape, generator=g), torch.randn(c.shape, generator=g)
        loss = (a * wa).sum() + (torch.nan_to_num(c) * wc).sum()
        loss.backward()
        out.update({tag + "_ft": ft.detach().numpy(), tag + "_sc": sc.numpy(), tag + "_act": a.detach().numpy(),
                    tag + "_comp": c.detach().numpy(), tag + "_wa": wa.numpy(), tag + "_wc": wc.numpy(),
                    tag + "_dft": ft.grad.numpy(), tag + "_mult": np.int64(mod.feat_multiplier)})
    np.savez_compressed(os.path.join(GOLD, "stpp.npz"), **out)

    # ---- 3. losses --------------------------------------------------------------------------
    out = {}
    K, Bv = 5, 4
    pred = torch.randn(Bv * 7, K, generator=g, requires_grad=True)
    labels = torch.randint(1, K + 1, (Bv * 7,), generator=g)
    for tag, pos, ratio, gs in (("pos", 1, 1.0, 1), ("neg", -1, 0.17, 7), ("half", -1, 0.5, 4)):
        p = pred.detach().clone().requires_grad_(True)
        l = R.OHEMHingeLoss.apply(p, labels, pos, ratio, gs)
        (l * 1.7).sum().backward()
        out.update({"ohem_" + tag + "_loss": l.detach().numpy(), "ohem_" + tag + "_grad": p.grad.numpy()})
    # label 0 wraps to the last class (ops/ssn_ops.py:186 negative index)
    lab0 = labels.clone(); lab0[::3] = 0
    p = pred.detach().clone().requires_grad_(True)
    l = R.OHEMHingeLoss.apply(p, lab0, -1, 0.3, 7)
    l.sum().backward()
    out.update(ohem_wrap_loss=l.detach().numpy(), ohem_wrap_grad=p.grad.numpy(), ohem_wrap_labels=lab0.numpy())
    p = pred.detach().clone().requires_grad_(True)
    cl = R.CompletenessLoss()(p, labels, 1, 7)
    cl.sum().backward()
    out.update(ohem_pred=pred.detach().numpy(), ohem_labels=labels.numpy(), comp_loss=cl.detach().numpy(),
               comp_grad=p.grad.numpy())
    rp = torch.randn(Bv, K, 2, generator=g, requires_grad=True)
    rl = torch.randint(1, K + 1, (Bv,), generator=g)
    rt = torch.randn(Bv, 2, generator=g) * 2
    l = R.ClassWiseRegressionLoss()(rp, rl, rt)
    l.backward()
    out.update(reg_pred=rp.detach().numpy(), reg_labels=rl.numpy(), reg_targets=rt.numpy(),
               reg_loss=l.detach().numpy(), reg_grad=rp.grad.numpy())
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **out)

    # ---- 4. STPPReorgainzed + prepare_test_fc -------------------------------------------------
    out = {}
    for tag, cfg in (("flat", (1, 1, 1)), ("pyr", (1, (1, 2), 1))):
        K, T, N = 3, 37, 24
        mult = sum(O.parse_stage_config(c)[1] for c in cfg)
        D = (K + 1) + mult * K + mult * 2 * K
        scores = torch.randn(T, D, generator=g)
        ticks = torch.sort(torch.randint(-6, T + 8, (N, 4), generator=g), dim=1)[0]
        ticks[0] = torch.tensor([-9, -5, 3, 6]); ticks[1] = torch.tensor([30, 35, 36, 60])
        ticks[2] = torch.tensor([5, 5, 5, 5]); ticks[3] = torch.tensor([0, 0, T, T])
        ticks[:, 1] = ticks[:, 1].clamp(0, T - 1)        # course start must index a real row (:155-160)
        ticks[:, 2] = torch.maximum(ticks[:, 2], ticks[:, 1])
        ticks[:, 3] = torch.maximum(ticks[:, 3], ticks[:, 2])
        sc = torch.rand(N, 2, generator=g)
        st = R.STPPReorgainzed(D, K + 1, K, 2 * K, True, True, stpp_cfg=cfg)
        a, c, r = st.forward(scores, ticks, sc)
        out.update({tag + "_scores": scores.numpy(), tag + "_ticks": ticks.numpy(), tag + "_sc": sc.numpy(),
                    tag + "_act": a.numpy(), tag + "_comp": c.numpy(), tag + "_reg": r.numpy()})
    K, M, D = 3, 5, 16
    fake = types.SimpleNamespace(activity_fc=torch.nn.Linear(D, K + 1), completeness_fc=torch.nn.Linear(D * M, K),
                                 regressor_fc=torch.nn.Linear(D * M, 2 * K), with_regression=True,
                                 stpp=types.SimpleNamespace(feat_multiplier=M))
    ssn_models.SSN.prepare_test_fc(fake)
    for nm in ("activity_fc", "completeness_fc", "regressor_fc"):
        out["tfc_" + nm + "_w"] = getattr(fake, nm).weight.detach().numpy()
        out["tfc_" + nm + "_b"] = getattr(fake, nm).bias.detach().numpy()
    out["tfc_w"], out["tfc_b"] = fake.test_fc.weight.detach().numpy(), fake.test_fc.bias.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "test_path.npz"), **out)

    # ---- 5. whole SSN fwd+bwd on synthetic weights (B=2 videos, K=4) --------------------------
    out = {}
    for modality, C, do_bwd in (("RGB", 3, True), ("Flow", 10, False)):
        K = 4
        with contextlib.redirect_stdout(io.StringIO()):
            model = ssn_models.SSN(K, 2, 5, 2, modality, base_model="BNInception", dropout=0,
                                   stpp_cfg=(1, (1, 2), 1), bn_mode="frozen")
        bb = synth.synth_backbone(C, seed=0)
        hd = synth.synth_heads(K, model.stpp.feat_multiplier, seed=0, std=0.02, bias_std=0.1)
        sd = model.state_dict()
        for k, v in bb.items():
            assert sd["base_model." + k].shape == v.shape, k
            sd["base_model." + k].copy_(v)
        for k, v in hd.items():
            sd[k].copy_(v)
        model.train()
        x, sc, tgt, rtgt, ptype = synth.synth_batch(2, K, C, seed=0)
        outs = model(x, sc, tgt, rtgt, ptype)
        act, act_t, comp, comp_t, reg, reg_l, reg_t = outs
        la = torch.nn.CrossEntropyLoss()(act, act_t)
        lc = R.CompletenessLoss()(comp, comp_t, 1, 7)
        lr = R.ClassWiseRegressionLoss()(reg, reg_l, reg_t)
        loss = la + 0.1 * lc + 0.1 * lr
        t = modality.lower() + "_"
        out.update({t + "act": act.detach().numpy(), t + "act_t": act_t.numpy(), t + "comp": comp.detach().numpy(),
                    t + "comp_t": comp_t.numpy(), t + "reg": reg.detach().numpy(), t + "reg_l": reg_l.numpy(),
                    t + "reg_t": reg_t.numpy(),
                    t + "losses": np.array([la.item(), lc.item(), lr.item(), loss.item()], np.float64)})
        if do_bwd:
            loss.backward()
            names, gsum, gabs = [], [], []
            for n_, p_ in model.named_parameters():
                if p_.grad is None:
                    continue
                names.append(n_); gsum.append(p_.grad.double().sum().item()); gabs.append(p_.grad.double().abs().sum().item())
            out.update({t + "grad_names": np.array(names), t + "grad_sum": np.array(gsum), t + "grad_abs": np.array(gabs),
                        t + "g_conv1_w": model.base_model.conv1_7x7_s2.weight.grad.numpy(),
                        t + "g_conv1_b": model.base_model.conv1_7x7_s2.bias.grad.numpy(),
                        t + "g_3c_3x3_w": model.base_model.inception_3c_3x3.weight.grad[:8].numpy(),
                        t + "g_5b_1x1_w": model.base_model.inception_5b_1x1.weight.grad[:4].numpy(),
                        t + "g_act_w": model.activity_fc.weight.grad.numpy(),
                        t + "g_comp_b": model.completeness_fc.bias.grad.numpy()})
        with torch.no_grad():
            feats = model.base_model(x.view(-1, C, 224, 224)[:18])
        out[t + "base_out18"] = feats.numpy()
    np.savez_compressed(os.path.join(GOLD, "ssn_e2e.npz"), **out)
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
