"""Pin the CPU oracle (oracle/ssn_oracle.py) against golden vectors produced by the real reference
(oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ssn_oracle as O
from oracle import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_graph_matches_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "bninception_graph.json")))
    ref_ops = g["ops"]
    mine = O.bninception_ops(3)
    # expand my (conv = conv+bn+relu) ops into the reference's sequence
    seq = []
    for kind, id_, out, ins, a in mine:
        if kind == "conv":
            seq.append(("Convolution", id_, a))
            seq.append(("BN", id_ + "_bn", None))
            seq.append(("ReLU", None, None))
        elif kind == "pool":
            seq.append(("Pooling", id_, a))
        elif kind == "concat":
            seq.append(("Concat", id_, ins))
        else:
            seq.append(("InnerProduct", id_, a))
    assert len(seq) == len(ref_ops) == 231
    for (op, id_, a), r in zip(seq, ref_ops):
        assert op == r["op"]
        if op == "Convolution":
            assert id_ == r["id"] and r["bias"] is True
            assert (a["cin"], a["cout"], a["k"], a["stride"], a["pad"]) == (r["cin"], r["cout"], r["k"], r["stride"], r["pad"])
        elif op == "BN":
            assert id_ == r["id"] and r["eps"] == 1e-5
        elif op == "Pooling":
            assert id_ == r["id"] and r["ceil"] is True
            assert (a["mode"], a["k"], a["stride"], a["pad"]) == (r["mode"], r["k"], r["stride"], r["pad"])
        elif op == "Concat":
            assert a == r["in"]
    keys = [k for k in g["state_dict_keys"] if not k.startswith("fc.") and "num_batches_tracked" not in k]
    assert keys == O.backbone_param_names(3)


@pytest.mark.parametrize("tag,cfg,seg", [("pyr", (1, (1, 2), 1), (2, 5, 2)), ("flat", [1, 1, 1], (2, 5, 2)),
                                         ("seg3", [1, 1, 1], (1, 1, 1)), ("seg3nan", (1, (1, 2), 1), (1, 1, 1)),
                                         ("deep", ((1, 2), (1, 2, 4), 2), (4, 8, 4))])
def test_stpp(golden_dir, tag, cfg, seg):
    z = _load(golden_dir, "stpp.npz")
    ft = torch.tensor(z[tag + "_ft"], requires_grad=True)
    S = sum(seg)
    a, c = O.stpp_forward(ft, torch.tensor(z[tag + "_sc"]), [seg[0], seg[0] + seg[1], S], cfg)
    np.testing.assert_array_equal(a.detach().numpy(), z[tag + "_act"])
    np.testing.assert_array_equal(c.detach().numpy(), z[tag + "_comp"])     # NaN positions included
    assert c.shape[1] == int(z[tag + "_mult"]) * ft.shape[1]
    loss = (a * torch.tensor(z[tag + "_wa"])).sum() + (torch.nan_to_num(c) * torch.tensor(z[tag + "_wc"])).sum()
    loss.backward()
    np.testing.assert_allclose(ft.grad.numpy(), z[tag + "_dft"], rtol=1e-6, atol=1e-7)


def test_losses(golden_dir):
    z = _load(golden_dir, "losses.npz")
    pred = torch.tensor(z["ohem_pred"])
    labels = torch.tensor(z["ohem_labels"])
    for tag, pos, ratio, gs in (("pos", 1, 1.0, 1), ("neg", -1, 0.17, 7), ("half", -1, 0.5, 4)):
        p = pred.clone().requires_grad_(True)
        l = O.OHEMHingeLoss.apply(p, labels, pos, ratio, gs)
        (l * 1.7).sum().backward()
        np.testing.assert_allclose(l.detach().numpy(), z["ohem_" + tag + "_loss"], rtol=1e-6)
        np.testing.assert_allclose(p.grad.numpy(), z["ohem_" + tag + "_grad"], rtol=1e-6)
    p = pred.clone().requires_grad_(True)
    l = O.OHEMHingeLoss.apply(p, torch.tensor(z["ohem_wrap_labels"]), -1, 0.3, 7)
    l.sum().backward()
    np.testing.assert_allclose(l.detach().numpy(), z["ohem_wrap_loss"], rtol=1e-6)
    np.testing.assert_array_equal(p.grad.numpy(), z["ohem_wrap_grad"])
    p = pred.clone().requires_grad_(True)
    cl = O.completeness_loss(p, labels, 1, 7)
    cl.sum().backward()
    np.testing.assert_allclose(cl.detach().numpy(), z["comp_loss"], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), z["comp_grad"], rtol=1e-6)
    rp = torch.tensor(z["reg_pred"], requires_grad=True)
    l = O.classwise_regression_loss(rp, torch.tensor(z["reg_labels"]), torch.tensor(z["reg_targets"]))
    l.backward()
    np.testing.assert_allclose(l.item(), z["reg_loss"], rtol=1e-6)
    np.testing.assert_allclose(rp.grad.numpy(), z["reg_grad"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("tag,cfg", [("flat", (1, 1, 1)), ("pyr", (1, (1, 2), 1))])
def test_stpp_reorganized(golden_dir, tag, cfg):
    z = _load(golden_dir, "test_path.npz")
    K = 3
    a, c, r = O.stpp_reorganized(torch.tensor(z[tag + "_scores"]), torch.tensor(z[tag + "_ticks"]),
                                 torch.tensor(z[tag + "_sc"]), K + 1, K, 2 * K, cfg)
    np.testing.assert_allclose(a.numpy(), z[tag + "_act"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(c.numpy(), z[tag + "_comp"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(r.numpy(), z[tag + "_reg"], rtol=1e-6, atol=1e-7)


def test_prepare_test_fc(golden_dir):
    z = _load(golden_dir, "test_path.npz")
    head = {}
    for nm in ("activity_fc", "completeness_fc", "regressor_fc"):
        head[nm + ".weight"] = torch.tensor(z["tfc_" + nm + "_w"])
        head[nm + ".bias"] = torch.tensor(z["tfc_" + nm + "_b"])
    w, b = O.prepare_test_fc(head, 5)
    np.testing.assert_array_equal(w.numpy(), z["tfc_w"])
    np.testing.assert_array_equal(b.numpy(), z["tfc_b"])


def _e2e(golden_dir, modality, C, do_bwd):
    z = _load(golden_dir, "ssn_e2e.npz")
    t = modality.lower() + "_"
    K = 4
    bb = synth.synth_backbone(C, seed=0)
    hd = synth.synth_heads(K, 5, seed=0, std=0.02, bias_std=0.1)
    if do_bwd:
        for d in (bb, hd):
            for k in d:
                if "running" not in k and "_bn." not in k:
                    d[k].requires_grad_(True)
    x, sc, tgt, rtgt, ptype = synth.synth_batch(2, K, C, seed=0)
    outs = O.ssn_train_forward(bb, hd, x, sc, tgt, rtgt, ptype, in_channels=C)
    loss, (la, lc, lr) = O.total_loss(outs)
    for name, o in zip(("act", "act_t", "comp", "comp_t", "reg", "reg_l", "reg_t"), outs):
        ref = z[t + name]
        if ref.dtype.kind == "f":
            np.testing.assert_allclose(o.detach().numpy(), ref, rtol=2e-4, atol=2e-5)
        else:
            np.testing.assert_array_equal(o.numpy(), ref)
    np.testing.assert_allclose([la.item(), lc.item(), lr.item(), loss.item()], z[t + "losses"], rtol=2e-5)
    if do_bwd:
        loss.backward()
        np.testing.assert_allclose(bb["conv1_7x7_s2.weight"].grad.numpy(), z[t + "g_conv1_w"], rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(bb["inception_3c_3x3.weight"].grad[:8].numpy(), z[t + "g_3c_3x3_w"], rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(hd["activity_fc.weight"].grad.numpy(), z[t + "g_act_w"], rtol=2e-4, atol=1e-7)
        names = [str(s) for s in z[t + "grad_names"]]
        for n_, gs, ga in zip(names, z[t + "grad_sum"], z[t + "grad_abs"]):
            p = bb[n_[len("base_model."):]] if n_.startswith("base_model.") else hd[n_]
            assert abs(p.grad.double().abs().sum().item() - ga) <= 2e-4 * ga + 1e-9, n_


def test_ssn_e2e_rgb(golden_dir):
    _e2e(golden_dir, "RGB", 3, True)


def test_ssn_e2e_flow_forward(golden_dir):
    _e2e(golden_dir, "Flow", 10, False)


def test_detect_oracle_matches_reference(golden_dir):
    """oracle/detect_oracle.py against tests/golden/detect.npz (real reference: ops.utils.softmax / temporal_nms and
    eval_detection_results.perform_regression): combined scores, NMS survivors and regressed detections, bit for bit."""
    from oracle import detect_oracle as D
    z = np.load(os.path.join(golden_dir, "detect.npz"))
    for tag in "abcd":
        props, act, comp, reg = z[tag + "_props"], z[tag + "_act"], z[tag + "_comp"], z[tag + "_reg"]
        thr = float(z[tag + "_thr"])
        K = comp.shape[1]
        np.testing.assert_array_equal(D.softmax(act)[:, 1:] * np.exp(comp), z[tag + "_combined"])
        dets = D.video_detections(props, act, comp, reg, thr)
        raw = D.video_detections(props, act, comp, reg, thr, regress=False)
        for c in range(K):
            np.testing.assert_array_equal(raw[c], z["%s_nms_%d" % (tag, c)])
            np.testing.assert_array_equal(dets[c], z["%s_det_%d" % (tag, c)])
