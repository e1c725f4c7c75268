"""BNInception with the reference's module surface (model_zoo/bninception/pytorch_load.py:8-61),
executed by libssn_b200's engine instead of an op-by-op PyTorch interpreter.

The nn.Conv2d / nn.BatchNorm2d children exist as parameter holders only: they give the module the
reference's state_dict keys (`<yaml id>.weight|bias`, `<id>_bn.weight|bias|running_mean|
running_var`, `fc.*`) and make SSN.get_optim_policies (ssn_models.py:203-251) work unchanged; their
own forward() is never called.  forward(x[N,C,224,224]) -> fc(global_pool features [N,1024]).
"""
import torch
from torch import nn

from ssn_b200 import _lib
from ssn_b200.engine import BackboneEngine, BackboneFunction, conv_table


class BNInception(nn.Module):
    def __init__(self, model_path=None, num_classes=101, weight_url=None, in_channels=3):
        super(BNInception, self).__init__()
        # model_path / weight_url are accepted for signature compatibility; the graph is built into
        # the library and there is no network access for pretrained weights.
        self._conv_names = []
        for (name, cin, cout, k, stride, pad) in conv_table(in_channels):
            setattr(self, name, nn.Conv2d(cin, cout, k, stride, pad, bias=True))
            setattr(self, name + "_bn", nn.BatchNorm2d(cout, momentum=0.1))
            self._conv_names.append(name)
        self.fc = nn.Linear(1024, 1000)
        self.last_layer_name = "fc"
        self.precision = _lib.EXACT_FP32
        self.grad_scale = 1.0
        self._engines = {}

    # ---- engine management --------------------------------------------------------------------
    def set_precision(self, precision, grad_scale=None):
        """precision: ssn_b200.EXACT_FP32 (fp32 SIMT), ssn_b200.FAST_FP16 (tcgen05, fp16 operands) or
        ssn_b200.EXACT_TC (tcgen05, error-compensated split fp16 operands: fp32-grade results)."""
        if precision == _lib.FAST_FP16 and grad_scale is None and self.grad_scale == 1.0:
            raise ValueError("FAST_FP16 stores gradients in fp16: pass an explicit power-of-two grad_scale (e.g. 4096) so small "
                             "gradients do not underflow, and poll engine.grad_overflow() to catch overflow")
        self.precision = precision
        if grad_scale is not None:
            self.grad_scale = float(grad_scale)
        self._engines = {}

    def _convs(self):
        return [getattr(self, n) for n in self._conv_names]

    def _bns(self):
        return [getattr(self, n + "_bn") for n in self._conv_names]

    def in_channels(self):
        return getattr(self, self._conv_names[0]).in_channels

    def grad_overflow(self, clear=True):
        """True when a gradient left the fp16 range under grad_scale in any engine since the last call (device sync)."""
        return any([e.grad_overflow(clear) for e in self._engines.values()])

    def _weights_version(self):
        v = 0
        for c, b in zip(self._convs(), self._bns()):
            v += c.weight._version + c.bias._version + b.weight._version + b.bias._version \
                + b.running_mean._version + b.running_var._version
        return (v, id(self._convs()[0].weight), self._convs()[0].weight.data_ptr())

    def invalidate_packed(self):
        """The kernels read BN-folded, re-laid-out copies of the weights.  They are refreshed automatically when a parameter's
        Tensor._version moves (optimizer.step(), in-place ops); writes that bypass the version counter -- `p.data.copy_()`,
        the fused SGD kernel, a raw pointer -- need this call."""
        for eng in self._engines.values():
            eng.packed_version = None

    def bn1_training(self):
        """bn_mode='partial' (ssn_models.py:95-105): the first BatchNorm2d stays in training mode, every other one is frozen.
        Returns True for that pattern, False when all are frozen, raises for anything else ('full')."""
        bns = self._bns()
        if any(b.training for b in bns[1:]):
            raise NotImplementedError("only bn_mode='frozen' and 'partial' are accelerated: BatchNorm2d layers after the first must be in "
                                      "eval mode ('full' is listed as a next step in DESIGN.md)")
        return bool(bns[0].training)

    def engine_for(self, frames, training, device, bn1_train=False):
        if bn1_train and self.precision == _lib.FAST_FP16:
            raise NotImplementedError("bn_mode='partial' runs in EXACT_FP32 / EXACT_TC precision (fp32 activations), not FAST_FP16")
        key = (frames, bool(training), self.precision, self.in_channels(), str(device), bool(bn1_train))
        eng = self._engines.get(key)
        if eng is None:
            eng = BackboneEngine(self.in_channels(), frames, self.precision, training, self.grad_scale, device, bn1_train=bn1_train)
            self._engines[key] = eng
        ver = self._weights_version()
        if eng.packed_version != ver:
            cs, bs = self._convs(), self._bns()
            eng.pack([c.weight.data for c in cs], [c.bias.data for c in cs], [b.weight.data for b in bs],
                     [b.bias.data for b in bs], [b.running_mean for b in bs], [b.running_var for b in bs])
            eng.packed_version = ver
        return eng

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("BNInception(B200) runs on CUDA only: move the model and input to the GPU "
                               "(libssn_b200 has no CPU path)")
        bn1_train = self.bn1_training()
        cs = self._convs()
        params = [c.weight for c in cs] + [c.bias for c in cs]
        bn1 = self._bns()[0]
        if bn1_train:
            params += [bn1.weight, bn1.bias]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        eng = self.engine_for(input.shape[0], need_grad, input.device, bn1_train)
        if bn1_train:
            eng.set_bn1(bn1)                         # batch statistics + running-stat update happen inside the forward
            if bn1.num_batches_tracked is not None:
                bn1.num_batches_tracked.add_(1)
        if need_grad:
            feat = BackboneFunction.apply(input, eng, len(cs), bn1 if bn1_train else None, *params)
        else:
            feat = eng.forward(input)
        return self.fc(feat)
