"""Fused SGD-momentum over flat fp32 buffers (replaces torch.optim.SGD as configured by ssn_train.py:141-144 with the
parameter groups of SSN.get_optim_policies and the per-group lr_mult / decay_mult of adjust_learning_rate, :391-398).

All parameters live in ONE flat buffer (p.data are views), all gradients in another (p.grad are views: one ncclAllReduce
exchanges them, ssn_b200/dp.py), the momentum in a third; a step is ONE kernel launch (ssnb_sgd_step_groups) instead of
~6 foreach launches per group.  Same arithmetic as torch.optim.SGD(momentum, weight_decay, dampening=0, nesterov=False)."""
import ctypes as C

import torch

from ._lib import lib, check


class FusedSGD:
    def __init__(self, policies, lr, momentum=0.9, weight_decay=5e-4, on_step=None, order=None):
        """policies: list of dicts with 'params', 'lr_mult', 'decay_mult' (SSN.get_optim_policies()).  on_step: callables run
        after every step (e.g. BNInception.invalidate_packed: the kernels' packed weight copies are stale).  order: the
        parameters in the order they should occupy the flat buffers (default: group by group; model.parameters() order
        puts each layer's weight and bias side by side, which ssn_b200.dp.GradSync's buckets need)."""
        self.param_groups = []
        for g in policies:
            ps = [p for p in g["params"] if p.requires_grad]
            if ps:
                d = dict(g)
                d["params"] = ps
                d["lr"] = lr * g.get("lr_mult", 1)
                d["weight_decay"] = weight_decay * g.get("decay_mult", 1)
                self.param_groups.append(d)
        self.momentum = float(momentum)
        self.on_step = list(on_step or [])
        params = [p for g in self.param_groups for p in g["params"]]
        self._group_of = {id(p): g for g in self.param_groups for p in g["params"]}
        if order is not None:
            order = [p for p in order if id(p) in self._group_of]
            assert len(order) == len(params), "order must list exactly the parameters of the groups"
            params = order
        self.params = params
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), "FusedSGD: fp32 CUDA parameters"
        dev = params[0].device
        self.device = dev
        n = sum(p.numel() for p in params)
        self.flat_param = torch.empty(n, device=dev)
        self.flat_grad = torch.zeros(n, device=dev)
        self.flat_mom = torch.zeros(n, device=dev)
        self.views = []
        off = 0
        for p in params:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            self.views.append((p, off, k))
            off += k
        self.n = n
        self._seg_end = torch.tensor([o + k for (_p, o, k) in self.views], dtype=torch.int64, device=dev)
        self._seg_lr = torch.empty(len(self.views), device=dev)
        self._seg_wd = torch.empty(len(self.views), device=dev)
        self.refresh_groups()

    def refresh_groups(self):
        """re-read lr / weight_decay of every group (call after adjust_learning_rate changed them)"""
        lr = [float(self._group_of[id(p)]["lr"]) for p in self.params]
        wd = [float(self._group_of[id(p)]["weight_decay"]) for p in self.params]
        self._seg_lr.copy_(torch.tensor(lr)); self._seg_wd.copy_(torch.tensor(wd))

    def rebind_grads(self):
        """p.grad must alias the flat gradient buffer (zero_grad(set_to_none=True) or an assignment breaks it): re-attach,
        keeping whatever gradient the parameter holds"""
        for p, off, k in self.views:
            view = self.flat_grad[off:off + k].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        self.rebind_grads()

    def step(self, grad_mult=1.0):
        self.rebind_grads()
        with torch.cuda.device(self.device):
            check(lib.ssnb_sgd_step_groups(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.flat_mom.data_ptr(), self.n,
                                           self._seg_end.data_ptr(), self._seg_lr.data_ptr(), self._seg_wd.data_ptr(), len(self.views),
                                           self.momentum, float(grad_mult), C.c_void_p(torch.cuda.current_stream().cuda_stream)), None, "sgd_step_groups")
        for f in self.on_step:
            f()
