"""ExtraAdam (extragradient Adam) — drop-in for ``climategan/optim.py`` (``ExtraAdam`` :199-291, ``Extragradient``
:137-197, ``get_optimizer`` :54-127, ``get_scheduler`` :10-51) with the update fused into ONE kernel launch per
parameter group over flat fp32 buffers (``cgb_extra_adam``) instead of ~10 elementwise launches per parameter tensor.

Parameters of a group are re-pointed to views of one flat buffer and so are their ``.grad`` tensors (autograd
accumulates into the views in place; ``zero_grad`` is one memset).  The flat gradient buffer is also what the
data-parallel all-reduce runs on (``flat_grads``), so no pack/unpack copies exist anywhere in the step.

Known deviation: the fused kernel updates EVERY element of a group with one global step count ``t``.  The
reference skips parameters whose ``.grad`` is None (``tutils.zero_grad`` sets them to None; ``optim.py:243-245``): no moment decay,
no per-parameter ``state['step']`` increment.  A parameter that has never received a gradient behaves identically here (g = 0 and
m = v = 0 give a zero update); one that STARTS receiving gradients late (the painter after ``train.kitti.pretrain``, a decoder
absent from some domains) sees bias corrections computed with the global ``t`` instead of its own count — its first updates are
up to (1 - b1) / sqrt(1 - b2) times larger than the reference's — and one that STOPS receiving gradients keeps moving on its
momentum.  The step fixtures (every parameter gets a gradient on every step) are not affected; a run that switches tasks mid-way is.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim import Optimizer, lr_scheduler

from . import _lib
from ._lib import check


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class ExtraAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is never enabled by the reference configs")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid ExtraAdam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._flat = None
        self._have_copy = False
        self._steps = 0

    # -- flat storage --------------------------------------------------------------------------------
    def _flatten(self):
        flats = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                flats.append(None)
                continue
            dev = ps[0].device
            al = 64   # every parameter starts on a 256-byte boundary: kernels read parameters with 16-byte vector loads
            n = sum((p.numel() + al - 1) // al * al for p in ps)
            fp = torch.zeros(n, dtype=torch.float32, device=dev)   # the alignment gaps stay zero under the update (g = 0)
            fg = torch.zeros(n, dtype=torch.float32, device=dev)
            off = 0
            for p in ps:
                k = p.numel()
                fp[off:off + k].copy_(p.data.reshape(-1))
                p.data = fp[off:off + k].view_as(p)
                if p.grad is not None:
                    fg[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = fg[off:off + k].view_as(p)
                off += (k + al - 1) // al * al
            flats.append(dict(p=fp, g=fg, m=torch.zeros_like(fp), v=torch.zeros_like(fp), c=torch.empty_like(fp), n=n))
        self._flat = flats

    @property
    def flat_grads(self):
        """One flat fp32 gradient buffer per parameter group (the tensors a DDP all-reduce should reduce)."""
        if self._flat is None:
            self._flatten()
        return [f["g"] for f in self._flat if f is not None]

    def zero_grad(self, set_to_none: bool = False):
        if self._flat is None:
            self._flatten()
        for f in self._flat:
            if f is not None:
                f["g"].zero_()

    def _update(self, mode):
        if self._flat is None:
            self._flatten()
        from .ops import invalidate_weight_cache

        invalidate_weight_cache()   # parameters are about to change under autograd's version counter
        self._steps += 1
        lib = _lib.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            b1, b2 = group["betas"]
            check(lib.cgb_extra_adam(_ptr(f["p"]), _ptr(f["g"]), _ptr(f["m"]), _ptr(f["v"]), _ptr(f["c"]), f["n"],
                                     float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                     float(group["weight_decay"]), self._steps, mode,
                                     0 if (mode == 1 or self._have_copy) else 1, st), "extra_adam")

    def extrapolation(self):
        """optim.py:153-170: save the parameters (first extrapolation only) and move to the look-ahead point."""
        self._update(0)
        self._have_copy = True

    def step(self, closure=None):
        """optim.py:172-197: apply the update computed at the look-ahead point to the saved parameters."""
        if not self._have_copy:
            raise RuntimeError("Need to call extrapolation before calling step.")
        loss = closure() if closure is not None else None
        self._update(1)
        self._have_copy = False
        return loss


    # -- checkpointing ----------------------------------------------------------------------------------
    def state_dict(self):
        """Flat moments / look-ahead copy per parameter group.  (Not interchangeable with the reference optimiser's
        per-parameter state; the model state_dicts are.)"""
        if self._flat is None:
            self._flatten()
        return {"steps": self._steps, "have_copy": self._have_copy,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "flat": [None if f is None else {"m": f["m"].clone(), "v": f["v"].clone(), "c": f["c"].clone()} for f in self._flat]}

    def load_state_dict(self, state):
        if self._flat is None:
            self._flatten()
        self._steps, self._have_copy = int(state["steps"]), bool(state["have_copy"])
        for g, sg in zip(self.param_groups, state["param_groups"]):
            g.update(sg)
        for f, sf in zip(self._flat, state["flat"]):
            if f is not None and sf is not None:
                for k in ("m", "v", "c"):
                    f[k].copy_(sf[k])


def get_scheduler(optimizer, hyperparameters, iterations=-1):
    """optim.py:10-51."""
    policy = hyperparameters.get("lr_policy")
    if policy is None or policy == "constant":
        return None
    if policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=hyperparameters.get("lr_step_size"),
                                   gamma=hyperparameters.get("lr_gamma"), last_epoch=iterations)
    if policy == "multi_step":
        milestones = hyperparameters.get("lr_milestones")
        if isinstance(milestones, int):
            last = 1000 if iterations == -1 else iterations
            milestones = list(range(milestones, last, hyperparameters["lr_step_size"]))
        return lr_scheduler.MultiStepLR(optimizer, milestones=milestones, gamma=hyperparameters.get("lr_gamma"),
                                        last_epoch=iterations)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % policy)


def get_optimizer(net, opt_conf, tasks=None, is_disc=False, iterations=-1):
    """optim.py:54-127 for the optimisers built here (ExtraAdam; torch Adam otherwise)."""
    lr_names = []
    lr = opt_conf.lr
    if tasks is None or isinstance(lr, float) or len(lr) == 1:
        lr_default = lr if isinstance(lr, float) else lr.default
        params = list(net.parameters())
        lr_names.append("full")
    else:
        lr_default = lr.default
        params = []
        for task in tasks:
            task_lr = lr.get(task, lr_default)
            parameters = None
            if not is_disc:
                if task == "m":
                    params.append({"params": list(net.encoder.parameters()), "lr": task_lr})
                    lr_names.append("encoder")
                if task == "p":
                    if hasattr(net, "painter"):
                        parameters = list(net.painter.parameters())
                        lr_names.append("painter")
                else:
                    parameters = list(net.decoders[task].parameters())
                    lr_names.append(f"decoder_{task}")
            elif task in net:
                parameters = list(net[task].parameters())
                lr_names.append(f"disc_{task}")
            if parameters is not None:
                params.append({"params": parameters, "lr": task_lr})
    if opt_conf.optimizer.lower() == "extraadam":
        opt = ExtraAdam(params, lr=lr_default, betas=(opt_conf.beta1, 0.999))
    elif opt_conf.optimizer.lower() == "adam":
        opt = torch.optim.Adam(params, lr=lr_default, betas=(opt_conf.beta1, 0.999))
    else:
        raise NotImplementedError(f"optimizer {opt_conf.optimizer} is not built (ExtraAdam / Adam)")
    return opt, get_scheduler(opt, opt_conf, iterations), lr_names
