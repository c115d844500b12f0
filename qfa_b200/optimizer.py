"""B200-native drop-in for reference QFA/optimizer.py (Adam, step_scheduler).

`Adam.update(params, g)` keeps the reference's dict-in / dict-out signature
(optimizer.py:37-52).  When `params` and `g` are views of packed buffers handed
out by qfa_b200.model.QFA (the normal case) the whole update -- L2 term, moment
EMAs, bias correction by the EPOCH counter (quirk Q8), parameter step and the
clipping of model.py:233-241 -- is ONE kernel launch (qfa_adam_clip_step).
Dicts of ordinary tensors take a torch-op path with the same arithmetic.
"""
import ctypes
from typing import Callable, Dict

import torch

from . import _lib
from .model import PackedDict, _ptr, _KEYS


class Adam(object):

    def __init__(self, params: Dict[str, torch.Tensor], device: torch.device, scheduler=None,
                 learning_rate: float = 1e-2, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
                 weight_decay: float = 1e-3) -> None:
        """reference optimizer.py:13-35"""
        self.learning_rate = learning_rate
        self.b1 = b1
        self.b2 = b2
        self.eps = eps
        self.device = torch.device(device)
        self.weight_decay = weight_decay
        self.scheduler = scheduler
        self.reset(params)

    def reset(self, params):
        """reference optimizer.py:54-63"""
        packed = getattr(params, "packed", None)
        if packed is not None:
            self._m = torch.zeros_like(packed, dtype=torch.float32, device=self.device)
            self._v = torch.zeros_like(packed, dtype=torch.float32, device=self.device)
            self.m = self._views(params, self._m)
            self.v = self._views(params, self._v)
        else:
            self._m = self._v = None
            self.m = {k: torch.zeros_like(params[k], dtype=torch.float32).to(self.device) for k in params}
            self.v = {k: torch.zeros_like(params[k], dtype=torch.float32).to(self.device) for k in params}
        self.i = 0

    @staticmethod
    def _views(params, packed):
        """views of `packed` with the same offsets/shapes as the views in `params`."""
        base = params.packed.data_ptr()
        out = {}
        for k, v in params.items():
            off = (v.data_ptr() - base) // 4
            out[k] = packed[off:off + v.numel()].view(v.shape)
        return out

    def step(self):
        """reference optimizer.py:65-69"""
        self.i += 1

    @property
    def scheduled_lr(self):
        """reference optimizer.py:71-76"""
        if callable(self.scheduler):
            return self.scheduler(self.i, self.learning_rate)
        return self.learning_rate

    # -- fused path ----------------------------------------------------------
    def _fused(self, model, params_packed, acc=None, grads_packed=None):
        L = _lib.lib()
        bias1 = 1. - self.b1 ** (self.i + 1)
        bias2 = 1. - self.b2 ** (self.i + 1)
        st = ctypes.c_void_p(torch.cuda.current_stream(params_packed.device).cuda_stream)
        with torch.cuda.device(params_packed.device):
            _lib.check(L.qfa_adam_clip_step(_ptr(params_packed), _ptr(self._m), _ptr(self._v), _ptr(acc),
                                            _ptr(grads_packed), model.Nb, model.Nr, model.Nh, model._prec,
                                            self.scheduled_lr, self.b1, self.b2, self.eps, self.weight_decay,
                                            bias1, bias2, model.min_value, model.max_value, st),
                       "qfa_adam_clip_step")

    def update_from_acc(self, model, acc):
        """forward -> (all-reduce) -> update without materialising the gradient:
        reads sum/count straight from the accumulation buffer."""
        if self._m is None:
            self.reset(model.parameters)
        self._fused(model, model._params, acc=acc)

    def sync_hyper(self, model):
        """Device copy of the epoch-dependent scalars {lr_i, 1 - b1^(i+1), 1 - b2^(i+1)} (optimizer.py:50-52,98) that
        qfa_adam_clip_step_dev reads: refreshed once per epoch, so a captured CUDA graph never goes stale."""
        if self._m is None:
            self.reset(model.parameters)
        vals = [self.scheduled_lr, 1. - self.b1 ** (self.i + 1), 1. - self.b2 ** (self.i + 1)]
        if getattr(self, "_hyper", None) is None or self._hyper.device != model._params.device:
            self._hyper = torch.zeros(3, dtype=torch.float32, device=model._params.device)
        self._hyper.copy_(torch.tensor(vals, dtype=torch.float32))
        return self._hyper

    def update_from_acc_dev(self, model, acc, loss_sum=None, loss_scale=0.0, cursor=None, cursor_step=0):
        """update_from_acc with the scalars read from the device (graph-capturable); optionally accumulates the step's
        mean NLL * loss_scale into `loss_sum` (double[1]) and advances the loader's device cursor."""
        if getattr(self, "_hyper", None) is None:
            self.sync_hyper(model)
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream(model._params.device).cuda_stream)
        with torch.cuda.device(model._params.device):
            _lib.check(L.qfa_adam_clip_step_dev(_ptr(model._params), _ptr(self._m), _ptr(self._v), _ptr(acc), model.Nb,
                                                model.Nr, model.Nh, model._prec, _ptr(self._hyper), self.b1, self.b2,
                                                self.eps, self.weight_decay, model.min_value, model.max_value,
                                                _ptr(loss_sum), float(loss_scale), _ptr(cursor), int(cursor_step), st),
                       "qfa_adam_clip_step_dev")

    def update(self, params, g):
        """reference optimizer.py:37-52 (functional signature: returns the updated dict)."""
        pp, gp = getattr(params, "packed", None), getattr(g, "packed", None)
        model = getattr(params, "model", None)
        if (pp is not None and gp is not None and self._m is not None and pp.is_cuda and model is not None):
            self._fused(model, pp, grads_packed=gp)
            out = PackedDict(params)
            out.packed, out.clipped = pp, True
            out.model = model
            return out
        # dicts of ordinary tensors (API compatibility; the hot path is the fused kernel above): same arithmetic as
        # oracle.qfa_dense.adam_update -- L2 term, moment EMAs, bias correction by the epoch counter, step
        t = self.i + 1
        c1, c2 = 1. - self.b1 ** t, 1. - self.b2 ** t
        lr = self.scheduled_lr
        out = {}
        for name, p in params.items():
            grad = g[name] + self.weight_decay * p
            self.m[name] = (1. - self.b1) * grad + self.b1 * self.m[name]
            self.v[name] = (1. - self.b2) * (grad * grad) + self.b2 * self.v[name]
            out[name] = p - lr * (self.m[name] / c1) / (torch.sqrt(self.v[name] / c2) + self.eps)
        return out


def step_scheduler(alpha: float, step: int) -> Callable[[int, float], float]:
    """reference optimizer.py:79-98: lr_i = lr * alpha ** ((i+1)//step)"""
    def scheduler(i, lr):
        return lr * alpha ** ((i + 1) // step)
    return scheduler
