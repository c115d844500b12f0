"""B200-native drop-in for reference QFA/model.py (class QFA).

Same constructor, attributes, methods, argument meaning and return shapes as the
reference class (file:line cited per method); the per-spectrum Python loop and
the dense n x n linear algebra are replaced by the sm_100a kernels behind the C
ABI of include/qfa_b200.h.  There is NO CPU fallback for the likelihood /
gradient / prediction path: on a non-CUDA device those methods raise QfaError.

State layout: the six trainable tensors live back to back in ONE float32 device
buffer `self._params` = [F | Psi | omega | tau0 | c0 | beta]; `self.F`, ...,
`self.beta` are views into it, so the fused Adam+clip kernel updates all of them
with one launch and `save_to_npz` keeps the reference's .npz layout.

Precision modes (constructor argument `precision`):
  'fp32'  (DEFAULT) float arithmetic on the CUDA cores with the Nh x Nh algebra in double -- the
          reference is a float32 program, this mode is at least as accurate as what it replaces and its
          results do not depend on the batch size.
  'fp64'  double everywhere: the parity mode (<= 1e-5 against the fp64-promoted reference).
  'mixed' opt-in speed mode: above a per-path batch size (predict 1280, train step 800 / 192 spectra) the
          contractions run on the tensor cores with TF32-rounded operands (continuum <= 1e-3); below it the
          'fp32' kernels run.  NOTE: the same model therefore gives slightly different numbers for different
          batch sizes (e.g. the last partial batch of an epoch), and the TF32 error grows with
          cond(I + F^T W F) (high-S/N spectra).  'tf32' = always tensor cores; 'tf32x3' = tensor cores with
          3xTF32 operand splitting (error ~1e-6).
"""
import contextlib
import ctypes
import os
import time
import warnings
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import (QfaError, QfaModelStruct, PRECISIONS, FLAG_ZERO_ACC, FLAG_FORCE_TENSOR, FLAG_SOLVE_FP64,
                   FLAG_TF32X3)
from .utils import default_tau, resolve_tau_law

log2pi = 1.8378770664093453  # reference model.py:20
_KEYS = ("F", "Psi", "omega", "tau0", "c0", "beta")


class PackedDict(dict):
    """dict of tensor views that remembers the packed buffer they alias."""
    packed: Optional[torch.Tensor] = None
    clipped: bool = False
    model = None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class QFA(object):

    def __init__(self, Nb: int, Nr: int, Nh: int, device: torch.device, tau: Callable = default_tau,
                 model_params: Dict[str, np.ndarray] = None, precision: str = "fp32") -> None:
        """reference model.py:26-55 (+ `precision`: 'fp32' (default) | 'fp64' | 'mixed' | 'tf32' | 'tf32x3', see the module
        docstring)."""
        self.Nb = int(Nb)
        self.Nr = int(Nr)
        self.Nh = int(Nh)
        self.device = torch.device(device)
        self.Npix = self.Nb + self.Nr
        self.Nparams = self.Npix * self.Nh + self.Npix + self.Nb + 3
        self.tau = tau
        self.tau_law = resolve_tau_law(tau)
        self.min_value = 1e-3
        self.max_value = 2.
        if precision not in PRECISIONS:
            raise QfaError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        if not 1 <= self.Nh <= 32:
            raise QfaError(f"Nh={Nh} unsupported by the kernels (1..32)")
        self._params = torch.zeros(self.Nparams, dtype=torch.float32, device=self.device)
        self._mu = None
        self._ws = None            # workspace cache (uint8)
        self._acc = None           # accumulation buffer cache
        self._grads = None
        self._loss = None
        self.process_group = None  # set by enable_data_parallel()
        self._dp = False
        self._peer = None
        self.allreduce_kind = "none (one rank)"
        self.solve_fp64 = False    # mixed mode, 8 < Nh <= 32: per-spectrum Cholesky in double instead of float
        self.use_cuda_graph = True  # train(): replay the whole step as ONE captured CUDA graph when the loader allows it
        self._graph = None
        self._graph_key = None
        self._loss_sum = None
        if model_params is not None:
            for k in _KEYS:
                self._view(k).copy_(torch.as_tensor(np.asarray(model_params[k]), dtype=torch.float32))
        else:
            self.random_init_func()

    # ------------------------------------------------------------------ state
    def _offsets(self):
        PH = self.Npix * self.Nh
        return {"F": (0, PH), "Psi": (PH, PH + self.Npix), "omega": (PH + self.Npix, PH + self.Npix + self.Nb),
                "tau0": (self.Nparams - 3, self.Nparams - 2), "c0": (self.Nparams - 2, self.Nparams - 1),
                "beta": (self.Nparams - 1, self.Nparams)}

    def _view_of(self, packed, key):
        a, b = self._offsets()[key]
        v = packed[a:b]
        if key == "F":
            return v.view(self.Npix, self.Nh)
        if key in ("tau0", "c0", "beta"):
            return v.view(())
        return v

    def _view(self, key):
        return self._view_of(self._params, key)

    def _as_dict(self, packed) -> PackedDict:
        d = PackedDict({k: self._view_of(packed, k) for k in _KEYS})
        d.packed = packed
        d.model = self
        return d

    F = property(lambda s: s._view("F"), lambda s, v: s._view("F").copy_(torch.as_tensor(v)))
    Psi = property(lambda s: s._view("Psi"), lambda s, v: s._view("Psi").copy_(torch.as_tensor(v)))
    omega = property(lambda s: s._view("omega"), lambda s, v: s._view("omega").copy_(torch.as_tensor(v)))
    tau0 = property(lambda s: s._view("tau0"), lambda s, v: s._view("tau0").copy_(torch.as_tensor(v)))
    c0 = property(lambda s: s._view("c0"), lambda s, v: s._view("c0").copy_(torch.as_tensor(v)))
    beta = property(lambda s: s._view("beta"), lambda s, v: s._view("beta").copy_(torch.as_tensor(v)))

    @property
    def mu(self):
        return self._mu

    @mu.setter
    def mu(self, v):
        self._mu = None if v is None else torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous()

    def random_init_func(self) -> None:
        """reference model.py:57-72 (the code's constants, not its docstring: quirk Q9)."""
        self.F = torch.rand((self.Npix, self.Nh), dtype=torch.float32) - 0.5
        self._view("Psi").fill_(1.0)
        self._view("omega").fill_(1.0)
        self._view("tau0").fill_(0.02)
        self._view("c0").fill_(0.3)
        self._view("beta").fill_(2.0)

    @property
    def parameters(self):
        """reference model.py:297-306"""
        return self._as_dict(self._params)

    @parameters.setter
    def parameters(self, params_dict):
        """reference model.py:308-316: assign, then clip."""
        packed = getattr(params_dict, "packed", None)
        if packed is not None and packed.data_ptr() == self._params.data_ptr():
            if getattr(params_dict, "clipped", False):
                return                      # fused Adam+clip already did both in place
        else:
            for k in _KEYS:
                self._view(k).copy_(torch.as_tensor(params_dict[k]).to(torch.float32))
        self.clip()

    # ------------------------------------------------------------------ plumbing
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise QfaError("the QFA likelihood/gradient/prediction kernels need a CUDA (sm_100a) device; "
                           "there is no CPU fallback")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _on_device(self):
        """Every C-ABI call runs with the model's device current: the library's launches, its per-device shared-memory
        opt-ins and the stream handle all belong to it (a model on cuda:1 must work while cuda:0 is current)."""
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def _struct(self, need_mu=False):
        if need_mu and self._mu is None:
            raise QfaError("model.mu is not set (load_from_npz or train first)")
        return QfaModelStruct(self.Nb, self.Nr, self.Nh, self.tau_law, self._params.data_ptr(),
                              0 if self._mu is None else self._mu.data_ptr())

    @property
    def _prec(self):
        return PRECISIONS[self.precision]

    @property
    def _flags(self):
        return ((FLAG_FORCE_TENSOR if self.precision in ("tf32", "tf32x3") else 0)
                | (FLAG_SOLVE_FP64 if self.solve_fp64 else 0) | (FLAG_TF32X3 if self.precision == "tf32x3" else 0))

    @property
    def _tdtype(self):
        return torch.float64 if self.precision == "fp64" else torch.float32

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self._ws

    def _prep_inputs(self, x, error, zabs, mask):
        """contiguous float32 / uint8 device tensors with the reference's shapes."""
        dev = self.device

        def f32(t):
            t = torch.as_tensor(t)
            if t.dtype != torch.float32:
                t = t.to(torch.float32)
            return t.to(dev, non_blocking=True).contiguous()
        x, error, zabs = f32(x), f32(error), f32(zabs)
        mask = torch.as_tensor(mask)
        if mask.dtype != torch.bool and mask.dtype != torch.uint8:
            mask = mask != 0
        mask = mask.to(dev, non_blocking=True).contiguous()
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        if x.dim() == 1:
            x, error, zabs, mask = x[None], error[None], zabs[None], mask[None]
        B = x.shape[0]
        if x.shape != (B, self.Npix) or error.shape != (B, self.Npix) or mask.shape != (B, self.Npix) \
                or zabs.shape != (B, self.Nb):
            raise QfaError(f"shape mismatch: expected (B,{self.Npix}) spectra and (B,{self.Nb}) zabs, got "
                           f"{tuple(x.shape)}, {tuple(error.shape)}, {tuple(zabs.shape)}, {tuple(mask.shape)}")
        return x, error, zabs, mask, B

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None, peer_allreduce=True):
        """Shard spectra over ranks (one process per GPU); `acc` is all-reduced before the division.
        peer_allreduce: on CUDA, sum `acc` with the library's ONE-SHOT kernel over peer-mapped memory (qfa_peer_allreduce:
        every rank publishes its accumulator in torch symmetric memory, flags its peers over NVLink and sums the world's
        buffers in rank order -- 8(f) row 4) instead of ncclAllReduce.  If peer memory cannot be set up on every rank the
        step keeps NCCL; `self.allreduce_kind` says which one runs and why."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise QfaError("torch.distributed is not initialised")
        self.process_group = process_group
        self._dp = dist.get_world_size(process_group) > 1
        self._peer = None
        self._graph = None
        self.allreduce_kind = "torch.distributed all_reduce" if self._dp else "none (one rank)"
        if self._dp and peer_allreduce and self.device.type == "cuda":
            self._setup_peer_allreduce()

    def _setup_peer_allreduce(self):
        import torch.distributed as dist
        L = _lib.lib()
        group = self.process_group if self.process_group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = int(L.qfa_acc_len(self.Nb, self.Nr, self.Nh))
        peer, why = None, ""
        try:
            nbytes = int(L.qfa_peer_buffer_bytes(n, PRECISIONS["fp64"], world))      # sized for a double accumulator
            if nbytes == 0:
                raise QfaError(f"world size {world} not supported by qfa_peer_allreduce")
            import torch.distributed._symmetric_memory as symm
            with self._on_device():
                buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
                hdl = symm.rendezvous(buf, group)
                buf.zero_()
                state = torch.zeros(2, dtype=torch.int32, device=self.device)
                torch.cuda.synchronize(self.device)
            peer = dict(buf=buf, hdl=hdl, base_dev=int(hdl.buffer_ptrs_dev), state=state, world=world, rank=rank, n=n)
        except Exception as e:                                   # no peer access, no symmetric-memory support, ...
            why = f"{type(e).__name__}: {e}"
        # every rank must take the same path; the reduction is also the barrier behind the zero fill of the flags
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            self._peer = peer
            self.allreduce_kind = "qfa_peer_allreduce (one-shot over symmetric memory)"
        else:
            self.allreduce_kind = "nccl all_reduce (peer memory unavailable: %s)" % (why or "on another rank")

    def _allreduce(self, acc):
        if not self._dp:
            return
        if self._peer is not None and acc.is_cuda:
            p = self._peer
            if acc.numel() != p["n"]:
                raise QfaError(f"accumulator of {acc.numel()} elements, peer buffers were sized for {p['n']}")
            prec = PRECISIONS["fp64"] if acc.dtype == torch.float64 else PRECISIONS["fp32"]
            with self._on_device():
                _lib.check(_lib.lib().qfa_peer_allreduce(_ptr(acc), acc.numel(), prec, p["base_dev"], _ptr(p["state"]),
                                                         p["world"], p["rank"], self._stream()), "qfa_peer_allreduce")
            return
        import torch.distributed as dist
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.process_group)

    # ------------------------------------------------------------------ hot path
    def accumulate(self, delta, error, zabs, mask, zero=True, nll_out=None):
        """Adds the un-normalised gradient sums, counts and NLL sum of a batch into the `acc`
        buffer (layout: include/qfa_b200.h) and returns it.  Kernel side of model.py:98-103."""
        self._require_cuda()
        delta, error, zabs, mask, B = self._prep_inputs(delta, error, zabs, mask)
        L = _lib.lib()
        n_acc = L.qfa_acc_len(self.Nb, self.Nr, self.Nh)
        if self._acc is None or self._acc.dtype != self._tdtype or self._acc.numel() != n_acc:
            self._acc = torch.zeros(n_acc, dtype=self._tdtype, device=self.device)
            zero = True
        nbytes = L.qfa_train_workspace_bytes(self.Nb, self.Nr, self.Nh, B, self._prec)
        ws = self._workspace(nbytes)
        st = self._struct()
        with self._on_device():
            _lib.check(L.qfa_train_accumulate(ctypes.byref(st), _ptr(delta), _ptr(error), _ptr(zabs), _ptr(mask), B,
                                              _ptr(ws), ws.numel(), _ptr(self._acc), _ptr(nll_out), self._prec,
                                              (FLAG_ZERO_ACC if zero else 0) | self._flags, self._stream()),
                       "qfa_train_accumulate")
        return self._acc

    def finalize(self, acc=None):
        """loss (1,1) and the packed float32 gradient from `acc` (model.py:100,104)."""
        acc = self._acc if acc is None else acc
        L = _lib.lib()
        if self._grads is None:
            self._grads = torch.empty(self.Nparams, dtype=torch.float32, device=self.device)
            self._loss = torch.empty(1, dtype=torch.float32, device=self.device)
        with self._on_device():
            _lib.check(L.qfa_grads_finalize(_ptr(acc), self.Nb, self.Nr, self.Nh, self._prec, _ptr(self._grads),
                                            _ptr(self._loss), self._stream()), "qfa_grads_finalize")
        return self._loss.view(1, 1), self._as_dict(self._grads)

    def forward(self, delta: torch.Tensor, error: torch.Tensor, zabs: torch.Tensor, mask: torch.Tensor):
        """reference model.py:74-105: (batch-mean NLL of shape (1,1), dict of gradients)."""
        acc = self.accumulate(delta, error, zabs, mask, zero=True)
        self._allreduce(acc)
        loss, grads = self.finalize(acc)
        grads.acc = acc
        return loss, grads

    def loglikelihood_and_gradient_for_single_spectra(self, delta, error, zabs, mask):
        """reference model.py:107-158: NLL (1,1) and the un-normalised partials of ONE spectrum."""
        acc = self.accumulate(delta, error, zabs, mask, zero=True).to(torch.float32)
        g = self._as_dict(acc[:self.Nparams].clone())
        o_nll = self.Nparams + self.Npix + 3
        return acc[o_nll].view(1, 1), g

    def predict_batch(self, flux, error, zabs, mask, want=("nll", "hmean", "hcov", "cont", "unc")):
        """Batched model.py:160-180. Returns a dict with the requested outputs:
        nll (B,), hmean (B,Nh), hcov (B,Nh,Nh), cont (B,Npix), unc (B,Npix)."""
        self._require_cuda()
        flux, error, zabs, mask, B = self._prep_inputs(flux, error, zabs, mask)
        dt, dev = self._tdtype, self.device
        out = {"nll": torch.empty(B, dtype=dt, device=dev)}
        if "hmean" in want:
            out["hmean"] = torch.empty(B, self.Nh, dtype=dt, device=dev)
        if "hcov" in want:
            out["hcov"] = torch.empty(B, self.Nh, self.Nh, dtype=dt, device=dev)
        if "cont" in want:
            out["cont"] = torch.empty(B, self.Npix, dtype=dt, device=dev)
        if "unc" in want:
            out["unc"] = torch.empty(B, self.Npix, dtype=dt, device=dev)
        self.predict_into(flux, error, zabs, mask, out)
        return out

    def predict_into(self, flux, error, zabs, mask, out):
        """predict_batch on prepared device tensors writing into preallocated outputs."""
        L = _lib.lib()
        B = flux.shape[0]
        st = self._struct(need_mu=True)
        ws = self._workspace(L.qfa_predict_workspace_bytes(self.Nb, self.Nr, self.Nh, B, self._prec))
        with self._on_device():
            _lib.check(L.qfa_predict(ctypes.byref(st), _ptr(flux), _ptr(error), _ptr(zabs), _ptr(mask), B, _ptr(ws),
                                     ws.numel(), _ptr(out["nll"]), _ptr(out.get("hmean")), _ptr(out.get("hcov")),
                                     _ptr(out.get("cont")), _ptr(out.get("unc")), self._prec, self._flags, self._stream()),
                       "qfa_predict")

    # ------------------------------------------------------------------ host-buffer (end-to-end) paths
    def _host_pipeline(self, arrays, chunk, body):
        """Streams `arrays` (CPU tensors, ideally pinned) through the GPU in chunks of `chunk`
        spectra: H2D on a copy stream, `body(dev_inputs, lo, hi, slot)` on the current stream,
        double-buffered so that copies overlap compute.  Returns the list of per-slot events the
        caller may wait on."""
        dev = self.device
        B = arrays[0].shape[0]
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h2d"):
            self._h2d = torch.cuda.Stream(dev)
            self._d2h = torch.cuda.Stream(dev)
        key = (chunk,) + tuple((a.shape[1:], a.dtype) for a in arrays)
        if getattr(self, "_stage_key", None) != key:
            self._stage = [[torch.empty((chunk,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev) for a in arrays]
                           for _ in range(2)]
            self._stage_key = key
        free = [None, None]          # compute finished reading slot
        for ci, lo in enumerate(range(0, B, chunk)):
            hi = min(lo + chunk, B)
            slot = ci & 1
            with torch.cuda.stream(self._h2d):
                if free[slot] is not None:
                    self._h2d.wait_event(free[slot])
                ins = []
                for a, st in zip(arrays, self._stage[slot]):
                    st[:hi - lo].copy_(a[lo:hi], non_blocking=True)
                    ins.append(st[:hi - lo])
                ready = torch.cuda.Event()
                ready.record(self._h2d)
            cur.wait_event(ready)
            body(ins, lo, hi, slot)
            free[slot] = torch.cuda.Event()
            free[slot].record(cur)

    def forward_host(self, delta, error, zabs, mask, chunk=8192):
        """QFA.forward for HOST tensors: inputs are streamed H2D in chunks (overlapped with the
        kernels), the packed gradient and the loss are copied back D2H.  Returns CPU tensors."""
        self._require_cuda()
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        first = [True]

        def body(ins, lo, hi, slot):
            self.accumulate(ins[0], ins[1], ins[2], ins[3], zero=first[0])
            first[0] = False
        if delta.shape[0] == 0:
            self.accumulate(delta.to(self.device), error.to(self.device), zabs.to(self.device), mask.to(self.device))
        else:
            self._host_pipeline([delta, error, zabs, mask], chunk, body)
        self._allreduce(self._acc)
        loss, grads = self.finalize(self._acc)
        if not hasattr(self, "_host_grads"):
            self._host_grads = torch.empty(self.Nparams + 1, dtype=torch.float32).pin_memory()
        self._host_grads[:self.Nparams].copy_(grads.packed, non_blocking=True)
        self._host_grads[self.Nparams:].copy_(loss.view(1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_grads[self.Nparams:].view(1, 1), self._as_dict(self._host_grads[:self.Nparams])

    def predict_host(self, flux, error, zabs, mask, out=None, want=("nll", "hmean", "hcov", "cont", "unc"),
                     chunk=8192):
        """predict_batch for HOST tensors (reference main.py:94-98 moves every spectrum H2D and every
        result D2H): chunked, double-buffered H2D -> kernel -> D2H pipeline on three streams.
        `out`: optional dict of (pinned) CPU tensors to fill."""
        self._require_cuda()
        B = flux.shape[0]
        dt = self._tdtype
        shapes = {"nll": (B,), "hmean": (B, self.Nh), "hcov": (B, self.Nh, self.Nh), "cont": (B, self.Npix),
                  "unc": (B, self.Npix)}
        names = ["nll"] + [k for k in ("hmean", "hcov", "cont", "unc") if k in want]
        if out is None:
            out = {k: torch.empty(shapes[k], dtype=dt).pin_memory() for k in names}
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        okey = (chunk, dt, tuple(names))
        if getattr(self, "_ostage_key", None) != okey:
            self._ostage = [{k: torch.empty((chunk,) + shapes[k][1:], dtype=dt, device=self.device) for k in names}
                            for _ in range(2)]
            self._ostage_key = okey
        drained = [None, None]
        cur = torch.cuda.current_stream(self.device)

        def body(ins, lo, hi, slot):
            n = hi - lo
            if drained[slot] is not None:
                cur.wait_event(drained[slot])
            o = {k: v[:n] for k, v in self._ostage[slot].items()}
            self.predict_into(ins[0], ins[1], ins[2], ins[3], o)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done)
                for k in names:
                    out[k][lo:hi].copy_(o[k], non_blocking=True)
                drained[slot] = torch.cuda.Event()
                drained[slot].record(self._d2h)
        if B > 0:
            self._host_pipeline([flux, error, zabs, mask], chunk, body)
            self._d2h.synchronize()
        return out

    def nll_batch(self, flux, error, zabs, mask):
        """Per-spectrum NEGATIVE log-likelihood only (likelihood / out-of-distribution scoring)."""
        return self.predict_batch(flux, error, zabs, mask, want=("nll",))["nll"]

    def prediction_for_single_spectra(self, flux, error, zabs, mask):
        """reference model.py:160-180: (nll (1,1), hmean (Nh,1), hcov (Nh,Nh), cont (Npix,), unc (Npix,))."""
        o = self.predict_batch(flux, error, zabs, mask)
        return (o["nll"].view(1, 1), o["hmean"].view(self.Nh, 1), o["hcov"].view(self.Nh, self.Nh),
                o["cont"].view(self.Npix), o["unc"].view(self.Npix))

    # ------------------------------------------------------------------ training loop
    def train(self, optimizer, dataloader, n_epochs, output_dir="./result", save_interval=5, smooth_interval=5,
              quiet=False, logger=None):
        """reference model.py:183-231 (same epoch/smooth/save/early-stop schedule and log line)."""
        rank0 = True
        if self._dp:
            import torch.distributed as dist
            rank0 = dist.get_rank(self.process_group) == 0
        if rank0:
            os.makedirs(output_dir, exist_ok=True)
        output_dir = os.path.join(output_dir, 'checkpoints')
        if rank0:
            os.makedirs(output_dir, exist_ok=True)
        self.mu = torch.as_tensor(np.asarray(dataloader.mu) if not torch.is_tensor(dataloader.mu) else dataloader.mu,
                                  dtype=torch.float32)
        Niter = dataloader.data_size // dataloader.batch_size      # quirk Q8: under-counts a partial batch
        fused = hasattr(optimizer, "update_from_acc")
        graphed = (fused and self.use_cuda_graph and self.device.type == "cuda" and hasattr(dataloader, "graph_batch")
                   and hasattr(optimizer, "update_from_acc_dev") and Niter > 0)

        def step(i):
            dataloader.rewind()
            start_time = time.time()
            if graphed:
                total_loss = self._graphed_epoch(optimizer, dataloader, Niter)
            elif fused:
                total = torch.zeros((), dtype=torch.float64, device=self.device)
                while dataloader.have_next_batch():
                    d, e, z, m = dataloader.next_batch()
                    acc = self.accumulate(d, e, z, m, zero=True)
                    self._allreduce(acc)
                    total += self._loss_from_acc(acc) / Niter   # stays on device: no per-batch sync
                    optimizer.update_from_acc(self, acc)
                total_loss = float(total.item())
            else:
                total_loss = 0.
                while dataloader.have_next_batch():
                    d, e, z, m = dataloader.next_batch()
                    loss, grads = self.forward(d, e, z, m)
                    total_loss += loss.item() / Niter
                    self.parameters = optimizer.update(self.parameters, grads)
            optimizer.step()
            end_time = time.time()
            msg = "epoch: {:03d}/{:03d}  ;  loss:  {:.2f}  ;  time:  {:.2f} s ".format(i, n_epochs, total_loss,
                                                                                    end_time - start_time)
            if not quiet and rank0:
                print(msg)
            if logger is not None and rank0:
                logger.info(msg)
            return total_loss

        for epoch in range(n_epochs):
            loss = step(epoch)
            if loss < 0.:
                self.smooth()
                if rank0:
                    self.save_to_npz(output_dir, 'model_parameters_epoch_%02i.npz' % (epoch + 1))
                break
            if (epoch + 1) % smooth_interval == 0:
                self.smooth()
            if (epoch + 1) % save_interval == 0 and rank0:
                self.save_to_npz(output_dir, 'model_parameters_epoch_%02i.npz' % (epoch + 1))

    def _loss_from_acc(self, acc):
        o = self.Nparams + self.Npix + 3
        return (acc[o] / acc[o + 1]).to(torch.float64)

    # ------------------------------------------------------------------ CUDA-graph train step
    def capture_train_step(self, optimizer, dataloader, Niter):
        """Captures ONE train step -- gather + delta of the next shuffled batch (qfa_gather_prepare, device cursor), the
        accumulation kernels, the all-reduce of `acc` (NCCL, when data parallel), the fused Adam + clip update, the loss
        bookkeeping and the cursor advance -- into a CUDA graph.  Nothing in it takes a host-side argument: the
        epoch-dependent scalars are read from the optimizer's device buffer, so the same graph serves every epoch.
        Reference loop being replaced: model.py:206-215 + optimizer.py:47-52."""
        bufs, cursor = dataloader.graph_batch()
        B = int(bufs[0].shape[0])
        key = (B, self.precision, id(optimizer), id(dataloader), Niter, self._dp)
        if self._graph is not None and self._graph_key == key:
            return self._graph
        if self._loss_sum is None:
            self._loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        optimizer.sync_hyper(self)
        # warm-up outside the capture: allocates workspace / acc and performs the library's one-off per-device set-up
        cur0 = cursor.clone()
        dataloader.graph_fill()
        self.accumulate(*bufs, zero=True)
        cursor.copy_(cur0)
        torch.cuda.synchronize(self.device)
        L = _lib.lib()
        n0 = L.qfa_launch_count()
        g = torch.cuda.CUDAGraph()
        with self._on_device(), torch.cuda.graph(g):
            dataloader.graph_fill()
            acc = self.accumulate(*bufs, zero=True)
            self._allreduce(acc)
            optimizer.update_from_acc_dev(self, acc, self._loss_sum, 1.0 / Niter, cursor, B)
        self.graph_launches_per_step = int(L.qfa_launch_count() - n0)
        self._graph, self._graph_key = g, key
        return g

    def _graphed_epoch(self, optimizer, dataloader, Niter):
        """One epoch with the captured step; a trailing partial batch runs eagerly through the same kernels."""
        g = self.capture_train_step(optimizer, dataloader, Niter)
        optimizer.sync_hyper(self)
        self._loss_sum.zero_()
        n_full = dataloader.graph_full_batches_left()
        for _ in range(n_full):
            g.replay()
        dataloader.graph_advance(n_full)
        while dataloader.have_next_batch():
            d, e, z, m = dataloader.next_batch()
            acc = self.accumulate(d, e, z, m, zero=True)
            self._allreduce(acc)
            optimizer.update_from_acc_dev(self, acc, self._loss_sum, 1.0 / Niter, None, 0)
        return float(self._loss_sum.item())

    # ------------------------------------------------------------------ consumers of the prediction
    def ood_select(self, nll, threshold=None, k=0, cap=None):
        """Out-of-distribution scoring on per-spectrum NLLs (device tensor, e.g. from nll_batch): entirely on the device.
        Returns dict(count = #(nll > threshold) [0-d int tensor], above = indices of up to `cap` of them (sorted),
        top_idx / top_val = the k largest NLLs, descending)."""
        self._require_cuda()
        nll = torch.as_tensor(nll).to(self.device, torch.float32).contiguous()
        B = nll.numel()
        L = _lib.lib()
        dev = self.device
        out = {}
        count = torch.zeros(1, dtype=torch.int32, device=dev) if threshold is not None else None
        cap = (B if cap is None else int(cap)) if threshold is not None else 0
        thr_idx = torch.empty(max(cap, 1), dtype=torch.int32, device=dev) if threshold is not None else None
        k = min(int(k), B)
        top_idx = torch.empty(max(k, 1), dtype=torch.int32, device=dev) if k > 0 else None
        top_val = torch.empty(max(k, 1), dtype=torch.float32, device=dev) if k > 0 else None
        with self._on_device():
            _lib.check(L.qfa_ood_select(_ptr(nll), B, float(threshold) if threshold is not None else 0.0, k, cap,
                                        _ptr(count), _ptr(thr_idx), _ptr(top_idx), _ptr(top_val), self._stream()),
                       "qfa_ood_select")
        if threshold is not None:
            out["count"] = count[0]
            n = torch.clamp(count[0], max=cap)
            out["above"] = torch.sort(thr_idx[:int(n.item())]).values.to(torch.int64)
        if k > 0:
            out["top_idx"], out["top_val"] = top_idx[:k].to(torch.int64), top_val[:k]
        return out

    def sample_posterior(self, hmean, hcov, n_samples=20, seed=0, want=("h", "cont")):
        """nb/predict.ipynb cell 11 on the device, batched: h ~ N(hmean, hcov) and the continuum samples mu + F h.
        hmean (B,Nh), hcov (B,Nh,Nh) as predict_batch returns them.  Returns dict(h (B,S,Nh), cont (B,S,Npix), z)."""
        self._require_cuda()
        hmean = torch.as_tensor(hmean).to(self.device, torch.float32).reshape(-1, self.Nh).contiguous()
        hcov = torch.as_tensor(hcov).to(self.device, torch.float32).reshape(-1, self.Nh, self.Nh).contiguous()
        B, S = hmean.shape[0], int(n_samples)
        out = {"z": torch.empty(B, S, self.Nh, device=self.device), "h": torch.empty(B, S, self.Nh, device=self.device)}
        if "cont" in want:
            out["cont"] = torch.empty(B, S, self.Npix, device=self.device)
        st = self._struct(need_mu="cont" in want)
        with self._on_device():
            _lib.check(_lib.lib().qfa_sample_posterior(ctypes.byref(st), _ptr(hmean), _ptr(hcov), B, S, int(seed),
                                                       _ptr(out["z"]), _ptr(out["h"]), _ptr(out.get("cont")), self._stream()),
                       "qfa_sample_posterior")
        return out

    def predict_to_npz(self, path, flux, error, zabs, mask, names=None, chunk=8192):
        """Batched replacement of the predict loop of reference main.py:94-98 (one .npz PER SPECTRUM with keys ll, hmean,
        hcov, cont, uncertainty): the host arrays stream through predict_host and ONE columnar .npz is written with the
        same five keys, each with a leading spectrum axis (+ `names` if given)."""
        o = self.predict_host(torch.as_tensor(flux), torch.as_tensor(error), torch.as_tensor(zabs), torch.as_tensor(mask),
                              chunk=chunk)
        cols = {"ll": o["nll"].numpy(), "hmean": o["hmean"].numpy(), "hcov": o["hcov"].numpy(), "cont": o["cont"].numpy(),
                "uncertainty": o["unc"].numpy()}
        if names is not None:
            cols["names"] = np.asarray(names)
        np.savez(path, **cols)
        return cols

    # ------------------------------------------------------------------ housekeeping
    def clip(self):
        """reference model.py:233-241"""
        if self.device.type == "cuda":
            L = _lib.lib()
            with self._on_device():
                _lib.check(L.qfa_clip(_ptr(self._params), self.Nb, self.Nr, self.Nh, self.min_value, self.max_value,
                                      self._stream()), "qfa_clip")
        else:   # host-side bookkeeping only (tests of the container logic); not a compute fallback
            self._view("omega").clamp_(self.min_value, self.max_value)
            self._view("Psi").clamp_(self.min_value, self.max_value)
            self._view("tau0").clamp_(0., 1.)
            self._view("beta").clamp_(0.1, 5.)
            self._view("c0").clamp_(-5., 5.)

    def smooth(self):
        """reference model.py:243-252"""
        if self.device.type == "cuda":
            L = _lib.lib()
            out = torch.empty_like(self._params)
            with self._on_device():
                _lib.check(L.qfa_smooth(_ptr(self._params), _ptr(out), self.Nb, self.Nr, self.Nh, self._stream()),
                           "qfa_smooth")
            self._params.copy_(out)
        else:
            import torch.nn.functional as Fn
            om = Fn.avg_pool1d(self.omega.reshape(1, -1), 15, 1, 7, count_include_pad=False).squeeze()
            ps = Fn.avg_pool1d(self.Psi.reshape(1, -1), 15, 1, 7, count_include_pad=False).squeeze()
            Fs = Fn.avg_pool2d(self.F.reshape(1, self.Npix, self.Nh), (31, 1), (1, 1), (15, 0),
                               count_include_pad=False).squeeze(0)
            self.omega, self.Psi, self.F = om, ps, Fs

    def save_to_npz(self, output_dir: str, file_name: str):
        """reference model.py:254-280: keys mu,F,Psi,omega,tau0,c0,beta, all float32 -- plus one extra key `qfa_b200`
        (the ABI version) that marks the file as written by this package: load_from_npz then restores the stored c0
        instead of applying the reference's c0 <- beta quirk, so checkpoints round-trip.  The reference's own loader
        reads by key and ignores the extra entry."""
        mu = self._mu.cpu().detach().numpy()
        host = {k: self._view(k).cpu().detach().numpy() for k in _KEYS}
        if not os.path.exists(output_dir):
            os.mkdir(output_dir)
        np.savez(os.path.join(output_dir, file_name), mu=mu, qfa_b200=np.int32(_lib.ABI_VERSION), **host)

    def load_from_npz(self, path: str, reference_c0_bug: Optional[bool] = None):
        """reference model.py:282-295.  The reference assigns c0 <- file['beta'] (model.py:295, quirk Q1) and its
        shipped golden vector only reproduces with that behaviour.  reference_c0_bug=None (default): files written by
        the REFERENCE (no `qfa_b200` marker, e.g. data/model_parameters.npz) are loaded exactly like the reference loads
        them -- with the quirk, and a warning when it changes the value -- while files written by this package's
        save_to_npz / train() restore their stored c0.  True / False force either behaviour."""
        file = np.load(path)
        if reference_c0_bug is None:
            reference_c0_bug = "qfa_b200" not in file.files
            if reference_c0_bug and float(file["c0"]) != float(file["beta"]):
                warnings.warn(f"{path}: loaded like the reference does (model.py:295): c0 <- beta = {float(file['beta']):.6g}, "
                              f"the stored c0 = {float(file['c0']):.6g} is ignored; pass reference_c0_bug=False to use it")
        self.mu = torch.tensor(file['mu'], dtype=torch.float32)
        for k in ("F", "omega", "Psi", "tau0", "beta"):
            self._view(k).copy_(torch.tensor(file[k], dtype=torch.float32))
        self._view("c0").copy_(torch.tensor(file['beta' if reference_c0_bug else 'c0'], dtype=torch.float32))
