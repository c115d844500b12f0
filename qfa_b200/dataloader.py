"""Device-resident counterpart of reference QFA/dataloader.py (class Dataloader) for arrays that are already in memory.

File reading, catalog filtering and the yacs config of the reference loader are out of scope (host I/O); what this
class mirrors is everything that happens AFTER the spectra are in memory (dataloader.py:95-138), on the device:

    zabs  = (1 + zqso) * wav[:Nb] / 1215.67 - 1                                   dataloader.py:102
    mu    = smooth( sum(flux * exp(+tau) * mask) / sum(flux != -999), 16 )        dataloader.py:110-112
    delta = flux - mu * exp(-tau)       per batch                                 dataloader.py:135-136

and the six members QFA.train uses (model.py:204-211): mu, data_size, batch_size, rewind(), have_next_batch(),
next_batch().  With `rank`/`world` the spectra are sharded by contiguous ranges and the two sums of `mu` are
all-reduced, so every rank holds the same mean spectrum.  Only the Ly-alpha optical depth is
applied (exact for grids that start redward of Ly-beta, like the reference's default 1030-1600 A grid).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .utils import LAW_CONSTANTS, LYA

_LYB = 1025.7222


def _tau(z, which):
    t0, be, C, zn = LAW_CONSTANTS[which]
    return t0 * ((1.0 + z) / zn) ** be + C


def smooth(s: torch.Tensor, window_len: int = 32) -> torch.Tensor:
    """reference utils.py:206-219: reflect-pad box filter (same output length as the reference's slicing gives)."""
    s = s.to(torch.float64)
    w = int(window_len)
    padded = torch.cat([s[1:w].flip(0), s, s[-w:-1].flip(0)])
    y = torch.nn.functional.conv1d(padded.view(1, 1, -1), torch.full((1, 1, w), 1.0 / w, dtype=torch.float64)).view(-1)
    return y[int(w / 2 - 1):-int(w / 2)]


class DeviceDataloader(object):

    def __init__(self, flux, error, zqso, mask, wav_grid, batch_size=500, device="cuda", tau="becker",
                 window_length_for_mu=16, rank=0, world=1, seed=0, shuffle=True, process_group=None):
        dev = torch.device(device)
        wav = np.asarray(wav_grid, dtype=np.float64)
        if wav[0] < _LYB:
            raise NotImplementedError("grid starts blueward of Ly-beta: the multi-series optical depth of "
                                      "reference utils.py:174-203 is not implemented on the device")
        if tau not in LAW_CONSTANTS:
            raise NotImplementedError(f"unknown mean optical depth law {tau!r}")
        self.which = tau
        self.Nb = int(np.sum(wav < LYA))
        self.Npix = len(wav)
        self.Nr = self.Npix - self.Nb
        n = int(np.shape(flux)[0])
        per = n // world
        sl = slice(rank * per, (rank + 1) * per)
        f32 = lambda a: torch.as_tensor(np.asarray(a)[sl] if not torch.is_tensor(a) else a[sl], dtype=torch.float32).to(dev).contiguous()
        self.flux, self.error, self.zqso = f32(flux), f32(error), f32(zqso)
        m = torch.as_tensor(np.asarray(mask)[sl] if not torch.is_tensor(mask) else mask[sl]).to(dev)
        self.mask = (m != 0).contiguous()
        self.wav = torch.tensor(wav, dtype=torch.float32, device=dev)
        self.device = dev
        self.data_size = n                      # global numbers, like synth.SyntheticLoader
        self.batch_size = int(batch_size)
        self.local_batch = max(1, self.batch_size // world)
        self.local_size = per
        self.cur = 0
        self.shuffle = shuffle
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self._perm = torch.arange(per)
        # zabs, dataloader.py:102 (float64 on the host side of the reference, float32 once it reaches the model)
        self.zabs = ((self.zqso.double() + 1.0)[:, None] * self.wav[:self.Nb].double()[None, :] / LYA - 1.0).float().contiguous()
        # mean spectrum, dataloader.py:110-112
        s = torch.ones(per, self.Npix, dtype=torch.float64, device=dev)
        s[:, :self.Nb] = torch.exp(_tau(self.zabs.double(), self.which))
        num = (self.flux.double() * s * self.mask).sum(0)
        den = (self.flux != -999.0).sum(0).double()
        if world > 1:
            import torch.distributed as dist
            both = torch.stack([num, den])
            dist.all_reduce(both, op=dist.ReduceOp.SUM, group=process_group)
            num, den = both[0], both[1]
        self._mu = smooth((num / den).cpu(), window_len=window_length_for_mu).float()
        self._mu_dev = self._mu.to(dev).contiguous()

    @property
    def mu(self):
        return self._mu.numpy()

    def rewind(self):
        self.cur = 0
        if self.shuffle:
            self._perm = torch.randperm(self.local_size, generator=self._gen)

    def have_next_batch(self):
        return self.cur < self.local_size

    def next_batch(self):
        """(delta, error, zabs, mask) of the next batch, all on the device (dataloader.py:124-138)."""
        end = min(self.cur + self.local_batch, self.local_size)
        if self.shuffle:
            ii = self._perm[self.cur:end].to(self.device)
            flux, err, zq, zabs, mask = self.flux[ii], self.error[ii], self.zqso[ii], self.zabs[ii], self.mask[ii]
        else:
            sl = slice(self.cur, end)
            flux, err, zq, zabs, mask = self.flux[sl], self.error[sl], self.zqso[sl], self.zabs[sl], self.mask[sl]
        self.cur = end
        B = flux.shape[0]
        if self.device.type == "cuda":
            delta = torch.empty_like(flux)
            L = _lib.lib()
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            flux = flux.contiguous()
            _lib.check(L.qfa_prepare_batch(p(flux), p(zq.contiguous()), p(self.wav), p(self._mu_dev), B, self.Nb, self.Nr,
                                           _lib.TAU_LAWS[self.which], None, p(delta), st), "qfa_prepare_batch")
        else:   # data preparation is not the hot path: plain torch on a CPU device (tests of the host logic)
            A = torch.ones_like(flux)
            A[:, :self.Nb] = torch.exp(-_tau(zabs, self.which))
            delta = flux - self._mu_dev[None, :] * A
        return delta, err, zabs, mask
