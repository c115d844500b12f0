"""Device-resident counterpart of reference QFA/dataloader.py (class Dataloader) for arrays that are already in memory.

File reading, catalog filtering and the yacs config of the reference loader are out of scope (host I/O); what this
class mirrors is everything that happens AFTER the spectra are in memory (dataloader.py:95-138), on the device:

    zabs  = (1 + zqso) * wav[:Nb] / 1215.67 - 1                                   dataloader.py:102
    mu    = smooth( sum(flux * exp(+tau_total) * mask) / sum(flux != -999), 16 )  dataloader.py:110-112
    delta = flux - mu * exp(-tau_total)     per batch                             dataloader.py:135-136
    rewind(): a fresh permutation of the spectra every epoch                      dataloader.py:154-167

with `tau_total` the multi-series Lyman optical depth of utils.py:174-203 (all 30 lines of Lyman_series.csv, evaluated
inside the kernels; on grids redward of Ly-beta, like the default 1030-1600 A one, it is Ly-alpha only), and the six
members QFA.train uses (model.py:204-211): mu, data_size, batch_size, rewind(), have_next_batch(), next_batch().

Everything per batch is ONE kernel (qfa_gather_prepare): the batch rows are gathered through a device-resident
permutation starting at a device-resident cursor, so a captured CUDA graph of the train step can replay it without any
host-side argument (QFA.train uses that when the loader offers `graph_batch()`).  With `rank`/`world` the spectra are
sharded by contiguous ranges and the two column sums of `mu` are all-reduced, so every rank holds the same mean
spectrum.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .utils import LAW_CONSTANTS, LYA, tau_total


def smooth(s: torch.Tensor, window_len: int = 32) -> torch.Tensor:
    """reference utils.py:206-219: reflect-pad box filter (same output length as the reference's slicing gives)."""
    s = s.to(torch.float64)
    w = int(window_len)
    padded = torch.cat([s[1:w].flip(0), s, s[-w:-1].flip(0)])
    y = torch.nn.functional.conv1d(padded.view(1, 1, -1), torch.full((1, 1, w), 1.0 / w, dtype=torch.float64)).view(-1)
    return y[int(w / 2 - 1):-int(w / 2)]


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class DeviceDataloader(object):

    def __init__(self, flux, error, zqso, mask, wav_grid, batch_size=500, device="cuda", tau="becker",
                 window_length_for_mu=16, rank=0, world=1, seed=0, shuffle=True, process_group=None):
        dev = torch.device(device)
        wav = np.asarray(wav_grid, dtype=np.float64)
        if tau not in LAW_CONSTANTS:
            raise NotImplementedError(f"unknown mean optical depth law {tau!r}")
        self.which = tau
        self.Nb = int(np.sum(wav < LYA))
        self.Npix = len(wav)
        self.Nr = self.Npix - self.Nb
        n = int(np.shape(flux)[0])
        per = n // world
        sl = slice(rank * per, (rank + 1) * per)

        def f32(a):
            return torch.as_tensor(np.asarray(a)[sl] if not torch.is_tensor(a) else a[sl], dtype=torch.float32).to(dev).contiguous()
        self.flux, self.error, self.zqso = f32(flux), f32(error), f32(zqso)
        m = torch.as_tensor(np.asarray(mask)[sl] if not torch.is_tensor(mask) else mask[sl]).to(dev)
        self.mask = (m != 0).contiguous()
        self.wav = torch.tensor(wav, dtype=torch.float32, device=dev)
        self._wav64 = wav
        self.device = dev
        self.data_size = n                      # global numbers, like synth.SyntheticLoader
        self.batch_size = int(batch_size)
        self.local_batch = max(1, self.batch_size // world)
        self.local_size = per
        self.cur = 0
        self.shuffle = shuffle
        self.seed = int(seed)
        self.epoch = -1
        # the permutation and the batch cursor live on the device (int64): one kernel gathers + prepares a batch
        self._gen = torch.Generator(device=dev).manual_seed(self.seed)
        self._perm = torch.arange(per, dtype=torch.int64, device=dev)
        self._cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self._static = None                     # batch buffers handed to a captured graph
        # zabs of the whole shard is never materialised any more: qfa_gather_prepare writes it per batch (dataloader.py:102)
        # mean spectrum, dataloader.py:110-112
        if dev.type == "cuda":
            sums = torch.empty(2 * self.Npix, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().qfa_mean_spectrum_sums(_p(self.flux), _p(self.mask.view(torch.uint8)), _p(self.zqso),
                                                             _p(self.wav), per, self.Nb, self.Nr, _lib.TAU_LAWS[tau],
                                                             _p(sums), self._stream()), "qfa_mean_spectrum_sums")
            both = sums.view(2, self.Npix)
        else:   # host logic only (CPU tests): same arithmetic with torch
            s = torch.ones(per, self.Npix, dtype=torch.float64)
            s[:, :self.Nb] = torch.exp(torch.as_tensor(tau_total(wav, self.zqso.double().numpy(), which=tau)))
            both = torch.stack([(self.flux.double() * s * self.mask).sum(0), (self.flux != -999.0).sum(0).double()])
        if world > 1:
            import torch.distributed as dist
            both = both.contiguous()
            dist.all_reduce(both, op=dist.ReduceOp.SUM, group=process_group)
        self._mu = smooth((both[0] / both[1]).cpu(), window_len=window_length_for_mu).float()
        self._mu_dev = self._mu.to(dev).contiguous()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def mu(self):
        return self._mu.numpy()

    def __len__(self):
        return self.local_size

    def rewind(self):
        """dataloader.py:154-167: reshuffle (a seeded device-side permutation; same on every rank) and reset."""
        self.cur = 0
        self.epoch += 1
        self._cursor.zero_()
        if self.shuffle:
            torch.randperm(self.local_size, generator=self._gen, device=self.device, out=self._perm)

    def have_next_batch(self):
        return self.cur < self.local_size

    # ---- one batch = one kernel
    def _buffers(self, B):
        dev = self.device
        return (torch.empty(B, self.Npix, dtype=torch.float32, device=dev), torch.empty(B, self.Npix, dtype=torch.float32, device=dev),
                torch.empty(B, self.Nb, dtype=torch.float32, device=dev), torch.empty(B, self.Npix, dtype=torch.bool, device=dev))

    def _gather(self, B, bufs, cursor):
        delta, err, zabs, mask = bufs
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().qfa_gather_prepare(
                _p(self.flux), _p(self.error), _p(self.mask.view(torch.uint8)), _p(self.zqso), _p(self.wav), _p(self._mu_dev),
                _p(self._perm), _p(cursor), B, self.Nb, self.Nr, _lib.TAU_LAWS[self.which], _p(zabs), _p(delta), _p(err),
                _p(mask.view(torch.uint8)), self._stream()), "qfa_gather_prepare")

    def next_batch(self):
        """(delta, error, zabs, mask) of the next batch, all on the device (dataloader.py:124-138)."""
        end = min(self.cur + self.local_batch, self.local_size)
        B = end - self.cur
        if self.device.type == "cuda":
            bufs = self._buffers(B)
            self._cursor.fill_(self.cur)
            self._gather(B, bufs, self._cursor)
            self.cur = end
            self._cursor.fill_(end)
            return bufs
        # data preparation is not the hot path: plain torch on a CPU device (tests of the host logic)
        ii = self._perm[self.cur:end]
        self.cur = end
        flux, err, zq, mask = self.flux[ii], self.error[ii], self.zqso[ii], self.mask[ii]
        zabs = ((zq.double() + 1.0)[:, None] * torch.as_tensor(self._wav64[:self.Nb])[None, :] / LYA - 1.0).float()
        A = torch.ones(B, self.Npix, dtype=torch.float64)
        A[:, :self.Nb] = torch.exp(-torch.as_tensor(tau_total(self._wav64, zq.double().numpy(), which=self.which)))
        delta = (flux.double() - self._mu_dev.double()[None, :] * A).float()
        return delta, err, zabs, mask

    # ---- CUDA-graph protocol (used by QFA.train): static batch buffers, device cursor, no host argument per step
    def graph_batch(self):
        """Static (delta, error, zabs, mask) buffers of one full local batch + the device cursor.  `graph_fill()` (captured
        into the graph by the caller) gathers rows perm[cursor .. cursor + B) into them; the captured update kernel advances
        the cursor; `graph_advance()` keeps the host-side position in step."""
        if self._static is None:
            self._static = self._buffers(self.local_batch)
        return self._static, self._cursor

    def graph_fill(self):
        self._gather(self.local_batch, self._static, self._cursor)

    def graph_full_batches_left(self):
        return (self.local_size - self.cur) // self.local_batch

    def graph_advance(self, nsteps=1):
        self.cur += nsteps * self.local_batch
