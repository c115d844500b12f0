"""Seeded synthetic SDSS / DESI / L32-shaped spectra drawn from the QFA generative model
(SURVEY.md section 8d).  Pure torch; runs on the CPU (tests, CPU baseline) or directly on
the GPU (benchmarks: the dataset is created resident in HBM).

    S = A(z) * (mu + F h + sqrt(Psi) xi) + sqrt(omega * zdep(z)) xi_blue + sigma xi
(reference README / model.py:125-131), sigma = abs(mu)/snr * U(0.8,1.2), snr ~ U(2,20),
zq ~ U(2,3.5) (reference config.py:35-36), zabs as reference dataloader.py:102,
mask = i.i.d. fraction + contiguous runs; masked pixels are set to -999 in flux and
error exactly like the reference's files (dataloader.py:26-28).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from .utils import LAW_CONSTANTS, LYA


@dataclass
class GridSpec:
    name: str
    lam_min: float
    dloglam: float
    Npix: int

    def wav(self):
        return self.lam_min * 10 ** (self.dloglam * np.arange(self.Npix))

    @property
    def Nb(self):
        return int(np.sum(self.wav() < LYA))

    @property
    def Nr(self):
        return self.Npix - self.Nb


# SDSS: reference config.py:38-40 (1030-1600 A, dloglam 1e-4) -> 1913 / 720 / 1193
# L32 : LOGLAM_DELTA=1.913e-4 -> 1000 / 377 / 623 (BASELINE config 5)
# DESI: shape of data/model_parameters_desi.npz (9243 / 2238); grid inferred
GRIDS = {
    "sdss": GridSpec("sdss", 1030.0, 1e-4, 1913),
    "l32": GridSpec("l32", 1030.0, 1.913e-4, 1000),
    "desi": GridSpec("desi", 910.0, 5.62e-5, 9243),
}


def smooth_random_params(grid: GridSpec, Nh: int, seed: int = 0, device="cpu"):
    """Ground-truth-like parameters for shapes that have no pretrained model: F = Nh
    Gaussian-smoothed random curves, Psi, omega ~ U(0.05,0.5)*scale, tau0 .15, beta 1.33, c0 .24,
    mu = smooth positive continuum with a Ly-alpha bump."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    P, Nb = grid.Npix, grid.Nb
    x = torch.linspace(-1, 1, P)
    raw = torch.randn(Nh, P + 200, generator=g)
    k = torch.exp(-0.5 * (torch.arange(-100, 101) / 25.0) ** 2)
    k = k / k.sum()
    Fm = torch.nn.functional.conv1d(raw[:, None, :], k[None, None, :]).squeeze(1)   # (Nh, P)
    Fm = 0.15 * Fm / Fm.std(dim=1, keepdim=True)
    wav = torch.tensor(grid.wav(), dtype=torch.float32)
    mu = 1.0 + 0.3 * (wav / 1300.0) ** -1.5 + 1.2 * torch.exp(-0.5 * ((wav - LYA) / 12.0) ** 2)
    params = {
        "F": Fm.T.contiguous().to(torch.float32),
        "Psi": (0.002 + 0.01 * torch.rand(P, generator=g)).to(torch.float32),
        "omega": (0.05 + 0.45 * torch.rand(Nb, generator=g)).to(torch.float32),
        "tau0": torch.tensor(0.15), "c0": torch.tensor(0.24), "beta": torch.tensor(1.33),
    }
    return {k_: v.to(device) for k_, v in params.items()}, mu.to(torch.float32).to(device)


def make_spectra(params, mu, grid: GridSpec, B: int, seed: int, device="cpu", law="becker",
                 mask_iid=0.05, mask_runs=3, run_len=(20, 200), ood_frac=0.0, chunk=8192):
    """Returns dict(flux, error, zabs, mask(bool), delta, zqso) of B spectra on `device`.

    delta = flux - mu*A is what the reference's dataloader feeds QFA.forward
    (dataloader.py:135-136; on these grids tau_total is Ly-alpha only)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    P, Nb = grid.Npix, grid.Nb
    wav = torch.tensor(grid.wav(), dtype=torch.float32, device=dev)
    F = params["F"].to(dev, torch.float32)
    Psi = params["Psi"].to(dev, torch.float32)
    omega = params["omega"].to(dev, torch.float32)
    tau0, c0, beta = (float(params[k]) for k in ("tau0", "c0", "beta"))
    mu = mu.to(dev, torch.float32)
    Nh = F.shape[1]
    t0, be, C, zn = LAW_CONSTANTS[law]
    out = {k: [] for k in ("flux", "error", "zabs", "mask", "delta", "zqso")}
    idx = torch.arange(P, device=dev)
    for s in range(0, B, chunk):
        n = min(chunk, B - s)

        def U(*shape):
            return torch.rand(*shape, generator=g, device=dev)

        def N(*shape):
            return torch.randn(*shape, generator=g, device=dev)
        zq = 2.0 + 1.5 * U(n)
        zabs = (1 + zq)[:, None] * wav[None, :Nb] / LYA - 1                      # dataloader.py:102
        A = torch.ones(n, P, device=dev)
        A[:, :Nb] = torch.exp(-(t0 * ((1 + zabs) / zn) ** be + C))
        zdep = (1 - c0 - torch.exp(-tau0 * (1 + zabs) ** beta)) ** 2
        h = N(n, Nh)
        cont = mu[None] + h @ F.T + torch.sqrt(Psi)[None] * N(n, P)
        snr = 2.0 + 18.0 * U(n)
        sigma = mu.abs()[None] / snr[:, None] * (0.8 + 0.4 * U(n, P))
        flux = A * cont + sigma * N(n, P)
        flux[:, :Nb] += torch.sqrt(omega[None] * zdep) * N(n, Nb)
        if ood_frac > 0:   # out-of-distribution spectra: broad absorption troughs redward of Ly-alpha
            bad = U(n) < ood_frac
            c = Nb + (U(n) * (P - Nb) * 0.6).long()
            wdt = 30 + (U(n) * 120).long()
            trough = (idx[None] >= c[:, None]) & (idx[None] < (c + wdt)[:, None]) & bad[:, None]
            flux = torch.where(trough, 0.15 * flux, flux)
        mask = U(n, P) >= mask_iid
        for _ in range(mask_runs):
            on = U(n) < 0.5
            st = (U(n) * P).long()
            ln = run_len[0] + (U(n) * (run_len[1] - run_len[0])).long()
            hole = (idx[None] >= st[:, None]) & (idx[None] < (st + ln)[:, None]) & on[:, None]
            mask &= ~hole
        flux = torch.where(mask, flux, torch.full_like(flux, -999.0))
        sigma = torch.where(mask, sigma, torch.full_like(sigma, -999.0))
        delta = flux - mu[None] * A
        for k, v in (("flux", flux), ("error", sigma), ("zabs", zabs), ("mask", mask), ("delta", delta), ("zqso", zq)):
            out[k].append(v.to(torch.float32) if v.dtype.is_floating_point else v)
    return {k: torch.cat(v) for k, v in out.items()}


class SyntheticLoader(object):
    """GPU-resident stand-in for reference QFA/dataloader.py exposing exactly the members
    QFA.train uses (model.py:204-211): mu, data_size, batch_size, rewind(), have_next_batch(),
    next_batch().  `rank`/`world` shard the data set by contiguous ranges;
    the shuffle is the same seeded permutation on every rank."""

    def __init__(self, data, mu, batch_size, rank=0, world=1, seed=0, shuffle=True):
        n = data["delta"].shape[0]
        per = n // world
        sl = slice(rank * per, (rank + 1) * per)
        self.delta, self.error = data["delta"][sl], data["error"][sl]
        self.zabs, self.mask = data["zabs"][sl], data["mask"][sl]
        self.mu = mu.detach().cpu().numpy() if torch.is_tensor(mu) else np.asarray(mu)
        self.data_size = n                       # global size (Niter = N // B uses global numbers)
        self.batch_size = batch_size             # global batch; each rank takes batch_size // world
        self.local_batch = max(1, batch_size // world)
        self.local_size = per
        self.cur = 0
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self._perm = torch.arange(per)
        self.shuffle = shuffle

    def rewind(self):
        self.cur = 0
        if self.shuffle:
            self._perm = torch.randperm(self.local_size, generator=self._gen)

    def have_next_batch(self):
        return self.cur < self.local_size

    def next_batch(self):
        end = min(self.cur + self.local_batch, self.local_size)
        if self.shuffle:
            ii = self._perm[self.cur:end].to(self.delta.device)
            batch = (self.delta[ii], self.error[ii], self.zabs[ii], self.mask[ii])
        else:
            s = slice(self.cur, end)
            batch = (self.delta[s], self.error[s], self.zabs[s], self.mask[s])
        self.cur = end
        return batch
