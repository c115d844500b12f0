"""Host-side helpers mirroring the public names of reference QFA/utils.py.

The CUDA kernels evaluate the optical-depth laws, `tauHI` and `omega_func`
themselves (qfa_b200/csrc/qfa_common.cuh); the torch functions here exist so
that user code written against the reference (`partial(tau, which='becker')`,
notebooks, data loaders) keeps working, and so that the model can recognise
WHICH law a callable stands for (the kernels take an enum, not a callable).
"""
from functools import partial
from typing import Optional

import numpy as np
import torch

from ._lib import TAU_LAWS, QfaError

LYA = 1215.67  # reference dataloader.py:15

# (t0, beta, C, z-normalisation): tau = t0*((1+z)/zn)**beta + C   reference utils.py:105,119,133,141
LAW_CONSTANTS = {
    "becker": (0.751, 2.90, -0.132, 4.5),
    "fg": (0.0018, 3.92, 0.0, 1.0),
    "kamble": (5.54 * 1e-3, 3.182, 0.0, 1.0),
    "mock": (0.2231435513142097, 3.2, 0.0, 3.25),
}

# Lyman series (oscillator strength f, wavelength in Angstrom): the atomic data of reference QFA/Lyman_series.csv
# (30 lines, Ly-alpha .. n = 31); coefficient = lambda*f / (lambda*f)_alpha (utils.py:146-147).  Same table as the
# kernels' (csrc/qfa_aux.cuh).
_LYMAN = [
    (4.1620e-01, 1215.6701), (7.9140e-02, 1025.7222), (2.9010e-02, 972.5367), (1.3950e-02, 949.7430),
    (7.8030e-03, 937.8034), (4.8160e-03, 930.7482), (3.1850e-03, 926.2256), (2.2170e-03, 923.1503),
    (1.6060e-03, 920.9630), (1.2010e-03, 919.3513), (9.2190e-04, 918.1293), (7.2310e-04, 917.1805),
    (5.7770e-04, 916.4291), (4.6890e-04, 915.8238), (3.8580e-04, 915.3289), (3.2120e-04, 914.9192),
    (2.7030e-04, 914.5762), (2.2970e-04, 914.2861), (1.9680e-04, 914.0385), (1.6990e-04, 913.8256),
    (1.4770e-04, 913.6411), (1.2930e-04, 913.4803), (1.1370e-04, 913.3391), (1.0060e-04, 913.2146),
    (8.9360e-05, 913.1042), (7.9780e-05, 913.0059), (7.1480e-05, 912.9179), (6.4350e-05, 912.8389),
    (5.8120e-05, 912.7676), (5.2640e-05, 912.7032)]


def series_coeff(series: int) -> float:
    if not 1 <= series <= len(_LYMAN):
        raise NotImplementedError(f"Lyman series {series} not tabulated (1..{len(_LYMAN)})")
    f0, l0 = _LYMAN[0]
    f, l = _LYMAN[series - 1]
    return (l * f) / (l0 * f0)


def tau(z, which: Optional[str] = "becker", series: Optional[int] = 1):
    """Mean optical depth, same call signature as reference utils.py:149-171."""
    if which not in LAW_CONSTANTS:
        raise NotImplementedError("currently available mean optical depth function: "
                                  "['becker', 'fg', 'kamble', 'mock']")
    t0, be, C, zn = LAW_CONSTANTS[which]
    return (t0 * ((1 + z) / zn) ** be + C) * series_coeff(series)


default_tau = partial(tau, which="becker")  # reference model.py:21


def tau_total(wav_grid, zqso, which: Optional[str] = "becker"):
    """Total Lyman-series optical depth on the blue side, same call signature and result as reference
    utils.py:174-203 (numpy in, numpy (N, Nb) out).  Host-side helper; the device path is qfa_gather_prepare."""
    wav_grid = np.asarray(wav_grid, dtype=np.float64)
    zqso = np.asarray(zqso, dtype=np.float64).reshape(-1)
    if not wav_grid[0] < _LYMAN[0][1]:
        raise ValueError("Wavelength grid does not cover Lyman series lines")
    Nb = int(np.sum(wav_grid < _LYMAN[0][1]))
    taus = np.zeros((len(zqso), Nb))
    for s, (_, lam) in enumerate(_LYMAN):
        if not wav_grid[0] < lam:
            break
        nb = int(np.sum(wav_grid < lam))
        zabs = (zqso + 1).reshape(-1, 1) * wav_grid[:nb] / lam - 1
        taus[:, :nb] += tau(zabs, which=which, series=s + 1)
    return taus


def tauHI(z, tau0, beta):
    """reference utils.py:57-72"""
    return tau0 * torch.pow((1.0 + z), beta)


def omega_func(z, tau0, beta, c0):
    """reference utils.py:75-92"""
    root = 1.0 - c0 - torch.exp(-1.0 * tauHI(z, tau0, beta))
    return root * root


def resolve_tau_law(tau_arg) -> int:
    """Map what the reference accepts as `tau=` onto the kernel enum.

    Accepted: a law name; `functools.partial(tau, which=...)` of the reference's
    or this module's `tau` (that is what reference main.py:77 / model.py:21 pass).
    Anything else is an error -- there is no CPU fallback that could call back
    into arbitrary Python from the kernels.
    """
    if isinstance(tau_arg, str):
        name = tau_arg
    elif isinstance(tau_arg, partial):
        series = tau_arg.keywords.get("series", 1)
        if series != 1:
            raise QfaError("only the Ly-alpha mean optical depth (series=1) is supported by the kernels")
        name = tau_arg.keywords.get("which", "becker")
    elif tau_arg is tau or getattr(tau_arg, "__name__", "") == "tau":
        name = "becker"
    else:
        raise QfaError("tau must be a law name or functools.partial(tau, which=<law>); arbitrary callables "
                       "cannot be evaluated inside the CUDA kernels")
    if name not in TAU_LAWS:
        raise QfaError(f"unknown mean optical depth law {name!r}; available: {sorted(TAU_LAWS)}")
    return TAU_LAWS[name]


def wavelength_grid(lam_min=1030.0, lam_max=1600.0, dloglam=1e-4):
    """Rest-frame grid of reference dataloader.py:61-63 -> (wav, Nb, Nr)."""
    wav = 10 ** np.arange(np.log10(lam_min), np.log10(lam_max), dloglam)
    Nb = int(np.sum(wav < LYA))
    return wav, Nb, len(wav) - Nb


def MatrixInverse(M: torch.Tensor, D: torch.Tensor, device=None) -> torch.Tensor:
    """API-compat only (reference utils.py:12-32). The kernels never form this n x n matrix."""
    Dinv = 1.0 / D
    Nh = M.shape[1]
    core = torch.linalg.inv(torch.eye(Nh, dtype=M.dtype, device=M.device) + (M.T * Dinv) @ M)
    DM = Dinv[:, None] * M
    return torch.diag(Dinv) - DM @ core @ DM.T


def MatrixLogDet(M: torch.Tensor, D: torch.Tensor, device=None) -> torch.Tensor:
    """API-compat only (reference utils.py:35-54)."""
    Nh = M.shape[1]
    return torch.sum(torch.log(D)) + torch.logdet(torch.eye(Nh, dtype=M.dtype, device=M.device) + (M.T / D) @ M)
