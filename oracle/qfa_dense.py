"""TEST INFRASTRUCTURE ONLY -- CPU "port" oracle of the reference's DENSE algorithm.

Nothing under qfa_b200/ may import this file. Allowed users: tests/,
__graft_entry__.smoke() (as the checker) and bench.py's cpu_baseline /
`--impl reference` legs (as the thing timed on the host cores).

This is a restatement, in our own words, of what the reference does per
spectrum: it gathers the unmasked pixels and then works with dense n x n
matrices (O(n^3)), exactly like the reference, so that (a) in float32 it
reproduces the reference's rounding behaviour closely and (b) timing it is a
fair stand-in for "the reference's own CPU path" on a box where
/root/reference does not exist.  Each function cites the reference lines it
follows (paths relative to /root/reference).

Pinned by tests/test_oracle.py against
  * the reference's shipped known-answer vector (data/spec-4321-55504-0114.npz
    keys ll,h,our / ll_red,h_red,our_red -- SURVEY.md section 4), and
  * goldens produced by running the real reference in the build container
    (oracle/make_golden.py -> tests/golden/*.npz).

All functions take a `dt` (torch.float32 to mimic the shipped reference,
torch.float64 to mimic the fp64-promoted reference of SURVEY.md section 8c).
"""
import math

import torch

LOG2PI = 1.8378770664093453  # model.py:20

# optical-depth laws, utils.py:105,119,133,141 (series=1 -> coefficient 1.0, utils.py:146-147,160)
TAU_LAWS = ("becker", "fg", "kamble", "mock")


def mean_tau(z, which="becker"):
    """utils.py:95-171 with series=1."""
    if which == "becker":
        return 0.751 * ((1 + z) / (1 + 3.5)) ** 2.90 + (-0.132)
    if which == "fg":
        return 0.0018 * (1 + z) ** 3.92
    if which == "kamble":
        return 5.54 * 1e-3 * (1 + z) ** 3.182
    if which == "mock":
        return 0.2231435513142097 * ((1 + z) / 3.25) ** 3.2
    raise NotImplementedError(which)


def lowrank_inverse(Mt, D, dt):
    """utils.py:29-32: dense Woodbury inverse of Mt Mt^T + diag(D)."""
    Dinv = torch.diag(1.0 / D)
    eye = torch.eye(Mt.shape[1], dtype=dt)
    core = torch.linalg.inv(eye + Mt.T @ Dinv @ Mt)
    return Dinv - Dinv @ Mt @ core @ Mt.T @ Dinv


def lowrank_logdet(Mt, D, dt):
    """utils.py:51-54: matrix determinant lemma."""
    Dinv = torch.diag(1.0 / D)
    eye = torch.eye(Mt.shape[1], dtype=dt)
    return torch.sum(torch.log(D)) + torch.log(torch.linalg.det(eye + Mt.T @ Dinv @ Mt))


def _absorbed_pieces(P, zabs, mask, Nb, which, dt):
    """Shared front half of model.py:121-131 and model.py:161-172."""
    bmask = mask[:Nb]
    nb = int(bmask.sum())
    nr = int(mask[Nb:].sum())
    zb = zabs[bmask]
    A = torch.hstack((torch.exp(-1.0 * mean_tau(zb, which)), torch.ones(nr, dtype=dt)))
    Ft = torch.diag(A) @ P["F"][mask, :]                       # model.py:126-127
    Psi_t = A * P["Psi"][mask] * A                              # model.py:128
    tauhi = P["tau0"] * torch.pow(1.0 + zb, P["beta"])          # utils.py:72
    root_exp = 1.0 - P["c0"] - torch.exp(-1.0 * tauhi)          # utils.py:91
    zdep = root_exp * root_exp                                  # utils.py:92
    om = torch.hstack((P["omega"][bmask] * zdep, torch.zeros(nr, dtype=dt)))  # model.py:130
    return bmask, nb, nr, zb, A, Ft, Psi_t, zdep, om


def nll_and_grad_single(P, delta, error, zabs, mask, Nb, which="becker", dt=torch.float32, fp32_logpi=True):
    """model.py:107-158. Returns (nll (1,1), dict of full-size partials).

    fp32_logpi: quirk Q5 -- `Npix*log2pi` is an int64 tensor times a Python
    float, i.e. evaluated in the default dtype; the shipped reference runs with
    default float32. The fp64-promoted oracle sets the default dtype to float64,
    so there the product is float64 (pass fp32_logpi=False).
    """
    Npix_full, Nh = P["F"].shape
    bmask, nb, nr, zb, A, Ft, Psi_t, zdep, om = _absorbed_pieces(P, zabs, mask, Nb, which, dt)
    n = nb + nr
    d = delta[mask]
    e = error[mask]
    D = Psi_t + om + e * e                                       # model.py:131
    Sinv = lowrank_inverse(Ft, D, dt)                            # model.py:132
    logdet = lowrank_logdet(Ft, D, dt)                           # model.py:133
    d = d[:, None]
    npi = torch.tensor(float(n), dtype=torch.float32 if fp32_logpi else torch.float64) * LOG2PI   # quirk Q5
    nll = 0.5 * (d.mT @ Sinv @ d + npi + logdet)                 # model.py:135
    G = 0.5 * (Sinv - Sinv @ d @ d.mT @ Sinv)                    # model.py:136 (left-assoc)
    dA = torch.diag(A)
    pF = 2 * dA @ G @ dA @ Ft                                    # model.py:137 (quirk Q2)
    g = torch.diag(G)
    pPsi = A * g * A                                             # model.py:139
    pOm = g[:nb] * zdep                                          # model.py:140
    root_lin = 1.0 - P["tau0"] * torch.pow(1.0 + zb, P["beta"]) - P["c0"]  # model.py:141 (quirk Q3)
    common = g[:nb] * om[:nb] * zdep * 2.0 * root_lin
    pT0 = -1.0 * torch.sum(common * torch.pow(1.0 + zb, P["beta"]))        # model.py:142
    pBe = -1.0 * torch.sum(common * (P["tau0"] * torch.pow(1 + zb, P["beta"]) * torch.log(1 + zb)))  # :143
    pC0 = -1.0 * torch.sum(common)                                          # model.py:144
    gF = torch.zeros((Npix_full, Nh), dtype=dt)
    gF[mask, :] = pF
    gOm = torch.zeros((Nb,), dtype=dt)
    gOm[bmask] = pOm
    gPsi = torch.zeros((Npix_full,), dtype=dt)
    gPsi[mask] = pPsi
    return nll, {"F": gF, "Psi": gPsi, "omega": gOm, "tau0": pT0, "c0": pC0, "beta": pBe}


def forward(P, delta, error, zabs, mask, Nb, which="becker", dt=torch.float32, fp32_logpi=True):
    """model.py:74-105: batch mean NLL, grads divided by per-element non-zero counts (quirk Q4)."""
    B = delta.shape[0]
    keys = ("F", "Psi", "omega", "tau0", "c0", "beta")
    acc = {k: torch.zeros_like(P[k], dtype=dt) for k in keys}
    cnt = {k: torch.zeros_like(P[k], dtype=dt) for k in keys}
    loss = 0.0
    for b in range(B):
        nll, g = nll_and_grad_single(P, delta[b], error[b], zabs[b], mask[b], Nb, which, dt, fp32_logpi)
        loss = loss + nll / B
        for k in keys:
            acc[k] += g[k]
            cnt[k] += (g[k] != 0.0)
    return loss, {k: acc[k] / cnt[k] for k in keys}


def predict_single(P, mu, flux, error, zabs, mask, Nb, which="becker", dt=torch.float32, fp32_logpi=True):
    """model.py:160-180. Returns (nll (1,1), hmean (Nh,1), hcov (Nh,Nh), cont (Npix,), unc (Npix,))."""
    Nh = P["F"].shape[1]
    bmask, nb, nr, zb, A, Ft, Psi_t, zdep, om = _absorbed_pieces(P, zabs, mask, Nb, which, dt)
    n = nb + nr
    d = flux[mask] - mu[mask] * A                                # model.py:166
    e = error[mask]
    D = Psi_t + om + e * e
    Sinv = lowrank_inverse(Ft, D, dt)
    logdet = lowrank_logdet(Ft, D, dt)
    d = d[:, None]
    npi = torch.tensor(float(n), dtype=torch.float32 if fp32_logpi else torch.float64) * LOG2PI   # quirk Q5
    nll = 0.5 * (d.mT @ Sinv @ d + npi + logdet)                 # model.py:176
    Se = torch.diag(1.0 / D)                                     # model.py:177
    hcov = torch.linalg.inv(torch.eye(Nh, dtype=dt) + Ft.T @ Se @ Ft)   # model.py:178
    hmean = hcov @ Ft.T @ Se @ d                                 # model.py:179
    F = P["F"]
    return nll, hmean, hcov, (F @ hmean).squeeze() + mu, torch.diag(F @ hcov @ F.T) ** 0.5  # model.py:180


# ---------------------------------------------------------------------------
# parameter housekeeping and optimiser (boundary rows a8-a10)
# ---------------------------------------------------------------------------

def clip_params(P, lo=1e-3, hi=2.0):
    """model.py:233-241."""
    Q = dict(P)
    Q["omega"] = torch.clip(P["omega"], min=lo, max=hi)
    Q["Psi"] = torch.clip(P["Psi"], min=lo, max=hi)
    Q["tau0"] = torch.clip(P["tau0"], min=0.0, max=1.0)
    Q["beta"] = torch.clip(P["beta"], min=0.1, max=5.0)
    Q["c0"] = torch.clip(P["c0"], min=-5.0, max=5.0)
    return Q


def smooth_params(P):
    """model.py:243-252: box filters that ignore the zero padding (count_include_pad=False)."""
    import torch.nn.functional as Fn
    Q = dict(P)
    Q["omega"] = Fn.avg_pool1d(P["omega"].reshape(1, -1), 15, 1, 7, count_include_pad=False).squeeze()
    Q["Psi"] = Fn.avg_pool1d(P["Psi"].reshape(1, -1), 15, 1, 7, count_include_pad=False).squeeze()
    Npix, Nh = P["F"].shape
    Q["F"] = Fn.avg_pool2d(P["F"].reshape(1, Npix, Nh), (31, 1), (1, 1), (15, 0), count_include_pad=False).squeeze()
    return Q


def scheduled_lr(lr0, alpha, step, i):
    """optimizer.py:98."""
    return lr0 * alpha ** ((i + 1) // step)


def adam_update(P, g, m, v, i, lr, b1=0.9, b2=0.999, eps=1e-8, wd=1e-3):
    """optimizer.py:47-52 (bias correction indexed by the EPOCH counter i, quirk Q8).
    Returns (new_params, new_m, new_v)."""
    g = {k: g[k] + wd * P[k] for k in g}
    m = {k: (1 - b1) * g[k] + b1 * m[k] for k in g}
    v = {k: (1 - b2) * g[k] * g[k] + b2 * v[k] for k in g}
    mh = {k: m[k] / (1.0 - b1 ** (i + 1)) for k in g}
    vh = {k: v[k] / (1.0 - b2 ** (i + 1)) for k in g}
    newP = {k: P[k] - lr * mh[k] / (torch.sqrt(vh[k]) + eps) for k in P}
    return newP, m, v


def params_from_npz(path, dt=torch.float32, c0_bug=True):
    """model.py:282-295, including `c0 <- file['beta']` (quirk Q1) unless c0_bug=False."""
    import numpy as np
    f = np.load(path)
    P = {k: torch.tensor(f[k], dtype=dt) for k in ("F", "Psi", "omega", "tau0", "beta")}
    P["c0"] = torch.tensor(f["beta"] if c0_bug else f["c0"], dtype=dt)
    mu = torch.tensor(f["mu"], dtype=dt)
    return P, mu
