"""TEST INFRASTRUCTURE ONLY -- fp64 NumPy batched restatement in LOW-RANK form.

Nothing under qfa_b200/ may import this file (see oracle/qfa_dense.py header).

The dense port (oracle/qfa_dense.py) is O(n^3) per spectrum and cannot check
DESI-sized or 10^4-spectrum cases in seconds.  This file restates the same
quantities through the Woodbury / determinant-lemma algebra of SURVEY.md
section 7.1 on the FULL pixel grid (masked pixels get weight 0), vectorised
over the batch, in float64.  It is pinned against the fp64-promoted reference
goldens (tests/golden/fwd_*_f64.npz) and the shipped known-answer vector.

Reference lines restated: model.py:121-158 (likelihood + partials),
model.py:98-104 (batch reduction, quirk Q4), model.py:161-180 (prediction),
utils.py:29-32,51-54 (Woodbury, determinant lemma), utils.py:72,91-92
(tauHI, omega_func), utils.py:105-141 (optical-depth laws).

Count semantics (quirk Q4): the reference divides every gradient element by the
number of spectra whose partial is `!= 0`.  In exact arithmetic with
generic-position data that is: F/Psi -> #{b: mask[b,i]}, omega -> same on the
blue side, tau0/c0 -> #{b: spectrum has >=1 unmasked blue pixel}, beta -> the
same unless tau0 == 0 (every term carries a factor tau0).  This file and the
CUDA kernels use that closed form; the dense port counts literally.  They agree
on all committed goldens; the difference is confined to exact-float-zero events.
"""
import numpy as np

LOG2PI = 1.8378770664093453
LAW_CONSTANTS = {  # tau(z) = t0 * ((1+z)/zn)^be + C      utils.py:105,119,133,141
    "becker": (0.751, 2.90, -0.132, 4.5),
    "fg": (0.0018, 3.92, 0.0, 1.0),
    "kamble": (5.54 * 1e-3, 3.182, 0.0, 1.0),
    "mock": (0.2231435513142097, 3.2, 0.0, 3.25),
}


def mean_tau(z, which="becker"):
    t0, be, C, zn = LAW_CONSTANTS[which]
    return t0 * ((1.0 + z) / zn) ** be + C


def _per_pixel(P, zabs, mask, err, Nb, which):
    """A, zdep, D, w on the full grid. Shapes (B, Npix); red side A=1, zdep=0."""
    B, Npix = mask.shape
    z = np.asarray(zabs, np.float64)
    A = np.ones((B, Npix))
    A[:, :Nb] = np.exp(-mean_tau(z, which))                       # model.py:125
    zdep = np.zeros((B, Npix))
    opz_b = (1.0 + z) ** float(P["beta"])
    zdep[:, :Nb] = (1.0 - float(P["c0"]) - np.exp(-float(P["tau0"]) * opz_b)) ** 2   # utils.py:91-92
    om = np.zeros(Npix)
    om[:Nb] = np.asarray(P["omega"], np.float64)
    D = A * A * np.asarray(P["Psi"], np.float64)[None, :] + om[None, :] * zdep + err * err   # model.py:128-131
    m = mask.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(mask, 1.0 / D, 0.0)
        logD = np.where(mask, np.log(np.where(mask, D, 1.0)), 0.0)
    return A, zdep, om, D, w, logD, m


def _solve_core(F, A, w, resid):
    """M = I + sum_i w A^2 f f^T ; b = sum_i w A resid f. Returns M, b."""
    s2 = w * A * A
    M = np.einsum("bi,ik,il->bkl", s2, F, F) + np.eye(F.shape[1])[None]
    bvec = np.einsum("bi,ik->bk", w * A * resid, F)
    return M, bvec


def nll_batch(P, delta, error, zabs, mask, Nb, which="becker"):
    """Per-spectrum NEGATIVE log-likelihood (model.py:135), shape (B,)."""
    F = np.asarray(P["F"], np.float64)
    delta = np.asarray(delta, np.float64)
    err = np.asarray(error, np.float64)
    A, zdep, om, D, w, logD, m = _per_pixel(P, zabs, mask, err, Nb, which)
    M, bvec = _solve_core(F, A, w, delta)
    a = np.linalg.solve(M, bvec[..., None])[..., 0]
    _, logdetM = np.linalg.slogdet(M)
    n = m.sum(1)
    return 0.5 * ((w * delta * delta).sum(1) - (bvec * a).sum(1) + n * LOG2PI + logD.sum(1) + logdetM)


def forward(P, delta, error, zabs, mask, Nb, which="becker", return_sums=False):
    """Batched restatement of model.py:74-158.

    Returns (loss, grads) with the reference's normalisation, plus -- when
    return_sums -- a dict holding the un-normalised sums, counts, per-spectrum
    NLLs and the extra d NLL / d mu (quirk Q6: -A_i u_i summed over spectra;
    the reference has no such gradient, so it is pinned only against autograd
    in tests/test_oracle.py).
    """
    F = np.asarray(P["F"], np.float64)
    Npix, Nh = F.shape
    delta = np.asarray(delta, np.float64)
    err = np.asarray(error, np.float64)
    z = np.asarray(zabs, np.float64)
    B = delta.shape[0]
    tau0, beta, c0 = float(P["tau0"]), float(P["beta"]), float(P["c0"])
    A, zdep, om, D, w, logD, m = _per_pixel(P, zabs, mask, err, Nb, which)
    s2 = w * A * A
    s3 = s2 * A
    M, bvec = _solve_core(F, A, w, delta)
    M2 = np.einsum("bi,ik,il->bkl", s3, F, F)
    b2 = np.einsum("bi,ik->bk", s2 * delta, F)
    Minv = np.linalg.inv(M)
    a = np.einsum("bkl,bl->bk", Minv, bvec)
    K = Minv @ M2
    _, logdetM = np.linalg.slogdet(M)
    n = m.sum(1)
    nll = 0.5 * ((w * delta * delta).sum(1) - (bvec * a).sum(1) + n * LOG2PI + logD.sum(1) + logdetM)
    fa = a @ F.T                                               # (B, Npix): f_i . a_b
    u = w * (delta - A * fa)                                   # (Sigma^-1 delta)_i
    q = np.einsum("ik,bkl,il->bi", F, Minv, F)
    g = 0.5 * (w - w * w * A * A * q - u * u) * m              # model.py:136,138
    c = b2 - np.einsum("bkl,bl->bk", M2, a)                    # sum_i u_i A_i^2 f_i
    fK = np.einsum("ik,bkl->bil", F, K)
    dF = s3[..., None] * F[None] - s2[..., None] * fK - (A * u)[..., None] * c[:, None, :]   # model.py:137 (Q2)
    dPsi = A * A * g                                           # model.py:139
    dOm = (g * zdep)[:, :Nb]                                   # model.py:140
    opz = 1.0 + z
    powb = opz ** beta
    root_lin = 1.0 - tau0 * powb - c0                          # model.py:141 (Q3)
    t = g[:, :Nb] * (om[None, :Nb] * zdep[:, :Nb]) * zdep[:, :Nb] * 2.0 * root_lin
    dT0 = -(t * powb).sum(1)                                   # model.py:142
    dBe = -(t * tau0 * powb * np.log(opz)).sum(1)              # model.py:143
    dC0 = -t.sum(1)                                            # model.py:144
    sums = {"F": dF.sum(0), "Psi": dPsi.sum(0), "omega": dOm.sum(0),
            "tau0": dT0.sum(), "c0": dC0.sum(), "beta": dBe.sum()}
    pix_cnt = m.sum(0)
    has_blue = (m[:, :Nb].sum(1) > 0).sum().astype(np.float64)
    counts = {"F": np.repeat(pix_cnt[:, None], Nh, 1), "Psi": pix_cnt, "omega": pix_cnt[:Nb],
              "tau0": has_blue, "c0": has_blue, "beta": has_blue if tau0 != 0.0 else 0.0}
    with np.errstate(divide="ignore", invalid="ignore"):
        grads = {k: np.asarray(sums[k]) / np.asarray(counts[k]) for k in sums}
    loss = nll.sum() / B
    if return_sums:
        extra = {"sums": sums, "counts": counts, "nll": nll, "dmu": -(A * u).sum(0),
                 "hmean": a, "Minv": Minv}
        return loss, grads, extra
    return loss, grads


def predict_batch(P, mu, flux, error, zabs, mask, Nb, which="becker"):
    """Batched restatement of model.py:160-180.
    Returns nll (B,), hmean (B,Nh), hcov (B,Nh,Nh), cont (B,Npix), unc (B,Npix)."""
    F = np.asarray(P["F"], np.float64)
    mu = np.asarray(mu, np.float64)
    flux = np.asarray(flux, np.float64)
    err = np.asarray(error, np.float64)
    A, zdep, om, D, w, logD, m = _per_pixel(P, zabs, mask, err, Nb, which)
    delta = flux - mu[None, :] * A                              # model.py:166
    M, bvec = _solve_core(F, A, w, delta)
    hcov = np.linalg.inv(M)                                     # model.py:178
    hmean = np.einsum("bkl,bl->bk", hcov, bvec)                 # model.py:179
    _, logdetM = np.linalg.slogdet(M)
    n = m.sum(1)
    nll = 0.5 * ((w * delta * delta).sum(1) - (bvec * hmean).sum(1) + n * LOG2PI + logD.sum(1) + logdetM)
    cont = hmean @ F.T + mu[None, :]                            # model.py:180
    unc = np.sqrt(np.einsum("ik,bkl,il->bi", F, hcov, F))
    return nll, hmean, hcov, cont, unc
