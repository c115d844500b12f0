"""TEST INFRASTRUCTURE ONLY -- golden vectors for the data-preparation row, produced by the
REAL reference functions `QFA.utils.tau_total`, `QFA.utils.smooth` and the three statements of reference
QFA/dataloader.py:102,109-112,135 (the Dataloader class itself needs spectrum files and a yacs config; its arithmetic
is re-issued here call by call with the reference's own functions).

    python oracle/make_golden_prep.py     ->  tests/golden/prep_desi_like.npz, tests/golden/prep_sdss_like.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402


def make(name, lam_min, lam_max, dloglam, n, seed, which):
    _, _, _, ru = load_reference()
    rng = np.random.default_rng(seed)
    wav = 10 ** np.arange(np.log10(lam_min), np.log10(lam_max), dloglam)          # dataloader.py:61
    Nb = int(np.sum(wav < 1215.67))
    Nr = len(wav) - Nb
    zqso = rng.uniform(2.0, 3.5, size=n)
    flux = rng.normal(1.0, 0.3, size=(n, len(wav)))
    error = rng.uniform(0.05, 0.3, size=(n, len(wav)))
    bad = rng.uniform(size=flux.shape) < 0.1
    flux[bad] = -999.0
    error[bad] = -999.0
    mask = (flux != -999.0) & (error != -999.0)                                     # dataloader.py:28
    zabs = (zqso + 1).reshape(-1, 1) * wav[:Nb] / 1215.67 - 1                        # dataloader.py:102
    taus = ru.tau_total(wav, zqso, which=which)                                      # utils.py:174-203
    s = np.hstack((np.exp(1 * taus), np.ones((n, Nr), dtype=float)))                 # dataloader.py:109
    mu = np.sum(flux * s * mask, axis=0) / np.sum(flux != -999., axis=0)             # dataloader.py:110
    mu = ru.smooth(mu, window_len=16)                                                # dataloader.py:111
    s2 = np.hstack((np.exp(-1 * taus), np.ones((n, Nr), dtype=float)))               # dataloader.py:134
    delta = (flux - mu * s2).astype(np.float32)                                      # dataloader.py:135
    np.savez(os.path.join(ROOT, "tests", "golden", name), wav=wav, zqso=zqso, flux=flux.astype(np.float32),
             error=error.astype(np.float32), mask=mask, zabs=zabs.astype(np.float32), taus=taus, mu=mu, delta=delta,
             law=np.array(which), Nb=np.int64(Nb))
    print(name, "Npix", len(wav), "Nb", Nb, "series on the first pixel:", int(np.sum(wav[0] < ru.lyseries['lambda'])))


if __name__ == "__main__":
    make("prep_desi_like.npz", 910.0, 1400.0, 8e-4, 24, 11, "becker")     # starts blueward of every Lyman line: all 30 series
    make("prep_sdss_like.npz", 1030.0, 1600.0, 1e-3, 16, 12, "kamble")    # redward of Ly-beta: Ly-alpha only
