"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
reference-shaped Python API -> ctypes -> C ABI, against
  * the committed goldens produced by the real reference (fp64-promoted and fp32),
  * the reference's shipped known-answer vector,
  * the fp64 low-rank oracle on larger seeded inputs,
  * size-independent properties at the BASELINE sizes.
Tolerances (BASELINE north_star): 1e-5 relative in fp64 mode on NLL, gradients and continua;
1e-3 on the continuum in fp32 / mixed mode.
"""
import os

import numpy as np
import pytest
import torch

from conftest import CASES, GOLD, load_case, relerr

pytestmark = pytest.mark.gpu
KEYS = ("F", "Psi", "omega", "tau0", "c0", "beta")
TOL64 = 1e-5          # the stated bar
TIGHT64 = 2e-8        # what double arithmetic should actually deliver (regression guard)


def dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def split_acc(m, acc):
    a = acc.detach().double().cpu().numpy()
    n = m.Nparams
    return dict(sums=a[:n], pix_cnt=a[n:n + m.Npix], scal_cnt=a[n + m.Npix:n + m.Npix + 3],
                nll_sum=a[n + m.Npix + 3], nsp=a[n + m.Npix + 4], dmu=a[n + m.Npix + 5:])


# ----------------------------------------------------------------------------- goldens, fp64 mode
@pytest.mark.parametrize("name", CASES)
def test_forward_fp64_matches_reference_golden(name, cuda_model_factory):
    c, g = load_case(name, "f64")
    m = cuda_model_factory(c, "fp64")
    nll = torch.empty(c["delta"].shape[0], dtype=torch.float64, device="cuda")
    m.accumulate(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]), nll_out=nll)
    assert relerr(nll.cpu().numpy(), g["nll"]) < TIGHT64
    loss, grads = m.forward(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert loss.shape == (1, 1)
    assert relerr(loss.cpu().numpy(), g["loss"]) < 1e-6           # loss/grads are emitted as float32
    for k in KEYS:
        assert grads[k].shape == tuple(np.shape(g["grad_" + k])), k
        assert relerr(grads[k].cpu().numpy(), g["grad_" + k]) < 1e-6, k      # incl. NaN placement (0/0)
    # double-precision view of the same numbers: sums / counts straight from the acc buffer
    a = split_acc(m, grads.acc)
    with np.errstate(divide="ignore", invalid="ignore"):
        gF = a["sums"][:m.Npix * m.Nh].reshape(m.Npix, m.Nh) / a["pix_cnt"][:, None]
    assert relerr(gF, g["grad_F"]) < TIGHT64
    assert TIGHT64 < TOL64


@pytest.mark.parametrize("name", CASES)
def test_single_spectrum_partials_fp64(name, cuda_model_factory):
    c, g = load_case(name, "f64")
    m = cuda_model_factory(c, "fp64")
    b = int(g["single_index"])
    nll, part = m.loglikelihood_and_gradient_for_single_spectra(dev(c["delta"][b]), dev(c["error"][b]),
                                                               dev(c["zabs"][b]), dev(c["mask"][b]))
    assert nll.shape == (1, 1)
    assert abs(float(nll) - g["nll"][b]) <= 1e-6 * max(1.0, abs(g["nll"][b]))
    for k in KEYS:
        assert relerr(part[k].cpu().numpy(), g["single_" + k]) < 1e-6, k


@pytest.mark.parametrize("name", CASES)
def test_predict_fp64_matches_reference_golden(name, cuda_model_factory):
    c, g = load_case(name, "f64")
    m = cuda_model_factory(c, "fp64")
    o = m.predict_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert relerr(o["nll"].cpu().numpy(), g["pred_nll"]) < TIGHT64
    assert relerr(o["hmean"].cpu().numpy(), g["pred_hmean"]) < 1e-7
    assert relerr(o["hcov"].cpu().numpy(), g["pred_hcov"]) < 1e-7
    assert relerr(o["cont"].cpu().numpy(), g["pred_cont"]) < TIGHT64
    assert relerr(o["unc"].cpu().numpy(), g["pred_unc"]) < 1e-7
    # reference call shape for one spectrum (model.py:180)
    ll, hm, hc, cont, unc = m.prediction_for_single_spectra(dev(c["flux"][0]), dev(c["error"][0]),
                                                            dev(c["zabs"][0]), dev(c["mask"][0]))
    assert ll.shape == (1, 1) and hm.shape == (m.Nh, 1) and hc.shape == (m.Nh, m.Nh)
    assert cont.shape == (m.Npix,) and unc.shape == (m.Npix,)


# ----------------------------------------------------------------------------- goldens, float modes
@pytest.mark.parametrize("precision", ["fp32", "mixed"])
@pytest.mark.parametrize("name", CASES)
def test_float_modes_against_fp64_golden(name, precision, cuda_model_factory):
    c, g = load_case(name, "f64")
    m = cuda_model_factory(c, precision)
    o = m.predict_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert relerr(o["cont"].cpu().numpy(), g["pred_cont"]) < 1e-3              # the stated mixed-mode bar
    assert relerr(o["unc"].cpu().numpy(), g["pred_unc"]) < 1e-3
    assert relerr(o["hmean"].cpu().numpy(), g["pred_hmean"]) < 2e-3
    # NLL in the float modes: the stated bar covers the continuum only; we hold the NLL to 1e-3 PER UNMASKED
    # PIXEL (it is a sum of n O(1) terms with cancellation).  The reference's own float32 run of these cases
    # is off by up to 0.1 per spectrum (tiny12) or overflows to inf (l32).
    npx = np.maximum(1, c["mask"].sum(1))
    nll = o["nll"].cpu().numpy()
    assert np.all(np.abs(nll - g["pred_nll"]) <= 1e-3 * npx)
    loss, grads = m.forward(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert abs(float(loss) - float(np.squeeze(g["loss"]))) <= 1e-3 * npx.mean()
    for k in KEYS:
        assert relerr(grads[k].cpu().numpy(), g["grad_" + k]) < 5e-2, k


# ----------------------------------------------------------------------------- tensor-core kernels
TC_CASES = tuple(n for n in CASES if n not in ("l32", "tiny12", "tiny16"))      # Nh <= 8


@pytest.mark.parametrize("name", TC_CASES)
def test_tensor_core_path_against_fp64_golden(name, cuda_model_factory):
    """precision="tf32" forces the tcgen05 kernels (k_tc_gram / k_tc_grad) even for these tiny batches.
    Operands are rounded to TF32 (2^-11 relative), accumulation is fp32 in TMEM.  On the reference's own model
    ('sdss': pretrained parameters, 1913 pixels) the stated mixed-mode bar holds: continuum <= 1e-3.  The tiny
    cases are random unit-scale factor matrices on 96 pixels: far fewer pixels to average the operand rounding
    over and a worse-conditioned M = I + F^T D^-1 F, so the error is amplified; they pin correctness of the
    kernels (edge shapes, padding of Nh, masks), with bounds that still catch any indexing / protocol bug."""
    c, g = load_case(name, "f64")
    m = cuda_model_factory(c, "tf32")
    real = name == "sdss"
    o = m.predict_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert relerr(o["cont"].cpu().numpy(), g["pred_cont"]) < (1e-3 if real else 5e-3)
    assert relerr(o["unc"].cpu().numpy(), g["pred_unc"]) < (1e-3 if real else 3e-2)
    assert relerr(o["hmean"].cpu().numpy(), g["pred_hmean"]) < (2e-3 if real else 4e-2)
    assert relerr(o["hcov"].cpu().numpy(), g["pred_hcov"]) < (2e-3 if real else 4e-2)
    npx = np.maximum(1, c["mask"].sum(1))
    assert np.all(np.abs(o["nll"].cpu().numpy() - g["pred_nll"]) <= (1e-3 if real else 1e-2) * npx)
    nll_only = m.nll_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert torch.equal(nll_only, o["nll"])                     # scoring mode = same kernel without phase O
    loss, grads = m.forward(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    assert abs(float(loss) - float(np.squeeze(g["loss"]))) <= (1e-3 if real else 1e-2) * npx.mean()
    for k in KEYS:
        # incl. NaN placement (0/0).  Measured: 8.6e-4 (gradF) / 1.2e-4 on the reference's own model, <= 2.2e-2 on the
        # badly conditioned 96-pixel cases (12 spectra: no averaging of the operand rounding)
        assert relerr(grads[k].cpu().numpy(), g["grad_" + k]) < (3e-3 if real else 5e-2), k


def test_tensor_core_path_ragged_tiles_and_properties():
    """SDSS shape (1913 pixels = 59.8 K-blocks = 14.9 pixel tiles), 300 + 5 spectra (2 full 128-tiles + 44, then a
    5-spectrum tile): tensor-core results per spectrum do not depend on the tile they fall in, masked garbage is
    ignored, repetition is bitwise identical, and everything agrees with the fp64 kernels."""
    from qfa_b200 import QFA
    grid, P, mu, d = _synthetic("sdss", 8, 305, 99, _sdss_pretrained())
    Pn = {k: v.numpy() for k, v in P.items()}
    mt = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32")
    m64 = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="fp64")
    mt.mu = mu; m64.mu = mu
    X, D, E, Z, M = (d[k].cuda() for k in ("flux", "delta", "error", "zabs", "mask"))
    o = mt.predict_batch(X, E, Z, M)
    r = m64.predict_batch(X, E, Z, M)
    assert relerr(o["cont"].cpu().numpy(), r["cont"].cpu().numpy()) < 1e-3
    assert relerr(o["unc"].cpu().numpy(), r["unc"].cpu().numpy()) < 1e-3
    assert relerr(o["hmean"].cpu().numpy(), r["hmean"].cpu().numpy()) < 2e-3
    npx = np.maximum(1, M.sum(1).cpu().numpy())
    assert np.all(np.abs(o["nll"].cpu().numpy() - r["nll"].cpu().numpy()) <= 1e-3 * npx)
    # a spectrum's outputs are independent of its position in the batch / tile
    o5 = mt.predict_batch(X[300:], E[300:], Z[300:], M[300:])
    for k in ("nll", "hmean", "cont", "unc"):
        assert torch.equal(o5[k], o[k][300:]), k
    full = mt.accumulate(D, E, Z, M).clone()
    ref = m64.accumulate(D, E, Z, M).double().cpu().numpy()
    a, b = split_acc(mt, full), split_acc(m64, torch.as_tensor(ref))
    assert np.array_equal(a["pix_cnt"], b["pix_cnt"]) and np.array_equal(a["scal_cnt"], b["scal_cnt"])
    assert abs(a["nll_sum"] - b["nll_sum"]) <= 1e-4 * abs(b["nll_sum"])
    n_f = mt.Npix * mt.Nh
    assert relerr(a["sums"][:n_f], b["sums"][:n_f]) < 2e-2
    assert relerr(a["sums"][n_f:n_f + mt.Npix], b["sums"][n_f:n_f + mt.Npix]) < 5e-3          # Psi
    assert relerr(a["dmu"], b["dmu"]) < 5e-3
    D2, E2 = D.clone(), E.clone()
    D2[~M] = float("nan")
    E2[~M] = float("inf")
    assert torch.equal(mt.accumulate(D2, E2, Z, M), full)
    assert torch.equal(mt.accumulate(D, E, Z, M), full)
    h = 128
    a0 = mt.accumulate(D[:h], E[:h], Z[:h], M[:h]).clone()
    a1 = mt.accumulate(D[h:], E[h:], Z[h:], M[h:]).clone()
    assert relerr((a0 + a1).cpu().numpy(), full.cpu().numpy()) < 2e-4


# ----------------------------------------------------------------------------- shipped known answer
@pytest.mark.parametrize("precision", ["fp64", "fp32", "mixed", "tf32", "tf32x3"])
def test_shipped_known_answer_vector(precision, tmp_path):
    """BASELINE config 1: data/spec-4321-55504-0114.npz with data/model_parameters.npz
    (nb/predict.ipynb cells 4, 9, 10) -> ll = -510.229248, ll_red = -791.925537."""
    from qfa_b200 import QFA
    k = dict(np.load(os.path.join(GOLD, "kat_sdss.npz")))
    path = tmp_path / "model_parameters.npz"
    np.savez(path, **{key[6:]: k[key] for key in k if key.startswith("param_")})
    m = QFA(720, 1193, 8, torch.device("cuda:0"), precision=precision)
    m.load_from_npz(str(path))                         # includes the c0 <- beta quirk
    wav = 10 ** np.arange(np.log10(1030), np.log10(1600), 1e-4)
    zabs = dev(wav[:720] * (1 + float(k["z"])) / 1215.67 - 1, torch.float32)
    flux, error = dev(k["flux"], torch.float32), dev(k["error"], torch.float32)
    tol = 1e-5 if precision in ("fp64", "tf32x3") else 1e-3        # the 3xTF32 tensor-core mode holds the fp64-mode bar here
    for sfx in ("", "_red"):
        mask = k["mask"].copy()
        if sfx:
            mask[:720] = False
        ll, hm, hc, cont, unc = m.prediction_for_single_spectra(flux, error, zabs, dev(mask))
        assert abs(float(ll) - float(k["ll" + sfx])) <= max(tol, 2e-6) * abs(float(k["ll" + sfx]))
        assert np.abs(hm.squeeze().cpu().numpy() - k["h" + sfx]).max() < max(tol, 1e-4)
        assert relerr(cont.cpu().numpy(), k["our" + sfx]) < max(tol, 5e-6)
    m.load_from_npz(str(path), reference_c0_bug=False)   # true c0: SURVEY section 4 quotes -714.22
    ll = m.prediction_for_single_spectra(flux, error, zabs, dev(k["mask"]))[0]
    assert abs(float(ll) - (-714.22)) < 0.05


# ----------------------------------------------------------------------------- oracle on larger inputs
def _synthetic(grid_name, Nh, B, seed, pretrained=None, **kw):
    from qfa_b200 import synth
    grid = synth.GRIDS[grid_name]
    if pretrained is None:
        P, mu = synth.smooth_random_params(grid, Nh, seed=seed)
    else:
        P, mu = pretrained
    data = synth.make_spectra(P, mu, grid, B, seed=seed, **kw)
    return grid, P, mu, data


def _sdss_pretrained():
    k = np.load(os.path.join(GOLD, "kat_sdss.npz"))
    P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
    P["c0"] = P["beta"].clone()
    return P, torch.tensor(k["param_mu"])


def _desi_pretrained():
    k = np.load(os.path.join(GOLD, "desi_params.npz"))
    P = {key: torch.tensor(k[key], dtype=torch.float32) for key in KEYS}
    return P, torch.tensor(k["mu"])


@pytest.mark.parametrize("shape", ["sdss", "l32", "desi"])
def test_fp64_against_lowrank_oracle_seeded(shape):
    from oracle import qfa_lowrank
    from qfa_b200 import QFA
    if shape == "sdss":
        grid, P, mu, d = _synthetic("sdss", 8, 160, 1234, _sdss_pretrained())
    elif shape == "l32":
        grid, P, mu, d = _synthetic("l32", 32, 96, 1237, mask_iid=0.15, run_len=(40, 160))
    else:
        grid, P, mu, d = _synthetic("desi", 8, 6, 1235, _desi_pretrained(), mask_iid=0.1)
    Pn = {k: v.numpy() for k, v in P.items()}
    m = QFA(grid.Nb, grid.Nr, Pn["F"].shape[1], torch.device("cuda:0"), model_params=Pn, precision="fp64")
    m.mu = mu
    loss, grads = m.forward(d["delta"].cuda(), d["error"].cuda(), d["zabs"].cuda(), d["mask"].cuda())
    a = split_acc(m, grads.acc)
    rl, rg, ex = qfa_lowrank.forward(Pn, d["delta"].numpy(), d["error"].numpy(), d["zabs"].numpy(), d["mask"].numpy(),
                                     grid.Nb, "becker", True)
    assert abs(a["nll_sum"] / a["nsp"] - rl) <= 1e-9 * abs(rl)
    ref_sums = np.concatenate([ex["sums"]["F"].ravel(), ex["sums"]["Psi"], ex["sums"]["omega"],
                               [ex["sums"]["tau0"], ex["sums"]["c0"], ex["sums"]["beta"]]])
    assert relerr(a["sums"], ref_sums) < TIGHT64
    assert np.array_equal(a["pix_cnt"], ex["counts"]["Psi"])
    assert relerr(a["dmu"], ex["dmu"]) < TIGHT64                 # extra output: d NLL / d mu
    for k in KEYS:
        assert relerr(grads[k].cpu().numpy(), rg[k]) < 1e-6, k
    o = m.predict_batch(d["flux"].cuda(), d["error"].cuda(), d["zabs"].cuda(), d["mask"].cuda())
    rn, rh, rc, rcont, runc = qfa_lowrank.predict_batch(Pn, mu.numpy(), d["flux"].numpy(), d["error"].numpy(),
                                                        d["zabs"].numpy(), d["mask"].numpy(), grid.Nb)
    assert relerr(o["nll"].cpu().numpy(), rn) < TIGHT64
    assert relerr(o["cont"].cpu().numpy(), rcont) < TIGHT64
    assert relerr(o["unc"].cpu().numpy(), runc) < 1e-7
    assert relerr(o["hcov"].cpu().numpy(), rc) < 1e-7
    # float modes on the same inputs: continuum bar 1e-3
    for prec in ("fp32", "mixed") + (("tf32",) if shape != "l32" else ()):
        m.precision = prec
        m._acc = None
        of = m.predict_batch(d["flux"].cuda(), d["error"].cuda(), d["zabs"].cuda(), d["mask"].cuda())
        assert relerr(of["cont"].cpu().numpy(), rcont) < 1e-3
        assert np.all(np.abs(of["nll"].cpu().numpy() - rn) <= 1e-3 * np.maximum(1, d["mask"].sum(1).numpy()))


# ----------------------------------------------------------------------------- properties at full size
@pytest.mark.parametrize("precision", ["fp64", "fp32", "mixed"])
def test_properties_at_baseline_size(precision):
    """SDSS shape, 6000 spectra (spans several sub-batches): results do not depend on how the batch is
    split (data-parallel == single GPU), on the order of the spectra, or on what is stored
    in masked pixels; counts are exact integers."""
    from qfa_b200 import QFA
    grid, P, mu, d = _synthetic("sdss", 8, 6000, 77, _sdss_pretrained())
    Pn = {k: v.numpy() for k, v in P.items()}
    m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision=precision)
    m.mu = mu
    D, E, Z, M = (d[k].cuda() for k in ("delta", "error", "zabs", "mask"))
    tol = 1e-10 if precision == "fp64" else 2e-4
    full = m.accumulate(D, E, Z, M).clone()
    # (1) two "ranks": shard, accumulate separately, sum == all-reduce
    h = 3000
    a0 = m.accumulate(D[:h], E[:h], Z[:h], M[:h]).clone()
    a1 = m.accumulate(D[h:], E[h:], Z[h:], M[h:]).clone()
    assert relerr((a0 + a1).cpu().numpy(), full.cpu().numpy()) < tol
    # (2) accumulate without zeroing == one call
    m.accumulate(D[:h], E[:h], Z[:h], M[:h], zero=True)
    a01 = m.accumulate(D[h:], E[h:], Z[h:], M[h:], zero=False).clone()
    assert relerr(a01.cpu().numpy(), full.cpu().numpy()) < tol
    # (3) permutation of the spectra
    perm = torch.randperm(6000, generator=torch.Generator().manual_seed(3)).cuda()
    ap = m.accumulate(D[perm], E[perm], Z[perm], M[perm]).clone()
    assert relerr(ap.cpu().numpy(), full.cpu().numpy()) < tol
    # (4) garbage (NaN / inf) in masked pixels is ignored
    D2, E2 = D.clone(), E.clone()
    D2[~M] = float("nan")
    E2[~M] = float("inf")
    ag = m.accumulate(D2, E2, Z, M).clone()
    assert torch.equal(ag, full)
    # (5) counts: integers equal to the mask column sums; #spectra
    a = split_acc(m, full)
    assert np.array_equal(a["pix_cnt"], M.sum(0).double().cpu().numpy())
    assert a["nsp"] == 6000 and a["scal_cnt"][0] == float((M[:, :grid.Nb].sum(1) > 0).sum())
    # (6) determinism: bitwise identical on repetition
    again = m.accumulate(D, E, Z, M).clone()
    assert torch.equal(again, full)
    # (7) predict: NLL of predict on flux equals NLL of the train path on delta = flux - mu*A
    o = m.predict_batch(d["flux"][:512].cuda(), E[:512], Z[:512], M[:512], want=("nll",))
    nll = torch.empty(512, dtype=m._tdtype, device="cuda")
    m.accumulate(D[:512], E[:512], Z[:512], M[:512], nll_out=nll)
    # (delta is stored as float32, so the two paths see inputs that differ by one float32 rounding)
    assert relerr(o["nll"].cpu().numpy(), nll.cpu().numpy()) < (1e-6 if precision == "fp64" else 2e-3)
    if precision == "mixed":      # 6000 >= QFA_TC_MIN_BATCH: this ran on the tensor-core kernels; tie it to fp64
        m64 = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="fp64")
        r = split_acc(m64, m64.accumulate(D, E, Z, M))
        n_f = m.Npix * m.Nh
        assert relerr(a["sums"][:n_f], r["sums"][:n_f]) < 2e-2 and relerr(a["dmu"], r["dmu"]) < 5e-3
        assert abs(a["nll_sum"] - r["nll_sum"]) <= 1e-4 * abs(r["nll_sum"])


def test_empty_and_ragged_batches():
    from qfa_b200 import QFA
    c = load_case("tiny5")
    m = QFA(c["Nb"], c["F"].shape[0] - c["Nb"], 5, torch.device("cuda:0"), model_params=c, precision="fp64")
    m.mu = c["mu"]
    z = lambda n, w, dt=torch.float32: torch.zeros(n, w, dtype=dt, device="cuda")
    acc = m.accumulate(z(0, m.Npix), z(0, m.Npix), z(0, m.Nb), z(0, m.Npix, torch.bool))
    assert float(acc.abs().sum()) == 0.0
    o = m.predict_batch(z(0, m.Npix), z(0, m.Npix), z(0, m.Nb), z(0, m.Npix, torch.bool))
    assert o["nll"].shape == (0,) and o["cont"].shape == (0, m.Npix)
    # all-masked spectrum: NLL 0, hmean 0, hcov I
    o = m.predict_batch(dev(c["flux"][:1]), dev(c["error"][:1]), dev(c["zabs"][:1]), z(1, m.Npix, torch.bool))
    assert float(o["nll"]) == 0.0 and float(o["hmean"].abs().max()) == 0.0
    assert torch.equal(o["hcov"][0], torch.eye(5, dtype=torch.float64, device="cuda"))
    # B = 1 ... 7 (ragged vs the 128-pixel / sub-batch tiling) agree with prefix sums of per-spectrum calls
    from qfa_b200._lib import QfaError
    with pytest.raises(QfaError):
        m.forward(dev(c["delta"][:, :-1]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
    run = None
    for b in range(7):
        one = m.accumulate(dev(c["delta"][b:b + 1]), dev(c["error"][b:b + 1]), dev(c["zabs"][b:b + 1]),
                           dev(c["mask"][b:b + 1])).clone()
        run = one if run is None else run + one
        many = m.accumulate(dev(c["delta"][:b + 1]), dev(c["error"][:b + 1]), dev(c["zabs"][:b + 1]),
                            dev(c["mask"][:b + 1])).clone()
        assert relerr(many.cpu().numpy(), run.cpu().numpy()) < 1e-12


# ----------------------------------------------------------------------------- optimiser and train()
def test_fused_adam_clip_and_train_match_reference_golden(tmp_path, cuda_model_factory):
    from qfa_b200 import Adam, step_scheduler
    g = dict(np.load(os.path.join(GOLD, "train_tiny_f32.npz")))
    c = load_case("train")

    def fresh():
        m = cuda_model_factory(c, "fp64")
        m.Psi = torch.ones(m.Npix)
        m.omega = torch.ones(m.Nb)
        return m
    D, E, Z, M = (dev(c[k]) for k in ("delta", "error", "zabs", "mask"))
    # (a) reference call sequence: forward -> optimizer.update -> parameters setter (model.py:212-214)
    m = fresh()
    opt = Adam(params=m.parameters, device=m.device, scheduler=step_scheduler(0.9, 2), learning_rate=1e-2,
               weight_decay=0.1)
    loss, grads = m.forward(D[:6], E[:6], Z[:6], M[:6])
    m.parameters = opt.update(m.parameters, grads)
    assert relerr(loss.cpu().numpy(), g["step_loss"]) < 1e-5
    for k in KEYS:
        assert relerr(m.parameters[k].cpu().numpy(), g["step_" + k]) < 1e-5, k
        assert relerr(opt.m[k].cpu().numpy(), g["step_m_" + k]) < 1e-4, k
        assert relerr(opt.v[k].cpu().numpy(), g["step_v_" + k]) < 1e-4, k

    # (b) 6-epoch train() crossing the smooth/save interval at epoch 5 (model.py:222-231)
    class Loader:
        def __init__(s):
            s.mu = c["mu"]; s.data_size = 12; s.batch_size = 6; s.cur = 0
        def rewind(s): s.cur = 0
        def have_next_batch(s): return s.cur < s.data_size
        def next_batch(s):
            a, b = s.cur, min(s.cur + s.batch_size, s.data_size); s.cur = b
            return D[a:b], E[a:b], Z[a:b], M[a:b]
    class ReferenceInterfaceOnly:            # hides update_from_acc -> QFA.train takes the reference's loop
        def __init__(s, inner): s.inner = inner
        def update(s, p, g_): return s.inner.update(p, g_)
        def step(s): s.inner.step()

    for use_fused_loop in (True, False):
        m = fresh()
        opt = Adam(params=m.parameters, device=m.device, scheduler=step_scheduler(0.9, 2), learning_rate=1e-2,
                   weight_decay=0.1)
        out = tmp_path / ("fused" if use_fused_loop else "generic")
        m.train(opt if use_fused_loop else ReferenceInterfaceOnly(opt), Loader(), 6, output_dir=str(out),
                save_interval=5, smooth_interval=5, quiet=True)
        ck = np.load(out / "checkpoints" / "model_parameters_epoch_05.npz")
        assert sorted(ck.files) == sorted(["mu", "F", "Psi", "omega", "tau0", "c0", "beta", "qfa_b200"])
        for k in KEYS:
            assert ck[k].dtype == np.float32
            assert relerr(ck[k], g["ckpt5_" + k]) < 5e-4, k
            assert relerr(m.parameters[k].cpu().numpy(), g["final_" + k]) < 5e-4, k


def test_smooth_clip_prepare_kernels(cuda_model_factory):
    import ctypes
    from qfa_b200 import _lib, synth
    g = dict(np.load(os.path.join(GOLD, "train_tiny_f32.npz")))
    c = load_case("train")
    m = cuda_model_factory(c, "fp32")
    m.smooth()
    for k in ("F", "Psi", "omega", "tau0"):
        assert relerr(m.parameters[k].cpu().numpy(), g["smooth_" + k]) < 1e-6, k
    m.Psi = torch.full((m.Npix,), 7.0)
    m.tau0 = torch.tensor(-1.0)
    m.clip()
    assert float(m.Psi.max()) == 2.0 and float(m.tau0) == 0.0
    # device-side batch preparation (dataloader.py:102,135-136) against the synthetic generator
    grid = synth.GridSpec("tiny", 1150.0, 6e-4, 96)
    P, mu = synth.smooth_random_params(grid, 4, seed=31)
    d = synth.make_spectra(P, mu, grid, 33, seed=9)
    zabs = torch.empty(33, grid.Nb, device="cuda")
    delta = torch.empty(33, grid.Npix, device="cuda")
    wav = torch.tensor(grid.wav(), dtype=torch.float32).cuda()
    L = _lib.lib()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    flux, zq, mu_d = d["flux"].cuda(), d["zqso"].cuda(), mu.cuda()
    _lib.check(L.qfa_prepare_batch(p(flux), p(zq), p(wav), p(mu_d), 33, grid.Nb, grid.Nr, 0, p(zabs), p(delta), None))
    torch.cuda.synchronize()
    assert relerr(zabs.cpu().numpy(), d["zabs"].numpy()) < 1e-6
    ok = d["mask"].numpy()
    assert np.abs(delta.cpu().numpy()[ok] - d["delta"].numpy()[ok]).max() < 1e-5


# ----------------------------------------------------------------------------- tensor-core train path for 8 < Nh <= 32
def test_tensor_core_nh32_train_path(cuda_model_factory):
    """8 < Nh <= 32 (zero-padded to 32): k_tc_gram32 (three tcgen05 passes over the 528 Khatri-Rao columns) + k_solve32 +
    k_tc_grad32 (per-spectrum stacked MMAs).
    (a) reference goldens forced onto the tensor-core path: the Npix 1000 / Nh 32 case at the usual tensor-core bound, and
        the 96-pixel Nh = 12 / 16 cases (random unit-scale factors: 8 pixels per factor, badly conditioned M) which pin the
        padding logic -- their gradF carries the amplified TF32 operand rounding (measured 0.18 on tiny12, where the
        REFERENCE'S OWN float32 run is 0.42 away from its float64 run: tests/golden/case_tiny12_f32 vs _f64), everything
        else <= 5e-2; the default float Cholesky and the opt-in double one (`solve_fp64`) give the same figures, because
        every step after the factorisation works with the triangular factors (cond(L) = sqrt(cond(M)));
    (b) ragged 3 001-spectra synthetic batches (1000 pixels; Nh = 12 and 32) against the float CUDA-core path of the same
        library (itself pinned to the goldens).  Single-pass TF32 operands (2^-11) and cond(M) ~ 1e3: the per-spectrum NLL
        (a cancelling difference, model.py:135) scatters by up to ~5.5e-3 per unmasked pixel (median 2.6e-4) WITHOUT bias --
        the batch loss agrees to 1e-4 per pixel; gradF to 3.3e-2 max-norm, the other gradients to <= 6.5e-3."""
    for name, tolF, f64 in (("l32", 5e-2, False), ("tiny12", 3e-1, False), ("tiny16", 3e-1, False), ("l32", 5e-2, True),
                            ("tiny12", 3e-1, True), ("tiny16", 3e-1, True)):
        c, g = load_case(name, "f64")
        m = cuda_model_factory(c, "tf32")
        m.solve_fp64 = f64
        loss, grads = m.forward(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
        npx = np.maximum(1, c["mask"].sum(1))
        assert abs(float(loss) - float(np.squeeze(g["loss"]))) <= 1e-2 * npx.mean(), name
        print(name, "double Cholesky" if f64 else "float Cholesky", {k: "%.1e" % relerr(grads[k].cpu().numpy(), g["grad_" + k]) for k in KEYS})
        for k in KEYS:
            assert relerr(grads[k].cpu().numpy(), g["grad_" + k]) < (tolF if k == "F" else 5e-2), (name, k)
    from qfa_b200 import QFA, synth
    grid = synth.GRIDS["l32"]
    for Nh in (12, 32):
        P, mu = synth.smooth_random_params(grid, Nh, seed=1237)
        d = synth.make_spectra(P, mu, grid, 3001, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
        Pn = {k: v.numpy() for k, v in P.items()}
        out = {}
        for prec in ("fp32", "tf32", "tf32-dchol"):
            mm = QFA(grid.Nb, grid.Nr, Nh, torch.device("cuda:0"), model_params=Pn, precision=prec.split("-")[0])
            mm.solve_fp64 = prec.endswith("dchol")
            nll = torch.empty(3001, device="cuda")
            mm.accumulate(d["delta"], d["error"], d["zabs"], d["mask"], nll_out=nll)
            l, gr = mm.forward(d["delta"], d["error"], d["zabs"], d["mask"])
            out[prec] = (float(l), {k: gr[k].cpu().numpy() for k in KEYS}, nll.cpu().numpy())
        npx = np.maximum(1, d["mask"].sum(1).cpu().numpy())
        # the ORACLE (fp64 low-rank restatement of the reference, oracle/qfa_lowrank.py) on the same 3 001 spectra: the
        # tensor-core path is tied to the reference directly, not only through the float kernels of this library
        from oracle import qfa_lowrank
        cpu = {k: d[k].cpu().numpy() for k in ("delta", "error", "zabs", "mask")}
        ol, og = qfa_lowrank.forward(Pn, cpu["delta"], cpu["error"], cpu["zabs"], cpu["mask"], grid.Nb)
        on = qfa_lowrank.nll_batch(Pn, cpu["delta"], cpu["error"], cpu["zabs"], cpu["mask"], grid.Nb)
        for prec in ("fp32", "tf32", "tf32-dchol"):
            dno = np.abs(out[prec][2] - on) / npx
            geo = {k: relerr(out[prec][1][k], og[k]) for k in KEYS}
            print("Nh %d %s vs fp64 oracle: loss diff / px %.2e, per-spectrum NLL / px max %.2e median %.2e, grads %s" % (
                Nh, prec, abs(out[prec][0] - ol) / npx.mean(), dno.max(), np.median(dno), {k: "%.1e" % v for k, v in geo.items()}))
            tc_path = prec != "fp32"
            assert abs(out[prec][0] - ol) <= (1e-4 if tc_path else 1e-6) * npx.mean(), prec
            assert dno.max() <= (1e-2 if tc_path else 5e-5) and np.median(dno) <= (1e-3 if tc_path else 2e-6), prec
            for k in KEYS:
                assert geo[k] < ((5e-2 if k == "F" else 1e-2) if tc_path else 1e-3), (prec, k)
        dn = np.abs(out["tf32"][2] - out["fp32"][2]) / npx
        ge = {k: relerr(out["tf32"][1][k], out["fp32"][1][k]) for k in KEYS}
        print("Nh %d tf32 vs fp32: loss diff / px %.2e, per-spectrum NLL / px max %.2e median %.2e, grads %s" % (
            Nh, abs(out["tf32"][0] - out["fp32"][0]) / npx.mean(), dn.max(), np.median(dn), {k: "%.1e" % v for k, v in ge.items()}))
        assert abs(out["tf32"][0] - out["fp32"][0]) <= 1e-4 * npx.mean()
        assert dn.max() <= 1e-2 and np.median(dn) <= 1e-3
        for k in KEYS:
            assert ge[k] < (5e-2 if k == "F" else 1e-2), k
        # opt-in double Cholesky: same bounds
        dn = np.abs(out["tf32-dchol"][2] - out["fp32"][2]) / npx
        ge = {k: relerr(out["tf32-dchol"][1][k], out["fp32"][1][k]) for k in KEYS}
        print("Nh %d double Cholesky: NLL / px max %.2e median %.2e, grads %s" % (Nh, dn.max(), np.median(dn),
                                                                                 {k: "%.1e" % v for k, v in ge.items()}))
        assert dn.max() <= 1e-2 and np.median(dn) <= 1e-3
        for k in KEYS:
            assert ge[k] < (5e-2 if k == "F" else 1e-2), k


def test_tensor_core_nh32_ragged_batches():
    """k_tc_gram32 / k_solve32 / k_tc_grad32 on batch sizes that exercise every ragged edge of their pipelines: fewer
    spectra than one step of 3, fewer steps than ring stages (5) or TMEM buffers (2), fewer steps than CTAs per pixel
    tile, a partial last tile of the Gram kernel (B % 120 != 0), B % 3 != 0.  Compared with the float CUDA-core path."""
    from qfa_b200 import QFA, synth
    grid = synth.GRIDS["l32"]
    P, mu = synth.smooth_random_params(grid, 32, seed=77)
    Pn = {k: v.numpy() for k, v in P.items()}
    dev0 = torch.device("cuda:0")
    ref = QFA(grid.Nb, grid.Nr, 32, dev0, model_params=Pn, precision="fp32")
    tc = QFA(grid.Nb, grid.Nr, 32, dev0, model_params=Pn, precision="tf32")
    dall = synth.make_spectra(P, mu, grid, 1300, seed=5, device=dev0, mask_iid=0.15, run_len=(40, 160))
    for B in (1, 2, 3, 4, 5, 7, 16, 31, 121, 359, 1300):
        a = [dall[k][:B].contiguous() for k in ("delta", "error", "zabs", "mask")]
        n0 = torch.empty(B, device="cuda"); n1 = torch.empty(B, device="cuda")
        ref.accumulate(*a, nll_out=n0); l0, g0 = ref.forward(*a)
        tc.accumulate(*a, nll_out=n1); l1, g1 = tc.forward(*a)
        npx = np.maximum(1, a[3].sum(1).cpu().numpy())
        assert torch.isfinite(n1).all(), B
        assert (np.abs((n1 - n0).cpu().numpy()) / npx).max() <= 1e-2, B
        assert abs(float(l1) - float(l0)) <= 2e-3 * npx.mean(), B
        for k in KEYS:
            x, y = g1[k].cpu().numpy(), g0[k].cpu().numpy()
            assert np.array_equal(np.isnan(x), np.isnan(y)), (B, k)          # 0/0 pixels (no unmasked spectrum) agree
            assert relerr(x, y) < (0.25 if k == "F" else 5e-2), (B, k)       # few spectra: little averaging of TF32 noise
        if B in (7, 1300):                                                   # fixed reduction orders: bitwise repeatable
            acc1 = tc.accumulate(*a).clone()
            assert torch.equal(tc.accumulate(*a), acc1), B


@pytest.mark.gpu
def test_tensor_core_nh32_cluster_gram_experiment(monkeypatch):
    """k_tc_gram32c (qfa_tc_gram32c.cuh): the Nh 32 Grams on a 4-CTA thread-block cluster -- columns split over the cluster,
    operand tiles broadcast through distributed shared memory, M2 accumulated as the difference M - Md.  Measured slower than
    the three-pass kernel and therefore only selected by QFA_GRAM32_CLUSTER=1; its results must agree with the default path
    (same TF32 operands; M2 / b2 / E are summed in a different order), on full, partial and sub-tile batches."""
    from qfa_b200 import QFA, synth
    grid = synth.GRIDS["l32"]
    P, mu = synth.smooth_random_params(grid, 32, seed=78)
    Pn = {k: v.numpy() for k, v in P.items()}
    dev0 = torch.device("cuda:0")
    tc = QFA(grid.Nb, grid.Nr, 32, dev0, model_params=Pn, precision="tf32")
    dall = synth.make_spectra(P, mu, grid, 2500, seed=6, device=dev0, mask_iid=0.15, run_len=(40, 160))
    for B in (5, 121, 2500):
        a = [dall[k][:B].contiguous() for k in ("delta", "error", "zabs", "mask")]
        monkeypatch.setenv("QFA_GRAM32_CLUSTER", "0")
        n0 = torch.empty(B, device="cuda"); tc.accumulate(*a, nll_out=n0); l0, g0 = tc.forward(*a)
        monkeypatch.setenv("QFA_GRAM32_CLUSTER", "1")
        n1 = torch.empty(B, device="cuda"); tc.accumulate(*a, nll_out=n1); l1, g1 = tc.forward(*a)
        monkeypatch.setenv("QFA_GRAM32_CLUSTER", "0")
        npx = np.maximum(1, a[3].sum(1).cpu().numpy())
        assert torch.isfinite(n1).all(), B
        assert (np.abs((n1 - n0).cpu().numpy()) / npx).max() <= 2e-3, B
        assert abs(float(l1) - float(l0)) <= 1e-3 * npx.mean(), B
        for k in KEYS:
            x, y = g1[k].cpu().numpy(), g0[k].cpu().numpy()
            assert np.array_equal(np.isnan(x), np.isnan(y)), (B, k)
            assert relerr(x, y) < 2e-2, (B, k)


# ----------------------------------------------------------------------------- device dataloader feeding QFA.train
def test_device_dataloader_on_gpu_and_train(tmp_path):
    """DeviceDataloader on the GPU (delta through qfa_prepare_batch) == the same loader on the CPU (torch restatement of
    dataloader.py:135-136), and QFA.train runs two epochs from it (model.py:183-231)."""
    from qfa_b200 import QFA, Adam, step_scheduler, DeviceDataloader
    from qfa_b200 import utils as U
    wav, Nb, Nr = U.wavelength_grid(1030.0, 1600.0, 2e-3)
    rng = np.random.default_rng(5)
    n, P = 96, len(wav)
    zq = rng.uniform(2.0, 3.5, n)
    flux = rng.normal(1.0, 0.2, (n, P)); err = rng.uniform(0.05, 0.2, (n, P))
    mask = rng.uniform(size=(n, P)) > 0.1
    flux[~mask] = -999.0; err[~mask] = -999.0
    lc = DeviceDataloader(flux, err, zq, mask, wav, batch_size=32, device="cpu", shuffle=False)
    lg = DeviceDataloader(flux, err, zq, mask, wav, batch_size=32, device="cuda", shuffle=False)
    assert relerr(lg.mu, lc.mu) < 1e-6
    lc.rewind(); lg.rewind()
    while lc.have_next_batch():
        dc, ec, zc, mc = lc.next_batch()
        dg, eg, zg, mg = lg.next_batch()
        assert relerr(dg.cpu().numpy()[mc.numpy()], dc.numpy()[mc.numpy()]) < 2e-6
        assert torch.equal(mg.cpu(), mc) and relerr(zg.cpu().numpy(), zc.numpy()) < 1e-6
    assert not lg.have_next_batch()
    torch.manual_seed(0)
    m = QFA(Nb, Nr, 4, torch.device("cuda:0"), precision="fp32")
    opt = Adam(params=m.parameters, device=torch.device("cuda:0"), scheduler=step_scheduler(0.9, 10), learning_rate=1e-3,
               weight_decay=0.1)
    m.train(opt, lg, 2, output_dir=str(tmp_path), save_interval=1, smooth_interval=5, quiet=True)
    ck = np.load(os.path.join(str(tmp_path), "checkpoints", "model_parameters_epoch_02.npz"))
    assert ck["F"].shape == (P, 4) and np.isfinite(ck["F"]).all() and relerr(ck["mu"], lc.mu) < 1e-6


# ----------------------------------------------------------------------------- 3xTF32 tensor-core mode
@pytest.mark.parametrize("name", TC_CASES)
def test_tf32x3_mode_against_fp64_golden(name, cuda_model_factory):
    """precision="tf32x3": the Grams, the 8 x 8 solve inputs and the continuum / sigma GEMM on the tensor cores with 3xTF32
    operand splitting (k_tc_gram_x3), per-spectrum algebra in double, gradient contraction by the float CUDA-core kernel.
    ONE set of bounds for every case (no per-case relaxation): what bounds the error now is float arithmetic of the
    per-cell physics and the fp32 accumulation -- the same floor the reference's own float32 program has (SURVEY 7.2:
    NLL 3.4e-6, gradF 1.6e-5, scalar gradients up to 1e-4)."""
    c, g = load_case(name, "f64")
    npx = np.maximum(1, c["mask"].sum(1))
    allerrs = {}
    for prec in ("tf32x3", "fp32", "tf32"):
        m = cuda_model_factory(c, prec)
        o = m.predict_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
        errs = {k: relerr(o[k].cpu().numpy(), g["pred_" + k]) for k in ("cont", "unc", "hmean", "hcov")}
        errs["nll/px"] = float((np.abs(o["nll"].cpu().numpy() - g["pred_nll"]) / npx).max())
        if prec == "tf32x3":
            nll_only = m.nll_batch(dev(c["flux"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
            assert torch.equal(nll_only, o["nll"])
        loss, grads = m.forward(dev(c["delta"]), dev(c["error"]), dev(c["zabs"]), dev(c["mask"]))
        errs["loss/px"] = abs(float(loss) - float(np.squeeze(g["loss"]))) / npx.mean()
        for k in KEYS:
            errs["g" + k] = relerr(grads[k].cpu().numpy(), g["grad_" + k])
        print(name, "%-6s" % prec, {k: "%.1e" % v for k, v in errs.items()})
        allerrs[prec] = errs
    errs = allerrs["tf32x3"]
    # continuum and likelihood: the 1e-5 bar of the fp64 mode, from the tensor cores (measured <= 4.5e-7 / 4.3e-6)
    assert errs["cont"] < 2e-6 and errs["nll/px"] < 1e-5 and errs["loss/px"] < 1e-5
    # quantities that go through M^-1 carry the float rounding of the Gram entries times cond(M): the float floor, the
    # same as the CUDA-core 'fp32' mode (printed above) and the reference's own float32 program (measured <= 1.2e-5,
    # gradF <= 4.1e-5 where the 'fp32' mode has 6.2e-5)
    assert errs["unc"] < 5e-5 and errs["hmean"] < 5e-5 and errs["hcov"] < 5e-5
    assert errs["gF"] < 1e-4 and errs["gPsi"] < 5e-5 and errs["gomega"] < 5e-5
    for k in ("tau0", "c0", "beta"):
        assert errs["g" + k] < 5e-5, k
    for k, v in errs.items():                                  # float-level: never far from the float CUDA-core mode
        assert v <= 3.0 * allerrs["fp32"][k] + 1e-5, (k, v, allerrs["fp32"][k])
    for k, v in errs.items():                                  # and always far better than single-pass TF32 (x 5 at least)
        assert v <= 0.2 * allerrs["tf32"][k] + 1e-6, (k, v, allerrs["tf32"][k])


def test_tf32x3_ragged_tiles_match_fp64_kernels():
    """SDSS shape, 300 + 5 spectra (partial tiles, P not a multiple of 16 or 128): 3xTF32 path vs the fp64 kernels."""
    from qfa_b200 import QFA, synth
    k = np.load(os.path.join(GOLD, "kat_sdss.npz"))
    P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
    P["c0"] = P["beta"].clone()
    mu = torch.tensor(k["param_mu"])
    grid = synth.GRIDS["sdss"]
    dev0 = torch.device("cuda:0")
    d = synth.make_spectra(P, mu, grid, 305, seed=3, device=dev0)
    Pn = {kk: v.numpy() for kk, v in P.items()}
    ref = QFA(grid.Nb, grid.Nr, 8, dev0, model_params=Pn, precision="fp64"); ref.mu = mu
    x3 = QFA(grid.Nb, grid.Nr, 8, dev0, model_params=Pn, precision="tf32x3"); x3.mu = mu
    for B in (305, 5, 1, 121):
        a = [d[kk][:B].contiguous() for kk in ("flux", "error", "zabs", "mask")]
        o0, o1 = ref.predict_batch(*a), x3.predict_batch(*a)
        for kk in ("cont", "unc", "hmean", "hcov"):
            assert relerr(o1[kk].cpu().numpy(), o0[kk].cpu().numpy()) < (1e-5 if kk == "cont" else 1e-4), (B, kk)
        npx = np.maximum(1, a[3].sum(1).cpu().numpy())
        assert (np.abs((o1["nll"].double() - o0["nll"]).cpu().numpy()) / npx).max() < 1e-5, B
        t = [d[kk][:B].contiguous() for kk in ("delta", "error", "zabs", "mask")]
        l0, g0 = ref.forward(*t)
        l1, g1 = x3.forward(*t)
        assert abs(float(l1) - float(l0)) < 1e-5 * npx.mean()
        for kk in KEYS:
            assert relerr(g1[kk].cpu().numpy(), g0[kk].cpu().numpy()) < (5e-4 if kk in ("tau0", "c0", "beta") else 1e-4), (B, kk)
    acc1 = x3.accumulate(*t).clone()
    assert torch.equal(x3.accumulate(*t), acc1)                 # deterministic


# ----------------------------------------------------------------------------- prediction for 8 < Nh <= 32 on the tensor cores
def test_tensor_core_nh32_predict_path(cuda_model_factory):
    """model.py:160-180 at the config-5 shape through k_tc_gram32<PRED> x 3 (3xTF32 Grams) + k_solve32<PRED> + k_out32
    (precision="tf32" forces the tensor-core kernels).  (a) reference goldens (Npix 1000 / Nh 32, and the badly conditioned
    96-pixel Nh 12 / 16 cases that pin the zero padding of Nh); (b) 3 001 ragged synthetic spectra (Nh 12 and 32) against
    the fp64 ORACLE (oracle/qfa_lowrank.py); NLL-only scoring = same kernels without k_out32.  Measured (scripts/
    p32_errors.py): continuum <= 4.8e-4 (the stated mixed-mode bar is 1e-3), hmean / hcov / sigma <= 1.2e-3, where single-pass
    TF32 Grams gave 1.6e-2 on the continuum of the Nh = 32 model (cond(M) ~ 1e3)."""
    for name in ("l32", "tiny12", "tiny16"):
        c, g = load_case(name, "f64")
        m = cuda_model_factory(c, "tf32")
        a = [dev(c[k]) for k in ("flux", "error", "zabs", "mask")]
        o = m.predict_batch(*a)
        npx = np.maximum(1, c["mask"].sum(1))
        e = {k: relerr(o[k].cpu().numpy(), g["pred_" + k]) for k in ("cont", "unc", "hmean", "hcov")}
        e["nll/px"] = float((np.abs(o["nll"].cpu().numpy() - g["pred_nll"]) / npx).max())
        print(name, "predict tf32 (Nh > 8):", {k: "%.1e" % v for k, v in e.items()})
        assert e["cont"] < 1e-3                                         # the stated mixed-mode bar, on every case
        assert e["unc"] < 3e-3 and e["hmean"] < 3e-3 and e["hcov"] < 3e-3 and e["nll/px"] < 5e-4
        assert torch.equal(m.nll_batch(*a), o["nll"])
        assert o["hcov"].shape == (len(npx), m.Nh, m.Nh) and o["cont"].shape == (len(npx), m.Npix)
    from qfa_b200 import QFA, synth
    from oracle import qfa_lowrank
    grid = synth.GRIDS["l32"]
    dev0 = torch.device("cuda:0")
    for Nh in (12, 32):
        P, mu = synth.smooth_random_params(grid, Nh, seed=1237)
        d = synth.make_spectra(P, mu, grid, 3001, seed=11, device=dev0, mask_iid=0.15, run_len=(40, 160))
        Pn = {k: v.numpy() for k, v in P.items()}
        a = [d[k] for k in ("flux", "error", "zabs", "mask")]
        cpu = [t.cpu().numpy() for t in a]
        rn, rh, rc, rcont, runc = qfa_lowrank.predict_batch(Pn, mu.numpy(), *cpu, grid.Nb)
        npx = np.maximum(1, cpu[3].sum(1))
        for prec in ("fp32", "tf32"):
            mm = QFA(grid.Nb, grid.Nr, Nh, dev0, model_params=Pn, precision=prec)
            mm.mu = mu
            o = mm.predict_batch(*a)
            e = {"cont": relerr(o["cont"].cpu().numpy(), rcont), "unc": relerr(o["unc"].cpu().numpy(), runc),
                 "hmean": relerr(o["hmean"].cpu().numpy(), rh), "hcov": relerr(o["hcov"].cpu().numpy(), rc),
                 "nll/px": float((np.abs(o["nll"].cpu().numpy() - rn) / npx).max())}
            print("Nh %d predict %s vs fp64 oracle:" % (Nh, prec), {k: "%.1e" % v for k, v in e.items()})
            tc_path = prec == "tf32"
            assert e["cont"] < (1e-3 if tc_path else 1e-4)
            assert e["unc"] < (3e-3 if tc_path else 1e-4) and e["hmean"] < (3e-3 if tc_path else 3e-4)
            assert e["hcov"] < (3e-3 if tc_path else 3e-4) and e["nll/px"] < (1e-3 if tc_path else 5e-5)
        # ragged batch sizes through the tensor-core kernels (forced): fewer spectra than a step, a ring, a tile
        tc = QFA(grid.Nb, grid.Nr, Nh, dev0, model_params=Pn, precision="tf32"); tc.mu = mu
        full = tc.predict_batch(*a)
        for B in (1, 2, 4, 7, 121, 359):
            ob = tc.predict_batch(*[t[:B].contiguous() for t in a])
            for k in ("nll", "hmean", "hcov", "cont", "unc"):
                assert torch.isfinite(ob[k]).all(), (B, k)
                assert relerr(ob[k].cpu().numpy(), full[k][:B].cpu().numpy()) < 3e-3, (Nh, B, k)   # tile heights differ: not bitwise
        # 'mixed' switches to the tensor cores above the measured cross-over (640 spectra for Nh > 16, 4480 for Nh <= 16)
        mx = QFA(grid.Nb, grid.Nr, Nh, dev0, model_params=Pn, precision="mixed"); mx.mu = mu
        om = mx.predict_batch(*a)
        same_as_tc = torch.equal(om["cont"], full["cont"])
        assert same_as_tc == (Nh > 16)


# ----------------------------------------------------------------------------- ill-conditioned inputs
def test_high_signal_to_noise_spectra_all_float_modes():
    """Ill-conditioned inputs: Psi and omega at their clip floor 1e-3 (model.py:237-238), quoted errors 20x
    smaller than usual (S/N 40-400) -- large weights 1/D -- and a factor matrix with two nearly collinear pairs of columns
    (with the pretrained F alone cond(I + F^T W F) saturates at ~30 however large the weights): cond(M) ~ 1e4, and every float
    path loses accuracy in proportion.  The float CUDA-core mode (the default) and
    the 3xTF32 tensor-core mode must still hold the continuum to 1e-4; single-pass TF32 ('tf32', and 'mixed' above its
    cross-over) is the opt-in speed mode and its error on such data is measured and bounded here, not hidden."""
    from qfa_b200 import QFA, synth
    P, mu = _sdss_pretrained()
    F2 = P["F"].clone()
    F2[:, 1] = F2[:, 0] + 0.03 * F2[:, 1]                # two nearly collinear factors: the Gram itself becomes ill-conditioned
    F2[:, 3] = F2[:, 2] - 0.05 * F2[:, 3]
    P = dict(P, F=F2, Psi=torch.full_like(P["Psi"], 1e-3), omega=torch.full_like(P["omega"], 1e-3))
    grid = synth.GRIDS["sdss"]
    dev0 = torch.device("cuda:0")
    d = synth.make_spectra(P, mu, grid, 600, seed=21, device=dev0)
    err = torch.where(d["mask"], d["error"] * 0.05, d["error"])
    Pn = {k: v.numpy() for k, v in P.items()}
    a = [d["flux"], err, d["zabs"], d["mask"]]
    ref = QFA(grid.Nb, grid.Nr, 8, dev0, model_params=Pn, precision="fp64"); ref.mu = mu
    o0 = ref.predict_batch(*a)
    cond = torch.linalg.cond(torch.linalg.inv(o0["hcov"])).cpu().numpy()
    print("cond(M): median %.1e max %.1e" % (np.median(cond), cond.max()))
    assert np.median(cond) > 1e3, cond
    res = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        m = QFA(grid.Nb, grid.Nr, 8, dev0, model_params=Pn, precision=prec); m.mu = mu
        o = m.predict_batch(*a)
        res[prec] = {k: relerr(o[k].cpu().numpy(), o0[k].cpu().numpy()) for k in ("cont", "unc", "hmean", "hcov")}
        print("high S/N", prec, {k: "%.1e" % v for k, v in res[prec].items()})
    # measured at cond(M) ~ 9e3: fp32 and 3xTF32 sit on the same float floor (continuum 1e-5, hmean 1e-3, hcov 2e-3 ~ eps x
    # cond); single-pass TF32 still holds the continuum to 7.5e-4 but hmean / hcov / sigma are 6e-2 / 1.4e-1 / 1.8e-1 off
    for prec in ("fp32", "tf32x3"):
        assert res[prec]["cont"] < 1e-4 and res[prec]["hmean"] < 5e-3 and res[prec]["hcov"] < 1e-2, prec
    assert res["tf32"]["cont"] < 1e-2 and res["tf32"]["hmean"] < 0.3   # degrades with cond(M): the documented price of the opt-in mode


def test_data_parallel_sum_on_two_gpus():
    """N > 1 on hardware (runs where two devices are visible, e.g. `gpurun --gpus 2`): two shards accumulated on two
    devices, summed, equal the single-device accumulation of the whole batch -- counts exactly, sums to round-off.
    (The NCCL path itself is asserted inside every multi-GPU bench run: bench.py dp_parity.)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from qfa_b200 import QFA, synth
    P, mu = _sdss_pretrained()
    grid = synth.GRIDS["sdss"]
    Pn = {k: v.numpy() for k, v in P.items()}
    d = synth.make_spectra(P, mu, grid, 2400, seed=8, device=torch.device("cuda:0"))
    keys = ("delta", "error", "zabs", "mask")
    m0 = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32")
    m1 = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:1"), model_params=Pn, precision="tf32")
    whole = m0.accumulate(*[d[k] for k in keys]).clone()
    a0 = m0.accumulate(*[d[k][:1200].contiguous() for k in keys]).clone()
    a1 = m1.accumulate(*[d[k][1200:].to("cuda:1").contiguous() for k in keys]).to("cuda:0")
    s = a0 + a1
    n = m0.Nparams
    assert torch.equal(s[n:n + grid.Npix + 3], whole[n:n + grid.Npix + 3]) and float(s[n + grid.Npix + 4]) == 2400.0
    assert relerr(s[:n].cpu().numpy(), whole[:n].cpu().numpy()) < 1e-5
