"""CPU: host-side mirror of the reference API (state container, npz layout, tau-law resolution,
optimizer bookkeeping).  The likelihood/gradient/prediction path must refuse to run without CUDA."""
import os
from functools import partial

import numpy as np
import pytest
import torch

from conftest import GOLD, load_case, relerr
from qfa_b200 import QFA, Adam, step_scheduler, QfaError
from qfa_b200 import utils as U
from qfa_b200 import synth

cpu = torch.device("cpu")


def test_constructor_matches_reference_defaults():
    m = QFA(720, 1193, 8, cpu)
    assert (m.Npix, m.Nparams) == (1913, 1913 * 8 + 1913 + 720 + 3)            # model.py:41-42
    assert m.F.shape == (1913, 8) and m.Psi.shape == (1913,) and m.omega.shape == (720,)
    assert m.tau0.shape == () and float(m.tau0) == pytest.approx(0.02)         # quirk Q9: code, not docstring
    assert float(m.c0) == pytest.approx(0.3) and float(m.beta) == pytest.approx(2.0)
    assert float(m.F.min()) >= -0.5 and float(m.F.max()) <= 0.5
    assert set(m.parameters) == {"F", "Psi", "omega", "tau0", "c0", "beta"}
    assert m.mu is None and m.min_value == 1e-3 and m.max_value == 2.


def test_parameters_setter_clips_like_reference():
    m = QFA(5, 7, 2, cpu)
    p = {k: v.clone() for k, v in m.parameters.items()}
    p["Psi"][:] = 5.0; p["omega"][:] = -1.0
    p["tau0"] = torch.tensor(3.0); p["beta"] = torch.tensor(0.0); p["c0"] = torch.tensor(-9.0)
    m.parameters = p                                                            # model.py:308-316
    assert float(m.Psi.max()) == 2.0 and float(m.omega.min()) == pytest.approx(1e-3)
    assert float(m.tau0) == 1.0 and float(m.beta) == pytest.approx(0.1) and float(m.c0) == -5.0


def test_npz_roundtrip_and_c0_quirk(tmp_path):
    k = np.load(os.path.join(GOLD, "kat_sdss.npz"))
    path = tmp_path / "params.npz"
    np.savez(path, **{key[6:]: k[key] for key in k.files if key.startswith("param_")})
    m = QFA(720, 1193, 8, cpu)
    with pytest.warns(UserWarning, match="c0 <- beta"):                         # reference-written file: quirk, announced
        m.load_from_npz(str(path))
    assert m.F.dtype == torch.float32                                           # file's F is float64
    assert float(m.c0) == float(m.beta) == pytest.approx(float(k["param_beta"]))   # quirk Q1 (model.py:295)
    m2 = QFA(720, 1193, 8, cpu)
    m2.load_from_npz(str(path), reference_c0_bug=False)
    assert float(m2.c0) == pytest.approx(float(k["param_c0"]))
    m.save_to_npz(str(tmp_path), "out.npz")
    o = np.load(tmp_path / "out.npz")
    ref_keys = ["mu", "F", "Psi", "omega", "tau0", "c0", "beta"]                # model.py:280
    assert sorted(o.files) == sorted(ref_keys + ["qfa_b200"])                   # + the marker of this package
    assert all(o[f].dtype == np.float32 for f in ref_keys) and o["tau0"].shape == ()
    assert np.array_equal(o["F"], k["param_F"].astype(np.float32))
    # a checkpoint written by this package round-trips: the stored c0 comes back (no quirk, no warning) ...
    m2.save_to_npz(str(tmp_path), "own.npz")
    m3 = QFA(720, 1193, 8, cpu)
    import warnings as W
    with W.catch_warnings():
        W.simplefilter("error")
        m3.load_from_npz(str(tmp_path / "own.npz"))
    assert float(m3.c0) == float(m2.c0) != float(m3.beta)
    for key in ("F", "Psi", "omega", "tau0", "c0", "beta"):
        assert torch.equal(getattr(m3, key), getattr(m2, key))
    # ... and the quirk can still be forced on it
    m3.load_from_npz(str(tmp_path / "own.npz"), reference_c0_bug=True)
    assert float(m3.c0) == float(m3.beta)


def test_default_precision_is_the_reference_arithmetic():
    assert QFA(5, 7, 2, cpu).precision == "fp32"      # TF32 tensor-core modes are opt-in ('mixed', 'tf32', 'tf32x3')
    for p in ("fp64", "mixed", "tf32", "tf32x3"):
        assert QFA(5, 7, 2, cpu, precision=p).precision == p
    with pytest.raises(QfaError):
        QFA(5, 7, 2, cpu, precision="bf16")


def test_tau_total_matches_reference_golden():
    """utils.tau_total (host helper; the device path is qfa_gather_prepare) against the REAL reference's tau_total
    (oracle/make_golden_prep.py): all 30 Lyman lines on a grid that starts at 910 A, Ly-alpha only redward of Ly-beta."""
    for name in ("prep_desi_like", "prep_sdss_like"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        t = U.tau_total(g["wav"], g["zqso"], which=str(g["law"]))
        assert t.shape == g["taus"].shape and np.abs(t - g["taus"]).max() < 1e-14
    with pytest.raises(ValueError):
        U.tau_total(np.array([1300.0, 1400.0]), np.array([2.5]))                # utils.py:191-192


def test_device_dataloader_host_logic_matches_reference_golden():
    """DeviceDataloader on a CPU device (torch arithmetic, the host mirror of the kernels): mu and delta of the reference's
    dataloader.py:102,109-112,135 with the multi-series optical depth, no shuffle."""
    from qfa_b200 import DeviceDataloader
    for name in ("prep_desi_like", "prep_sdss_like"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        n = len(g["zqso"])
        dl = DeviceDataloader(g["flux"], g["error"], g["zqso"], g["mask"], g["wav"], batch_size=n, device="cpu",
                              tau=str(g["law"]), shuffle=False)
        assert (dl.Nb, dl.data_size, len(dl)) == (int(g["Nb"]), n, n)
        assert relerr(dl.mu, g["mu"]) < 2e-6                                    # float32 flux / zqso vs the float64 reference
        dl.rewind()
        d, e, z, m = dl.next_batch()
        assert not dl.have_next_batch()
        ok = g["mask"]
        assert np.abs(d.numpy() - g["delta"])[ok].max() < 5e-6 * np.abs(g["delta"][ok]).max()
        assert relerr(z.numpy(), g["zabs"]) < 1e-6 and np.array_equal(m.numpy(), g["mask"])


def test_tau_law_resolution():
    assert U.resolve_tau_law(U.default_tau) == 0
    assert U.resolve_tau_law(partial(U.tau, which="kamble")) == 2
    assert U.resolve_tau_law("mock") == 3
    with pytest.raises(QfaError):
        U.resolve_tau_law(lambda z: z)
    with pytest.raises(QfaError):
        U.resolve_tau_law(partial(U.tau, which="nope"))
    z = torch.tensor([2.0, 3.0])
    assert torch.allclose(U.tau(z, "becker"), 0.751 * ((1 + z) / 4.5) ** 2.9 - 0.132)
    wav, Nb, Nr = U.wavelength_grid()
    assert (len(wav), Nb, Nr) == (1913, 720, 1193)                              # dataloader.py:61-63
    for g, shape in (("sdss", (1913, 720)), ("l32", (1000, 377)), ("desi", (9243, 2238))):
        assert (synth.GRIDS[g].Npix, synth.GRIDS[g].Nb) == shape


def test_no_cpu_fallback_for_hot_path():
    c = load_case("tiny5")
    m = QFA(c["Nb"], c["F"].shape[0] - c["Nb"], 5, cpu, model_params=c)
    m.mu = c["mu"]
    T = torch.tensor
    with pytest.raises(QfaError):
        m.forward(T(c["delta"]), T(c["error"]), T(c["zabs"]), T(c["mask"]))
    with pytest.raises(QfaError):
        m.prediction_for_single_spectra(T(c["flux"][0]), T(c["error"][0]), T(c["zabs"][0]), T(c["mask"][0]))
    with pytest.raises(QfaError):
        QFA(10, 10, 33, cpu)


def test_adam_dict_path_and_scheduler_match_reference_golden():
    """optimizer.py:47-52,98 through the torch-op path (dicts of ordinary tensors)."""
    g = dict(np.load(os.path.join(GOLD, "train_tiny_f32.npz")))
    c = load_case("train")
    keys = ("F", "Psi", "omega", "tau0", "c0", "beta")
    sch = step_scheduler(0.9, 2)
    assert sch(0, 1.0) == 1.0 and sch(1, 1.0) == 0.9 and sch(3, 1.0) == pytest.approx(0.81)
    from oracle import qfa_dense
    P = {k: torch.tensor(c[k]) for k in keys}
    P["Psi"] = torch.ones_like(P["Psi"]); P["omega"] = torch.ones_like(P["omega"])
    T = torch.tensor
    _, grads = qfa_dense.forward(P, T(c["delta"][:6]), T(c["error"][:6]), T(c["zabs"][:6]), T(c["mask"][:6]),
                                 c["Nb"], c["law"])
    opt = Adam(params=P, device=cpu, scheduler=sch, learning_rate=1e-2, weight_decay=0.1)
    m = QFA(c["Nb"], c["F"].shape[0] - c["Nb"], 4, cpu)
    m.parameters = opt.update(P, grads)
    for k in keys:
        assert relerr(m.parameters[k].numpy(), g["step_" + k]) < 1e-5, k
    opt.step()
    assert opt.i == 1 and opt.scheduled_lr == pytest.approx(1e-2 * 0.9)
    opt.reset(P)
    assert opt.i == 0 and float(opt.m["F"].abs().max()) == 0.0


def test_smooth_host_matches_reference_golden():
    g = dict(np.load(os.path.join(GOLD, "train_tiny_f32.npz")))
    c = load_case("train")
    m = QFA(c["Nb"], c["F"].shape[0] - c["Nb"], 4, cpu, model_params=c)
    m.smooth()
    for k in ("F", "Psi", "omega"):
        assert relerr(m.parameters[k].numpy(), g["smooth_" + k]) < 1e-6


def test_synthetic_loader_protocol_and_sharding():
    grid = synth.GridSpec("t", 1150.0, 6e-4, 96)
    P, mu = synth.smooth_random_params(grid, 4, seed=1)
    data = synth.make_spectra(P, mu, grid, 40, seed=2)
    assert data["mask"].dtype == torch.bool and data["zabs"].shape == (40, grid.Nb)
    assert bool((data["flux"][~data["mask"]] == -999).all())
    seen = []
    for rank in range(2):
        ld = synth.SyntheticLoader(data, mu, batch_size=8, rank=rank, world=2, seed=5)
        assert ld.data_size == 40 and ld.batch_size == 8 and len(ld.mu) == 96
        ld.rewind()
        n = 0
        while ld.have_next_batch():
            d, e, z, m = ld.next_batch()
            assert d.shape[0] <= 4 and z.shape[1] == grid.Nb
            n += d.shape[0]
            seen.append(d)
        assert n == 20
    allrows = torch.cat(seen)
    assert allrows.shape[0] == 40
    assert torch.equal(torch.sort(allrows[:, 0])[0], torch.sort(data["delta"][:, 0])[0])


# ----------------------------------------------------------------------------- data parallel host logic (gloo, 2 ranks)
def _pack_acc(model, extra, nsp):
    """oracle sums / counts -> the library's `acc` layout (include/qfa_b200.h)."""
    s, c = extra["sums"], extra["counts"]
    parts = [np.asarray(s["F"]).ravel(), np.asarray(s["Psi"]), np.asarray(s["omega"]),
             np.array([s["tau0"], s["c0"], s["beta"]]), np.asarray(c["Psi"]),
             np.array([c["tau0"], c["c0"], c["beta"]], dtype=np.float64),
             np.array([extra["nll"].sum(), float(nsp)]), np.asarray(extra["dmu"])]
    acc = torch.tensor(np.concatenate([np.asarray(p, np.float64).ravel() for p in parts]))
    assert acc.numel() == model.Nparams + model.Npix + 3 + 2 + model.Npix
    return acc


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import qfa_lowrank          # checker only
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        grid = synth.GridSpec("t", 1150.0, 6e-4, 96)
        P, mu = synth.smooth_random_params(grid, 4, seed=1)
        data = synth.make_spectra(P, mu, grid, 24, seed=2)
        data["mask"][:, 5] = False                       # a pixel masked everywhere: 0/0 -> NaN on every rank alike
        Pn = {k: np.asarray(v) for k, v in P.items()}
        m = QFA(grid.Nb, grid.Nr, 4, cpu, model_params=Pn)
        m.enable_data_parallel()
        assert m._dp
        # the peer-memory kernel is a CUDA path: a CPU model keeps the torch.distributed collective, and says so
        assert m._peer is None and m.allreduce_kind == "torch.distributed all_reduce"
        ld = synth.SyntheticLoader(data, mu, batch_size=24, rank=rank, world=world, seed=3, shuffle=False)
        ld.rewind()
        d, e, z, k = ld.next_batch()                     # this rank's half of the global batch
        assert d.shape[0] == 12 and not ld.have_next_batch()
        _, _, ex = qfa_lowrank.forward(Pn, d.numpy(), e.numpy(), z.numpy(), k.numpy(), grid.Nb, return_sums=True)
        acc = _pack_acc(m, ex, d.shape[0])
        m._allreduce(acc)                                # the ONE collective of a train step
        n = m.Nparams
        with np.errstate(divide="ignore", invalid="ignore"):
            a = acc.numpy()
            gF = a[:m.Npix * m.Nh].reshape(m.Npix, m.Nh) / a[n:n + m.Npix, None]     # divide AFTER the reduce
            gt0 = a[n - 3] / a[n + m.Npix]
        loss = float(m._loss_from_acc(acc))
        _, gref, _ = qfa_lowrank.forward(Pn, data["delta"].numpy(), data["error"].numpy(), data["zabs"].numpy(),
                                         data["mask"].numpy(), grid.Nb, return_sums=True)
        lref, _ = qfa_lowrank.forward(Pn, data["delta"].numpy(), data["error"].numpy(), data["zabs"].numpy(),
                                      data["mask"].numpy(), grid.Nb)
        q.put((rank, relerr(gF, gref["F"]), abs(gt0 - gref["tau0"]) / abs(gref["tau0"]), abs(loss - lref) / abs(lref),
               float(a[n + m.Npix + 4])))
    finally:
        dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2():
    """N>1 host path on CPU: shard by rank, all-reduce the packed sums/counts over gloo, divide after the
    reduce (quirk Q4) == single-process result on the global batch."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, eF, et0, el, nsp in res:
        assert eF < 1e-12 and et0 < 1e-12 and el < 1e-12, (rank, eF, et0, el)
        assert nsp == 24.0


# ----------------------------------------------------------------------------- device dataloader (host logic on CPU)
def test_device_dataloader_matches_reference_formulas():
    """zabs, mu and delta of qfa_b200.dataloader.DeviceDataloader against a NumPy restatement of reference
    QFA/dataloader.py:102,110-112,135-136 and QFA/utils.py:206-219 (smooth)."""
    from qfa_b200.dataloader import DeviceDataloader
    wav, Nb, Nr = U.wavelength_grid(1030.0, 1600.0, 2e-3)
    rng = np.random.default_rng(3)
    n, P = 64, len(wav)
    zq = rng.uniform(2.0, 3.5, n)
    flux = rng.normal(1.0, 0.2, (n, P))
    err = rng.uniform(0.05, 0.2, (n, P))
    mask = rng.uniform(size=(n, P)) > 0.1
    flux[~mask] = -999.0
    ld = DeviceDataloader(flux, err, zq, mask, wav, batch_size=16, device="cpu", window_length_for_mu=16, shuffle=False)
    # ---- NumPy restatement
    zabs = (zq + 1).reshape(-1, 1) * wav[:Nb] / 1215.67 - 1                                   # dataloader.py:102
    tau = 0.751 * ((1 + zabs) / 4.5) ** 2.90 - 0.132                                          # utils.py:105-106 (series 1)
    s = np.hstack((np.exp(tau), np.ones((n, Nr))))
    mu = np.sum(flux * s * mask, axis=0) / np.sum(flux != -999., axis=0)                      # dataloader.py:111
    w = 16
    sp = np.r_[mu[w - 1:0:-1], mu, mu[-2:-w - 1:-1]]                                          # utils.py:216-219
    mu_s = np.convolve(np.ones(w) / w, sp, mode='valid')[int(w / 2 - 1):-int(w / 2)]
    assert ld.mu.shape == (P,) and relerr(ld.mu, mu_s) < 1e-6
    ld.rewind()
    d, e, z, m = ld.next_batch()
    assert relerr(z.numpy(), zabs[:16]) < 1e-6
    s_b = np.hstack((np.exp(-tau[:16]), np.ones((16, Nr))))
    assert relerr(d.numpy(), flux[:16] - mu_s * s_b) < 1e-5                                   # dataloader.py:135-136
    assert m.dtype == torch.bool and z.shape == (16, Nb) and e.shape == (16, P)
    cnt = 16
    while ld.have_next_batch():
        cnt += ld.next_batch()[0].shape[0]
    assert cnt == n


def test_separable_optical_depth_identity_of_the_gather_kernel():
    """k_gather_prepare<TABLE> (csrc/qfa_aux.cuh) evaluates  tau_total(i, row) = (1 + zq)^be * T1[i] + T0[i]  with per-pixel tables
    T1[i] = t0 * sum_s c_s (wav_i / (lambda_s zn))^be,  T0[i] = C * sum_s c_s  over the Lyman lines redward of pixel i.  Pinned here
    on the CPU: (a) the identity against tau_total (utils.py:174-203 of the reference, restated in qfa_b200/utils.py and itself pinned
    to the reference's output by tests/golden/prep_*.npz) in float64, all four laws, a DESI-like grid whose bluest pixels lie
    below the Lyman limit; (b) the same two formulas in float32 arithmetic, as the kernels evaluate them, agree to a few ulp of
    exp(-tau) -- the bound the GPU test asserts between the table and the direct kernel."""
    from qfa_b200.utils import LAW_CONSTANTS, _LYMAN, series_coeff, tau_total
    wav = 910.0 * 10 ** (5.62e-5 * np.arange(2238))
    zq = np.array([2.0, 2.7, 3.5])
    lam = np.array([l for _, l in _LYMAN])
    coef = np.array([series_coeff(s + 1) for s in range(len(_LYMAN))])
    for which, (t0, be, C, zn) in LAW_CONSTANTS.items():
        ref = tau_total(wav, zq, which=which)                                   # (3, Nb)
        nb = ref.shape[1]
        red = wav[:nb, None] < lam[None, :]                                      # lines redward of each pixel
        T1 = t0 * np.sum(np.where(red, coef[None, :] * (wav[:nb, None] / (lam[None, :] * zn)) ** be, 0.0), axis=1)
        T0 = C * np.sum(np.where(red, coef[None, :], 0.0), axis=1)
        sep = (1.0 + zq)[:, None] ** be * T1[None, :] + T0[None, :]
        assert np.abs(sep - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), which
        # float32, in the kernels' order of operations
        f = np.float32
        w32, l32, c32 = wav[:nb].astype(f), lam.astype(f), coef.astype(f)
        direct = np.zeros((len(zq), nb), f)
        t1 = np.zeros(nb, f); t0s = np.zeros(nb, f)
        for s in range(len(lam)):
            on = w32 < l32[s]
            if not on.any():
                break
            for r, z in enumerate(zq):
                z1 = f(1.0 + z) * w32 / l32[s]
                direct[r] += np.where(on, (f(t0) * np.power(z1 / f(zn), f(be)) + f(C)) * c32[s], f(0))
            t1 += np.where(on, c32[s] * np.power(w32 / (l32[s] * f(zn)), f(be)), f(0))
            t0s += np.where(on, c32[s], f(0))
        for r, z in enumerate(zq):
            g = np.power(f(1.0 + z), f(be))
            table = g * (f(t0) * t1) + f(C) * t0s
            assert np.abs(np.exp(-table) - np.exp(-direct[r])).max() < 1e-6, (which, z)
