"""GPU: the rows either side of the hot path -- device data preparation with the multi-series
optical depth, device shuffle, the CUDA-graph train step, OOD selection, posterior sampling, the columnar predict writer
-- each against the reference-generated goldens (oracle/make_golden_prep.py) or an independent torch/numpy statement."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_case, relerr

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["prep_desi_like", "prep_sdss_like"])
def test_gather_prepare_and_mean_spectrum_match_reference_golden(name):
    """qfa_mean_spectrum_sums + qfa_gather_prepare (float32, in-kernel tau_total over all 30 Lyman lines) against the REAL
    reference's tau_total / smooth / dataloader arithmetic (float64)."""
    from qfa_b200 import DeviceDataloader
    dev = _dev()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n = len(g["zqso"])
    dl = DeviceDataloader(g["flux"], g["error"], g["zqso"], g["mask"], g["wav"], batch_size=n, device=dev,
                          tau=str(g["law"]), shuffle=False)
    assert relerr(dl.mu, g["mu"]) < 5e-6
    dl.rewind()
    d, e, z, m = dl.next_batch()
    ok = g["mask"]
    assert np.abs(d.cpu().numpy() - g["delta"])[ok].max() < 1e-5 * np.abs(g["delta"][ok]).max()
    assert relerr(z.cpu().numpy(), g["zabs"]) < 1e-6
    assert np.array_equal(m.cpu().numpy(), g["mask"]) and np.array_equal(e.cpu().numpy(), g["error"])


def test_gather_prepare_separable_table_path_matches_direct_path():
    """Batches of >= 4 rows per CTA take k_gather_prepare<TABLE>: tau_total = (1 + zq)^be T1[i] + T0[i] with per-pixel tables in
    shared memory instead of a powf per cell and line.  Same data through the table path (one batch of 3 700 rows) and the
    direct path (batches of 925, itself pinned to the reference's goldens above), on the DESI-shaped grid whose bluest pixels
    lie below Ly-beta .. Ly-5 (multi-line sums): delta within a few ulp of the continuum, zabs to 3e-7, error / mask exact."""
    from qfa_b200 import DeviceDataloader, synth
    dev = _dev()
    grid = synth.GRIDS["desi"]
    P, mu = synth.smooth_random_params(grid, 4, seed=3)
    n = 3700
    d = synth.make_spectra(P, mu, grid, n, seed=21, device=dev)
    outs = {}
    for bs in (n, 925):
        dl = DeviceDataloader(d["flux"], d["error"], d["zqso"], d["mask"], grid.wav(), batch_size=bs, device=dev, shuffle=False)
        dl.rewind()
        parts = []
        while dl.have_next_batch():
            parts.append([t.clone() for t in dl.next_batch()])
        outs[bs] = [torch.cat([p[j] for p in parts]) for j in range(4)]
        mu_dev = dl._mu_dev
    (d1, e1, z1, m1), (d0, e0, z0, m0) = outs[n], outs[925]
    assert torch.equal(e1, e0) and torch.equal(m1, m0)
    assert float((z1 - z0).abs().max() / z0.abs().max()) < 3e-7
    rel = (d1 - d0).abs() / mu_dev.abs().clamp_min(1e-3)[None, :]
    assert float(rel[m0].max()) < 3e-6                   # unmasked pixels (masked ones hold flux = -999: one ulp there is 6e-5)
    assert float((d1 - d0).abs()[m0].max()) > 0.0        # really two different code paths


def test_device_shuffle_visits_every_spectrum_once():
    from qfa_b200 import DeviceDataloader
    dev = _dev()
    g = np.load(os.path.join(GOLD, "prep_desi_like.npz"))
    n = len(g["zqso"])
    flux = g["flux"].copy()
    flux[:, -1] = np.arange(n)                       # tag every spectrum in its last (red, tau = 0) pixel
    mask = g["mask"].copy(); mask[:, -1] = True
    dl = DeviceDataloader(flux, g["error"], g["zqso"], mask, g["wav"], batch_size=5, device=dev, shuffle=True, seed=3)
    orders = []
    for _ in range(2):
        dl.rewind()
        seen = []
        while dl.have_next_batch():
            d, e, z, m = dl.next_batch()
            tags = (d[:, -1] + dl._mu_dev[-1]).round().long().cpu().tolist()
            seen += tags
            assert d.shape[0] in (5, n % 5)
        assert sorted(seen) == list(range(n))
        orders.append(seen)
    assert orders[0] != orders[1] and orders[0] != list(range(n))          # a fresh permutation every epoch
    dl2 = DeviceDataloader(flux, g["error"], g["zqso"], mask, g["wav"], batch_size=5, device=dev, shuffle=True, seed=3)
    dl2.rewind()
    assert torch.equal(dl2._perm.cpu(), torch.tensor(orders[0]))           # seeded: every rank draws the same order


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_cuda_graph_train_equals_eager_train(tmp_path, precision):
    """QFA.train with the captured step (gather -> accumulate -> Adam+clip -> loss/cursor, one graph launch per batch,
    epoch scalars from device memory) == the eager fused loop: same kernels, same order -> identical parameters."""
    from qfa_b200 import QFA, Adam, step_scheduler, DeviceDataloader
    from qfa_b200 import utils as U
    dev = _dev()
    wav, Nb, Nr = U.wavelength_grid(1030.0, 1600.0, 2e-3)
    rng = np.random.default_rng(5)
    n, P = 104, len(wav)                              # batch 32: three full batches + a partial one of 8 per epoch
    zq = rng.uniform(2.0, 3.5, n)
    flux = rng.normal(1.0, 0.2, (n, P)); err = rng.uniform(0.05, 0.2, (n, P))
    mask = rng.uniform(size=(n, P)) > 0.1
    flux[~mask] = -999.0; err[~mask] = -999.0
    out = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(0)
        m = QFA(Nb, Nr, 4, dev, precision=precision)
        m.use_cuda_graph = mode == "graph"
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.5, 2), learning_rate=1e-2, weight_decay=0.1)
        dl = DeviceDataloader(flux, err, zq, mask, wav, batch_size=32, device=dev, shuffle=True, seed=11)
        logs = []

        class Log:
            def info(self, msg):
                logs.append(float(msg.split("loss:")[1].split(";")[0]))
        m.train(opt, dl, 5, output_dir=str(tmp_path / mode), save_interval=5, smooth_interval=3, quiet=True, logger=Log())
        out[mode] = (m._params.clone(), opt._m.clone(), opt._v.clone(), logs, m)
    assert out["graph"][4]._graph is not None and out["eager"][4]._graph is None
    assert out["graph"][4].graph_launches_per_step >= 4
    for a, b in zip(out["eager"][:3], out["graph"][:3]):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    assert np.allclose(out["eager"][3], out["graph"][3], rtol=0, atol=0.006)       # the log line prints two decimals


def test_ood_select_matches_torch():
    from qfa_b200 import QFA
    dev = _dev()
    c = load_case("tiny5")
    m = QFA(c["Nb"], c["F"].shape[0] - c["Nb"], 5, dev, model_params=c)
    g = torch.Generator(device="cpu").manual_seed(1)
    for B in (1, 37, 5000, 200_003):
        nll = (torch.randn(B, generator=g) * 50 + 100).to(dev)
        if B > 100:
            nll[7] = nll[11] = nll[99] = 1e4                                  # ties at the top: ascending index order
            nll[13] = float("nan")                                             # NaN = most suspicious
        thr = 150.0
        r = m.ood_select(nll, threshold=thr, k=min(B, 64))
        bad = (nll > thr) | torch.isnan(nll)
        assert int(r["count"]) == int(bad.sum())
        assert torch.equal(r["above"], torch.nonzero(bad).flatten())
        k = min(B, 64)
        key = torch.where(torch.isnan(nll), torch.full_like(nll, float("inf")), nll)
        order = torch.sort(key, descending=True, stable=True).indices[:k]
        assert torch.equal(r["top_idx"], order), B
        assert torch.equal(torch.nan_to_num(r["top_val"], nan=-1.0), torch.nan_to_num(nll[order], nan=-1.0))
    r = m.ood_select(nll, threshold=None, k=2048)
    assert r["top_idx"].numel() == 2048 and "count" not in r
    r = m.ood_select(nll, threshold=1e9, k=0, cap=10)
    assert int(r["count"]) == 1 and r["above"].tolist() == [13]


def test_posterior_sampling_seeded():
    """h = hmean + chol(hcov) z with z from Philox (seeded), continuum samples mu + F h (nb/predict.ipynb cell 11)."""
    dev = _dev()
    c, ref = load_case("sdss", "f64")
    from qfa_b200 import QFA
    P_, Nh = c["F"].shape
    m = QFA(c["Nb"], P_ - c["Nb"], Nh, dev, model_params=c)
    m.mu = torch.tensor(c["mu"])
    T = lambda a: torch.tensor(a).to(dev)
    o = m.predict_batch(T(c["flux"]), T(c["error"]), T(c["zabs"]), T(c["mask"]))
    s = m.sample_posterior(o["hmean"], o["hcov"], n_samples=4000, seed=42)
    B = o["hmean"].shape[0]
    assert s["h"].shape == (B, 4000, Nh) and s["cont"].shape == (B, 4000, P_)
    Lc = torch.linalg.cholesky(o["hcov"].double())
    h_ref = o["hmean"].double()[:, None, :] + torch.einsum("bkc,bsc->bsk", Lc, s["z"].double())
    assert relerr(s["h"].cpu().numpy(), h_ref.cpu().numpy()) < 1e-5
    cont_ref = T(c["mu"]).double()[None, None, :] + torch.einsum("ik,bsk->bsi", T(c["F"]).double(), s["h"].double())
    assert relerr(s["cont"].cpu().numpy(), cont_ref.cpu().numpy()) < 1e-5
    z = s["z"].reshape(-1, Nh)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    emp = torch.einsum("bsk,bsl->bkl", s["h"] - o["hmean"][:, None], s["h"] - o["hmean"][:, None]) / 4000
    assert relerr(emp.cpu().numpy(), o["hcov"].cpu().numpy()) < 0.15
    s2 = m.sample_posterior(o["hmean"], o["hcov"], n_samples=4000, seed=42, want=("h",))
    s3 = m.sample_posterior(o["hmean"], o["hcov"], n_samples=4000, seed=43, want=("h",))
    assert torch.equal(s2["h"], s["h"]) and not torch.equal(s3["h"], s["h"]) and "cont" not in s2


def test_predict_to_npz_columnar_writer(tmp_path):
    """One columnar .npz with the reference's five keys (main.py:96-98 writes one file per spectrum)."""
    dev = _dev()
    c, ref = load_case("tiny5", "f64")
    from qfa_b200 import QFA
    P_, Nh = c["F"].shape
    m = QFA(c["Nb"], P_ - c["Nb"], Nh, dev, model_params=c, precision="fp64")
    m.mu = torch.tensor(c["mu"])
    path = str(tmp_path / "pred.npz")
    m.predict_to_npz(path, c["flux"], c["error"], c["zabs"], c["mask"], names=[f"s{i}" for i in range(len(c["flux"]))])
    o = np.load(path)
    assert sorted(o.files) == sorted(["ll", "hmean", "hcov", "cont", "uncertainty", "names"])
    assert relerr(o["ll"], ref["pred_nll"]) < 1e-5 and relerr(o["cont"], ref["pred_cont"]) < 1e-5
    assert relerr(o["uncertainty"], ref["pred_unc"]) < 1e-5 and relerr(o["hmean"], ref["pred_hmean"]) < 1e-5


def test_model_on_a_non_current_device():
    """the library's per-device state (shared-memory opt-ins, SM count) and the launches follow the model's
    device, not whatever device happens to be current."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from qfa_b200 import QFA
    c, ref = load_case("tiny5", "f64")
    P_, Nh = c["F"].shape
    outs = []
    for d in (1, 0):
        torch.cuda.set_device(0)
        dev = torch.device("cuda", d)
        for prec in ("fp32", "tf32"):
            m = QFA(c["Nb"], P_ - c["Nb"], Nh, dev, model_params=c, precision=prec)
            m.mu = torch.tensor(c["mu"])
            T = lambda a: torch.tensor(a).to(dev)
            loss, grads = m.forward(T(c["delta"]), T(c["error"]), T(c["zabs"]), T(c["mask"]))
            o = m.predict_batch(T(c["flux"]), T(c["error"]), T(c["zabs"]), T(c["mask"]))
            torch.cuda.synchronize(dev)
            assert abs(float(loss) - float(ref["loss"])) < 2e-3 * abs(float(ref["loss"]))
            outs.append(o["cont"].cpu())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])


@pytest.mark.gpu
def test_peer_allreduce_protocol_one_device():
    """qfa_peer_allreduce (the exchange step of 8(e)): 1, 2, 3 and 8 emulated ranks on one GPU -- separate
    buffers, states and streams, kernels that really wait for each other's flags -- float and double, vector body and scalar
    tail, a rank enqueued a step ahead (double-buffered publish).  Bit-exact against the rank-ordered torch sum.  Own
    process: see tests/workers/peer_one_device.py."""
    import subprocess
    import sys
    w = os.path.join(os.path.dirname(os.path.abspath(__file__)), "workers", "peer_one_device.py")
    r = subprocess.run([sys.executable, w], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, QFA_PEER_TIMEOUT_S="30"))
    assert r.returncode == 0 and "PEER-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
