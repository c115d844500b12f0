"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference
(/root/reference, unmodified) on CPU in the build container.

    python oracle/make_golden.py            # writes every fixture (fp32 + fp64-promoted)

The fp64-promoted run must own its process (it monkey-patches torch.float32, SURVEY.md
section 8c), so this script re-invokes itself with `--worker fp32|fp64`.
The GPU box has no /root/reference: the committed .npz files are what travels.

Fixtures
  kat_sdss.npz            reference data/model_parameters.npz + data/spec-4321-55504-0114.npz
                          (the reference's only known-answer vector) re-packed,
                          plus the outputs of the reference run here (fp32 and fp64)
  desi_params.npz         reference data/model_parameters_desi.npz re-packed (shape fixture)
  case_<name>.npz         seeded inputs + parameters of a synthetic case
  case_<name>_<f32|f64>.npz  reference outputs for that case: forward loss/grads, per-spectrum
                          NLLs, single-spectrum partials, predictions
  train_tiny_f32.npz      forward -> Adam.update -> clip step and a 6-epoch train() (smooth+save at 5)
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- case definitions
def build_cases():
    """Inputs are produced with qfa_b200.synth on the CPU and STORED, so the fixtures do not
    depend on RNG reproducibility across torch versions."""
    import torch
    from qfa_b200 import synth
    cases = {}

    def pack(name, params, mu, d, Nb, law="becker"):
        cases[name] = dict(
            F=params["F"].numpy().astype(np.float32), Psi=params["Psi"].numpy().astype(np.float32),
            omega=params["omega"].numpy().astype(np.float32), tau0=np.float32(params["tau0"]),
            c0=np.float32(params["c0"]), beta=np.float32(params["beta"]), mu=mu.numpy().astype(np.float32),
            flux=d["flux"].numpy(), error=d["error"].numpy(), zabs=d["zabs"].numpy(), mask=d["mask"].numpy(),
            delta=d["delta"].numpy(), Nb=np.int64(Nb), law=np.array(law))

    # SDSS shape, pretrained parameters (with the c0<-beta load quirk, as the reference would hold them)
    ref = np.load("/root/reference/data/model_parameters.npz")
    sd = synth.GRIDS["sdss"]
    p_sdss = {"F": torch.tensor(ref["F"], dtype=torch.float32), "Psi": torch.tensor(ref["Psi"]),
              "omega": torch.tensor(ref["omega"]), "tau0": torch.tensor(ref["tau0"]),
              "c0": torch.tensor(ref["beta"]), "beta": torch.tensor(ref["beta"])}
    mu_sdss = torch.tensor(ref["mu"])
    d = synth.make_spectra(p_sdss, mu_sdss, sd, 6, seed=1234)
    pack("sdss", p_sdss, mu_sdss, d, sd.Nb)

    # L32 shape (Npix 1000, Nh 32), ~30 % masked
    lg = synth.GRIDS["l32"]
    p32, mu32 = synth.smooth_random_params(lg, 32, seed=1237)
    d = synth.make_spectra(p32, mu32, lg, 4, seed=1237, mask_iid=0.15, mask_runs=3, run_len=(40, 160))
    pack("l32", p32, mu32, d, lg.Nb)

    # tiny grids: odd Nh values (zero padding to 8 / 16), other optical-depth laws, edge cases
    tiny = synth.GridSpec("tiny", 1150.0, 6e-4, 96)
    for name, Nh, law, seed in (("tiny5", 5, "becker", 11), ("tiny12", 12, "fg", 12), ("tiny3k", 3, "kamble", 13),
                                ("tiny8m", 8, "mock", 14), ("tiny16", 16, "becker", 15), ("tiny1", 1, "becker", 16)):
        pt, mut = synth.smooth_random_params(tiny, Nh, seed=seed)
        d = synth.make_spectra(pt, mut, tiny, 12, seed=seed, law=law, mask_iid=0.1, mask_runs=2, run_len=(5, 30))
        pack(name, pt, mut, d, tiny.Nb, law)

    # edge cases on the tiny grid
    pt, mut = synth.smooth_random_params(tiny, 4, seed=21)
    d = synth.make_spectra(pt, mut, tiny, 8, seed=21, mask_iid=0.1, mask_runs=1, run_len=(5, 20))
    m = d["mask"].clone()
    Nb = tiny.Nb
    m[0, :] = False                       # all-masked spectrum
    m[1, :Nb] = False                     # red-only
    m[2, Nb:] = False                     # blue-only
    m[3, :] = False
    m[3, Nb + 3] = True                   # single (red) pixel
    m[4, :] = False
    m[4, 2] = True                        # single (blue) pixel
    m[:, 7] = False                       # a pixel masked in every spectrum -> NaN gradient (0/0)
    m[:, Nb + 9] = False
    for k in ("flux", "error"):
        d[k] = torch.where(m, d[k], torch.full_like(d[k], -999.0))
    d["mask"] = m
    d["delta"] = torch.where(m, d["delta"], d["flux"] - mut[None])
    pack("edge", pt, mut, d, Nb)

    # training set for the optimiser / train() golden: every batch of 6 holds one fully observed
    # spectrum, so no gradient element is 0/0 (the reference's NaN would poison Adam, quirk Q4)
    pt, mut = synth.smooth_random_params(tiny, 4, seed=31)
    d = synth.make_spectra(pt, mut, tiny, 12, seed=31, mask_iid=0.0, mask_runs=0)
    gm = torch.Generator().manual_seed(31)
    m = torch.rand(12, tiny.Npix, generator=gm) >= 0.08
    m[0, :] = True
    m[6, :] = True
    for k in ("flux", "error"):
        d[k] = torch.where(m, d[k], torch.full_like(d[k], -999.0))
    d["mask"] = m
    pack("train", pt, mut, d, Nb)
    return cases


# ----------------------------------------------------------------------------- worker
def worker(kind):
    fp64 = kind == "fp64"
    from oracle.ref_loader import load_reference
    QFA, Adam, step_scheduler, rutils = load_reference(fp64=fp64)
    import torch
    from functools import partial
    torch.set_num_threads(8)
    dt = torch.float64 if fp64 else torch.float32   # NB: under promotion torch.float32 IS float64
    tag = "f64" if fp64 else "f32"
    cpu = torch.device("cpu")

    def model_for(c):
        Npix, Nh = c["F"].shape
        Nb = int(c["Nb"])
        m = QFA(Nb, Npix - Nb, Nh, cpu, tau=partial(rutils.tau, which=str(c["law"])))
        for k in ("F", "Psi", "omega", "tau0", "c0", "beta"):
            setattr(m, k, torch.tensor(c[k], dtype=dt))
        m.mu = torch.tensor(c["mu"], dtype=dt)
        return m

    def T(x):
        return torch.tensor(x, dtype=dt)

    for name in sorted(f[5:-4] for f in os.listdir(GOLD) if f.startswith("case_") and f.count("_") == 1):
        c = dict(np.load(os.path.join(GOLD, f"case_{name}.npz")))
        m = model_for(c)
        delta, error, zabs = T(c["delta"]), T(c["error"]), T(c["zabs"])
        flux, mask = T(c["flux"]), torch.tensor(c["mask"])
        out = {}
        loss, grads = m.forward(delta, error, zabs, mask)
        out["loss"] = loss.numpy()
        for k, v in grads.items():
            out["grad_" + k] = v.numpy()
        nlls = []
        for b in range(delta.shape[0]):
            ll, g = m.loglikelihood_and_gradient_for_single_spectra(delta[b], error[b], zabs[b], mask[b])
            nlls.append(float(ll))
            if b == min(5, delta.shape[0] - 1):
                for k, v in g.items():
                    out["single_" + k] = np.asarray(v.numpy())
                out["single_index"] = np.int64(b)
        out["nll"] = np.array(nlls)
        pn, ph, pc, pcont, punc = [], [], [], [], []
        for b in range(flux.shape[0]):
            ll, hm, hc, cont, unc = m.prediction_for_single_spectra(flux[b], error[b], zabs[b], mask[b])
            pn.append(float(ll)); ph.append(hm.squeeze(-1).numpy()); pc.append(hc.numpy())
            pcont.append(cont.numpy()); punc.append(unc.numpy())
        out.update(pred_nll=np.array(pn), pred_hmean=np.array(ph), pred_hcov=np.array(pc),
                   pred_cont=np.array(pcont), pred_unc=np.array(punc))
        np.savez_compressed(os.path.join(GOLD, f"case_{name}_{tag}.npz"), **out)
        print(f"[{tag}] case {name}: loss {float(loss):.6f}")

    # ---- shipped known-answer vector through the real reference (both dtypes)
    spec = np.load("/root/reference/data/spec-4321-55504-0114.npz")
    wav = 10 ** np.arange(np.log10(1030), np.log10(1600), 1e-4)
    m = QFA(720, 1193, 8, cpu)
    m.load_from_npz("/root/reference/data/model_parameters.npz")     # c0 <- beta quirk included
    if fp64:   # load_from_npz hard-codes dtype=torch.float32, which IS float64 under promotion
        pass
    zabs = T(wav[:720] * (1 + float(spec["z"])) / 1215.67 - 1)
    flux, error = T(spec["flux"]), T(spec["error"])
    res = {}
    for suffix, mask in (("", torch.tensor(spec["mask"])),
                         ("_red", torch.tensor(np.concatenate([np.zeros(720, bool), spec["mask"][720:]])))):
        ll, hm, hc, cont, unc = m.prediction_for_single_spectra(flux, error, zabs, mask)
        res.update({"ref_ll" + suffix: float(ll), "ref_h" + suffix: hm.squeeze(-1).numpy(),
                    "ref_hcov" + suffix: hc.numpy(), "ref_cont" + suffix: cont.numpy(),
                    "ref_unc" + suffix: unc.numpy()})
    np.savez_compressed(os.path.join(GOLD, f"kat_sdss_{tag}.npz"), **res)
    print(f"[{tag}] KAT ll {res['ref_ll']:.6f} (stored {float(spec['ll']):.6f}), red {res['ref_ll_red']:.6f}")

    if not fp64:
        # ---- one optimiser step + a 6-epoch train() on a tiny grid (boundary rows a8-a10)
        c = dict(np.load(os.path.join(GOLD, "case_train.npz")))
        m = model_for(c)
        delta, error, zabs, mask = T(c["delta"]), T(c["error"]), T(c["zabs"]), torch.tensor(c["mask"])

        class Loader:     # duck-typed dataloader (reference model.py:204-211), no shuffling
            def __init__(s):
                s.mu = c["mu"]; s.data_size = delta.shape[0]; s.batch_size = 6; s.cur = 0
            def rewind(s): s.cur = 0
            def have_next_batch(s): return s.cur < s.data_size
            def next_batch(s):
                a, b = s.cur, min(s.cur + s.batch_size, s.data_size); s.cur = b
                return delta[a:b], error[a:b], zabs[a:b], mask[a:b]

        def far_from_optimum(mm):   # Psi = omega = 1 as in random_init_func (model.py:68-69): NLL stays > 0,
            mm.Psi = torch.ones_like(mm.Psi)      # so train() does not take its `loss < 0` early exit (model.py:224)
            mm.omega = torch.ones_like(mm.omega)
            return mm
        m = far_from_optimum(m)
        opt = Adam(params=m.parameters, device=cpu, scheduler=step_scheduler(0.9, 2), learning_rate=1e-2,
                   weight_decay=0.1)
        loss, grads = m.forward(delta[:6], error[:6], zabs[:6], mask[:6])
        m.parameters = opt.update(m.parameters, grads)
        out = {"step_loss": loss.numpy()}
        for k, v in m.parameters.items():
            out["step_" + k] = v.numpy()
        for k in opt.m:
            out["step_m_" + k] = opt.m[k].numpy(); out["step_v_" + k] = opt.v[k].numpy()
        m = far_from_optimum(model_for(c))
        opt = Adam(params=m.parameters, device=cpu, scheduler=step_scheduler(0.9, 2), learning_rate=1e-2,
                   weight_decay=0.1)
        import tempfile, io, contextlib
        tmp = tempfile.mkdtemp()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            m.train(opt, Loader(), 6, output_dir=tmp, save_interval=5, smooth_interval=5)
        out["train_log"] = np.array(buf.getvalue())
        ck = np.load(os.path.join(tmp, "checkpoints", "model_parameters_epoch_05.npz"))
        for k in ck.files:
            out["ckpt5_" + k] = ck[k]
        for k, v in m.parameters.items():
            out["final_" + k] = v.numpy()
        sm = model_for(c)
        sm.smooth()
        for k, v in sm.parameters.items():
            out["smooth_" + k] = v.numpy()
        np.savez_compressed(os.path.join(GOLD, "train_tiny_f32.npz"), **out)
        print("[f32] train_tiny:", buf.getvalue().strip().splitlines()[-1])


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    os.makedirs(GOLD, exist_ok=True)
    for name, c in build_cases().items():
        np.savez_compressed(os.path.join(GOLD, f"case_{name}.npz"), **c)
    ref = np.load("/root/reference/data/model_parameters.npz")
    spec = np.load("/root/reference/data/spec-4321-55504-0114.npz")
    keep = ("ll", "h", "our", "our_uncertainty", "ll_red", "h_red", "our_red", "our_uncertainty_red",
            "flux", "error", "z", "mask")
    np.savez_compressed(os.path.join(GOLD, "kat_sdss.npz"), **{"param_" + k: ref[k] for k in ref.files},
                        **{k: spec[k] for k in keep})
    desi = np.load("/root/reference/data/model_parameters_desi.npz")
    np.savez_compressed(os.path.join(GOLD, "desi_params.npz"), **{k: desi[k] for k in desi.files})
    for kind in ("fp32", "fp64"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", kind], check=True, cwd=ROOT)


if __name__ == "__main__":
    main()
