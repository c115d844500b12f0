"""GPU scratch: error table of the Nh > 8 tensor-core prediction path vs goldens / oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case, relerr
from qfa_b200 import QFA, synth
from oracle import qfa_lowrank
dev0 = torch.device("cuda:0")
T = lambda a: torch.tensor(a).to(dev0)
for name in ("l32", "tiny12", "tiny16"):
    c, g = load_case(name, "f64")
    Npix, Nh = c["F"].shape
    for prec in ("fp32", "tf32"):
        m = QFA(c["Nb"], Npix - c["Nb"], Nh, dev0, tau=c["law"], model_params={k: c[k] for k in ("F", "Psi", "omega", "tau0", "c0", "beta")}, precision=prec)
        m.mu = torch.tensor(c["mu"])
        for sf in (False, True):
            m.solve_fp64 = sf
            o = m.predict_batch(T(c["flux"]), T(c["error"]), T(c["zabs"]), T(c["mask"]))
            npx = np.maximum(1, c["mask"].sum(1))
            e = {k: relerr(o[k].cpu().numpy(), g["pred_" + k]) for k in ("cont", "unc", "hmean", "hcov")}
            e["nll/px"] = float((np.abs(o["nll"].cpu().numpy() - g["pred_nll"]) / npx).max())
            print(name, prec, "dchol" if sf else "fchol", {k: "%.1e" % v for k, v in e.items()})
grid = synth.GRIDS["l32"]
for Nh in (12, 32):
    P, mu = synth.smooth_random_params(grid, Nh, seed=1237)
    d = synth.make_spectra(P, mu, grid, 3001, seed=11, device=dev0, mask_iid=0.15, run_len=(40, 160))
    Pn = {k: v.numpy() for k, v in P.items()}
    a = [d[k] for k in ("flux", "error", "zabs", "mask")]
    cpu = [t.cpu().numpy() for t in a]
    rn, rh, rc, rcont, runc = qfa_lowrank.predict_batch(Pn, mu.numpy(), *cpu, grid.Nb)
    npx = np.maximum(1, cpu[3].sum(1))
    for prec in ("fp32", "tf32"):
        mm = QFA(grid.Nb, grid.Nr, Nh, dev0, model_params=Pn, precision=prec); mm.mu = mu
        o = mm.predict_batch(*a)
        ce = np.abs(o["cont"].cpu().numpy() - rcont).max(1) / np.abs(rcont).max()
        e = {"cont": relerr(o["cont"].cpu().numpy(), rcont), "cont_med": float(np.median(ce)), "unc": relerr(o["unc"].cpu().numpy(), runc),
             "hmean": relerr(o["hmean"].cpu().numpy(), rh), "hcov": relerr(o["hcov"].cpu().numpy(), rc),
             "nll/px": float((np.abs(o["nll"].cpu().numpy() - rn) / npx).max())}
        print("Nh", Nh, prec, {k: "%.1e" % v for k, v in e.items()})
    import time
    for prec in ("fp32", "tf32"):
        mm = QFA(grid.Nb, grid.Nr, Nh, dev0, model_params=Pn, precision=prec); mm.mu = mu
        for B in (256, 1024, 3001):
            b = [t[:B].contiguous() for t in a]
            o = mm.predict_batch(*b)
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(10): mm.predict_into(b[0], b[1], b[2], b[3].view(torch.uint8), o)
            torch.cuda.synchronize(); print("  Nh", Nh, prec, "B", B, "%.1f us/call" % ((time.time() - t0) / 10 * 1e6))
