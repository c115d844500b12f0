"""GPU scratch: tensor-core (mixed) predict path vs the fp64 CUDA path + timing."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import CASES, load_case
from qfa_b200 import QFA, synth

def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float); ok = ~np.isnan(b)
    return np.abs(a[ok] - b[ok]).max() / max(np.abs(b[ok]).max(), 1e-300)

d = lambda x: torch.as_tensor(x).cuda()
for name in CASES:
    c, g = load_case(name, "f64")
    Npix, Nh = c["F"].shape
    if Nh > 8: continue
    m = QFA(c["Nb"], Npix - c["Nb"], Nh, torch.device("cuda:0"), tau=c["law"],
            model_params={k: c[k] for k in ("F", "Psi", "omega", "tau0", "c0", "beta")}, precision="tf32")
    m.mu = torch.tensor(c["mu"])
    o = m.predict_batch(d(c["flux"]), d(c["error"]), d(c["zabs"]), d(c["mask"]))
    torch.cuda.synchronize()
    npx = np.maximum(1, c["mask"].sum(1))
    e = np.abs(o["nll"].cpu().numpy() - g["pred_nll"])
    print(f"{name:8s} Nh={Nh} P={Npix} B={c['flux'].shape[0]} nll abs {e.max():.3e} (per px {np.max(e/npx):.2e}) cont {rel(o['cont'].cpu().numpy(), g['pred_cont']):.1e} "
          f"unc {rel(o['unc'].cpu().numpy(), g['pred_unc']):.1e} hm {rel(o['hmean'].cpu().numpy(), g['pred_hmean']):.1e} hcov {rel(o['hcov'].cpu().numpy(), g['pred_hcov']):.1e}", flush=True)

# SDSS synthetic, larger batch, vs fp64 CUDA path
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
res = {}
for prec in ("fp64", "mixed"):
    m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision=prec); m.mu = mu
    X, E, Z, M = data["flux"], data["error"], data["zabs"], data["mask"].view(torch.uint8)
    n = Bn if prec == "mixed" else min(Bn, 4000)
    o = m.predict_batch(X[:n], E[:n], Z[:n], M[:n]); torch.cuda.synchronize()
    res[prec] = {kk: v.double().cpu().numpy() for kk, v in o.items()}
    if prec == "mixed":
        for _ in range(3): m.predict_into(X, E, Z, M, o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m.predict_into(X, E, Z, M, o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"mixed predict: {ms:.3f} ms / {Bn} spectra = {Bn/ms*1e3/1e6:.2f} M spectra/s ; {Bn*35693/ms/1e6:.0f} GB/s algorithmic")
        e0.record()
        for _ in range(5): m.predict_into(X, E, Z, M, {"nll": o["nll"]})
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"mixed nll-only: {ms:.3f} ms = {Bn/ms*1e3/1e6:.2f} M spectra/s ; {Bn*20101/ms/1e6:.0f} GB/s algorithmic")
n = min(Bn, 4000)
npx = np.maximum(1, data["mask"][:n].sum(1).cpu().numpy())
e = np.abs(res["mixed"]["nll"][:n] - res["fp64"]["nll"])
print(f"sdss synth vs fp64: nll abs max {e.max():.3e} per px {np.max(e/npx):.2e} rel {np.max(e/np.abs(res['fp64']['nll'])):.2e}")
for kk in ("hmean", "hcov", "cont", "unc"):
    print(f"  {kk}: {rel(res['mixed'][kk][:n], res['fp64'][kk]):.2e}")
