"""GPU scratch: tensor-core (mixed) train path vs the fp64 CUDA path + timing."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import CASES, load_case
from qfa_b200 import QFA, synth
KEYS = ("F", "Psi", "omega", "tau0", "c0", "beta")
def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float); ok = ~np.isnan(b)
    if not ok.any(): return 0.0
    return np.abs(a[ok] - b[ok]).max() / max(np.abs(b[ok]).max(), 1e-300)
d = lambda x: torch.as_tensor(x).cuda()
for name in CASES:
    c, g = load_case(name, "f64")
    Npix, Nh = c["F"].shape
    if Nh > 8: continue
    m = QFA(c["Nb"], Npix - c["Nb"], Nh, torch.device("cuda:0"), tau=c["law"],
            model_params={k: c[k] for k in KEYS}, precision="tf32")
    m.mu = torch.tensor(c["mu"])
    loss, grads = m.forward(d(c["delta"]), d(c["error"]), d(c["zabs"]), d(c["mask"]))
    torch.cuda.synchronize()
    print(f"{name:8s} Nh={Nh} loss {float(loss):.4f} ref {float(np.squeeze(g['loss'])):.4f} | " +
          " ".join(f"{k}:{rel(grads[k].cpu().numpy(), g['grad_'+k]):.1e}" for k in KEYS), flush=True)

k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
X, E, Z, M = data["delta"], data["error"], data["zabs"], data["mask"].view(torch.uint8)
res = {}
for prec in ("fp64", "mixed"):
    m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision=prec); m.mu = mu
    n = Bn if prec == "mixed" else min(Bn, 8192)
    nn = min(Bn, 8192)
    acc = m.accumulate(X[:nn], E[:nn], Z[:nn], M[:nn]).double().cpu().numpy(); torch.cuda.synchronize()
    res[prec] = acc
    if prec == "mixed":
        for _ in range(3): m.accumulate(X, E, Z, M)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m.accumulate(X, E, Z, M)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"mixed train accumulate: {ms:.3f} ms / {Bn} spectra = {Bn/ms*1e3/1e6:.2f} M spectra/s ; {Bn*20101/ms/1e6:.0f} GB/s algorithmic")
n = grid.Npix * 8 + grid.Npix + grid.Nb + 3
a, b = res["mixed"], res["fp64"]
PH = grid.Npix * 8
print("sums F %.2e Psi %.2e omega %.2e scal %s" % (rel(a[:PH], b[:PH]), rel(a[PH:PH+grid.Npix], b[PH:PH+grid.Npix]),
      rel(a[PH+grid.Npix:PH+grid.Npix+grid.Nb], b[PH+grid.Npix:PH+grid.Npix+grid.Nb]), np.abs(a[n-3:n]-b[n-3:n])/np.abs(b[n-3:n])))
print("counts equal:", np.array_equal(a[n:n+grid.Npix+3], b[n:n+grid.Npix+3]), " nll sum rel %.2e" % (abs(a[n+grid.Npix+3]-b[n+grid.Npix+3])/abs(b[n+grid.Npix+3])),
      " dmu %.2e" % rel(a[n+grid.Npix+5:], b[n+grid.Npix+5:]))
