import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from qfa_b200 import QFA, synth
from conftest import relerr
KEYS = ("F", "Psi", "omega", "tau0", "c0", "beta")
grid = synth.GRIDS["l32"]
for Nh in (12, 16, 20):
    P, mu = synth.smooth_random_params(grid, Nh, seed=1237)
    d = synth.make_spectra(P, mu, grid, 3001, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
    Pn = {k: v.numpy() for k, v in P.items()}
    out = {}
    for prec in ("fp32", "tf32"):
        mm = QFA(grid.Nb, grid.Nr, Nh, torch.device("cuda:0"), model_params=Pn, precision=prec)
        l, gr = mm.forward(d["delta"], d["error"], d["zabs"], d["mask"])
        out[prec] = (float(l), {k: gr[k].cpu().numpy() for k in KEYS})
    print(Nh, "loss", out["fp32"][0], out["tf32"][0], {k: "%.1e" % relerr(out["tf32"][1][k], out["fp32"][1][k]) for k in KEYS})
