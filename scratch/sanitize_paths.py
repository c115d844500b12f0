"""GPU scratch: one small call of every tensor-core path (ragged sizes) -- meant to run under compute-sanitizer."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
dev = torch.device("cuda:0")
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
d = synth.make_spectra(P, mu, grid, 301, seed=3, device=dev)
m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision="tf32"); m.mu = mu
o = m.predict_batch(d["flux"], d["error"], d["zabs"], d["mask"])
n = m.nll_batch(d["flux"], d["error"], d["zabs"], d["mask"])
l, g = m.forward(d["delta"], d["error"], d["zabs"], d["mask"])
print("sdss nh8:", float(l), float(o["cont"].abs().mean()), float(n.mean()))
g32 = synth.GRIDS["l32"]
P32, mu32 = synth.smooth_random_params(g32, 32, seed=1237)
d32 = synth.make_spectra(P32, mu32, g32, 250, seed=5, device=dev, mask_iid=0.15, run_len=(40, 160))
m32 = QFA(g32.Nb, g32.Nr, 32, dev, model_params={a: b.numpy() for a, b in P32.items()}, precision="tf32")
l32, gg = m32.forward(d32["delta"], d32["error"], d32["zabs"], d32["mask"])
print("l32 nh32:", float(l32), float(gg["F"].abs().mean()))
torch.cuda.synchronize()
print("done")
