"""GPU scratch: a few train steps at one batch size (for an ncu launch list): python scratch/small_batch_one.py 8192"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, Adam, step_scheduler, synth
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
d = synth.make_spectra(P, mu, grid, B, seed=1, device=dev)
m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision="tf32"); m.mu = mu
opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
X, E, Z, M = d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8)
for _ in range(6):
    acc = m.accumulate(X, E, Z, M, zero=True)
    opt.update_from_acc(m, acc)
torch.cuda.synchronize()
print("done")
