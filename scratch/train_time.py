"""GPU scratch: time QFA.accumulate (tensor-core train path) on SDSS-shaped data; prints per-kernel split via CUDA events."""
import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 71040
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32"); m.mu = mu
E, Z, M, D = data["error"], data["zabs"], data["mask"].view(torch.uint8), data["delta"]
for _ in range(3): m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.accumulate(D, E, Z, M)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"QFA_GRAD_BLUE_COST={os.environ.get('QFA_GRAD_BLUE_COST','default')} accumulate: {ms:.3f} ms / {Bn} = {Bn/ms/1e3:.2f} M spectra/s")
