"""GPU scratch: the float CUDA-core train accumulation at the reference's batch of 500 (SDSS shape), for an ncu capture."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qfa_b200 import QFA, synth
dev = torch.device("cuda:0")
k = np.load(os.path.join(ROOT, "tests", "golden", "kat_sdss.npz"))
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 500
m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={kk: v.numpy() for kk, v in P.items()}, precision="mixed"); m.mu = mu
d = synth.make_spectra(P, mu, grid, B, seed=1234, device=dev)
D, E, Z, M = d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8)
for _ in range(6):
    m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
print("done")
