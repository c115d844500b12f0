"""GPU scratch: run the mixed predict path a few times (for ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from qfa_b200 import QFA, synth
k = np.load(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'tests', 'golden', 'kat_sdss.npz'))
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
mode = sys.argv[2] if len(sys.argv) > 2 else "predict"
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32"); m.mu = mu
X, E, Z, M = data["flux"], data["error"], data["zabs"], data["mask"].view(torch.uint8)
if mode == "predict":
    o = m.predict_batch(X, E, Z, M)
    for _ in range(3): m.predict_into(X, E, Z, M, o)
else:
    D = data["delta"]
    for _ in range(4): m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
print("done")
