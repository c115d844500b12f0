"""GPU scratch: one-line timings of the three tensor-core paths (for A/B of library builds on one box)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
dev = torch.device("cuda:0")
def t(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
B = 88800
d = synth.make_spectra(P, mu, grid, B, seed=1234, device=dev)
m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision="tf32"); m.mu = mu
X, E, Z, M, D = d["flux"], d["error"], d["zabs"], d["mask"].view(torch.uint8), d["delta"]
o = m.predict_batch(X, E, Z, M)
tp = t(lambda: m.predict_into(X, E, Z, M, o))
tn = t(lambda: m.predict_into(X, E, Z, M, {"nll": o["nll"]}))
tt = t(lambda: m.accumulate(D, E, Z, M))
print(f"mixed predict {tp:.3f} ms ({B/tp/1e3:.1f} M/s) | nll-only {tn:.3f} ms ({B/tn/1e3:.1f} M/s) | train {tt:.3f} ms ({B/tt/1e3:.1f} M/s)")
