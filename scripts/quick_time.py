"""GPU scratch: CUDA-event timings of the main tensor-core paths (SDSS predict, SDSS train accumulate, L32 accumulate)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qfa_b200 import QFA, synth
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["predict", "train", "l32"]
dev = torch.device("cuda:0")
def timeit(fn, n=8, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = []
if "predict" in which or "train" in which:
    k = np.load(os.path.join(ROOT, "tests", "golden", "kat_sdss.npz"))
    P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
    P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
    grid = synth.GRIDS["sdss"]
    Pn = {key: v.numpy() for key, v in P.items()}
    m = QFA(grid.Nb, grid.Nr, 8, dev, model_params=Pn, precision="tf32"); m.mu = mu
    if "predict" in which:
        d = synth.make_spectra(P, mu, grid, 100_000, seed=1234, device=dev)
        X, E, Z, M = d["flux"], d["error"], d["zabs"], d["mask"].view(torch.uint8)
        o = m.predict_batch(X, E, Z, M)
        ms = timeit(lambda: m.predict_into(X, E, Z, M, o))
        out.append("predict %.1f us (%.1f M/s)" % (ms * 1e3, 100_000 / ms / 1e3))
        del d, X, E, Z, M, o
    if "train" in which:
        d = synth.make_spectra(P, mu, grid, 71_040, seed=1234, device=dev)
        D, E, Z, M = d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8)
        ms = timeit(lambda: m.accumulate(D, E, Z, M))
        out.append("train %.1f us (%.1f M/s)" % (ms * 1e3, 71_040 / ms / 1e3))
        del d, D, E, Z, M
if "l32" in which:
    grid = synth.GRIDS["l32"]
    P, mu = synth.smooth_random_params(grid, 32, seed=1237)
    d = synth.make_spectra(P, mu, grid, 65_536, seed=11, device=dev, mask_iid=0.15, run_len=(40, 160))
    m = QFA(grid.Nb, grid.Nr, 32, dev, model_params={k: v.numpy() for k, v in P.items()}, precision="tf32")
    a = (d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8))
    ms = timeit(lambda: m.accumulate(*a), n=5)
    out.append("l32 %.1f us (%.1f M/s)" % (ms * 1e3, 65_536 / ms / 1e3))
print(" | ".join(out))
