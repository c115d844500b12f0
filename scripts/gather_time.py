"""GPU scratch: time of the fused gather + prepare of a shuffled batch (k_gather_prepare) alone, SDSS shape."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qfa_b200 import synth, DeviceDataloader
dev = torch.device("cuda:0")
k = np.load(os.path.join(ROOT, "tests", "golden", "kat_sdss.npz"))
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rows = 81920
d = synth.make_spectra(P, mu, grid, rows, seed=1, device=dev)
dl = DeviceDataloader(d["flux"], d["error"], d["zqso"], d["mask"], grid.wav(), batch_size=B, device=dev, shuffle=True, seed=5)
dl.rewind()
bufs, cursor = dl.graph_batch()
n = rows // B
for _ in range(3):
    dl.graph_fill()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    cursor.zero_()
    e0.record()
    for i in range(n):
        dl.graph_fill(); cursor += B
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / n * 1e3)
nbytes = B * (grid.Npix * 9 + grid.Npix * 9 + grid.Nb * 4)
print("gather+prepare B=%d QFA_GP_CTAS=%s: %.1f us  (%.2f TB/s of %d MB read+written; includes a cursor += B kernel)" %
      (B, os.environ.get("QFA_GP_CTAS", "default"), best, nbytes / best / 1e6, nbytes / 1e6))
