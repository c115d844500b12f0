"""GPU scratch: latency of the captured train step (gather -> accumulate -> Adam) at small batches, SDSS shape, Nh 8."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qfa_b200 import QFA, Adam, step_scheduler, synth, DeviceDataloader
dev = torch.device("cuda:0")
k = np.load(os.path.join(ROOT, "tests", "golden", "kat_sdss.npz"))
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
out = []
for prec in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["mixed", "tf32"]):
    for B in (250, 500, 1024, 8192):
        rows = max(B * 16, 20000)
        d = synth.make_spectra(P, mu, grid, rows, seed=1, device=dev)
        dl = DeviceDataloader(d["flux"], d["error"], d["zqso"], d["mask"], grid.wav(), batch_size=B, device=dev, shuffle=True, seed=5)
        m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={kk: v.numpy() for kk, v in P.items()}, precision=prec); m.mu = mu
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
        dl.rewind()
        g = m.capture_train_step(opt, dl, rows // B)
        n = rows // B - 2
        for _ in range(3): g.replay()
        dl._cursor.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): g.replay()
        e1.record(); torch.cuda.synchronize()
        out.append("%s B=%d %.1f us" % (prec, B, e0.elapsed_time(e1) / n * 1e3))
        del d, dl, m, g
print(" | ".join(out))
