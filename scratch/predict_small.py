"""GPU scratch: predict at small batches, float CUDA-core path vs forced tensor-core path (SDSS shape, Nh 8)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
dev = torch.device("cuda:0")
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
d = synth.make_spectra(P, mu, grid, 4096, seed=1, device=dev)
for prec in ("fp32", "tf32"):
    m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision=prec); m.mu = mu
    for B in (64, 128, 256, 512, 768, 1024, 2048, 4096):
        a = (d["flux"][:B], d["error"][:B], d["zabs"][:B], d["mask"][:B].view(torch.uint8))
        o = m.predict_batch(*a)
        for _ in range(3): m.predict_into(*a, o)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): m.predict_into(*a, o)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"{prec} B={B:5d}: {dt*1e6:8.1f} us  {B/dt/1e6:6.2f} M spectra/s")
