"""GPU scratch: Npix 1000 / Nh 32 accumulate at small batches, float CUDA-core path vs forced tensor-core path."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
grid = synth.GRIDS["l32"]
P, mu = synth.smooth_random_params(grid, 32, seed=1237)
d = synth.make_spectra(P, mu, grid, 4096, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
Pn = {k: v.numpy() for k, v in P.items()}
for prec in ("fp32", "tf32"):
    m = QFA(grid.Nb, grid.Nr, 32, torch.device("cuda:0"), model_params=Pn, precision=prec)
    for B in (32, 64, 128, 256, 512, 1024, 4096):
        a = (d["delta"][:B], d["error"][:B], d["zabs"][:B], d["mask"][:B].view(torch.uint8))
        for _ in range(3): m.accumulate(*a)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): m.accumulate(*a)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"{prec} B={B:5d}: {dt*1e6:8.1f} us  {B/dt/1e6:6.2f} M spectra/s")
