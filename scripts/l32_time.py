"""GPU scratch: time QFA.accumulate on the Npix 1000 / Nh 32 workload (tensor-core path)."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from qfa_b200 import QFA, synth
grid = synth.GRIDS["l32"]
P, mu = synth.smooth_random_params(grid, 32, seed=1237)
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = synth.make_spectra(P, mu, grid, Bn, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
Pn = {k: v.numpy() for k, v in P.items()}
m = QFA(grid.Nb, grid.Nr, 32, torch.device("cuda:0"), model_params=Pn, precision="tf32")
import os
m.solve_fp64 = bool(os.environ.get("QFA_SOLVE_FP64"))
args = (d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8))
for _ in range(3): m.accumulate(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.accumulate(*args)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"l32 accumulate: {ms:.3f} ms / {Bn} = {Bn/ms/1e3:.2f} M spectra/s")
