"""GPU scratch: predict_host (pinned host buffers in, pinned host results out) at several chunk sizes."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth
dev = torch.device("cuda:0")
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
B = 100000
d = synth.make_spectra(P, mu, grid, B, seed=1234, device=dev)
m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision="mixed"); m.mu = mu
hX, hE, hZ, hM = (t.cpu().pin_memory() for t in (d["flux"], d["error"], d["zabs"], d["mask"].view(torch.uint8)))
del d; torch.cuda.empty_cache()
hout = {"nll": torch.empty(B).pin_memory(), "cont": torch.empty(B, grid.Npix).pin_memory(), "unc": torch.empty(B, grid.Npix).pin_memory()}
h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM)); d2h = sum(t.numel() * t.element_size() for t in hout.values())
for chunk in [int(a) for a in sys.argv[1:]] or [8192, 4096, 2048, 16384]:
    for _ in range(2): m.predict_host(hX, hE, hZ, hM, out=hout, want=("nll", "cont", "unc"), chunk=chunk)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): m.predict_host(hX, hE, hZ, hM, out=hout, want=("nll", "cont", "unc"), chunk=chunk)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"chunk {chunk:6d}: {dt*1e3:7.2f} ms  {B/dt/1e6:.3f} M spectra/s  H2D {h2d/dt/1e9:.1f} GB/s  D2H {d2h/dt/1e9:.1f} GB/s")
