"""GPU scratch (build the library with QFA_ENABLE_TRACE=1 first: python -c "from qfa_b200 import _lib; _lib.build(force=True)"): clock64 trace of k_tc_grad CTA (0,0)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth, _lib
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 120 * 2
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32"); m.mu = mu
E, Z, M, D = data["error"], data["zabs"], data["mask"].view(torch.uint8), data["delta"]
NCH = 400
tr = torch.zeros(2 * NCH * 16 * 8, dtype=torch.int64, device="cuda")
L = _lib.lib()
for _ in range(2): m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
L.qfa_debug_set_trace_grad(ctypes.c_void_p(tr.data_ptr()))
m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
L.qfa_debug_set_trace_grad(None)
def analyse(t, nch):
  names = ["wait tm_full", "ldtm", "compute", "wait gen_empty", "sts+fence", "loads", "arrive+issue"]
  d = np.diff(t, axis=2)
  print("chunks traced:", nch, " total cycles:", t[-1, :, 7].max() - t[0, :, 0].min())
  print("period per chunk (mean):", (t[-1, :, 0].mean() - t[2, :, 0].mean()) / (nch - 3))
  for i, nm in enumerate(names):
    print(f"{nm:16s} mean {d[4:-2, :, i].mean():8.0f}  max {d[4:-2, :, i].max():8d}")
  print("per-warp sum of arrive+issue:", d[:, :, 6].sum(0))
  print("per-warp sum of waits:", (d[:, :, 0] + d[:, :, 3]).sum(0))

  t0 = t[10:13, :, :] - t[10, :, 0].min()
  print('chunks 10..12 raw (rows = warps 0,1,5,15; cols = stamps)')
  for c in range(3):
    for w in (0, 1, 5, 15): print(c + 10, w, list(t0[c, w]))
tall = tr.cpu().numpy().reshape(2, NCH, 16, 8)
for which, t in (("BLUE CTA 0", tall[0]), ("RED last CTA", tall[1])):
  print("=====", which)
  nch = int((t[:, 0, 0] != 0).sum())
  t = t[:nch]
  analyse(t, nch)

