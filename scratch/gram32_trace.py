"""GPU scratch (library built with QFA_ENABLE_TRACE=1): clock64 trace of k_tc_gram32, CTA 0, first tile, all three passes."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth, _lib
grid = synth.GRIDS["l32"]
P, mu = synth.smooth_random_params(grid, 32, seed=1237)
d = synth.make_spectra(P, mu, grid, 17760, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
m = QFA(grid.Nb, grid.Nr, 32, torch.device("cuda:0"), model_params={k: v.numpy() for k, v in P.items()}, precision="tf32")
args = (d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8))
for _ in range(2): m.accumulate(*args)
torch.cuda.synchronize()
nkb = (grid.Npix + 31) // 32
tr = torch.zeros(3 * nkb * 16 * 4, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.qfa_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
m.accumulate(*args)
torch.cuda.synchronize()
L.qfa_debug_set_trace(None)
t = tr.cpu().numpy().reshape(3, nkb, 16, 4)[:, :, :15, :]
for p in range(3):
    w = t[p, :, :, 1] - t[p, :, :, 0]; c = t[p, :, :, 2] - t[p, :, :, 1]; tl = t[p, :, :, 3] - t[p, :, :, 2]
    per = (t[p, -1, :, 0].mean() - t[p, 2, :, 0].mean()) / (nkb - 3)
    print(f"pass {p}: period {per:7.0f}  wait_empty {w[3:].mean():7.0f} (max {w[3:].max()})  compute+sts {c[3:].mean():7.0f}  fence+arrive+loads {tl[3:].mean():7.0f}")
    print("   red K-blocks (kb>=13): wait %.0f comp %.0f tail %.0f ; blue (kb<11): wait %.0f comp %.0f tail %.0f" % (
        w[13:].mean(), c[13:].mean(), tl[13:].mean(), w[3:11].mean(), c[3:11].mean(), tl[3:11].mean()))
print("pass spans:", [int(t[p, -1, :, 3].max() - t[p, 0, :, 0].min()) for p in range(3)], " between passes:", [int(t[p + 1, 0, :, 0].min() - t[p, -1, :, 3].max()) for p in range(2)])
