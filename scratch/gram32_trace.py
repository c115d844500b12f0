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
tr = torch.zeros(nkb * 16 * 4 + 2 * nkb * 8, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.qfa_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
m.accumulate(*args)
torch.cuda.synchronize()
L.qfa_debug_set_trace(None)
full = tr.cpu().numpy()
t = full[:nkb * 64].reshape(nkb, 16, 4)[:, :15, :]
rp = full[nkb * 64:].reshape(2, nkb, 8)
w = t[:, :, 1] - t[:, :, 0]; c = t[:, :, 2] - t[:, :, 1]; tl = t[:, :, 3] - t[:, :, 2]
per = (t[-1, :, 0].mean() - t[2, :, 0].mean()) / (nkb - 3)
print(f"pass 0: period {per:7.0f}  wait_empty {w[3:].mean():7.0f} (max {w[3:].max()})  compute+sts {c[3:].mean():7.0f}  fence+arrive+loads {tl[3:].mean():7.0f}")
print("   red K-blocks (kb>=13): wait %.0f comp %.0f tail %.0f ; blue (kb<11): wait %.0f comp %.0f tail %.0f" % (
    w[13:].mean(), c[13:].mean(), tl[13:].mean(), w[3:11].mean(), c[3:11].mean(), tl[3:11].mean()))
print("pass 0 span:", int(t[-1, :, 3].max() - t[0, :, 0].min()))
for p in range(2):
    r = rp[p]
    print("replay pass %d: span %d, period %.0f ; wait tiles %.0f, wait image %.0f, MMA issue %.0f, wait stage free + issue copies %.0f" % (
        p + 1, r[-1, 4] - r[0, 0], (r[-1, 0] - r[4, 0]) / (nkb - 5), (r[4:, 1] - r[4:, 0]).mean(), (r[4:, 2] - r[4:, 1]).mean(),
        (r[4:, 3] - r[4:, 2]).mean(), (r[4:, 4] - r[4:, 3]).mean()))
print("pass 0 start -> replay 1 start %d -> replay 2 start %d -> end %d" % (rp[0, 0, 0] - t[0, :, 0].min(), rp[1, 0, 0] - t[0, :, 0].min(), rp[1, -1, 4] - t[0, :, 0].min()))
