"""GPU scratch (library built with QFA_ENABLE_TRACE=1): clock64 trace of k_tc_grad32, CTA (0,0), first 256 steps.
stamps per step: control [0] top, [1] after tm_empty wait, [2] after image wait, [3] after MMA issue + commit;
worker warp 5 [4] top, [5] after tm_full wait, [6] after tm_empty arrive (end of the step's math)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth, _lib
grid = synth.GRIDS["l32"]
P, mu = synth.smooth_random_params(grid, 32, seed=1237)
d = synth.make_spectra(P, mu, grid, 65536, seed=11, device=torch.device("cuda:0"), mask_iid=0.15, run_len=(40, 160))
m = QFA(grid.Nb, grid.Nr, 32, torch.device("cuda:0"), model_params={k: v.numpy() for k, v in P.items()}, precision="tf32")
args = (d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8))
for _ in range(2): m.accumulate(*args)
torch.cuda.synchronize()
tr = torch.zeros(256 * 8, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.qfa_debug_set_trace_grad(ctypes.c_void_p(tr.data_ptr()))
m.accumulate(*args)
torch.cuda.synchronize()
L.qfa_debug_set_trace_grad(None)
t = tr.cpu().numpy().reshape(256, 8)[16:240]
per = (t[-1, 0] - t[0, 0]) / (len(t) - 1)
print(f"period {per:.0f} cycles/step")
print("control: wait tm_empty %.0f, issue copies + wait images %.0f, 12 MMAs + commit %.0f, loop %.0f" % (
    (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean(), (t[:, 3] - t[:, 2]).mean(), (t[1:, 0] - t[:-1, 3]).mean()))
print("worker : wait tm_full %.0f, TMEM loads + math %.0f, cell loads + loop %.0f" % (
    (t[:, 5] - t[:, 4]).mean(), (t[:, 6] - t[:, 5]).mean(), (t[1:, 4] - t[:-1, 6]).mean()))
print("commit(n) -> worker sees full(n): %.0f ; worker arrive(n) -> control sees empty (step n+2 [1]): %.0f" % (
    (t[:, 5] - t[:, 3]).mean(), (t[2:, 1] - t[:-2, 6]).mean()))
for n in range(100, 106): print(n, (t[n] - t[100, 0]).tolist())
