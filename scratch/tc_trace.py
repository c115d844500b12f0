"""GPU scratch (build the library with QFA_ENABLE_TRACE=1 first: python -c "from qfa_b200 import _lib; _lib.build(force=True)"): clock64 trace of k_tc_gram (CTA 0, first tile): where do the warps wait?"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, synth, _lib
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2
mode = sys.argv[2] if len(sys.argv) > 2 else "predict"
data = synth.make_spectra(P, mu, grid, Bn, seed=1234, device=torch.device("cuda:0"))
Pn = {key: v.numpy() for key, v in P.items()}
m = QFA(grid.Nb, grid.Nr, 8, torch.device("cuda:0"), model_params=Pn, precision="tf32"); m.mu = mu
X, E, Z, M = data["flux"], data["error"], data["zabs"], data["mask"].view(torch.uint8)
nkb = (grid.Npix + 31) // 32
tr = torch.zeros(nkb * 16 * 8 + 8, dtype=torch.int64, device="cuda")
L = _lib.lib()
if mode == "predict":
    o = m.predict_batch(X, E, Z, M)
    for _ in range(2): m.predict_into(X, E, Z, M, o)
    torch.cuda.synchronize()
    L.qfa_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    m.predict_into(X, E, Z, M, o)
else:
    D = data["delta"]
    for _ in range(2): m.accumulate(D, E, Z, M)
    torch.cuda.synchronize()
    L.qfa_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    m.accumulate(D, E, Z, M)
torch.cuda.synchronize()
L.qfa_debug_set_trace(None)
t = tr.cpu().numpy()
ph = t[nkb * 128:]
t = t[:nkb * 128].reshape(nkb, 16, 8)
t0 = ph[0]
print(f"phases: G {ph[1]-ph[0]} S {ph[2]-ph[1]} O {ph[3]-ph[2]} cycles")
tt = t - t0
wait = (t[:, :, 1] - t[:, :, 0]); comp = (t[:, :, 2] - t[:, :, 1]); tail = (t[:, :, 3] - t[:, :, 2])
fen = t[:, :, 4] - t[:, :, 2]; arr = t[:, :, 5] - t[:, :, 4]; lds = t[:, :, 6] - t[:, :, 5]; iss = t[:, :, 3] - t[:, :, 6]
print("tail split (mean over warps & K-blocks 24..55): fence %.0f arrive %.0f loads %.0f issue-duty %.0f (max %d)" % (
    fen[24:56].mean(), arr[24:56].mean(), lds[24:56].mean(), iss[24:56].mean(), iss[24:56].max()))
print("blue K-blocks 4..20: wait %.0f comp %.0f fence %.0f arrive %.0f loads %.0f issue %.0f" % (
    wait[4:21].mean(), comp[4:21].mean(), fen[4:21].mean(), arr[4:21].mean(), lds[4:21].mean(), iss[4:21].mean()))
print("per K-block means over warps: wait_empty, compute+sts, loads+arrive ; skew(enter) ; K-block period")
for kb in range(nkb):
    per = (tt[kb, :, 0].mean() - tt[kb - 1, :, 0].mean()) if kb else 0
    print(f"kb {kb:2d} wait {wait[kb].mean():7.0f} (max {wait[kb].max():6d}) comp {comp[kb].mean():7.0f} tail {tail[kb].mean():7.0f} "
          f"enter min {tt[kb,:,0].min():8d} max {tt[kb,:,0].max():8d} period {per:7.0f}")
print("per warp totals: wait, comp, tail")
for w in range(16):
    print(w, wait[:, w].sum(), comp[:, w].sum(), tail[:, w].sum())
