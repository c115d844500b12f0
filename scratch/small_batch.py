"""GPU scratch: train-step latency at the reference's default batch size (500) and a few others."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import QFA, Adam, step_scheduler, synth
dev = torch.device("cuda:0")
k = np.load('/root/repo/tests/golden/kat_sdss.npz')
P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
P["c0"] = P["beta"].clone(); mu = torch.tensor(k["param_mu"])
grid = synth.GRIDS["sdss"]
d = synth.make_spectra(P, mu, grid, 8192, seed=1, device=dev)
for prec in ("mixed", "tf32"):
    m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={a: b.numpy() for a, b in P.items()}, precision=prec); m.mu = mu
    opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
    for B in ([int(a) for a in sys.argv[1:]] or [500, 1024, 2048, 8192]):
        X, E, Z, M = d["delta"][:B], d["error"][:B], d["zabs"][:B], d["mask"][:B].view(torch.uint8)
        def step():
            acc = m.accumulate(X, E, Z, M, zero=True)
            opt.update_from_acc(m, acc)
        for _ in range(5): step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50): step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        print(f"{prec:6s} B={B:5d}: {dt*1e6:8.1f} us/step  {B/dt/1e6:7.2f} M spectra/s")
