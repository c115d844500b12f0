"""GPU scratch: HBM read bandwidth of the 2-D TMA tile path (qfa_bench_tma2d) for box widths 32 / 64 / 128 pixels."""
import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
from qfa_b200 import _lib
L = _lib.lib()
rows, npix, pitch = 142080, 1913, 1920          # 1.09 GB: 8 waves of 120-row tiles on 148 SMs
src = torch.randn(rows, pitch, device="cuda")
sink = torch.zeros(4, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for bw in (32, 64, 128):
    for _ in range(2): _lib.check(L.qfa_bench_tma2d(p(src), rows, npix, pitch, bw, p(sink), p(err), None), "bench")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.check(L.qfa_bench_tma2d(p(src), rows, npix, pitch, bw, p(sink), p(err), None), "bench")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"box {bw:3d} px: {ms:.3f} ms  {rows*npix*4/ms/1e6:.0f} GB/s of payload  (err flag {int(err.item())})")
for name, pit in (("dense pitch 1913", 1913), ("padded pitch 1920", 1920)):
    a = torch.randn(rows, pit, device="cuda")
    for _ in range(2): _lib.check(L.qfa_bench_ldg(p(a), rows, npix, pit, p(sink), None), "ldg")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.check(L.qfa_bench_ldg(p(a), rows, npix, pit, p(sink), None), "ldg")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"per-thread loads, {name}: {ms:.3f} ms  {rows*(npix//32*32)*4/ms/1e6:.0f} GB/s of payload")

