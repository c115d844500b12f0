"""qfa_peer_allreduce protocol check on ONE device: `world` ranks are emulated by `world` buffers and streams of the same GPU
(their kernels run concurrently and really wait for each other's flags).  Run in a process of its own by
tests/test_gpu_aux.py (with QFA_PEER_TIMEOUT_S=30): a protocol bug ends in the kernel's trap, which would take the caller's
CUDA context with it.
Prints PEER-OK on success."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qfa_b200 import _lib  # noqa: E402


def run(world, n, dtype, steps, skew):
    L = _lib.lib()
    dev = torch.device("cuda:0")
    prec = _lib.PRECISIONS["fp64"] if dtype == torch.float64 else _lib.PRECISIONS["fp32"]
    nbytes = L.qfa_peer_buffer_bytes(n, _lib.PRECISIONS["fp64"], world)
    assert nbytes >= 128 + 2 * n * 8
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    states = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
    base = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(world * 1000 + n)
    accs = [[torch.randn(n, generator=g, device=dev, dtype=dtype) for _ in range(world)] for _ in range(steps)]
    want = []
    for s in range(steps):
        t = accs[s][0].clone()
        for q in range(1, world):
            t = t + accs[s][q]                       # rank order, like the kernel
        want.append(t)
    torch.cuda.synchronize()
    order = [(s, q) for s in range(steps) for q in range(world)]
    if skew:                                         # rank 0 enqueues two steps before the others enqueue any
        order = [(s, q) for q in range(world) for s in range(steps)] if steps <= 2 else \
                [(0, 0), (1, 0)] + [(s, q) for s in range(steps) for q in range(world) if not (q == 0 and s < 2)]
    for s, q in order:
        rc = L.qfa_peer_allreduce(ctypes.c_void_p(accs[s][q].data_ptr()), n, prec, ctypes.c_void_p(base.data_ptr()),
                                  ctypes.c_void_p(states[q].data_ptr()), world, q, ctypes.c_void_p(streams[q].cuda_stream))
        _lib.check(rc, "qfa_peer_allreduce")
    torch.cuda.synchronize()
    for s in range(steps):
        for q in range(world):
            assert torch.equal(accs[s][q], want[s]), (world, n, dtype, s, q)
    for q in range(world):
        assert states[q].tolist() == [steps, 0], states[q].tolist()


if __name__ == "__main__":
    L = _lib.lib()
    # argument checks happen on the host, before any launch
    assert L.qfa_peer_allreduce(None, 4, 1, None, None, 1, 0, None) == -1
    d = torch.zeros(8, device="cuda")
    assert L.qfa_peer_allreduce(ctypes.c_void_p(d.data_ptr()), 4, 1, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(d.data_ptr()),
                                33, 0, None) == -2
    assert L.qfa_peer_buffer_bytes(10, 1, 33) == 0
    run(1, 22424, torch.float32, 3, False)           # one rank: the sum is the input
    run(2, 22424, torch.float32, 4, False)
    run(2, 22427, torch.float32, 5, True)            # scalar tail (n % 4 = 3), rank 0 a step ahead of rank 1
    run(3, 1001, torch.float64, 4, True)
    # the emulation needs the CTAs of ALL ranks co-resident on this one GPU (a real rank has the device to itself):
    run(8, 40000, torch.float32, 3, False)           # 8 ranks x 40 CTAs
    run(4, 105000, torch.float32, 3, True)           # DESI-sized accumulator (103 CTAs per rank)
    print("PEER-OK")
