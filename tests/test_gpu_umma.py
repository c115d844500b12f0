"""GPU: pins the tcgen05 conventions (smem descriptor, SWIZZLE_128B image, K-step advance, sub-tile
offset, TMEM lane mapping, bulk copy + mbarrier) that the tensor-core kernels rely on."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sw128_image(Bm):
    """[rows][32] float matrix -> flat image in the order qfa_umma.cuh::sw128_offset lays it out."""
    rows = Bm.shape[0]
    img = np.zeros(rows * 32, np.float32)
    for r in range(rows):
        for k in range(32):
            off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2)
            img[off // 4] = Bm[r, k]
    return img


def tf32_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("split", [0, 1])
def test_umma_selftest(split):
    from qfa_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    A = rng.standard_normal((128, 32)).astype(np.float32)
    Bm = rng.standard_normal((48, 32)).astype(np.float32)
    Bhi = tf32_round(Bm)
    Blo = (Bm - Bhi).astype(np.float32)
    dA, dBh, dBl = (torch.tensor(x).cuda() for x in (A, sw128_image(Bhi), sw128_image(Blo)))
    D = torch.zeros(128, 64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.qfa_selftest_umma(p(dA), p(dBh), p(dBl), p(D), split, p(err), None), "qfa_selftest_umma")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "mbarrier wait timed out"
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    got = D.cpu().numpy()
    tol = 2e-5 if split else 1e-2
    assert np.abs(got[:, :48] - ref).max() < tol * np.abs(ref).max()
    assert np.abs(got[:, 48:] - ref[:, 32:48]).max() < tol * np.abs(ref).max()
    if not split:   # and the plain product really is TF32: matches the rounded-operand product much tighter
        ref_tf = tf32_round(A).astype(np.float64) @ Bhi.astype(np.float64).T
        assert np.abs(got[:, :48] - ref_tf).max() < 2e-5 * np.abs(ref).max()


def test_tma2d_selftest_pitched_rows():
    """2-D TMA box loads from a padded-pitch float array (the building block of the planned TMA input path): interior box,
    a box that hangs over the right edge (pixels >= npix arrive as 0) and over the last row; and the dense odd-pitch layout
    of the reference (1913 floats per row) is rejected, which is why today's kernels load per thread."""
    from qfa_b200 import _lib
    L = _lib.lib()
    rows, npix, pitch = 300, 1913, 1920
    src = torch.arange(rows * pitch, dtype=torch.float32, device="cuda").reshape(rows, pitch)
    out = torch.empty(120, 32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for x0, y0 in ((64, 8), (1900, 0), (32, 250)):
        out.fill_(-1.0)
        _lib.check(L.qfa_selftest_tma2d(p(src), rows, npix, pitch, x0, y0, p(out), p(err), None), "qfa_selftest_tma2d")
        torch.cuda.synchronize()
        assert int(err.item()) == 0
        ref = torch.zeros(120, 32, device="cuda")
        r1, c1 = min(rows, y0 + 120), min(npix, x0 + 32)
        ref[:r1 - y0, :c1 - x0] = src[y0:r1, x0:c1]
        assert torch.equal(out, ref), (x0, y0)
    dense = torch.zeros(rows * npix, device="cuda")
    assert L.qfa_selftest_tma2d(p(dense), rows, npix, npix, 0, 0, p(out), p(err), None) == -6      # QFA_ERR_ALIGN

