"""CPU: libqfa_b200.so builds for sm_100a, loads, and exports every symbol include/qfa_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from qfa_b200 import _lib


def declared_symbols(header="qfa_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qfa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    _lib.build()
    names = declared_symbols()
    assert len(names) >= 12
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    # the test / design / trace entry points live in their own header, not in the reference-facing one
    dbg = declared_symbols("qfa_b200_debug.h")
    assert set(dbg) == set(_lib.DEBUG_SIGNATURES), set(dbg) ^ set(_lib.DEBUG_SIGNATURES)
    assert not set(dbg) & set(names)
    for n in dbg:
        assert hasattr(L, n), n
    assert not any("selftest" in n or "bench" in n or "debug" in n for n in names)


def test_host_side_entry_points_without_gpu():
    L = _lib.lib()
    assert L.qfa_abi_version() == 2
    assert L.qfa_param_len(720, 1193, 8) == 1913 * 8 + 1913 + 720 + 3          # reference model.py:42
    assert L.qfa_acc_len(720, 1193, 8) == 1913 * 8 + 1913 + 720 + 3 + 1913 + 3 + 2 + 1913
    assert L.qfa_train_workspace_bytes(720, 1193, 8, 500, 1) > 0
    # argument errors are reported on the host, before any launch, with a message
    rc = L.qfa_clip(None, 720, 1193, 8, 1e-3, 2.0, None)
    assert rc == -1 and b"NULL" in L.qfa_last_error_string()
    rc = L.qfa_clip(ctypes.c_void_p(16), 720, 1193, 64, 1e-3, 2.0, None)
    assert rc == -3
    with pytest.raises(_lib.QfaError):
        _lib.check(rc, "qfa_clip")
    # the production build keeps no trace pointer: the debug setters refuse
    assert L.qfa_debug_set_trace(None) == -8
    assert L.qfa_launch_count() == 0
    assert L.qfa_ood_select(None, 0, 0.0, 4096, 0, None, None, None, None, None) == -8
    # peer all-reduce: buffer = 128 B of flags + two 16-byte-rounded copies of the accumulator; argument errors before any launch
    n = L.qfa_acc_len(720, 1193, 8)
    assert L.qfa_peer_buffer_bytes(n, 1, 8) == 128 + 2 * ((4 * n + 15) // 16 * 16)
    assert L.qfa_peer_buffer_bytes(n, 0, 8) == 128 + 2 * ((8 * n + 15) // 16 * 16)          # QFA_PREC_FP64: double accumulator
    assert L.qfa_peer_buffer_bytes(n, 1, 33) == 0 and L.qfa_peer_buffer_bytes(0, 1, 2) == 0
    assert L.qfa_peer_allreduce(None, n, 1, None, None, 2, 0, None) == -1
    assert L.qfa_peer_allreduce(ctypes.c_void_p(16), n, 1, ctypes.c_void_p(16), ctypes.c_void_p(16), 2, 2, None) == -2
    assert L.qfa_peer_allreduce(ctypes.c_void_p(20), n, 1, ctypes.c_void_p(16), ctypes.c_void_p(16), 2, 0, None) == -6


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)
