"""ctypes binding of libqfa_b200.so (C ABI declared in include/qfa_b200.h).

The shared library is built IN-TREE by `qfa_b200._lib.build()` (called from
__graft_entry__.build()) with nvcc for sm_100a only.  There is no CPU fallback:
if the library is missing or a call fails, a QfaError is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libqfa_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "qfa_capi.cu")]
HEADERS = [os.path.join(_ROOT, "include", "qfa_b200.h"), os.path.join(_ROOT, "include", "qfa_b200_debug.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

PREC_FP64, PREC_FP32, PREC_TF32 = 0, 1, 2
# "mixed": tensor cores above the path's cross-over batch size (predict 1280, train 800 / 192; env QFA_TC_MIN_BATCH),
# float CUDA cores below; "tf32": always tensor cores
PRECISIONS = {"fp64": PREC_FP64, "fp32": PREC_FP32, "mixed": PREC_TF32, "tf32": PREC_TF32, "tf32x3": PREC_TF32}
TAU_LAWS = {"becker": 0, "fg": 1, "kamble": 2, "mock": 3}
FLAG_ZERO_ACC = 1
FLAG_FORCE_TENSOR = 2
FLAG_SOLVE_FP64 = 4
FLAG_TF32X3 = 8
ABI_VERSION = 2


class QfaError(RuntimeError):
    pass


class QfaModelStruct(ctypes.Structure):
    _fields_ = [("Nb", ctypes.c_int32), ("Nr", ctypes.c_int32), ("Nh", ctypes.c_int32),
                ("tau_law", ctypes.c_int32), ("params", ctypes.c_void_p), ("mu", ctypes.c_void_p)]


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = SOURCES + HEADERS + [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libqfa_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DQFA_ENABLE_TRACE"] if os.environ.get("QFA_ENABLE_TRACE") else []     # clock64 stamps (include/qfa_b200_debug.h)
    extra += os.environ.get("QFA_NVCC_EXTRA", "").split()                            # A/B variants (scripts/ab_variants.sh)
    cmd = [nvcc] + NVCC_FLAGS + extra + SOURCES + ["-o", LIB_PATH + ".tmp"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise QfaError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


_lib = None

_VP, _I, _F, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_D, _LL, _ULL = ctypes.c_double, ctypes.c_longlong, ctypes.c_ulonglong
_MP = ctypes.POINTER(QfaModelStruct)

# name -> (restype, argtypes); must list every symbol include/qfa_b200.h declares
SIGNATURES = {
    "qfa_abi_version": (_I, []),
    "qfa_last_error_string": (ctypes.c_char_p, []),
    "qfa_param_len": (_SZ, [_I, _I, _I]),
    "qfa_acc_len": (_SZ, [_I, _I, _I]),
    "qfa_train_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "qfa_predict_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "qfa_train_accumulate": (_I, [_MP, _VP, _VP, _VP, _VP, _I, _VP, _SZ, _VP, _VP, _I, _I, _VP]),
    "qfa_grads_finalize": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    "qfa_predict": (_I, [_MP, _VP, _VP, _VP, _VP, _I, _VP, _SZ, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
    "qfa_adam_clip_step": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _F, _F, _F, _F, _F, _F, _F, _F, _F, _VP]),
    "qfa_adam_clip_step_dev": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _F, _F, _F, _F, _F, _F, _VP, _D, _VP, _LL,
                                    _VP]),
    "qfa_clip": (_I, [_VP, _I, _I, _I, _F, _F, _VP]),
    "qfa_smooth": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "qfa_prepare_batch": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    "qfa_gather_prepare": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "qfa_mean_spectrum_sums": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _VP]),
    "qfa_ood_select": (_I, [_VP, _I, _F, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "qfa_sample_posterior": (_I, [_MP, _VP, _VP, _I, _I, _ULL, _VP, _VP, _VP, _VP]),
    "qfa_peer_buffer_bytes": (_SZ, [_LL, _I, _I]),
    "qfa_peer_allreduce": (_I, [_VP, _LL, _I, _VP, _VP, _I, _I, _VP]),
    "qfa_launch_count": (_ULL, []),
}
# include/qfa_b200_debug.h: hardware self-tests, design micro-benchmarks, trace hooks (not reference-facing)
DEBUG_SIGNATURES = {
    "qfa_selftest_umma": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP]),
    "qfa_selftest_tma2d": (_I, [_VP, _I, _I, _I, _I, _I, _VP, _VP, _VP]),
    "qfa_bench_tma2d": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    "qfa_bench_ldg": (_I, [_VP, _I, _I, _I, _VP, _VP]),
    "qfa_debug_set_trace": (_I, [_VP]),
    "qfa_debug_set_trace_grad": (_I, [_VP]),
}


def lib():
    """Load (once) and return the ctypes handle. Raises QfaError if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise QfaError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(no CPU fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().qfa_last_error_string().decode("utf-8", "replace")
        raise QfaError(f"{what} failed (code {rc}): {msg}")
