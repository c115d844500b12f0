"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (qfa_b200/).

Loads the *untouched* reference package from /root/reference (build container
only; the GPU box does not have it) so that golden vectors can be generated
and the restatements in this directory can be pinned against it.

Recipe follows SURVEY.md Appendix A:
  * `yacs` is absent in this image -> stub `yacs.config.CfgNode` with an
    attribute-dict (reference QFA/config.py:14-63 builds its defaults at import).
  * reference QFA/utils.py:144 opens './Lyman_series.csv' relative to the CWD
    -> chdir into <ref>/QFA for the duration of the import.
  * optional fp64 promotion of the very same code (own process!): the reference
    looks `torch.float32` up at call time (model.py:67-72,90-97,125,130,145-149;
    utils.py:31,53), and `Npix*log2pi` follows the default dtype (quirk Q5).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("QFA_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "QFA", "model.py"))


def load_reference(fp64: bool = False):
    """Returns (QFA_class, Adam_class, step_scheduler, utils_module)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if "yacs" not in sys.modules:
        y, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")

        class CN(dict):
            def __getattr__(s, k):
                try:
                    return s[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(s, k, v):
                s[k] = v

            def clone(s):
                return s

        yc.CfgNode = CN
        y.config = yc
        sys.modules["yacs"], sys.modules["yacs.config"] = y, yc
    import torch
    if fp64:
        torch.float32 = torch.float64
        torch.float = torch.float64
        torch.set_default_dtype(torch.float64)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF_ROOT, "QFA"))
    try:
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        from QFA.model import QFA
        from QFA.optimizer import Adam, step_scheduler
        import QFA.utils as rutils
    finally:
        os.chdir(cwd)
    return QFA, Adam, step_scheduler, rutils
