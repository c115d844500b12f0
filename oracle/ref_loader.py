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


def find_reference():
    """Root of an importable copy of the UNMODIFIED reference: $QFA_REF, baseline/_ref (the pip --target install that
    travels to the GPU box), or /root/reference (build container only).  None if there is none."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for root in (os.environ.get("QFA_REF"), os.path.join(here, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "QFA", "model.py")):
            return root
    return None


def _csv_dir(root):
    """Directory to chdir into for the import.  The reference reads './Lyman_series.csv' from the CWD (utils.py:144); its
    setup.py does not ship the file (package_data lists only README/LICENSE), so for a pip-installed copy the same table
    (atomic data, qfa_b200.utils._LYMAN) is written to a scratch directory in the reference's column layout."""
    d = os.path.join(root, "QFA")
    if os.path.isfile(os.path.join(d, "Lyman_series.csv")):
        return d
    import tempfile
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    from qfa_b200.utils import _LYMAN
    tmp = tempfile.mkdtemp(prefix="qfa_ref_cwd_")
    with open(os.path.join(tmp, "Lyman_series.csv"), "w") as fh:
        fh.write("name,f,lambda,coeff\n")
        for f, lam in _LYMAN:
            fh.write(f"HI_{int(lam)},{f:.4e},{lam:.4f},-1\n")
    return tmp


def load_reference(fp64: bool = False, root: str = None):
    """Returns (QFA_class, Adam_class, step_scheduler, utils_module)."""
    global REF_ROOT
    if root is not None:
        REF_ROOT = root
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
    os.chdir(_csv_dir(REF_ROOT))
    try:
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        from QFA.model import QFA
        from QFA.optimizer import Adam, step_scheduler
        import QFA.utils as rutils
    finally:
        os.chdir(cwd)
    return QFA, Adam, step_scheduler, rutils
