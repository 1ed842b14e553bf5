"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference modules.

Works only where the reference checkout exists (the build container:
``/root/reference``; never on the GPU box).  The reference imports
``matplotlib.pyplot`` without using it (``models/pooling/nfp.py:12``); that
package is not installed here, so an empty stub is injected before import.
Nothing is copied out of the reference tree.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference checkout (build container), else the copy of its two operator files that __graft_entry__.build()
# leaves under the git-ignored baseline/_ref/ (it travels to the GPU box with the repo snapshot)
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("NFP_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/models/pooling/nfp.py") else STAGED_ROOT)
OPERATOR_FILES = ("models/__init__.py", "models/pooling/nfp.py", "models/NFP_Pooling.py")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "pooling", "nfp.py"))


def stage_reference(src_root: str = "/root/reference") -> bool:
    """Copy the reference's two operator files (unmodified) to baseline/_ref/ so that bench.py can time the REAL
    reference on a box that has no /root/reference.  baseline/_ref/ is git-ignored: nothing enters the history."""
    import shutil
    if not os.path.isfile(os.path.join(src_root, "models", "pooling", "nfp.py")):
        return os.path.isfile(os.path.join(STAGED_ROOT, "models", "pooling", "nfp.py"))
    for rel in OPERATOR_FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
    return True


def load_reference():
    """Return ``(NFPPooling, nfp_pooling)`` classes of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if getattr(sys.modules.get("matplotlib"), "pyplot", None) is None:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # The drop-in package may have registered itself under the reference's import
    # paths (neighbour_feature_pooling_b200.install_dropin); make sure we get the
    # real reference here.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "models" or k.startswith("models.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from models.pooling.nfp import NFPPooling  # type: ignore
        from models.NFP_Pooling import nfp_pooling  # type: ignore
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    return NFPPooling, nfp_pooling
