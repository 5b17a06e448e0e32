"""Import shim that makes the UNMODIFIED reference (/root/reference, read-only)
importable in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (fixture generation, run
in the build container where /root/reference exists), by `bench.py --impl reference`
and by the drop-in tests (on the GPU box through the copy oracle/_ref that
oracle/make_ref.py makes) -- never by the product.
"""
import functools
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_COPY = os.path.join(os.path.dirname(_HERE), "_ref")        # oracle/_ref: made by oracle/make_ref.py, travels to the GPU box


def _default_root():
    if os.path.isdir("/root/reference/recbole"):
        return "/root/reference"
    return _REF_COPY


REFERENCE_ROOT = os.environ.get("FAIRREC_REFERENCE_ROOT") or _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "recbole"))


def install():
    """Apply the 4-part shim, then make `import recbole` resolve to the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (REFERENCE_ROOT, here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    import torch
    if not getattr(torch.load, "_fairrec_shim", False):
        patched = functools.partial(torch.load, weights_only=False)
        patched._fairrec_shim = True
        torch.load = patched
    import scipy.sparse as sp
    if not hasattr(sp.dok_matrix, "_update"):
        sp.dok_matrix._update = lambda self, d: [self.__setitem__(k, v) for k, v in d.items()]
    import recbole  # noqa: F401
    return recbole
