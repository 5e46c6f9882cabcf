"""Import the REAL reference (pfnet/pfrl) in the build container.

TEST INFRASTRUCTURE ONLY.  /root/reference is read-only and does not exist on
the GPU box, so this is used only by oracle/gen_golden.py and by the
``-m "not gpu"`` tests that re-validate the oracle live (skipped when the
reference is absent).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("PFRL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pfrl"))


def import_reference():
    """Return the reference's top-level ``pfrl`` module (or raise ImportError)."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the mount is read-only
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pfrl  # noqa: E402

    return pfrl
