"""CPU: the C-ABI shared library builds (nvcc cross-compiles without a GPU),
loads, and exports every entry point include/b2rl.h declares; the ctypes
signature table covers the same set.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "b2rl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2rl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pfrl_b200 import _lib

    path = _lib.build()
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libb2rl.so does not export %s" % name
    assert set(names) == set(_lib.SIGNATURES), (
        set(names) ^ set(_lib.SIGNATURES))
    L = _lib.load()
    assert L.b2rl_version().startswith(b"b2rl")


def test_gpu_classes_fail_loudly_without_a_device():
    """No CPU fallback: creating a store without a CUDA device is an error."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pfrl_b200 import _lib
    from pfrl_b200._lib import B2rlError, ReplayConfig

    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = ReplayConfig(capacity=8, part_capacity=8, part_bytes=16, stack=1, n_step=1,
                       action_bytes=8, prioritized=1, device=0, max_batch=8)
    rc = L.b2rl_replay_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"cuda" in L.b2rl_last_error().lower()
    with pytest.raises(B2rlError):
        _lib.check(rc)


def test_oracle_builds_and_is_not_imported_by_the_product():
    import oracle

    oracle.build()
    import subprocess
    import sys

    code = ("import sys; import pfrl_b200, pfrl_b200.agents, pfrl_b200.replay_buffers, "
            "pfrl_b200.experiments, pfrl_b200.envs; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), "
            "[m for m in sys.modules if m.startswith('oracle')]")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    hits = subprocess.run(["grep", "-rl", "--include=*.py", "-E", r"^\s*(from|import) oracle",
                           os.path.join(ROOT, "pfrl_b200")], capture_output=True, text=True)
    assert hits.stdout.strip() == ""


def test_host_priority_arithmetic_is_bit_identical_to_python_floats():
    """b2rl_host_priority_from_errors (libm pow inside the library) == the reference's list
    comprehension in Python floats (pfrl/replay_buffers/prioritized.py:47-55), bit for bit:
    this is what update_errors(host list) feeds the trees with."""
    import ctypes

    import numpy as np

    from pfrl_b200 import _lib

    L = _lib.load()
    rng = np.random.RandomState(0)
    err = np.concatenate([np.abs(rng.randn(20000)), rng.rand(20000) * 1e-3, rng.randn(5000) * 5,
                          np.array([0.0, 1.0, 1e-300, 0.5, 2.0, 123456.789])])
    for alpha, eps, lo, hi in ((0.5, 0.01, 0, 1), (0.6, 0.01, 0, 1), (0.7, 1e-6, None, None),
                               (1.0, 0.01, 0, None), (0.5, 0.01, None, 3.5)):
        want = []
        for d in err.tolist():
            if lo is not None:
                d = max(lo, d)
            if hi is not None:
                d = min(hi, d)
            want.append((d + eps) ** alpha if d + eps > 0 else float("nan"))
        got = np.empty_like(err)
        rc = L.b2rl_host_priority_from_errors(
            err.ctypes.data_as(ctypes.c_void_p), len(err), alpha, eps, int(lo is not None),
            0.0 if lo is None else float(lo), int(hi is not None), 0.0 if hi is None else float(hi),
            got.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        want = np.asarray(want, dtype=np.float64)
        ok = np.isfinite(want)
        assert np.array_equal(got[ok].view(np.uint64), want[ok].view(np.uint64)), (alpha, eps, lo, hi)
