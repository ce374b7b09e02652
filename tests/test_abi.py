"""The C-ABI library builds, loads and exports every symbol include/gpfo.h declares (no compute calls, no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpfo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpfo_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(libgpfo):
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(libgpfo, name), "libgpfo.so does not export %s" % name
    from gpflowopt_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding and header disagree"


def test_version_and_status_strings(libgpfo):
    assert libgpfo.gpfo_version() == 200
    assert libgpfo.gpfo_status_string(0) == b"ok"
    assert b"positive definite" in libgpfo.gpfo_status_string(1)
    assert libgpfo.gpfo_status_string(99) == b"unknown status"


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "gpfo.h")).read()
    assert "torch" not in text.lower().replace("pytorch tensors", "") or "at::" not in text
    assert 'extern "C"' in text


def test_enum_values_match_binding():
    from gpflowopt_b200 import _lib
    text = open(os.path.join(ROOT, "include", "gpfo.h")).read()
    for name, val in [("GPFO_OP_EI", _lib.OP_EI), ("GPFO_OP_POI", _lib.OP_POI), ("GPFO_OP_POF", _lib.OP_POF),
                      ("GPFO_OP_LCB", _lib.OP_LCB), ("GPFO_OP_HVPOI", _lib.OP_HVPOI), ("GPFO_OP_SUM", _lib.OP_SUM),
                      ("GPFO_OP_PROD", _lib.OP_PROD), ("GPFO_OP_SCALE", _lib.OP_SCALE), ("GPFO_OP_MEAN", _lib.OP_MEAN),
                      ("GPFO_OP_VAR", _lib.OP_VAR), ("GPFO_KERN_RBF", _lib.KERN_RBF),
                      ("GPFO_KERN_MATERN52", _lib.KERN_MATERN52), ("GPFO_KERN_MATERN32", _lib.KERN_MATERN32),
                      ("GPFO_ERR_NOT_SPD", _lib.ERR_NOT_SPD), ("GPFO_ERR_STATE", _lib.ERR_STATE)]:
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, text)
        assert m and int(m.group(1)) == val, name
    for name, val in [("GPFO_MAX_GPS", _lib.MAX_GPS), ("GPFO_MAX_SLOTS", _lib.MAX_SLOTS), ("GPFO_MAX_OPS", _lib.MAX_OPS),
                      ("GPFO_MAX_PARAMS", _lib.MAX_PARAMS), ("GPFO_MAX_DIM", _lib.MAX_DIM)]:
        m = re.search(r"#define\s+%s\s+(\d+)" % name, text)
        assert m and int(m.group(1)) == val, name


def test_product_fails_loudly_without_a_device(libgpfo):
    """No CPU fallback: without a CUDA device creating a context is an error, never a silent slow path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gpflowopt_b200 import _lib
    with pytest.raises(_lib.GpfoError):
        _lib.Context(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gpflowopt_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "gpflowopt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "oracle." not in src and "oracle/" not in src, os.path.join(dirpath, f)


def test_fast_normal_cdf_of_the_hvpoi_kernel(libgpfo):
    """csrc/ndtr_fast.cuh (same source on host and device) against scipy.special.ndtr -- itself the float64 erf/erfc forms TF's
    ndtr uses -- and against 50-digit mpmath in the tails: relative error of a few ulp wherever the value
    is a normal number, exact limits at +-inf."""
    import mpmath as mp
    from scipy.special import ndtr
    from gpflowopt_b200 import _lib
    z = np.concatenate((np.linspace(-37.0, 9.0, 200001), np.random.RandomState(0).randn(50000) * 3.0, [0.0, -0.0, 1e-300, -1e-300]))
    got = _lib.ndtr_fast_host(z)
    ref = ndtr(z)
    rel = np.abs(got - ref) / ref
    assert np.max(rel[z > -30]) < 1e-13, np.max(rel[z > -30])          # |z| <= 30: Phi down to 5e-198
    assert np.max(np.abs(got - ref)) < 5e-16                            # absolute: what the cell sums see
    assert np.all(np.diff(got[:200001]) >= 0.0)                         # monotone on the grid
    np.testing.assert_array_equal(_lib.ndtr_fast_host([np.inf, -np.inf]), [1.0, 0.0])
    mp.mp.dps = 40
    for zz in (-35.0, -20.0, -8.3, -1.0, -0.3, 0.3, 1.0, 5.0):
        true = float(mp.ncdf(mp.mpf(zz)))
        assert abs(_lib.ndtr_fast_host([zz])[0] - true) <= 2e-14 * true + 3e-17


def test_native_numpy_stream_design_rows_are_np_random_rand(libgpfo):
    """gpfo_mt19937_design_rows continues np.random's global MT19937 stream: the rows equal RandomDesign.generate() on the same
    seed bit for bit (any chunking, odd stream positions, a buffer handed in) and np.random resumes where NumPy itself would
    (gpflowopt/design.py:55-70, 83-94)."""
    from gpflowopt_b200.design import RandomDesign
    from gpflowopt_b200.domain import ContinuousParameter
    from gpflowopt_b200.optim import MCOptimizer
    dom = np.sum([ContinuousParameter("x%d" % i, -5 + i, 10 + 0.37 * i) for i in range(5)])
    for seed, sizes in ((0, (4000,)), (3, (1, 7, 1000, 2992)), (12345, (311, 312, 313, 3064))):
        np.random.seed(seed)
        np.random.rand(seed % 3 + 1)                         # odd / even position inside the 624-word block
        ref = RandomDesign(sum(sizes), dom).generate()
        after_ref = np.random.rand(4)
        np.random.seed(seed)
        np.random.rand(seed % 3 + 1)
        opt = MCOptimizer(dom, sum(sizes))
        buf = np.empty((sizes[0], 5))
        rows = [opt._get_eval_rows(0, sizes[0], out=buf).copy()] + [opt._get_eval_rows(0, k) for k in sizes[1:]]
        assert np.array_equal(np.vstack(rows), ref)
        assert np.array_equal(np.random.rand(4), after_ref)
        assert ref in dom
