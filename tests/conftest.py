import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def assert_parity(got, ref, rtol=1e-9, floor=1e-12, what=""):
    """SURVEY.md 8c criterion: |got-ref| <= rtol*|ref| + floor*max|ref|, plus a pure relative check on
    points that are not negligible (|ref| >= 1e-6 max|ref|)."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, "%s shape %s vs %s" % (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    assert np.all(np.isfinite(got)), what + ": non-finite values"
    m = np.max(np.abs(ref))
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + floor * m
    worst = int(np.argmax(err - bound))
    assert np.all(err <= bound), "%s: |d|=%g at %d (ref %g, max %g)" % (what, err.flat[worst], worst, ref.flat[worst], m)
    big = np.abs(ref) >= 1e-6 * m
    if np.any(big):
        rel = np.max(err[big] / np.abs(ref[big]))
        assert rel <= rtol, "%s: max rel err %g" % (what, rel)


@pytest.fixture(scope="session")
def cuda_lib():
    from radtxfr_b200 import _lib
    return _lib.lib()
