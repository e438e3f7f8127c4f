import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def nd():
    import ndarray_b200
    return ndarray_b200


@pytest.fixture(scope="session")
def ctx(nd):
    return nd.default_context()


def device_view_of(nd, base_dev, base_np: np.ndarray, view_np: np.ndarray):
    """The device view with the same (offset, shape, strides) over `base_dev` that `view_np` has over `base_np`."""
    item = base_np.itemsize
    off = (view_np.__array_interface__["data"][0] - base_np.__array_interface__["data"][0]) // item
    strides = [s // item for s in view_np.strides]
    return base_dev.raw_view_with(off, view_np.shape, strides)


def bits(a: np.ndarray) -> np.ndarray:
    """Bit pattern with NaNs canonicalised (x86 and NVIDIA default NaNs differ in payload/sign)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        u = a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64).copy()
        u[np.isnan(a)] = 0
        return u
    return a


def assert_bit_equal(got: np.ndarray, want: np.ndarray, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    g, w = bits(got), bits(want)
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {g.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}")
