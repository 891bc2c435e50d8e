"""pytest configuration: markers, import path, shared fixtures."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG_DIR = os.path.join(ROOT, "amd-versal-phylogenetic-likelihood-function_b200")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """Import the product package (its directory name is not a Python identifier)."""
    name = "plf_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def coracle():
    import oracle
    oracle.build()
    return oracle.COracle()


@pytest.fixture(scope="session")
def ref_cases():
    return np.load(os.path.join(GOLDEN, "ref_cases.npz"))


def bits(a):
    """Bit pattern view for exact comparisons: distinguishes -0.0 from +0.0 and compares NaNs as
    equal to each other.  NaN payload/sign is the one thing IEEE-754 leaves to the implementation
    (x86 SSE produces 0xFFC00000 for invalid operations and propagates input payloads; NVIDIA GPUs
    produce the canonical 0x7FFFFFFF), so every NaN is mapped to one pattern before comparing."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def first_mismatch(got, want):
    """Human-readable description of the first differing element (for assertion messages)."""
    g, w = bits(got).reshape(-1), bits(want).reshape(-1)
    idx = np.nonzero(g != w)[0]
    if idx.size == 0:
        return "identical"
    i = int(idx[0])
    return (f"{idx.size} of {g.size} elements differ; first at flat index {i} (site {i // 16}, elem {i % 16}): "
            f"got {np.asarray(got).reshape(-1)[i]!r} (0x{g[i]:08x}) want {np.asarray(want).reshape(-1)[i]!r} (0x{w[i]:08x})")
