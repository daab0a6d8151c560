import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rapidfuzz-rs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a hung kernel must not eat the GPU budget: every GPU test gets a hard per-test timeout (pytest-timeout)
    have_gpu = None
    for it in items:
        if not it.get_closest_marker("gpu"):
            continue
        if have_gpu is None:   # probed once, only when GPU tests were collected; no device => skip, not 300 red tests
            try:
                from rapidfuzz_b200 import _ffi
                have_gpu = _ffi.lib().rf_device_count() > 0
            except Exception:
                have_gpu = False
        if not have_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device (GPU parity tests run on the B200 box with -m gpu)"))
        elif not it.get_closest_marker("timeout"):
            it.add_marker(pytest.mark.timeout(180))
