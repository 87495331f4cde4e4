"""pytest configuration: `gpu` marker, import paths, shared helpers."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the reference's own suite is vendored unmodified; it is run in a subprocess by test_ref_suite_gpu.py
collect_ignore = ["ref_suite"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def iso():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import isoext_b200
    return isoext_b200


@pytest.fixture(scope="session")
def ref():
    """The reference's own CUDA build behind oracle/ref_shim.cu (prebuilt; travels with gpurun)."""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libisoext_ref.so not built")
    return r
