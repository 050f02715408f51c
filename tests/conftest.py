import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    from oracle import refrun
    skip_ref = pytest.mark.skip(reason="reference tree not present on this box")
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    have_gpu = None
    for item in items:
        if "reference" in item.keywords and not refrun.available():
            item.add_marker(skip_ref)
        if "gpu" in item.keywords:
            if have_gpu is None:
                have_gpu = _have_gpu()
            if not have_gpu:
                item.add_marker(skip_gpu)


@pytest.fixture(scope="session")
def built_library():
    from svim_asm_b200 import build
    return build.build_library()


@pytest.fixture(scope="session")
def oracle_clib():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    from oracle import port
    return port.clib()


@pytest.fixture(scope="session")
def engine(built_library):
    from svim_asm_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
