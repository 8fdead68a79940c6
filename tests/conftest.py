import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


from tests.scenes import build_scene  # noqa: E402


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/rm_oracle.c), built on demand. Checker only."""
    from oracle import build_oracle, refso
    build_oracle.build(verbose=False)
    return refso.load("oracle")


@pytest.fixture(scope="session")
def ref_strict():
    """The reference's own kernel text (oracle/_ref), when it has been built."""
    from oracle import refso
    if not refso.available("ref_strict"):
        pytest.skip("oracle/_ref/libref_strict.so not built (needs /root/reference at build time)")
    return refso.load("ref_strict")


@pytest.fixture(scope="session")
def gpu_renderer():
    from raymarchcl_b200 import _lib
    from raymarchcl_b200.renderer import Renderer
    try:
        lib = _lib.load()
    except (ImportError, OSError) as e:
        pytest.skip(f"libraymarch_b200.so not loadable: {e}")
    if lib.rm_device_count() == 0:
        pytest.skip("no CUDA device (run the gpu-marked tests on the B200 box)")
    try:
        r = Renderer(0)
    except _lib.RaymarchError as e:
        if e.code == -6:  # RM_ERR_NO_DEVICE: not an sm_100 device
            pytest.skip(str(e))
        raise
    yield r
    r.close()
