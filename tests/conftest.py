import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs on the box")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# The library's built-in values of the partition_kernel knobs (Tuning in csrc/lsbsort.cu), plus what
# LSB_TEST_TUNE="key=value,..." asks for: a candidate default is run through the WHOLE suite this way before it
# becomes the built-in value (tools/run_final_1gpu.sh).  Test infrastructure only: nothing in the library reads the
# environment.
LIB_DEFAULT_TUNE = {"pt_direct": 1, "pt_chunks": 2, "pt_variant": 1, "pt_pf_tiles": 0}


def session_tune():
    kv = dict(LIB_DEFAULT_TUNE)
    for item in filter(None, os.environ.get("LSB_TEST_TUNE", "").split(",")):
        k, v = item.split("=")
        kv[k.strip()] = int(v)
    return kv


def apply_session_tune():
    import distributed_lsb_b200 as lsb
    for k, v in session_tune().items():
        lsb.tune(k, v)


def pytest_sessionstart(session):
    if os.environ.get("LSB_TEST_TUNE"):
        apply_session_tune()


@pytest.fixture
def tune():
    """lsb.tune for one test; the session's values are back in place afterwards"""
    import distributed_lsb_b200 as lsb
    try:
        yield lsb.tune
    finally:
        apply_session_tune()
