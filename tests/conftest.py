import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
# The library picks some kernel forms by the number of streams (fewer than 32: single CTAs, plain leak sweep; from 32 on: CTA
# pairs where a layer allows them, leak sweep with skip bitmaps and the window sweep).  The tests run a handful of streams; unless
# a test says otherwise they take the many-stream sweep the benchmark runs (both forms are compared bit for bit in
# test_sweep_skipping_changes_no_bit; the pair kernels in test_pair_units_change_no_bit).
os.environ.setdefault("AEC_SWEEP_SKIP", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
