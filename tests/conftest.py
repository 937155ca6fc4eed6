import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_terminal_summary(terminalreporter):
    """The raw element-wise maxima behind the tolerance helpers of tests/common.py (no floors, no scales)."""
    from tests import common
    if not common.RAW:
        return
    terminalreporter.write_sep("-", "raw element-wise maxima seen by the parity helpers (before the documented floors)")
    for kind, (mabs, mrel, at, n) in sorted(common.RAW.items()):
        used = common.USED.get(kind)
        terminalreporter.write_line("%-12s max abs err %.3e   max element-wise rel err %.3e (at |ref| = %.3g)   [%d comparisons]%s" % (
            kind, mabs, mrel, at, n, "" if used is None else "   worst asserted error = %.2f of its tolerance" % used))


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle as po
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):  # the checker is plain C: build it on the spot (gcc only)
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return po.load("orc")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code compiled from /root/reference (prebuilt oracle/_ref travels to the GPU box)."""
    from oracle import pyoracle as po
    if not po.have_ref():
        pytest.skip("oracle/_ref/libvbref.so not built (needs /root/reference: make -C oracle ref)")
    try:
        return po.load("ref")
    except OSError as e:  # pragma: no cover
        pytest.skip("libvbref.so does not load here: %s" % e)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return dict(np.load(os.path.join(d, "htk_golden.npz"))), dict(np.load(os.path.join(d, "ref_golden.npz")))
