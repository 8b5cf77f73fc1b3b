import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
TOY_DIR = os.path.join(REF_DIR, "toy")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref (built by oracle/Makefile)")


def have_ref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_dump")) and os.path.exists(os.path.join(TOY_DIR, "index_toy", "cmer-m4r1-frac"))


needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (run make -C oracle ref toy where /root/reference exists)")
