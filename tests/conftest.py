import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def ex1_fa():
    return os.path.join(ROOT, "tests", "golden", "ex1.fa")


@pytest.fixture(scope="session")
def synth_fa(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    p = str(tmp_path_factory.mktemp("synth") / "synth.fa")
    make_golden.synth_fasta(p)
    return p
