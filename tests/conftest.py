import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "c-ray_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILT = os.path.join(ROOT, "scenes", "_built")
GOLDEN_SCENES = ["g_nodes", "g_legacy", "g_single", "g_meshmat"]
# + the fixture whose node graphs no JSON can express (built by the reference's C constructors in oracle/ref_harness.c):
# only its flat export exists, so the loader tests skip it
GOLDEN_FLAT = GOLDEN_SCENES + ["g_f4"]
SUMMARY_LINES = []   # headline parity figures (tests/test_zz_full_config.py): repeated in the terminal summary so the run's tail shows them


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_terminal_summary(terminalreporter):
    for line in SUMMARY_LINES:
        terminalreporter.write_line(line)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the shared libraries exist (cheap no-op when they are already built)."""
    import __graft_entry__ as g
    g.build(quiet=True)
