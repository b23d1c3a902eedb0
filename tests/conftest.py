import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Collection order = SURVEY.md §8 order: the headline path (rows a1-a19: particle, guiding centre, adaptive, user
# fields, size-independent properties, multi-GPU identity) first, then the (f) "next" rows (getters, quadrature, grid,
# bounce centre).  With `pytest -x` a marginal assert in a late row can then no longer hide the headline parity files.
_ORDER = ["test_gpu_particle", "test_gpu_gc", "test_gpu_adaptive", "test_gpu_userfield", "test_gpu_properties",
          "test_gpu_dist", "test_gpu_getters", "test_gpu_quad", "test_gpu_grid", "test_gpu_bc"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        mod = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _ORDER.index(mod) if mod in _ORDER else -1          # CPU files keep their place in front
    items.sort(key=key)                                            # stable: order inside a file is kept


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
