import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"

# the peer-exchange kernels wait for remote pushes with a watchdog (default: as patient as a
# collective library); in the test process a protocol bug must trap within seconds, never hang.
# The library reads the variable once, at its first exchange call.
os.environ.setdefault("IRR_EXCHANGE_TIMEOUT_MS", "5000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """libirr_b200.so must exist for every test (symbol checks on CPU, compute on GPU)."""
    from imageretrievalresearch_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def golden_losses():
    import numpy as np
    return dict(np.load(GOLDEN / "losses.npz"))


@pytest.fixture(scope="session")
def golden_retrieval():
    import numpy as np
    return dict(np.load(GOLDEN / "retrieval.npz"))


@pytest.fixture(scope="session")
def golden_autocast():
    import numpy as np
    return dict(np.load(GOLDEN / "autocast.npz"))


@pytest.fixture(scope="session")
def golden_pc():
    import numpy as np
    return dict(np.load(GOLDEN / "producer_consumer.npz"))
