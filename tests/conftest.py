import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: multi-minute CPU test (opt in with CC_SLOW=1)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("CC_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="slow; set CC_SLOW=1")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)
