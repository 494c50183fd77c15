import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The CUDA library must already be built in-tree (python __graft_entry__.py); build it if stale."""
    from prb_project_bearing_only_slam_b200 import build
    build.build()
    from prb_project_bearing_only_slam_b200 import capi
    return capi.lib()
