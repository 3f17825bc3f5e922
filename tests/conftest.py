import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def load_package():
    """Register the hyphenated product directory `hevc-hop_b200/` as module `hevc_hop_b200`."""
    if "hevc_hop_b200" in sys.modules:
        return sys.modules["hevc_hop_b200"]
    pkg_dir = os.path.join(ROOT, "hevc-hop_b200")
    spec = importlib.util.spec_from_file_location(
        "hevc_hop_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hevc_hop_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libhopref.so (the compiled reference)")


@pytest.fixture(scope="session")
def hop():
    return load_package()


@pytest.fixture(scope="session")
def ctx(hop):
    c = hop.HopContext(0)
    yield c
    c.close()
