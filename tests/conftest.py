import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def lib_built():
    """Builds libvptb200.so if needed (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "vision_pt_b200", "csrc"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("vpt_build", os.path.join(ROOT, "vision_pt_b200", "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="session")
def golden_blocks():
    """Reference-live vectors of the SDXL / CogView4 blocks and PoPE (tests/golden/make_golden_blocks.py)."""
    path = os.path.join(ROOT, "tests", "golden", "block_family_vectors.pt")
    return torch.load(path, map_location="cpu", weights_only=False)
