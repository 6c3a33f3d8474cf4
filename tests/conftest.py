import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cuda():
    import torch

    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    from llmq_b200 import lib

    lib.require_device()  # fails loudly if libb200q.so is missing or the GPU is not sm_100
    return torch.device("cuda", 0)
