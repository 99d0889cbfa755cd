import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", category=FutureWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")


BACKENDS = [pytest.param("hostsim", id="hostsim"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]


@pytest.fixture(params=BACKENDS)
def buffers(request):
    """Buffer provider under the product's Python host layer: the CUDA engine (gpu tests) or the CPU test
    double of the device backend (host-logic tests, tests/hostsim)."""
    if request.param == "cuda":
        import jfnk_b200

        return jfnk_b200.CudaBuffers()
    from tests.hostsim.sim import SimBuffers

    return SimBuffers()


@pytest.fixture
def cuda_buffers():
    import jfnk_b200

    return jfnk_b200.CudaBuffers()
