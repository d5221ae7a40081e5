import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ofdg():
    """The product package with its native library built."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ofdg_build", os.path.join(ROOT, "optical-flow-2d-data-generation_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    if not os.path.exists(b.LIB) or (b.needs_build() and os.path.exists("/usr/local/cuda/bin/nvcc")):
        b.build()
    import ofdg_b200
    ofdg_b200.lib()
    return ofdg_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def refimpl():
    """oracle/_ref/libofdg_ref.so: the reference's own sources compiled against oracle/shim (oracle/ref_build.sh). Built here
    from /root/reference; on the GPU box the prebuilt library that travelled with the snapshot is used."""
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref is not built and /root/reference is not present to build it from")
    ref_binding.build()
    ref_binding.lib()
    return ref_binding


@pytest.fixture(scope="session")
def textures8(ofdg):
    return ofdg.synth_textures(8, 1024, 768, seed=1)
