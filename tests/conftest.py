import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the oracle are built in-tree before any test runs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_kmc_build", ROOT / "kissmcmc.jl_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()
    from oracle import oracle
    oracle.lib()


@pytest.fixture(scope="session")
def km():
    import kissmcmc_b200
    return kissmcmc_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle
