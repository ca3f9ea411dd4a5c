import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
DATA = GOLDEN / "reference_data"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN / "fet_kats.json").read_text())


@pytest.fixture(scope="session")
def fixture_bytes():
    return {p.name: p.read_bytes() for p in DATA.iterdir()}


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from feht_b200 import build, capi
    if not (ROOT / "feht_b200" / "libfeht_cuda.so").exists():
        build.build_cuda()
    return capi.lib()


@pytest.fixture()
def ctx(cuda_lib):
    from feht_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
