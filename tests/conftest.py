"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here on CPU (oracle vs golden vectors, host logic, C-ABI symbols);
`-m gpu` runs on a B200 and compares the CUDA path, through the C-ABI, with the oracle.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    return pyoracle.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    r = pyoracle.ref()
    if r is None:
        pytest.skip("oracle/_ref (the reference's own object code) is not available here")
    return r


def load_package():
    """Import libcoolmic-dsp_b200/ (hyphenated directory) as module `libcoolmic_dsp_b200`."""
    import importlib.util
    name = "libcoolmic_dsp_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = ROOT / "libcoolmic-dsp_b200"
    spec = importlib.util.spec_from_file_location(name, pkg / "__init__.py", submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def cm():
    return load_package()
