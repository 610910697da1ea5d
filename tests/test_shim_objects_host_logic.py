"""The stand-alone reference-named objects (coolmic_transform_* / coolmic_vumeter_* / coolmic_iohandle_*) on CPU:
the SAME test bodies as tests/test_gpu_shim.py -- Appendix-B vectors, the seeded pipeline fixtures, random chunking
against the reference's own object code (oracle/_ref), the sine golden -- with the product's host shim
(csrc/host/*.c) linked against the CPU stand-in of the engine (tests/stub/cmgpu_stub.c) instead of the CUDA library.
What this covers without a GPU is the host logic of the objects: whole-frame reads and the partial-frame carry
(transform.c:126-165), the gain setter's adaptation (transform.c:195-222), the meter's physical read and its
1,024-byte staging (vumeter.c:112-187), result / reset semantics (vumeter.c:189-218), reference counting.
The functions are imported, not copied: the module-level `gpu` mark of test_gpu_shim.py stays behind, the
`shim` fixture they ask for resolves to the stub build here.
"""
import pytest

from tests.test_gpu_shim import (  # noqa: F401  (collected here as CPU tests)
    test_shim_objects_on_appendix_b,
    test_shim_second_result_is_inval,
    test_shim_sine_golden_16k,
    test_shim_transform_on_fuzz_goldens,
    test_shim_vs_reference_random_chunking,
)


@pytest.fixture(scope="module")
def shim():
    from tests.shimlib import ShimLib
    return ShimLib(stub=True)


@pytest.fixture(scope="module")
def checker(port):
    from oracle import pyoracle
    return pyoracle.ref() or port


def test_null_and_argument_checks_of_the_objects(shim):
    """NULL self / buffer -> COOLMIC_ERROR_FAULT, constructors refusing rate 0 / channels 0 / channels > 16, read of
    length 0, a handle without a read callback (iohandle.c:58-60,79-84; transform.c:69-75; vumeter.c:73-78)."""
    assert shim.lib.shimh_null_checks() == 0
