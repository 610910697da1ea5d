"""The drop-in (integration) build on CPU: the reference's own pull chain -- its unmodified iohandle.c, tee.c,
snddev.c, snddev_sine.c -- with src/transform.c and src/vumeter.c replaced by the product's host shim compiled as
libigloo objects (-DCOOLMIC_B200_WITH_IGLOO, against the reference's own headers), as oracle/Makefile's `dropin`
target builds it, but linked against the CPU stand-in of the engine (tests/stub/cmgpu_stub.c) instead of the CUDA
library (tests/shimlib.py: build_dropin_stub). The test bodies are those of tests/test_gpu_dropin.py, imported: on the
GPU they prove the drop-in end to end, here they prove its host half -- that the reference's tee and iohandle drive the
product's read / eof / free callbacks and reference counts exactly as they drive the reference's own.
Needs the reference's sources (/root/reference); skipped where they are absent.
"""
import pytest

from tests.test_gpu_dropin import (  # noqa: F401  (collected here as CPU tests; the module's `gpu` mark stays behind)
    test_dropin_pipeline_on_appendix_b,
    test_dropin_pipeline_on_fuzz_goldens,
    test_dropin_result_struct_is_the_references,
    test_dropin_sine_driver_chain,
    test_dropin_vs_all_reference_build_random,
)


@pytest.fixture(scope="module")
def dropin():
    from oracle import pyoracle
    from tests.shimlib import build_dropin_stub
    so = build_dropin_stub()
    if so is None:
        pytest.skip("the reference's sources are not here: the stub drop-in cannot be built")
    d = pyoracle.RefLib(so)
    d.kind = "dropin on the CPU stub engine"
    return d
