"""libcoolmic-dsp_b200: B200-native transform + vumeter hot path of libcoolmic-dsp.

The product is the C-ABI shared library built from csrc/ (include/cmgpu.h and the
coolmic_transform_* / coolmic_vumeter_* host shim). This Python package is only the ctypes
binding used by tests/, bench.py and __graft_entry__.py; it contains no arithmetic.

The directory name carries a hyphen (it mirrors the reference's name), so import it through
`load_package()` in /__graft_entry__.py or tests/conftest.py, which registers it as
`libcoolmic_dsp_b200`.
"""
from .binding import (  # noqa: F401
    LIB_PATH, Engine, CmgpuError, MeterState, Result, lib, build_library,
    PinnedArray, state_dict, Comm, Colors, RESULTS_DEVICE_DB, link_probe, FUSED, TRANSFORM, METER, SEPARATE_OUT, NO_PINNED, FORCE_GENERIC, PLANAR_F32, PLANAR, MIX_OUTPUT_METER_ONLY,
)
from . import sharding  # noqa: F401,E402
