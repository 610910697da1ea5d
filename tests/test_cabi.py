"""CPU-side checks of the boundary: the shared library loads, exports every symbol the two public
headers declare, refuses to work without a GPU instead of falling back, and the object shim's
argument handling matches the reference's error behaviour. No compute calls."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
INCLUDE = ROOT / "include"


def declared(header: str):
    text = (INCLUDE / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cmgpu|coolmic)_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["cmgpu.h", "coolmic_b200_shim.h"])
def test_every_declared_symbol_is_exported(cm, header):
    lib = cm.lib()
    names = declared(header)
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header} declares symbols the library does not export: {missing}"


def test_binding_covers_cmgpu_header(cm):
    from libcoolmic_dsp_b200 import binding
    assert set(declared("cmgpu.h")) == set(binding.SYMBOLS)


def test_no_cpu_fallback_without_device(cm):
    lib = cm.lib()
    if lib.cmgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cm.CmgpuError) as e:
        cm.Engine(2, 4, 16)
    assert "no CPU fallback" in str(e.value)


def test_ctx_argument_validation(cm):
    lib = cm.lib()
    for args in [(0, 0, 4, 1, 16, 0), (0, 17, 4, 1, 16, 0), (0, 2, 0, 1, 16, 0), (0, 2, 4, 0, 16, 0), (0, 2, 4, 1, 0, 0)]:
        assert not lib.cmgpu_ctx_create(*args)
    assert lib.cmgpu_process(None, 0, 3) == -9
    assert lib.cmgpu_sync(None) == -9
    assert lib.cmgpu_meter_reset(None, 0, 1) == -9


def test_finalise_is_pure_host_arithmetic(cm, port):
    """cmgpu_finalise restates vumeter.c:198-212; compare it with the oracle port on crafted states."""
    from oracle.pyoracle import Meter
    lib = cm.lib()
    cases = [(1, 12, [1073741824 * 12], [-32768]), (2, 3, [99, 107], [7, -7]), (2, 2, [0, 0], [0, 0]),
             (3, 1000, [5, 123456789012, 1], [1, -30000, 1]), (16, 7, list(range(16)), [0] * 16)]
    for ch, frames, power, peaks in cases:
        st = cm.MeterState()
        m = Meter()
        st.frames = m.frames = frames
        for c in range(ch):
            st.power[c] = m.power[c] = power[c]
            st.channel_peak[c] = m.channel_peak[c] = peaks[c]
        st.global_peak = m.global_peak = max(peaks, key=abs)
        res = cm.Result()
        assert lib.cmgpu_finalise(C.byref(st), 44100, ch, C.byref(res)) == 0
        want = port.finalise(m, 44100, ch)
        got = res.as_dict()
        assert got["frames"] == want["frames"] and got["channel_peak"] == want["channel_peak"]
        import numpy as np
        for a, b in zip([got["global_power"]] + got["channel_power"], [want["global_power"]] + want["channel_power"]):
            assert np.float64(a).tobytes() == np.float64(b).tobytes()
    st = cm.MeterState()
    assert lib.cmgpu_finalise(C.byref(st), 48000, 2, C.byref(cm.Result())) == -10     # frames == 0


def test_meter_row_decode(cm):
    """Rows as the kernel writes them -> decoded state, including the interleaved-order global peak."""
    lib = cm.lib()
    ch = 2

    def key(mag, pos, neg):
        return (mag << 47) | (((~pos) & ((1 << 46) - 1)) << 1) | neg

    # channel 0: +7 at frame 1; channel 1: -7 at frame 0 -> global peak is channel 1's (earlier frame)
    rows = (C.c_uint64 * 6)(key(7, 1, 0), key(7, 0, 1), 99, 107, 3, 0)
    st = (cm.MeterState * 1)()
    assert lib.cmgpu_meter_decode(rows, 1, ch, st) == 0
    assert (st[0].channel_peak[0], st[0].channel_peak[1], st[0].global_peak, st[0].frames) == (7, -7, -7, 3)
    # same frame: channel order decides
    rows = (C.c_uint64 * 6)(key(9, 5, 1), key(9, 5, 0), 1, 1, 8, 0)
    assert lib.cmgpu_meter_decode(rows, 1, ch, st) == 0
    assert st[0].global_peak == -9
    # |-32768| beats 32767
    rows = (C.c_uint64 * 6)(key(32767, 0, 0), key(32768, 100, 1), 1, 1, 200, 0)
    assert lib.cmgpu_meter_decode(rows, 1, ch, st) == 0
    assert (st[0].global_peak, st[0].channel_peak[1]) == (-32768, -32768)


def test_shim_harness_and_argument_checks(cm):
    from tests.shimlib import ShimLib
    shim = ShimLib()
    assert shim.lib.shimh_sizeof_result() == 192          # the reference's result struct on LP64
    assert shim.lib.shimh_null_checks() == 0


def test_dropin_build_exports_the_reference_api(cm):
    """oracle/_ref/libcoolmic_dropin.so = the reference's iohandle.c / tee.c / snddev*.c + the product's
    transform / vumeter shim compiled against the reference's own headers (oracle/Makefile). It must
    load, export the reference's names, and take every cmgpu_* call from the product library."""
    import subprocess
    from oracle import pyoracle
    d = pyoracle.dropin()
    if d is None:
        pytest.skip("oracle/_ref/libcoolmic_dropin.so is not available here")
    for name in ["coolmic_transform_new", "coolmic_transform_attach_iohandle", "coolmic_transform_get_iohandle",
                 "coolmic_transform_set_master_gain", "coolmic_vumeter_new", "coolmic_vumeter_reset",
                 "coolmic_vumeter_attach_iohandle", "coolmic_vumeter_read", "coolmic_vumeter_result",
                 "coolmic_iohandle_new", "coolmic_iohandle_read", "coolmic_iohandle_eof",
                 "coolmic_tee_new", "coolmic_tee_attach_iohandle", "coolmic_tee_get_iohandle", "coolmic_snddev_new"]:
        assert hasattr(d.lib, name), name
    assert d.sizeof_result() == 192
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(d.path)], capture_output=True, text=True).stdout
    wanted = sorted(set(re.findall(r"\bU (cmgpu_[a-z_0-9]+)", nm)))
    assert len(wanted) >= 10
    assert all(hasattr(cm.lib(), n) for n in wanted)
    # no GPU here: constructing the objects works (host-side), the first read that needs the device fails
    # loudly instead of computing anything on the CPU
    if cm.lib().cmgpu_device_count() == 0:
        import numpy as np
        out, results, rc = d.pipeline(np.arange(64, dtype=np.int16), 2, (2, 4, [3, 5]), result_every=0)
        assert out.size == 0 and results[-1].get("rc") == -10


@pytest.mark.gpu
def test_out_of_memory_is_reported_not_survived(cm):
    """A ring that cannot fit in HBM: NULL + COOLMIC_ERROR_NOMEM-style message, nothing falls back."""
    lib = cm.lib()
    ctx = lib.cmgpu_ctx_create(0, 16, 1 << 20, 64, 1 << 16, cm.NO_PINNED)      # 16 ch x 1 Mi streams x 64 slots x 2 MiB
    assert not ctx
    assert b"cudaMalloc" in lib.cmgpu_last_error()
    with cm.Engine(2, 4, 16) as eng:                                              # and the library still works afterwards
        eng.submit(0)
        eng.process(0)
        eng.sync()
