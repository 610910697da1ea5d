"""CPU tests of the host-side input generator (libcoolmic-dsp_b200/synth.py) that mirrors the device
tone / noise source, and of the communicator's argument handling (no GPU: nothing may succeed)."""
import ctypes as C

import numpy as np
import pytest


def test_tone_rows_follow_the_survey_formula(cm):
    from libcoolmic_dsp_b200 import synth
    period = synth.load_period(48000)
    assert period.size == 48
    first_stream, n, channels, frames, first_frame = 12345, 50, 3, 101, 987654321
    got = synth.tone_rows(period, first_stream, n, channels, frames, first_frame, 7, 5)
    for s in (0, 1, 17, 49):
        for f in (0, 1, 47, 48, 100):
            for c in range(channels):
                want = period[(first_frame + f + 7 * (first_stream + s) + 5 * c) % 48]
                assert got[s, f * channels + c] == want


def test_noise_rows_are_splitmix64(cm):
    from libcoolmic_dsp_b200 import synth

    def sm(x):
        m = (1 << 64) - 1
        x = (x + 0x9E3779B97F4A7C15) & m
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
        return x ^ (x >> 31)

    got = synth.noise_rows(70000, 3, 2, 9, first_frame=1 << 33)
    for s in range(3):
        for f in range(9):
            for c in range(2):
                v = sm(synth.NOISE_SEED ^ ((70000 + s) << 40) ^ (((1 << 33) + f) << 4) ^ c) & 0xffff
                assert got[s, f * 2 + c] == np.uint16(v).astype(np.int16)
    # every 16th stream of the bench data set is noise, the rest tone
    period = synth.load_period(16000)
    rows = synth.synth_rows(period, 0, 40, 1, 64, 0, 5, 0, 16, 5)
    assert np.array_equal(rows[5], synth.noise_rows(5, 1, 1, 64)[0])
    assert np.array_equal(rows[6], synth.tone_rows(period, 6, 1, 1, 64, 0, 5, 0)[0])


def test_sine_periods_are_the_drivers(cm, ref):
    """The committed periods equal what the reference's snddev_sine driver emits here."""
    from libcoolmic_dsp_b200 import synth
    for rate in (8000, 16000, 44100, 48000, 96000):
        period = synth.load_period(rate)
        pcm = np.frombuffer(ref.sine(rate, period.size * 2 * 2), dtype=np.int16)
        assert np.array_equal(pcm, np.tile(period, 2))


def test_comm_argument_validation(cm):
    lib = cm.lib()
    assert lib.cmgpu_comm_nccl_version() >= 22000          # NCCL is linked into the C library itself
    assert not lib.cmgpu_comm_create(0, 0, 0, (C.c_ubyte * 128)())
    assert not lib.cmgpu_comm_create(0, 2, 2, (C.c_ubyte * 128)())
    assert not lib.cmgpu_comm_create(0, 0, 1, None)
    assert not lib.cmgpu_comm_create_file(0, 0, 1, None, 10)
    assert not lib.cmgpu_comm_adopt(None, 0)
    assert lib.cmgpu_comm_rank(None) == -1 and lib.cmgpu_comm_size(None) == 0
    assert lib.cmgpu_gather_results(None, None, 0, 48000, 1, None, None, None, None) == -9
    assert lib.cmgpu_comm_barrier(None) == -9
    assert lib.cmgpu_meter_results(None, 0, 1, 48000, 1, 0, None, None, None) == -9
    assert lib.cmgpu_meter_colors(None, 0, 1, 1.0, 1.0, 1.0, None) == -9
    assert lib.cmgpu_tone_fill(None, 0, 0, 0, 7, 3) == -9
    if lib.cmgpu_device_count() == 0:
        # a rank other than 0 waits for the id file and gives up: no GPU work is attempted
        assert not lib.cmgpu_comm_create_file(0, 1, 2, b"/tmp/cmgpu_no_such_id_file", 50)
        assert b"timed out" in lib.cmgpu_last_error()


def test_library_links_nccl(cm):
    import subprocess
    from libcoolmic_dsp_b200 import binding
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(binding.LIB_PATH)], capture_output=True, text=True).stdout
    for sym in ("ncclCommInitRank", "ncclSend", "ncclRecv", "ncclGroupStart", "ncclAllReduce"):
        assert sym in nm, sym
