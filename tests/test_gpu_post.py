"""GPU tests of what sits either side of the fused tick (csrc/cmgpu_post.cu, csrc/cmgpu_comm.cu):
results for many streams in one round trip (and the optional on-device dB finaliser), meter colours
(reference src/util.c:59-139), the on-device tone / noise source (reference src/snddev_sine.c:118-150
read as a cyclic table), the NCCL result gather, and the ordering fixes of round 2.

Needs a B200: run with `pytest -m gpu`.
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.test_gpu_parity import make_gains, make_pcm, oracle_batch, same_result

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _run_ticks(cm, port, eng, rng, n_streams, channels, block, n_ticks=2, kind="full", silent=()):
    """n_ticks uploads + ticks on slot 0; returns the oracle's meters over all of them."""
    scale, gain = make_gains(rng, n_streams, channels)
    eng.set_gain_table(scale, gain)
    meters = None
    for _ in range(n_ticks):
        host = eng.host_slot(0)
        host[:] = make_pcm(rng, kind, host.shape)
        for s in silent:
            host[s, :] = 0
        src = host.copy()
        ref = src.copy()
        meters, _ = port.batch(ref, np.full(n_streams, block, np.uint32), channels, scale, gain, meters=meters)
        eng.submit(0)
        eng.process(0)
        eng.sync()
    return meters


@pytest.mark.parametrize("channels", [1, 2, 6, 8])
def test_meter_results_batch_equals_per_stream_result(cm, port, channels):
    rng = np.random.default_rng(100 + channels)
    n_streams, block = 97, 700
    with cm.Engine(channels, n_streams, block) as eng:
        meters = _run_ticks(cm, port, eng, rng, n_streams, channels, block, n_ticks=3, silent=(4,))
        # nothing reset yet: peek
        res0, st0, rcs0 = eng.results(44100, reset=False)
        res, st, rcs = eng.results(44100, reset=True)
        for s in range(n_streams):
            want = port.finalise(meters[s], 44100, channels)
            for r, rc in ((res0[s], rcs0[s]), (res[s], rcs[s])):
                assert rc == 0
                assert same_result(r.as_dict(), want), f"stream {s}"
            assert int(st[s].frames) == 3 * block
        # a silent stream still has frames -> a result with -inf powers, like the reference (Appendix B row 14)
        assert math.isinf(res[4].global_power) and res[4].global_power < 0
        # everything was reset in the same step: the reference's second result() -> INVAL (Appendix B row 18)
        res2, st2, rcs2 = eng.results(44100)
        assert all(rc == -10 for rc in rcs2) and all(int(s.frames) == 0 for s in st2)
        assert eng.result(0, 44100) == {"rc": -10}
        # cmgpu_meter_result is the count == 1 case: result + reset as one step
        meters = _run_ticks(cm, port, eng, rng, n_streams, channels, block, n_ticks=1)
        assert same_result(eng.result(7, 48000), port.finalise(meters[7], 48000, channels))
        assert eng.result(7, 48000) == {"rc": -10}
        assert same_result(eng.result(8, 48000), port.finalise(meters[8], 48000, channels))


def test_results_device_db_within_tolerance(cm, port):
    """The optional on-device finaliser (fp64 sqrt/log10 in the take kernel): north_star allows 1e-6
    relative; it lands within 1e-12 of the host's (reference libm) doubles, -inf and 0.0 included."""
    rng = np.random.default_rng(7)
    n_streams, channels, block = 513, 2, 960
    with cm.Engine(channels, n_streams, block) as eng:
        _run_ticks(cm, port, eng, rng, n_streams, channels, block, n_ticks=2, kind="gauss", silent=(0, 9))
        host, _, _ = eng.results(48000, reset=False)
        dev, _, rcs = eng.results(48000, reset=False, flags=cm.RESULTS_DEVICE_DB)
        worst = 0.0
        for s in range(n_streams):
            assert rcs[s] == 0
            a, b = host[s].as_dict(), dev[s].as_dict()
            for k in ("rate", "channels", "frames", "global_peak", "channel_peak"):
                assert a[k] == b[k]
            for x, y in zip([a["global_power"]] + a["channel_power"], [b["global_power"]] + b["channel_power"]):
                if math.isinf(x):
                    assert x == y
                else:
                    worst = max(worst, abs(x - y) / max(abs(x), 1e-300))
        assert worst <= 1e-12, worst          # tolerance stated by north_star: 1e-6 relative


def test_meter_colors_match_host_util(cm, port):
    """Device colours vs the host's coolmic_util_* (bit-exact with reference util.c, tests/test_util.py):
    hues within 1e-13 relative (the device's dB and sin() each differ from libm in the last places), ARGB
    words equal."""
    lib = cm.lib()
    lib.coolmic_util_power2hue.restype = C.c_double
    lib.coolmic_util_power2hue.argtypes = [C.c_double, C.c_char_p]
    lib.coolmic_util_peak2hue.restype = C.c_double
    lib.coolmic_util_peak2hue.argtypes = [C.c_int16, C.c_char_p]
    lib.coolmic_util_ahsv2argb.restype = C.c_uint32
    lib.coolmic_util_ahsv2argb.argtypes = [C.c_double] * 4
    rng = np.random.default_rng(11)
    n_streams, channels, block = 300, 2, 480
    with cm.Engine(channels, n_streams, block) as eng:
        scale = np.full(n_streams, 1000, np.uint16)
        # a wide spread of levels: -60 dB .. clipping
        gain = np.stack([np.linspace(1, 4000, n_streams), np.linspace(4000, 1, n_streams)], axis=1).astype(np.uint16)
        eng.set_gain_table(scale, gain)
        host = eng.host_slot(0)
        host[:] = make_pcm(rng, "gauss", host.shape)
        host[3, :] = 0
        eng.submit(0)
        eng.process(0)
        cols = eng.colors(alpha=1.0, saturation=1.0, value=1.0)
        res, _, _ = eng.results(48000, reset=False)
        flips = 0
        for s in range(n_streams):
            r = res[s]
            pairs = [(r.global_power, cols[s].global_power_hue, cols[s].global_power_argb)]
            pairs += [(r.channel_power[c], cols[s].channel_power_hue[c], cols[s].channel_power_argb[c]) for c in range(channels)]
            for power, hue, argb in pairs:
                want_hue = lib.coolmic_util_power2hue(power, b"default")
                assert abs(hue - want_hue) <= 1e-13 * abs(want_hue) + 1e-300, (s, power, hue, want_hue)
                flips += int(argb != lib.coolmic_util_ahsv2argb(1.0, want_hue, 1.0, 1.0))
            peaks = [(r.global_peak, cols[s].global_peak_argb)] + [(r.channel_peak[c], cols[s].channel_peak_argb[c]) for c in range(channels)]
            for peak, argb in peaks:
                assert argb == lib.coolmic_util_ahsv2argb(1.0, lib.coolmic_util_peak2hue(peak, b"default"), 1.0, 1.0)
        assert flips == 0, f"{flips} colour words differ from the host's"
        # no frames metered -> all-zero colours
        eng.reset_meters()
        z = eng.colors(0, 4)
        assert all(z[i].global_power_argb == 0 and z[i].global_peak_argb == 0 for i in range(4))


@pytest.mark.parametrize("channels,block,rate,sstep,cstep", [(2, 4800, 48000, 7, 3), (1, 320, 16000, 5, 0),
                                                            (8, 1001, 48000, 7, 5), (6, 777, 44100, 7, 3), (3, 5, 8000, 1, 1)])
def test_device_generator_equals_host_formula(cm, channels, block, rate, sstep, cstep):
    from libcoolmic_dsp_b200 import synth
    period = synth.load_period(rate)
    n_streams, first_stream, first_frame = 67, 1000003, 123456789
    with cm.Engine(channels, n_streams, block, ring_slots=2) as eng:
        frames = np.array([(block * (s % 5)) // 4 if s % 7 == 3 else block for s in range(n_streams)], np.uint32)
        eng.set_frames(1, frames)
        synth.device_fill(eng, 1, period, first_stream, first_frame, sstep, cstep, noise_every=4, noise_phase=1)
        eng.fetch(1)                 # identity context, in place: the download is the input ring
        eng.sync()
        got = eng.host_slot(1)
        want = synth.synth_rows(period, first_stream, n_streams, channels, block, first_frame, sstep, cstep, 4, 1)
        for s in range(n_streams):
            n = int(frames[s]) * channels
            assert np.array_equal(got[s, :n], want[s, :n]), f"stream {s}"
            assert not got[s, n:].any(), f"stream {s}: samples past the valid frames must be zero"


def test_sine_period_fixture_is_the_drivers_output(cm, ref):
    """tests/golden/sine.json's periods are what the reference's snddev_sine driver emits (read through
    the real driver here, as SURVEY.md section 0 item 3 asks) -- the bench inputs are built from them."""
    from libcoolmic_dsp_b200 import synth
    for rate in (8000, 16000, 44100, 48000, 96000):
        period = synth.load_period(rate)
        pcm = np.frombuffer(ref.sine(rate, period.size * 2 * 3), dtype=np.int16)
        assert np.array_equal(pcm, np.tile(period, 3)), rate


def test_fetch_planar_straight_after_a_tick_on_resident_data(cm, port):
    """submit, process(TRANSFORM), process(METER|PLANAR), fetch_planar with NO sync in
    between -- the plane download must wait for the tick that writes the planes."""
    rng = np.random.default_rng(5)
    channels, n_streams, block = 2, 512, 24000
    with cm.Engine(channels, n_streams, block, flags=cm.PLANAR_F32) as eng:
        scale, gain = make_gains(rng, n_streams, channels)
        eng.set_gain_table(scale, gain)
        host = eng.host_slot(0)
        host[:] = make_pcm(rng, "full", host.shape)
        want, _ = oracle_batch(port, host.copy(), np.full(n_streams, block, np.uint32), channels, scale, gain)
        for _ in range(3):
            eng.submit(0)
            eng.process(0, cm.TRANSFORM)
            eng.process(0, cm.METER | cm.PLANAR)
            planes = eng.fetch_planar(0)          # queued right behind the tick
            eng.sync()
            for c in range(channels):
                exp = want[:, c: block * channels: channels].astype(np.float32) / np.float32(32768.0)
                assert np.array_equal(planes[:, c, :block], exp)


def test_set_frames_copies_even_from_pinned_memory(cm, port):
    """cmgpu_slot_set_frames is documented as 'Copied': the caller may reuse a PAGE-LOCKED array at once."""
    rng = np.random.default_rng(9)
    channels, n_streams, block = 2, 4096, 2000
    pinned = cm.PinnedArray((n_streams * 2,))        # int16 view; reinterpret as uint32 counts
    counts = pinned.array.view(np.uint32)
    with cm.Engine(channels, n_streams, block, ring_slots=2, flags=cm.SEPARATE_OUT) as eng:
        scale, gain = make_gains(rng, n_streams, channels)
        eng.set_gain_table(scale, gain)
        want = []
        for slot in range(2):
            host = eng.host_slot(slot)
            host[:] = make_pcm(rng, "full", host.shape)
        fr = [rng.integers(0, block + 1, size=n_streams).astype(np.uint32) for _ in range(2)]
        meters = None
        for slot in range(2):
            ref = eng.host_slot(slot).copy()
            meters, _ = port.batch(ref, fr[slot], channels, scale, gain, meters=meters)
        # a long tick keeps the compute stream busy while the counts of both slots are set from ONE array
        eng.submit(0); eng.submit(1)
        eng.process(0, cm.TRANSFORM); eng.process(0, cm.TRANSFORM); eng.process(0, cm.TRANSFORM)
        counts[:] = fr[0]
        eng.set_frames(0, counts)
        counts[:] = fr[1]
        eng.set_frames(1, counts)
        counts[:] = 0
        eng.process(0)
        eng.process(1)
        snap = eng.snapshot()
        for s in range(n_streams):
            assert int(snap[s].frames) == int(fr[0][s]) + int(fr[1][s])
            for c in range(channels):
                assert int(snap[s].power[c]) == int(meters[s].power[c])
    pinned.free()


def test_tick_base_rebased_by_a_full_reset(cm, port):
    """Position keys keep 46 - pbits bits of tick number; a reset of every stream starts them again from
    zero, and peaks of the next window are ordered correctly (first occurrence across ticks)."""
    channels, n_streams, block = 1, 8, 64
    with cm.Engine(channels, n_streams, block) as eng:
        host = eng.host_slot(0)
        for rounds in range(3):
            for t in range(5):
                host[:] = 0
                host[:, 10] = 1000 if t in (1, 3) else 0      # equal peaks in ticks 1 and 3 -> tick 1 wins (positive)
                if t == 3:
                    host[:, 10] = -1000
                eng.submit(0); eng.process(0); eng.sync()
            res, st, rcs = eng.results(48000)                  # full range, reset -> rebase
            assert all(int(st[s].channel_peak[0]) == 1000 for s in range(n_streams))


def test_passthrough_slots_are_not_downloaded(cm, port):
    """The reference's default state is no gain at all (transform.c:107-108: the buffer is left alone).
    Such a slot is metered on the device and nothing comes back: the staging slot it was uploaded from
    already holds the result. As soon as a tick writes the slot's PCM the download happens again --
    also when a LATER tick of the same upload is a pass-through."""
    rng = np.random.default_rng(21)
    channels, n_streams, block = 2, 300, 4800
    with cm.Engine(channels, n_streams, block, ring_slots=2) as eng:
        host = eng.host_slot(0)
        host[:] = make_pcm(rng, "full", host.shape)
        src = host.copy()
        scale = np.zeros(n_streams, np.uint16)
        gain = np.zeros((n_streams, channels), np.uint16)
        fr = np.full(n_streams, block, np.uint32)
        eng.submit(0); eng.process(0); eng.fetch(0); eng.sync()
        up, down = eng.transfer_bytes()
        assert up == n_streams * eng.stride and down == 0
        assert np.array_equal(eng.host_slot(0), src)
        want, meters = oracle_batch(port, src, fr, channels, scale, gain)
        snap = eng.snapshot(reset=True)
        assert all(int(snap[s].power[0]) == int(meters[s].power[0]) and int(snap[s].frames) == block for s in range(n_streams))
        # one stream with a gain: the tick writes the slot, the download is back
        assert eng.set_gain(5, 2, 1000, [500, 2000]) == 0
        scale[5], gain[5] = 1000, [500, 2000]
        eng.submit(0); eng.process(0)
        # ... and a pass-through tick AFTER it (gain off again, no new upload) must not hide that
        assert eng.set_gain(5, 0, 0, None) == 0
        eng.process(0); eng.fetch(0); eng.sync()
        up2, down2 = eng.transfer_bytes()
        assert down2 == n_streams * eng.stride and up2 == 2 * up
        want, _ = oracle_batch(port, src, fr, channels, scale, gain)
        assert np.array_equal(eng.host_slot(0), want)
        # a caller's own buffer is always filled
        out = np.empty_like(src)
        eng.host_slot(1)[:] = src
        eng.submit(1); eng.process(1); eng.fetch(1, out); eng.sync()
        assert np.array_equal(out, src) and eng.transfer_bytes()[1] == 2 * down2


def test_claimed_work_items_under_overlapping_launches(cm, port):
    """Dynamic work claims (TickArgs::work): counters that only grow, one of eight per launch, the base
    handed over by the host. 160 ticks in random slot order with a separate output ring (so the chain of
    overlapping launches never closes) while the number of active streams -- and with it the grid, and
    whether a launch claims dynamically or stays static -- keeps changing. Every stream's meter must
    equal the oracle's over exactly the ticks it took part in."""
    rng = np.random.default_rng(77)
    channels, n_streams, block, ring = 2, 2400, 3000, 6
    with cm.Engine(channels, n_streams, block, ring_slots=ring, flags=cm.SEPARATE_OUT) as eng:
        scale, gain = make_gains(rng, n_streams, channels)
        eng.set_gain_table(scale, gain)
        data = make_pcm(rng, "ties", (ring, n_streams, block * channels))
        for slot in range(ring):
            eng.host_slot(slot)[:, : block * channels] = data[slot]
            eng.submit(slot)
        eng.sync()
        order = []
        for i in range(160):
            active = int(rng.choice([7, 60, 300, 1200, n_streams]))
            slot = int(rng.integers(0, ring))
            eng.set_active(active)
            eng.process(slot)
            order.append((slot, active))
        eng.set_active(n_streams)
        snap = eng.snapshot()
        # oracle: ticks in issue order (first-occurrence peaks depend on it)
        meters = None
        for slot, active in order:
            fr = np.zeros(n_streams, np.uint32)
            fr[:active] = block
            work = data[slot].copy()
            meters, _ = port.batch(work, fr, channels, scale, gain, meters=meters, threads=8)
        for s in range(n_streams):
            assert int(snap[s].frames) == int(meters[s].frames), f"stream {s}"
            for c in range(channels):
                assert int(snap[s].power[c]) == int(meters[s].power[c]), f"stream {s} ch {c}"
                assert int(snap[s].channel_peak[c]) == int(meters[s].channel_peak[c]), f"stream {s} ch {c}"
            assert int(snap[s].global_peak) == int(meters[s].global_peak)


def test_adopting_a_communicator_made_elsewhere(cm, port):
    """cmgpu_comm_adopt: an ncclComm_t the application already has (made here with libnccl directly) is
    used for the gather and NOT destroyed by cmgpu_comm_destroy."""
    nccl = C.CDLL("libnccl.so.2")
    uid = (C.c_ubyte * 128)()
    assert nccl.ncclGetUniqueId(uid) == 0

    class Uid(C.Structure):
        _fields_ = [("b", C.c_ubyte * 128)]

    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, Uid, C.c_int]
    u = Uid()
    C.memmove(u.b, uid, 128)
    with cm.Engine(2, 50, 480) as eng:                   # (creates the CUDA context on device 0 first)
        assert nccl.ncclCommInitRank(C.byref(comm), 1, u, 0) == 0
        lib = cm.lib()
        mine = lib.cmgpu_comm_adopt(comm, 0)
        assert mine and lib.cmgpu_comm_size(mine) == 1 and lib.cmgpu_comm_rank(mine) == 0
        rng = np.random.default_rng(3)
        meters = _run_ticks(cm, port, eng, rng, 50, 2, 480, n_ticks=2)
        res = (cm.Result * 50)()
        rcs = (C.c_int * 50)()
        counts = (C.c_uint * 1)()
        assert lib.cmgpu_gather_results(eng.ctx, mine, 0, 48000, 1, res, None, rcs, counts) == 0
        assert counts[0] == 50 and all(rc == 0 for rc in rcs)
        for s in range(50):
            assert same_result(res[s].as_dict(), port.finalise(meters[s], 48000, 2))
        lib.cmgpu_comm_destroy(mine)
        # still usable: the adopted communicator was not destroyed
        n = C.c_int(0)
        nccl.ncclCommCount.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        assert nccl.ncclCommCount(comm, C.byref(n)) == 0 and n.value == 1
        nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        assert nccl.ncclCommDestroy(comm) == 0


def _gather_ranks(tmp_path, nranks, total_streams, channels):
    path = tmp_path / "nccl.id"
    procs = []
    for r in range(nranks):
        env = dict(os.environ, PYTHONPATH=str(ROOT))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "gather_worker.py"), str(r), str(nranks),
                                       str(path), str(total_streams), str(channels), str(tmp_path / f"out{r}.json")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r]}"
    return json.loads((tmp_path / "out0.json").read_text())


@pytest.mark.parametrize("channels", [2, 5])
def test_nccl_gather_in_c(cm, tmp_path, channels):
    """cmgpu_gather_results over every GPU of the box (one process per GPU; a 1-GPU box runs the same
    code with a communicator of one rank): what arrives at rank 0 equals the oracle for all streams of
    all ranks, uneven shards and active-stream counts included."""
    n = min(cm.lib().cmgpu_device_count(), 8)
    out = _gather_ranks(tmp_path, n, 1000 + 3 * n + 1, channels)
    assert out["ok"], out
    assert out["ranks"] == n and out["streams_checked"] == out["total_streams"]


# ---- the completion word of a tick launch (TickArgs::done_flag) -------------------------------------

def _cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("no libcudart to read device memory behind the library's back")


@pytest.mark.parametrize("channels,n_streams,block,separate", [(1, 512, 320, False), (2, 64, 4800, True), (6, 8, 1000, False),
                                                               (8, 300, 64, False)])
def test_sync_by_completion_word_sees_the_finished_tick(cm, port, channels, n_streams, block, separate):
    """cmgpu_sync straight after cmgpu_process returns when the launch's last CTA has written the completion
    word to host memory -- no driver call. Everything the tick wrote must be visible by then: the output
    ring is read with a plain cudaMemcpy on another stream (nothing orders it after the tick except that
    the host saw the word), tick after tick with a different gain each time."""
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rng = np.random.default_rng(channels * 1000 + block)
    flags = cm.SEPARATE_OUT if separate else 0
    with cm.Engine(channels, n_streams, block, ring_slots=1, flags=flags) as eng:
        host = eng.host_slot(0)
        got = np.empty_like(host)
        meters = None
        for it in range(40):
            pcm = make_pcm(rng, "full", host.shape)
            scale, gain = make_gains(rng, n_streams, channels, "plain")
            scale[scale == 0] = 7
            eng.set_gain_table(scale, gain)
            host[:] = pcm
            eng.submit(0)
            eng.process(0)                      # consumes the upload: a streaming tick (it records its event)
            eng.sync()                          # (a launch gets a completion word only on a stream that was just waited for)
            fr = np.full(n_streams, block, np.uint32)
            want = pcm.copy()
            meters, _ = port.batch(want, fr, channels, scale, gain, meters=meters)
            again = pcm.copy() if separate else want        # in place the second tick works on the first one's output
            meters, _ = port.batch(again, fr, channels, scale, gain, meters=meters)
            want = again
            eng.process(0)                      # on resident data: a bare launch, the tail of the compute stream
            eng.sync()                          # <- by completion word
            assert rt.cudaMemcpy(got.ctypes.data, eng.device_out_slot(0), got.nbytes, 2) == 0
            n = block * channels
            assert np.array_equal(got[:, :n], want[:, :n]), f"tick {it}: output not complete when cmgpu_sync returned"
        # (one sync per iteration can end by the word; a wait that outlasts the library's 60 us of polling -- a busy
        #  box -- falls back to the driver and is not counted: leave some slack)
        assert eng.word_waits() >= 30, "cmgpu_sync never took the completion-word path"
        snap = eng.snapshot(0, n_streams)
        for s in range(n_streams):
            assert int(snap[s].frames) == int(meters[s].frames)
            for c in range(channels):
                assert int(snap[s].power[c]) == int(meters[s].power[c])
                assert int(snap[s].channel_peak[c]) == int(meters[s].channel_peak[c])


def test_completion_word_can_be_switched_off(cm, port, monkeypatch):
    """CMGPU_NO_DONE_WORD=1 (read at cmgpu_ctx_create): every wait asks the driver, results are the same."""
    monkeypatch.setenv("CMGPU_NO_DONE_WORD", "1")
    rng = np.random.default_rng(5)
    with cm.Engine(2, 33, 777) as eng:
        meters = _run_ticks(cm, port, eng, rng, 33, 2, 777, n_ticks=3)
        eng.process(0, cm.TRANSFORM)            # a bare launch on resident data, nothing metered
        eng.sync()
        assert eng.word_waits() == 0
        snap = eng.snapshot(0, 33)
        assert all(int(snap[s].frames) == int(meters[s].frames) for s in range(33))
        assert all(int(snap[s].power[0]) == int(meters[s].power[0]) for s in range(33))
