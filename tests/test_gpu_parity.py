"""Parity tests proper: the CUDA path, called through the C ABI (include/cmgpu.h), against the
oracle on identical inputs. Bar: transformed PCM, peaks, sums of squares and frame counts
bit-exact; dB values bit-exact too because the finaliser runs on the host with the reference's
expression and the same libm (north_star allows 1e-6 relative; we do not need it).

Needs a B200: run with `pytest -m gpu`.
"""
import ctypes as C
import json
import math
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden"
KATS = json.loads((GOLD / "kat_appendix_b.json").read_text())
SINE = json.loads((GOLD / "sine.json").read_text())


def f64_bits(x):
    return np.float64(x).tobytes()


def same_result(a, b):
    if a.get("rc", 0) != b.get("rc", 0):
        return False
    if a.get("rc", 0) != 0:
        return True
    for k in ("rate", "channels", "frames", "global_peak", "channel_peak"):
        if a[k] != b[k]:
            return False
    fa = [a["global_power"]] + list(a["channel_power"])
    fb = [b["global_power"]] + list(b["channel_power"])
    return all(f64_bits(x) == f64_bits(y) for x, y in zip(fa, fb))


def unhex(res):
    r = dict(res)
    if r.get("rc", 0) == 0 and "global_power" in r:
        r["global_power"] = float.fromhex(r["global_power"])
        r["channel_power"] = [float.fromhex(x) for x in r["channel_power"]]
    return r


def make_pcm(rng, kind, shape):
    if kind == "full":
        return rng.integers(-32768, 32768, size=shape).astype(np.int16)
    if kind == "ties":       # few distinct magnitudes, both signs: stresses first-occurrence order
        return rng.choice(np.array([-32768, -32767, -100, -1, 0, 1, 100, 32767], dtype=np.int16), size=shape)
    if kind == "small":
        return rng.integers(-50, 51, size=shape).astype(np.int16)
    if kind == "gauss":
        return rng.normal(0, 8000, size=shape).clip(-32768, 32767).astype(np.int16)
    raise ValueError(kind)


def make_gains(rng, n_streams, channels, kind="mixed"):
    scale = rng.integers(1, 65536, size=n_streams).astype(np.uint16)
    gain = rng.integers(0, 65536, size=(n_streams, channels)).astype(np.uint16)
    if kind == "mixed":
        for s in range(n_streams):
            m = s % 6
            if m == 0:
                scale[s] = 0                                  # disabled
            elif m == 1:
                gain[s, :] = scale[s]                         # unity
            elif m == 2:                                      # mild attenuation / boost around 1.0
                scale[s] = 1000 + s % 9000
                gain[s, :] = (int(scale[s]) * 3 // 4 + 37 * ((s + np.arange(channels)) % 64)).astype(np.uint16)
            elif m == 3:
                gain[s, 0] = 0                                # mute one channel
    return scale, gain


def oracle_batch(port, pcm, frames, channels, scale, gain):
    """pcm: [n][stride] int16 -> (transformed copy, meters)"""
    ref = pcm.copy()
    meters, _ = port.batch(ref, frames, channels, scale, gain)
    return ref, meters


def check_meters(cm, port, eng, meters, n_streams, channels, rate=48000, reset=True):
    snap = eng.snapshot(0, n_streams, reset=reset)
    for s in range(n_streams):
        got, want = snap[s], meters[s]
        assert int(got.frames) == int(want.frames), f"stream {s}: frames"
        for c in range(channels):
            assert int(got.power[c]) == int(want.power[c]), f"stream {s} ch {c}: power"
            assert int(got.channel_peak[c]) == int(want.channel_peak[c]), f"stream {s} ch {c}: peak"
        assert int(got.global_peak) == int(want.global_peak), f"stream {s}: global peak"
        r_got = eng.finalise(got, rate)
        r_want = port.finalise(want, rate, channels)
        assert same_result(r_got, r_want), f"stream {s}: {r_got} != {r_want}"


def run_case(cm, port, channels, n_streams, block_frames, frames, kind, seed, flags=0, gain_kind="mixed",
             process_flags=None):
    rng = np.random.default_rng(seed)
    process_flags = cm.FUSED if process_flags is None else process_flags
    with cm.Engine(channels, n_streams, block_frames, ring_slots=1, flags=flags) as eng:
        host = eng.host_slot(0)
        stride = host.shape[1]
        host[:] = make_pcm(rng, kind, host.shape)
        scale, gain = make_gains(rng, n_streams, channels, gain_kind)
        eng.set_gain_table(scale, gain)
        fr = np.full(n_streams, block_frames, dtype=np.uint32) if frames is None else np.asarray(frames, np.uint32)
        eng.set_frames(0, None if frames is None else fr)
        src = host.copy()
        if not (process_flags & cm.TRANSFORM):
            scale = np.zeros_like(scale)
        want, meters = oracle_batch(port, src, fr, channels, scale, gain)
        eng.submit(0)
        eng.process(0, process_flags)
        eng.fetch(0)
        eng.sync()
        got = eng.host_slot(0)
        if flags & cm.SEPARATE_OUT:
            # only valid frames are ever written to the second ring
            for s in range(n_streams):
                n = int(fr[s]) * channels
                assert np.array_equal(got[s, :n], want[s, :n]), f"stream {s}: " + _first_diff(got[s:s + 1, :n], want[s:s + 1, :n], channels)
        else:
            # valid frames must match the oracle; bytes past them must be untouched input
            assert np.array_equal(got, want), _first_diff(got, want, channels)
        if process_flags & cm.METER:
            check_meters(cm, port, eng, meters, n_streams, channels)
        else:
            snap = eng.snapshot(0, n_streams)
            assert all(int(snap[s].frames) == 0 for s in range(n_streams))
        return eng.kernel_name()


def _first_diff(got, want, channels):
    idx = np.argwhere(got != want)
    if idx.size == 0:
        return "equal"
    s, i = idx[0]
    return f"{len(idx)} samples differ; first at stream {s} sample {i} (frame {i // channels} ch {i % channels}): " \
           f"got {got[s, i]} want {want[s, i]}"


# ---- known answers (regenerated from the reference's object code) ----------

@pytest.mark.parametrize("case", [k for k in KATS if "out" in k], ids=lambda k: k["name"])
def test_appendix_b(cm, case):
    ch = case["channels"]
    nbytes = case["in_bytes"]
    frames = nbytes // (2 * ch)            # transform only ever emits whole frames (transform.c:129-165)
    samples = np.array(case["in"], dtype=np.int16)
    with cm.Engine(ch, 1, max(frames, 1)) as eng:
        host = eng.host_slot(0)
        host[0, : frames * ch] = samples[: frames * ch]
        if case["gain"] is not None:
            n, scale, gains = case["gain"]
            assert eng.set_gain(0, n, scale, gains) == case["gain_rc"]
        eng.set_frames(0, [frames])
        eng.submit(0)
        eng.process(0)
        eng.fetch(0)
        eng.sync()
        assert eng.host_slot(0)[0, : frames * ch].tolist() == case["out"]
        got = eng.result(0, 48000)
        assert same_result(got, unhex(case["results"][0])), (got, case["results"][0])
        # Appendix B row 18: a second result() straight after a successful one
        assert eng.result(0, 48000) == {"rc": -10}


@pytest.mark.parametrize("case", [s for s in SINE if "period" in s],
                         ids=lambda s: f"{s['rate']}Hz-{s['gain']}")
def test_sine_goldens(cm, port, case):
    """Includes BASELINE config 1: the sine driver at 44.1 kHz mono for 60 s, gain off,
    1/2 and 3/4, against what the reference's own snddev_sine -> transform -> tee -> vumeter produced."""
    src = np.resize(np.array(case["period"], dtype=np.int16), case["bytes"] // 2)
    with cm.Engine(1, 1, src.size) as eng:
        eng.host_slot(0)[0, : src.size] = src
        if case["gain"] is not None:
            assert eng.set_gain(0, *case["gain"]) == 0
        eng.submit(0)
        eng.process(0)
        eng.fetch(0)
        eng.sync()
        out = eng.host_slot(0)[0, : src.size]
        assert f"{port.fnv1a64(out):016x}" == case["out_fnv1a64"]
        assert same_result(eng.result(0, case["rate"]), unhex(case["result"]))


# ---- differential tests against the oracle port -----------------------------------------------

@pytest.mark.parametrize("channels", list(range(1, 17)))
@pytest.mark.parametrize("kind", ["full", "ties"])
def test_random_all_channel_counts(cm, port, channels, kind):
    rng = np.random.default_rng(channels * 7 + len(kind))
    n_streams = 37
    block_frames = int(rng.integers(1, 700))
    frames = rng.integers(0, block_frames + 1, size=n_streams)
    frames[0] = block_frames
    frames[1] = 0
    run_case(cm, port, channels, n_streams, block_frames, frames, kind, seed=channels)


@pytest.mark.parametrize("channels,block_frames", [(1, 320), (1, 5), (2, 4800), (2, 1), (4, 333), (8, 1024),
                                                   (8, 3), (16, 512), (16, 7), (2, 70001), (1, 300007), (6, 5000)])
def test_full_blocks(cm, port, channels, block_frames):
    name = run_case(cm, port, channels, 19, block_frames, None, "gauss", seed=block_frames)
    assert "tick" in name


@pytest.mark.parametrize("channels", [1, 2, 4, 8, 16])
def test_generic_kernel_agrees_with_fast(cm, port, channels):
    a = run_case(cm, port, channels, 11, 2049, None, "ties", seed=3)
    b = run_case(cm, port, channels, 11, 2049, None, "ties", seed=3, flags=cm.FORCE_GENERIC)
    assert a.startswith("fused_tick") and b.startswith("generic_tick")


@pytest.mark.parametrize("channels", [1, 2, 5, 8])
def test_separate_out_ring(cm, port, channels):
    run_case(cm, port, channels, 9, 1000, None, "full", seed=11, flags=cm.SEPARATE_OUT)


@pytest.mark.parametrize("channels", [2, 3])
def test_transform_only_and_meter_only(cm, port, channels):
    run_case(cm, port, channels, 9, 777, None, "full", seed=5, process_flags=cm.TRANSFORM)
    run_case(cm, port, channels, 9, 777, None, "full", seed=5, process_flags=cm.METER)
    run_case(cm, port, channels, 9, 777, None, "full", seed=5, process_flags=cm.METER, flags=cm.SEPARATE_OUT)


def test_small_buffer_regime(cm, port):
    # BASELINE config 3 shape: many mono streams, 320-frame (640-byte) stream-blocks
    name = run_case(cm, port, 1, 4096, 320, None, "gauss", seed=1, gain_kind="mixed")
    assert name == "fused_tick<C=1,G=8>"


def test_meter_window_spans_ticks_and_slots(cm, port):
    """Meter state persists from reset to result across ticks (vumeter.c:170,177), an equal
    magnitude later in time never replaces the older peak, and chunking does not matter."""
    rng = np.random.default_rng(42)
    channels, n_streams, block = 2, 23, 256
    ticks = 7
    data = make_pcm(rng, "ties", (ticks, n_streams, block * channels))
    scale, gain = make_gains(rng, n_streams, channels)
    with cm.Engine(channels, n_streams, block, ring_slots=3) as eng:
        eng.set_gain_table(scale, gain)
        meters = None
        outs = []
        for t in range(ticks):
            slot = t % 3
            eng.slot_wait(slot)
            eng.host_slot(slot)[:, : block * channels] = data[t]
            frames = rng.integers(0, block + 1, size=n_streams).astype(np.uint32)
            eng.set_frames(slot, frames)
            eng.submit(slot)
            eng.process(slot)
            eng.fetch(slot)
            ref = data[t].copy()
            meters, _ = port.batch(ref, frames, channels, scale, gain, meters=meters)
            eng.slot_wait(slot)
            assert np.array_equal(eng.host_slot(slot)[:, : block * channels], ref)
            if t == 3:            # take a result mid-way for half of the streams, like simple.c:486-499
                snap = eng.snapshot(0, n_streams)
                for s in range(0, n_streams, 2):
                    got = eng.result(s, 44100)
                    want = port.finalise(meters[s], 44100, channels)     # resets meters[s] on success
                    assert same_result(got, want)
        check_meters(cm, port, eng, meters, n_streams, channels)
        # everything was reset by the snapshot above: no frames -> INVAL
        assert eng.result(0, 48000) == {"rc": -10}


def test_one_long_block_equals_many_short_ticks(cm, port):
    rng = np.random.default_rng(9)
    channels, frames_total = 2, 4096
    pcm = make_pcm(rng, "ties", (1, frames_total * channels))
    scale = np.array([7], np.uint16)
    gain = np.array([[5, 9]], np.uint16)
    results = []
    for block in (frames_total, 512, 1):
        with cm.Engine(channels, 1, block) as eng:
            eng.set_gain_table(scale, gain)
            out = np.empty_like(pcm)
            for t in range(frames_total // block):
                lo, hi = t * block * channels, (t + 1) * block * channels
                eng.host_slot(0)[0, : block * channels] = pcm[0, lo:hi]
                eng.submit(0)
                eng.process(0)
                eng.fetch(0)
                eng.sync()
                out[0, lo:hi] = eng.host_slot(0)[0, : block * channels]
            results.append((out.copy(), eng.result(0, 48000)))
    for out, res in results[1:]:
        assert np.array_equal(out, results[0][0])
        assert same_result(res, results[0][1])
    ref = pcm.copy()
    meters, _ = port.batch(ref, [frames_total], channels, scale, gain)
    assert np.array_equal(results[0][0], ref)
    assert same_result(results[0][1], port.finalise(meters[0], 48000, channels))


def test_gain_setter_semantics(cm):
    # transform.c:195-222 through the C ABI
    with cm.Engine(3, 2, 8) as eng:
        assert eng.set_gain(0, 3, 10, [1, 2, 3]) == 0
        assert eng.get_gain(0) == (10, [1, 2, 3])
        assert eng.set_gain(0, 2, 2, [1, 3]) == -10          # cannot map 2 -> 3; state kept
        assert eng.get_gain(0) == (10, [1, 2, 3])
        assert eng.set_gain(0, 1, 4, [9]) == 0
        assert eng.get_gain(0) == (4, [9, 9, 9])
        assert eng.set_gain(0, 0, 4, [9]) == 0
        assert eng.get_gain(0)[0] == 0
        assert eng.set_gain(1, 3, 0, [1, 1, 1]) == 0
        assert eng.get_gain(1)[0] == 0
        assert eng.set_gain(5, 1, 1, [1]) == -10
    with cm.Engine(1, 1, 8) as eng:
        assert eng.set_gain(0, 2, 2, [1, 3]) == 0
        assert eng.get_gain(0) == (2, [2])


def test_full_size_config2_properties(cm, port):
    """BASELINE config 2 at full size: 1,024 stereo 48 kHz streams x 10 s in one tick. The oracle
    checks a sample of streams sample-for-sample; every stream is checked through properties that
    do not need the oracle: sum of squares and peak magnitude recomputed from the fetched PCM."""
    n_streams, channels, block = 1024, 2, 480000
    s_idx = np.arange(n_streams)
    scale = (1000 + s_idx % 9000).astype(np.uint16)
    gain = (scale[:, None].astype(np.int64) * 3 // 4 + 37 * ((s_idx[:, None] + np.arange(channels)) % 64)).astype(np.uint16)
    period = np.round(32766 * np.sin(2 * np.pi * np.arange(48) / 48)).astype(np.int16)
    with cm.Engine(channels, n_streams, block, flags=cm.NO_PINNED) as eng:
        eng.set_gain_table(scale, gain)
        f = np.arange(block)
        host = np.empty((n_streams, block * channels), dtype=np.int16)
        for s in range(n_streams):
            for c in range(channels):
                host[s, c::channels] = period[(f + 7 * s + 3 * c) % 48]
        # overwrite a few streams with full-range noise to stress the clamp and tie-breaks
        rng = np.random.default_rng(2)
        noisy = [3, 500, 1023]
        for s in noisy:
            host[s] = rng.integers(-32768, 32768, size=block * channels).astype(np.int16)
        eng.submit(0, host)
        eng.process(0)
        out = np.empty_like(host)
        eng.fetch(0, out)
        eng.sync()
        snap = eng.snapshot()
        check = sorted(set(noisy + [0, 1, 63, 64, 511, 777, 1022]))
        sub = host[check].copy()
        meters, _ = port.batch(sub, np.full(len(check), block, np.uint32), channels, scale[check], gain[check])
        assert np.array_equal(out[check], sub)
        for i, s in enumerate(check):
            for c in range(channels):
                assert int(snap[s].power[c]) == int(meters[i].power[c])
                assert int(snap[s].channel_peak[c]) == int(meters[i].channel_peak[c])
            assert int(snap[s].global_peak) == int(meters[i].global_peak)
        for s in range(0, n_streams, 8):
            y = out[s].astype(np.int64)
            for c in range(channels):
                yc = y[c::channels]
                assert int(snap[s].power[c]) == int((yc * yc).sum())
                assert abs(int(snap[s].channel_peak[c])) == int(np.abs(yc).max())
            assert int(snap[s].frames) == block


@pytest.mark.parametrize("channels,block,n_streams", [(2, 60000, 40), (8, 9000, 64), (1, 320, 3000), (2, 4800, 700)])
@pytest.mark.parametrize("no_pdl", [False, True])
def test_overlapping_ticks_on_resident_data(cm, port, monkeypatch, channels, block, n_streams, no_pdl):
    """Back-to-back ticks on data that is already on the device are bare kernel launches, each marked
    as a programmatic dependent launch of the one before when it reads nothing that one writes (other
    ring slots, or a separate output ring): the next tick's CTAs start while the previous tick's last
    work items drain. The meter must not notice: ticks are numbered by the host, keys carry (tick,
    frame), so a tie between ticks still goes to the earlier tick. Checked against the oracle run
    tick after tick, with the overlap on and off (CMGPU_NO_PDL)."""
    if no_pdl:
        monkeypatch.setenv("CMGPU_NO_PDL", "1")
    rng = np.random.default_rng(900 + channels + block)
    ring, rounds = 3, 4
    data = make_pcm(rng, "ties", (ring, n_streams, block * channels))
    scale, gain = make_gains(rng, n_streams, channels)
    with cm.Engine(channels, n_streams, block, ring_slots=ring, flags=cm.SEPARATE_OUT) as eng:
        eng.set_gain_table(scale, gain)
        for slot in range(ring):
            eng.host_slot(slot)[:, : block * channels] = data[slot]
            eng.submit(slot)
        eng.sync()
        order = []
        for r in range(rounds):
            for slot in ([0, 0, 1, 2, 2, 1] if r % 2 else [2, 1, 0]):
                eng.process(slot)               # the same slot twice in a row is fine too: the input ring is read-only
                order.append(slot)
        for slot in range(ring):
            eng.fetch(slot)
        eng.sync()
        assert eng.launch_count() == len(order)
        meters = None
        frames = np.full(n_streams, block, np.uint32)
        outs = {}
        for slot in order:
            work = data[slot].copy()
            meters, _ = port.batch(work, frames, channels, scale, gain, meters=meters)
            outs[slot] = work
        for slot in range(ring):
            assert np.array_equal(eng.host_slot(slot)[:, : block * channels], outs[slot])
        check_meters(cm, port, eng, meters, n_streams, channels)


def test_tick_numbering_across_plain_ticks_and_captured_cycles(cm, port, monkeypatch):
    """Plain ticks are numbered by the host (an offset from the device's tick counter), the ticks of a
    captured cycle count from the device counter itself, which is brought up to date before every
    replay. Mixed freely, later ticks must still lose ties against earlier ones."""
    monkeypatch.setenv("CMGPU_NO_SPAN", "1")          # force the CUDA-graph form of the cycle
    rng = np.random.default_rng(4242)
    channels, n_streams, block, ring = 2, 97, 640, 3
    data = make_pcm(rng, "ties", (ring, n_streams, block * channels))
    scale, gain = make_gains(rng, n_streams, channels)
    scale[:] = 0                                       # pass-through: the ring can be re-read as is
    plan = [("tick", 2), ("tick", 0), ("cycle", None), ("tick", 1), ("cycle", None), ("cycle", None), ("tick", 0)]
    with cm.Engine(channels, n_streams, block, ring_slots=ring) as eng:
        eng.set_gain_table(scale, gain)
        for slot in range(ring):
            eng.host_slot(slot)[:, : block * channels] = data[slot]
            eng.submit(slot)
        order = []
        for kind, slot in plan:
            if kind == "tick":
                eng.process(slot)
                order.append(slot)
            else:
                eng.process_cycle(0, ring)
                order += list(range(ring))
        eng.sync()
        meters = None
        frames = np.full(n_streams, block, np.uint32)
        for slot in order:
            meters, _ = port.batch(data[slot].copy(), frames, channels, scale, gain, meters=meters)
        check_meters(cm, port, eng, meters, n_streams, channels)


def test_upload_waits_for_ticks_on_resident_data(cm, port):
    """A tick on resident data records no event of its own; the slot's next upload asks for one. The
    upload must still not overwrite the slot while such a tick is reading it (in place)."""
    rng = np.random.default_rng(99)
    channels, n_streams, block = 2, 64, 200000                   # 51 MB per slot: ticks that take a while
    a_data = make_pcm(rng, "gauss", (n_streams, block * channels))
    b_data = make_pcm(rng, "full", (n_streams, block * channels))
    scale, gain = make_gains(rng, n_streams, channels)
    frames = np.full(n_streams, block, np.uint32)
    staging = cm.PinnedArray((n_streams, block * channels))
    with cm.Engine(channels, n_streams, block) as eng:
        eng.set_gain_table(scale, gain)
        eng.host_slot(0)[:, : block * channels] = a_data
        eng.submit(0)
        eng.process(0)                  # consumes the upload
        eng.process(0)                  # resident data, in place: transforms the transformed block again
        eng.process(0)
        staging.array[:] = b_data
        eng.submit(0, staging.array)    # must wait for all three ticks
        eng.process(0)
        eng.fetch(0)
        eng.sync()
        work = a_data.copy()
        meters = None
        for _ in range(3):
            meters, _ = port.batch(work, frames, channels, scale, gain, meters=meters)
        want = b_data.copy()
        meters, _ = port.batch(want, frames, channels, scale, gain, meters=meters)
        assert np.array_equal(eng.host_slot(0)[:, : block * channels], want)
        check_meters(cm, port, eng, meters, n_streams, channels)
    staging.free()


@pytest.mark.parametrize("channels,block,ragged,no_span", [
    (1, 320, False, False), (1, 320, False, True), (1, 320, True, False), (2, 4799, False, False),
    (2, 20000, True, False), (8, 257, False, False), (16, 100, True, False), (4, 333, False, True), (6, 500, False, False)])
def test_cycle_equals_individual_ticks(cm, port, monkeypatch, channels, block, ragged, no_span):
    """cmgpu_process_cycle runs a ring's ticks as ONE launch over all slots (a span: work items
    numbered (tick, stream, chunk)) or, for the kernels that cannot (6 channels here) and with
    CMGPU_NO_SPAN, as one CUDA graph of per-tick launches. Either way every tick must get its own
    place in the position order, and every replay a fresh one (the tick number lives on the device),
    otherwise an equal magnitude from a later tick could steal the peak. Same result as issuing the
    ticks one by one, and as the oracle."""
    rng = np.random.default_rng(77 + channels + block)
    n_streams, ring, cycles = 301 if block < 1000 else 23, 5, 3
    data = make_pcm(rng, "ties", (ring, n_streams, block * channels))
    # slot 0 opens with +100 everywhere, later slots repeat the magnitude with the other sign
    data[:, :7, :] = 0
    data[0, :7, 0] = 100
    data[1:, :7, 3 * channels] = -100
    scale, gain = make_gains(rng, n_streams, channels)
    scale[:7] = 0
    frames = np.full((ring, n_streams), block, np.uint32)
    if ragged:
        frames = rng.integers(0, block + 1, size=(ring, n_streams)).astype(np.uint32)
        frames[:, :7] = block
    if no_span:
        monkeypatch.setenv("CMGPU_NO_SPAN", "1")
    results = []
    for mode in ("cycle", "single"):
        with cm.Engine(channels, n_streams, block, ring_slots=ring) as eng:
            eng.set_gain_table(scale, gain)
            for slot in range(ring):
                eng.host_slot(slot)[:, : block * channels] = data[slot]
                if ragged:
                    eng.set_frames(slot, frames[slot])
                eng.submit(slot)
            for _ in range(cycles):
                if mode == "cycle":
                    eng.process_cycle(0, ring)
                else:
                    for slot in range(ring):
                        eng.process(slot)
            for slot in range(ring):
                eng.fetch(slot)
            eng.sync()
            outs = [eng.host_slot(slot)[:, : block * channels].copy() for slot in range(ring)]
            snap = eng.snapshot()
            results.append((outs, [cm.state_dict(snap[s], channels) for s in range(n_streams)]))
            spans = mode == "cycle" and not no_span and channels != 6
            assert eng.launch_count() == (cycles if spans else ring * cycles)
    # oracle: the in-place ring is transformed again on every cycle
    meters = None
    work = data.copy()
    for _ in range(cycles):
        for slot in range(ring):
            meters, _ = port.batch(work[slot], frames[slot], channels, scale, gain, meters=meters)
    for outs, states in results:
        for slot in range(ring):
            assert np.array_equal(outs[slot], work[slot])
        for s in range(n_streams):
            assert states[s]["frames"] == int(meters[s].frames)
            for c in range(channels):
                assert states[s]["power"][c] == int(meters[s].power[c])
                assert states[s]["channel_peak"][c] == int(meters[s].channel_peak[c]), f"stream {s} channel {c}"
            assert states[s]["global_peak"] == int(meters[s].global_peak)
    assert all(results[0][1][s]["channel_peak"][0] == 100 for s in range(7))


@pytest.mark.parametrize("channels,block,ragged,ring,separate", [
    (1, 320, False, 5, False), (1, 320, True, 7, False), (1, 320, False, 50, True), (2, 256, True, 4, True), (2, 255, False, 3, False),
    (4, 128, True, 6, False), (8, 64, True, 5, True), (1, 512, True, 2, False), (1, 5, True, 9, False), (2, 1, False, 4, False)])
def test_stream_major_span(cm, port, monkeypatch, channels, block, ragged, ring, separate):
    """The small-buffer span kernel (cmgpu_span.cuh): one 8-lane group owns a STREAM for all ticks of a
    span and publishes its meter partials once. Forced here for small stream counts
    (CMGPU_SPAN_BY_STREAM); equal peaks in different ticks and ragged frame counts included. Must equal
    the oracle run over the ticks in order, for two consecutive spans (the second continues the window)."""
    monkeypatch.setenv("CMGPU_SPAN_BY_STREAM", "1")
    rng = np.random.default_rng(500 + channels * 31 + block + ring)
    n_streams, cycles = 203, 2
    data = make_pcm(rng, "ties", (ring, n_streams, block * channels))
    data[:, :5, :] = 0
    data[1, :5, 0] = 77                       # first occurrence in tick 1 ...
    data[2:, :5, 0] = -77                     # ... repeated with the other sign later
    scale, gain = make_gains(rng, n_streams, channels)
    frames = np.full((ring, n_streams), block, np.uint32)
    if ragged:
        frames = rng.integers(0, block + 1, size=(ring, n_streams)).astype(np.uint32)
        frames[:, :5] = block
    flags = cm.SEPARATE_OUT if separate else 0
    with cm.Engine(channels, n_streams, block, ring_slots=ring, flags=flags) as eng:
        eng.set_gain_table(scale, gain)
        for slot in range(ring):
            eng.host_slot(slot)[:, : block * channels] = data[slot]
            if ragged:
                eng.set_frames(slot, frames[slot])
            eng.submit(slot)
        for _ in range(cycles):
            eng.process_cycle(0, ring)
        assert eng.launch_count() == cycles
        for slot in range(ring):
            eng.fetch(slot)
        eng.sync()
        outs = [eng.host_slot(slot)[:, : block * channels].copy() for slot in range(ring)]
        snap = eng.snapshot()
    meters = None
    work = data.copy()
    last = None
    for _ in range(cycles):
        src = data.copy() if separate else work           # a separate output ring leaves the input pristine
        for slot in range(ring):
            meters, _ = port.batch(src[slot], frames[slot], channels, scale, gain, meters=meters)
        last = src
        work = src
    for slot in range(ring):
        for s in range(n_streams):
            n = int(frames[slot][s]) * channels
            if separate:
                assert np.array_equal(outs[slot][s, :n], last[slot][s, :n]), f"slot {slot} stream {s}"
            else:
                assert np.array_equal(outs[slot][s], last[slot][s]), f"slot {slot} stream {s}"
    for s in range(n_streams):
        assert int(snap[s].frames) == int(meters[s].frames), f"stream {s}"
        for c in range(channels):
            assert int(snap[s].power[c]) == int(meters[s].power[c]), f"stream {s} ch {c}"
            assert int(snap[s].channel_peak[c]) == int(meters[s].channel_peak[c]), f"stream {s} ch {c}"
        assert int(snap[s].global_peak) == int(meters[s].global_peak), f"stream {s}"


@pytest.mark.parametrize("channels,block", [(1, 320), (2, 1000), (2, 4799), (4, 333), (8, 257), (16, 100), (6, 500), (3, 77)])
def test_planar_float_second_output(cm, port, channels, block):
    """the encoder-side sample-format stage (enc_vorbis.c:108-117: de-interleave,
    sample / 32768.f) as an optional second output of the fused pass. The checker is the oracle
    port's oracle_planar, which is pinned bit for bit to what the reference's OWN enc_vorbis.c writes
    (compiled unmodified against a libvorbis stand-in, oracle/Makefile target refenc;
    tests/golden/planar.json, tests/test_oracle.py) -- and, where that object code is present, it is
    asked directly as well."""
    from oracle import pyoracle
    enc = pyoracle.refenc()
    rng = np.random.default_rng(channels + block)
    n_streams = 21
    with cm.Engine(channels, n_streams, block, flags=cm.PLANAR_F32) as eng:
        host = eng.host_slot(0)
        host[:] = make_pcm(rng, "full", host.shape)
        scale, gain = make_gains(rng, n_streams, channels)
        eng.set_gain_table(scale, gain)
        frames = rng.integers(0, block + 1, size=n_streams).astype(np.uint32)
        frames[0] = block
        eng.set_frames(0, frames)
        want, meters = oracle_batch(port, host.copy(), frames, channels, scale, gain)
        eng.submit(0)
        eng.process(0, cm.FUSED | cm.PLANAR)
        eng.fetch(0)
        eng.sync()
        planes = eng.fetch_planar(0)
        eng.sync()
        assert np.array_equal(eng.host_slot(0), want)
        check_meters(cm, port, eng, meters, n_streams, channels)
        for s in range(n_streams):
            n = int(frames[s])
            ref = port.planar(want[s, : n * channels], channels)                # [channel][frame]
            got = planes[s, :, :n]
            assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"stream {s}"
            if enc is not None and n:
                assert np.array_equal(got.view(np.uint32), enc.planes(want[s, : n * channels], channels).view(np.uint32)), f"stream {s}"
    with cm.Engine(2, 2, 16) as eng:                                       # no plane ring: the flag is refused
        eng.submit(0)
        with pytest.raises(cm.CmgpuError):
            eng.process(0, cm.FUSED | cm.PLANAR)


PLANAR_GOLD = json.loads((GOLD / "planar.json").read_text()) if (GOLD / "planar.json").exists() else []


@pytest.mark.parametrize("case", PLANAR_GOLD, ids=lambda k: f"{k['channels']}ch_{k['frames']}f")
def test_planar_golden_from_reference_enc_vorbis(cm, case):
    """The float planes the reference's own enc_vorbis.c wrote for these inputs (tests/golden/planar.json),
    reproduced by the device's second output with the gain off (the encoder sees the transform's output;
    with no gain that is the input)."""
    channels, frames = case["channels"], case["frames"]
    pcm = np.array(case["pcm"], dtype=np.int16)
    with cm.Engine(channels, 3, frames, flags=cm.PLANAR_F32) as eng:
        host = eng.host_slot(0)
        host[:] = 0
        host[1, : frames * channels] = pcm
        eng.submit(0)
        eng.process(0, cm.FUSED | cm.PLANAR)
        planes = eng.fetch_planar(0)
        eng.sync()
        want = np.array(case["planes_u32"], dtype=np.uint32)
        assert np.array_equal(planes[1, :, :frames].view(np.uint32), want)


def test_opus_sized_blocks(cm, port):
    """ticks of exactly 2,880 frames (60 ms at 48 kHz), the packet size enc_opus.c:341
    hands to opus_encode -- each stream-block is then one ready-made, contiguous encoder input."""
    run_case(cm, port, 2, 64, 2880, None, "gauss", seed=2880)
    run_case(cm, port, 1, 64, 2880, None, "gauss", seed=2881)


@pytest.mark.parametrize("channels,block", [(2, 65536), (1, 131072 + 5), (2, 40001), (1, 200003), (4, 32768 + 3),
                                            (8, 16384 + 77), (16, 8192 + 5)])
@pytest.mark.parametrize("kind", ["ties", "full"])
def test_tma_staged_kernel(cm, port, channels, block, kind, monkeypatch):
    """The TMA-staged kernel (cmgpu_tma.cuh, CMGPU_TMA=1: bulk loads into a shared-memory ring,
    stores straight from registers) for long stream-blocks: ragged frame counts, a partial last
    vector, several ticks onto the same meter window -- same bit-exact bar as the default kernel."""
    monkeypatch.setenv("CMGPU_TMA", "1")
    rng = np.random.default_rng(block + channels)
    n_streams = 23
    frames = rng.integers(0, block + 1, size=n_streams)
    frames[0] = block
    frames[1] = 0
    frames[2] = block - 1
    frames[3] = 1
    name = run_case(cm, port, channels, n_streams, block, frames, kind, seed=block)
    assert name == f"tma_tick<C={channels}>"
    # several ticks into one window, in place
    data = make_pcm(rng, kind, (3, 5, block * channels))
    scale, gain = make_gains(rng, 5, channels)
    with cm.Engine(channels, 5, block) as eng:
        eng.set_gain_table(scale, gain)
        meters = None
        for t in range(3):
            eng.host_slot(0)[:, : block * channels] = data[t]
            fr = rng.integers(block // 2, block + 1, size=5).astype(np.uint32)
            eng.set_frames(0, fr)
            eng.submit(0)
            eng.process(0)
            eng.fetch(0)
            eng.sync()
            ref = data[t].copy()
            meters, _ = port.batch(ref, fr, channels, scale, gain, meters=meters)
            assert np.array_equal(eng.host_slot(0)[:, : block * channels], ref)
        check_meters(cm, port, eng, meters, 5, channels)


def test_tma_and_ldg_kernels_agree(cm, port, monkeypatch):
    b = run_case(cm, port, 2, 7, 50000, None, "ties", seed=5)
    monkeypatch.setenv("CMGPU_TMA", "1")
    a = run_case(cm, port, 2, 7, 50000, None, "ties", seed=5)
    assert a == "tma_tick<C=2>" and b == "fused_tick<C=2,G=32>"
