"""GPU parity at BASELINE.json's full sizes, and the division recipe proven ON THE DEVICE.

* configs 4a (4,096 x 8 ch x 96,000 frames) and 5 (65,536 stereo streams; seconds shortened, stream
  count kept -- that is where 32-bit item arithmetic could bite): the inputs are the SURVEY.md 8d data
  sets (the reference's snddev_sine period read cyclically + splitmix64 noise), written by the device
  generator; a sample of streams is re-derived on the host and checked sample for sample against the
  oracle, every stream is checked through properties that need no oracle (frame count, sum of squares
  and peak magnitude recomputed from the fetched PCM; idempotence of a repeated identical tick).
* all 65,536 inputs x the ~900 (gain, scale) pairs of tests/test_recipe.py through the kernels
  themselves -- fused_tick in MASKED and in ADDALL mode, mono and stereo, 32- and 8-lane groups,
  generic_tick and any_tick -- against the reference arithmetic (transform.c:110-123).

Needs a B200: run with `pytest -m gpu`.
"""
import numpy as np
import pytest

from tests.test_recipe import pairs, truth

pytestmark = pytest.mark.gpu


def bench_gains(first, n, channels):
    s = np.arange(first, first + n)
    scale = (1000 + s % 9000).astype(np.uint16)
    gain = (scale[:, None].astype(np.int64) * 3 // 4 + 37 * ((s[:, None] + np.arange(channels)) % 64)).astype(np.uint16)
    return scale, gain


def _full_size(cm, port, channels, n_streams, block, rate, sstep, cstep, oracle_streams, every_prop):
    from libcoolmic_dsp_b200 import synth
    period = synth.load_period(rate)
    scale, gain = bench_gains(0, n_streams, channels)
    with cm.Engine(channels, n_streams, block, flags=cm.NO_PINNED | cm.SEPARATE_OUT) as eng:
        eng.set_gain_table(scale, gain)
        synth.device_fill(eng, 0, period, 0, 0, sstep, cstep, 16, 5)
        eng.process(0)
        out = np.empty((n_streams, eng.stride // 2), dtype=np.int16)
        eng.fetch(0, out)
        eng.sync()
        res, st, rcs = eng.results(rate, reset=True)
        assert all(rc == 0 for rc in rcs)
        # (1) oracle on a sample of streams, inputs re-derived on the host
        for s in oracle_streams:
            pcm = np.ascontiguousarray(synth.synth_rows(period, s, 1, channels, block, 0, sstep, cstep, 16, 5))
            meters, _ = port.batch(pcm, np.array([block], np.uint32), channels, scale[s:s + 1], gain[s:s + 1])
            assert np.array_equal(out[s, : block * channels], pcm[0]), f"stream {s}: PCM"
            assert int(st[s].frames) == block and int(st[s].global_peak) == int(meters[0].global_peak)
            for c in range(channels):
                assert int(st[s].power[c]) == int(meters[0].power[c]), f"stream {s} ch {c}"
                assert int(st[s].channel_peak[c]) == int(meters[0].channel_peak[c]), f"stream {s} ch {c}"
            want = port.finalise(meters[0], rate, channels)
            got = res[s].as_dict()
            assert np.float64(got["global_power"]).tobytes() == np.float64(want["global_power"]).tobytes()
        # (2) every stream: frames; every `every_prop`-th: meter state recomputed from the fetched PCM
        frames = np.array([int(st[s].frames) for s in range(n_streams)])
        assert (frames == block).all()
        for s in range(0, n_streams, every_prop):
            y = out[s, : block * channels].astype(np.int64).reshape(block, channels)
            for c in range(channels):
                assert int(st[s].power[c]) == int((y[:, c] * y[:, c]).sum()), f"stream {s} ch {c}: power"
                k = int(np.argmax(np.abs(y[:, c])))                  # first occurrence
                assert int(st[s].channel_peak[c]) == int(y[k, c]), f"stream {s} ch {c}: peak"
        # (3) idempotence: the same tick again over the pristine input ring gives the same state
        eng.process(0)
        res2, st2, _ = eng.results(rate, reset=True)
        for s in range(0, n_streams, max(1, n_streams // 997)):
            assert bytes(st2[s]) == bytes(st[s]), f"stream {s}: a repeated tick must meter the same"
        return eng.kernel_name()


def test_full_size_config4a(cm, port):
    """BASELINE config 4 (parity mode 4a): 4,096 eight-channel 48 kHz streams x 2 s, input
    period[(f + 7s + 5c) mod 48] of the reference's snddev_sine table."""
    name = _full_size(cm, port, 8, 4096, 96000, 48000, 7, 5, oracle_streams=[0, 5, 2047, 4090, 4095], every_prop=64)
    assert name.startswith("fused_tick<C=8")


def test_full_size_config5_stream_count(cm, port):
    """BASELINE config 5's shape on ONE GPU: all 65,536 stereo streams, 0.25 s each (12,000 frames; the
    stream count is what stresses the item arithmetic, the seconds only repeat it)."""
    name = _full_size(cm, port, 2, 65536, 12000, 48000, 7, 3,
                      oracle_streams=[0, 5, 21, 32767, 32768, 65529, 65535], every_prop=257)
    assert name.startswith("fused_tick<C=2")


def test_config5_full_second_on_a_shard(cm, port):
    """Config 5's full 1 s stream-blocks (48,000 frames) on the 8,192-stream shard one of 8 GPUs owns."""
    _full_size(cm, port, 2, 8192, 48000, 48000, 7, 3, oracle_streams=[0, 5, 4095, 8191], every_prop=129)


def test_config3_full_size_cycle(cm, port):
    """BASELINE config 3 at full size: 16,384 mono 16 kHz streams, 320-frame blocks, 50 ticks as one
    span launch over a 50-slot ring; input period[(f + 5s) mod 16] of the reference's 16 kHz table."""
    from libcoolmic_dsp_b200 import synth
    n_streams, channels, block, ring, rate = 16384, 1, 320, 50, 16000
    period = synth.load_period(rate)
    scale, gain = bench_gains(0, n_streams, channels)
    with cm.Engine(channels, n_streams, block, ring_slots=ring, flags=cm.NO_PINNED | cm.SEPARATE_OUT) as eng:
        eng.set_gain_table(scale, gain)
        eng.tone_table(period)
        for t in range(ring):
            synth.device_fill(eng, t, None, 0, t * block, 5, 0, 16, 5)
        eng.process_cycle(0, ring)
        out = np.empty((ring, n_streams, eng.stride // 2), dtype=np.int16)
        for t in range(ring):
            eng.fetch(t, out[t])
        eng.sync()
        res, st, rcs = eng.results(rate)
        check = [0, 5, 21, 8191, 16383]
        for s in check:
            pcm = np.ascontiguousarray(synth.synth_rows(period, s, 1, channels, block * ring, 0, 5, 0, 16, 5))
            meters, _ = port.batch(pcm, np.array([block * ring], np.uint32), channels, scale[s:s + 1], gain[s:s + 1])
            got = np.concatenate([out[t, s, :block] for t in range(ring)])
            assert np.array_equal(got, pcm[0]), f"stream {s}"
            assert int(st[s].power[0]) == int(meters[0].power[0]) and int(st[s].channel_peak[0]) == int(meters[0].channel_peak[0])
        y = out[:, :, :block].astype(np.int64)                        # [tick][stream][frame]
        power = (y * y).sum(axis=(0, 2))
        assert all(int(st[s].power[0]) == int(power[s]) and int(st[s].frames) == block * ring for s in range(n_streams))
        mags = np.abs(y).transpose(1, 0, 2).reshape(n_streams, -1)
        first = mags.argmax(axis=1)
        vals = y.transpose(1, 0, 2).reshape(n_streams, -1)[np.arange(n_streams), first]
        assert all(int(st[s].channel_peak[0]) == int(vals[s]) for s in range(n_streams))


# ---- the division recipe on the device, exhaustively ------------------------------------------------

ALL_X = np.arange(-32768, 32768, dtype=np.int16)


def _device_tables(cm, channels, block, grouped, flags=0):
    """`grouped`: (g, d) pairs, every run of `channels` consecutive ones sharing d -> one stream each,
    pair k of the run on channel k; every channel of every stream sees all 65,536 inputs (blocks shorter
    than that spread a stream's inputs over 65536 / block sub-streams with the same gains). Returns the
    transformed PCM as [len(grouped)][65536] in input order, and the kernel's name."""
    assert len(grouped) % channels == 0 and 65536 % block == 0
    n_streams = len(grouped) // channels
    per_stream = 65536 // block
    total = n_streams * per_stream
    with cm.Engine(channels, total, block, flags=flags | cm.NO_PINNED) as eng:
        scale = np.repeat(np.array([grouped[s * channels][1] for s in range(n_streams)], np.uint16), per_stream)
        gain = np.repeat(np.array([[grouped[s * channels + c][0] for c in range(channels)] for s in range(n_streams)],
                                  np.uint16), per_stream, axis=0)
        host = np.zeros((total, eng.stride // 2), dtype=np.int16)
        for k in range(per_stream):
            host[k::per_stream, : block * channels] = np.repeat(ALL_X[k * block:(k + 1) * block], channels)[None, :]
        eng.set_gain_table(scale, gain)
        eng.submit(0, host)
        eng.process(0)
        got = np.empty_like(host)
        eng.fetch(0, got)
        eng.sync()
        kernel = eng.kernel_name()
    g3 = got[:, : block * channels].reshape(n_streams, per_stream, block, channels)
    tables = g3.transpose(0, 3, 1, 2).reshape(n_streams * channels, 65536)
    return tables, kernel


@pytest.mark.parametrize("mode", ["masked", "addall"])
@pytest.mark.parametrize("channels,block,flags,expect", [
    (1, 65536, 0, "fused_tick<C=1,G=32>"), (2, 65536, 0, "fused_tick<C=2,G=32>"), (2, 65536, "separate", "fused_tick<C=2,G=32>"),
    (1, 256, 0, "fused_tick<C=1,G=8>"), (2, 128, 0, "fused_tick<C=2,G=8>"), (8, 65536, 0, "fused_tick<C=8,G=32>"),
    (2, 65536, "generic", "generic_tick"), (3, 65536, 0, "any_tick"), (6, 65536, "separate", "any_tick")])
def test_division_recipe_exhaustive_on_device(cm, mode, channels, block, flags, expect):
    """All 65,536 inputs x every (gain, scale) pair of tests/test_recipe.py through the device code that
    the headline kernels run: IMAD.HI + I2IP saturation in MASKED mode (all pairs) and in ADDALL mode
    (pairs with gain/scale >= 1/2 only, so that the host picks the add-all kernel)."""
    ps = pairs()
    ps = [(g, d) for g, d in ps if g != d]                      # unity rows would make a stream 'identity'
    if mode == "addall":
        ps = [(g, d) for g, d in ps if 2 * g >= d]
    # the channels of one stream share a scale: group pairs by scale, pad groups with a repeat
    by_d = {}
    for g, d in ps:
        by_d.setdefault(d, []).append(g)
    grouped = []
    for d, gs in sorted(by_d.items()):
        while len(gs) % channels:
            gs.append(gs[-1])
        grouped += [(g, d) for g in gs]
    f = {"separate": cm.SEPARATE_OUT, "generic": cm.FORCE_GENERIC}.get(flags, 0)
    tables, kernel = _device_tables(cm, channels, block, grouped, f)
    assert kernel.startswith(expect), kernel
    bad = []
    for i, (g, d) in enumerate(grouped):
        if not np.array_equal(tables[i], truth(g, d)):
            bad.append((g, d))
    assert not bad, f"{kernel} ({mode}): device result differs from trunc(x*g/d) for {bad[:8]} ({len(bad)} pairs)"
