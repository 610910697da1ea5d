"""EXTENSION (parity unpinned): the N -> M integer downmix of BASELINE config 4. libcoolmic-dsp has
no downmix, so there is no reference output to compare with: the checker is
our own CPU restatement oracle_mix_process(), written in the reference's arithmetic style
(int64 accumulate, truncating division, saturation) and metered with the vumeter rules on both
the input and the output channels. GPU only."""
import numpy as np
import pytest

from oracle.pyoracle import Meter

pytestmark = pytest.mark.gpu


def run_mix(cm, port, cin, cout, n_streams, block_frames, seed, flags=0, ticks=1, kind="full", in_meter=True):
    rng = np.random.default_rng(seed)
    with cm.Engine(cin, n_streams, block_frames, out_channels=cout, flags=flags) as eng:
        scales = rng.integers(1, 65536, size=n_streams)
        scales[:4] = [1, 2, 65535, 32768][: min(4, n_streams)]
        weights = rng.integers(0, 65536, size=(n_streams, cout, cin)).astype(np.uint16)
        if n_streams > 5:
            weights[4] = 0
            weights[5] = 65535
        for s in range(n_streams):
            assert eng.set_mix(s, int(scales[s]), weights[s]) == 0
        m_in = [Meter() for _ in range(n_streams)]
        m_out = [Meter() for _ in range(n_streams)]
        for t in range(ticks):
            host = eng.host_slot(0)
            if kind == "full":
                host[:] = rng.integers(-32768, 32768, size=host.shape).astype(np.int16)
            else:
                host[:] = rng.choice(np.array([-32768, -3, 0, 3, 32767], dtype=np.int16), size=host.shape)
            frames = rng.integers(0, block_frames + 1, size=n_streams).astype(np.uint32)
            frames[0] = block_frames
            eng.set_frames(0, frames)
            src = host.copy()
            eng.submit(0)
            eng.process(0)
            eng.fetch(0)
            eng.sync()
            got = eng.host_out_slot(0)
            for s in range(n_streams):
                n = int(frames[s])
                want = port.mix(src[s, : n * cin], n, cin, cout, int(scales[s]), weights[s], m_in[s], m_out[s])
                assert np.array_equal(got[s, : n * cout], want), f"tick {t} stream {s}"
        so = eng.snapshot()
        si = eng.input_snapshot()
        if not in_meter:
            assert all(int(si[s].frames) == 0 and not any(si[s].power) and not any(si[s].channel_peak) for s in range(n_streams))
        for s in range(n_streams):
            for st, want, ch in ((so[s], m_out[s], cout), (si[s], m_in[s], cin))[: 2 if in_meter else 1]:
                assert int(st.frames) == int(want.frames)
                assert int(st.global_peak) == int(want.global_peak)
                for c in range(ch):
                    assert int(st.power[c]) == int(want.power[c]), f"stream {s} ch {c}"
                    assert int(st.channel_peak[c]) == int(want.channel_peak[c]), f"stream {s} ch {c}"
        return eng.kernel_name()


def test_downmix_8_to_2(cm, port):
    assert run_mix(cm, port, 8, 2, 37, 1000, seed=1) == "mix8to2_tick"
    assert run_mix(cm, port, 8, 2, 11, 5000, seed=2, ticks=3, kind="ties") == "mix8to2_tick"


def test_downmix_8_to_2_outputs_metered_only(cm, port):
    """CMGPU_MIX_OUTPUT_METER_ONLY: same PCM and output meters, the 8 input channels are not metered."""
    assert run_mix(cm, port, 8, 2, 37, 1000, seed=3, flags=cm.MIX_OUTPUT_METER_ONLY, in_meter=False).startswith("mix8to2_tick")
    assert run_mix(cm, port, 8, 2, 9, 7000, seed=4, ticks=2, kind="ties", flags=cm.MIX_OUTPUT_METER_ONLY,
                   in_meter=False) == "mix8to2_tick<outputs metered>"


def test_downmix_generic_kernel_agrees(cm, port):
    assert run_mix(cm, port, 8, 2, 9, 700, seed=3, flags=cm.FORCE_GENERIC) == "mix_tick<generic>"


@pytest.mark.parametrize("cin,cout", [(2, 1), (6, 2), (16, 16), (1, 2), (5, 3)])
def test_downmix_other_shapes(cm, port, cin, cout):
    run_mix(cm, port, cin, cout, 7, 333, seed=cin * 17 + cout, ticks=2)


def test_mix_argument_checks(cm):
    with cm.Engine(8, 2, 16, out_channels=2) as eng:
        assert eng.set_mix(0, 0, np.ones((2, 8))) == -10         # scale 0 is not a valid mix
        assert eng.set_mix(2, 1, np.ones((2, 8))) == -10
    with cm.Engine(2, 2, 16) as eng:
        assert eng.L.cmgpu_stream_set_mix(eng.ctx, 0, 1, np.ones(4, np.uint16).ctypes.data_as(
            __import__("ctypes").POINTER(__import__("ctypes").c_uint16))) == -10   # not a mix context
