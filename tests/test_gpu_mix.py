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


def test_downmix_sync_by_completion_word(cm, port):
    """Downmix launches carry the completion word like every other tick (MixArgs::done_flag): cmgpu_sync straight after a
    tick on resident data returns when the launch's last CTA has written the word, and everything the tick wrote is
    visible by then -- the output ring is read with a plain cudaMemcpy that nothing but the host's wait orders after the
    tick; the weights change from tick to tick."""
    import ctypes as C
    from tests.test_gpu_post import _cudart
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rng = np.random.default_rng(77)
    cin, cout, n_streams, block = 8, 2, 96, 700
    with cm.Engine(cin, n_streams, block, out_channels=cout) as eng:
        host = eng.host_slot(0)
        got = np.empty_like(eng.host_out_slot(0))
        m_in = [Meter() for _ in range(n_streams)]
        m_out = [Meter() for _ in range(n_streams)]
        for it in range(20):
            scales = rng.integers(1, 65536, size=n_streams)
            weights = rng.integers(0, 65536, size=(n_streams, cout, cin)).astype(np.uint16)
            for s in range(n_streams):
                assert eng.set_mix(s, int(scales[s]), weights[s]) == 0
            host[:] = rng.integers(-32768, 32768, size=host.shape).astype(np.int16)
            src = host.copy()
            eng.submit(0)
            eng.process(0)                      # consumes the upload (a streaming tick)
            eng.sync()                          # a launch gets a word only on a stream that was just waited for
            eng.process(0)                      # on resident data: a bare launch, the tail of the compute stream
            eng.sync()                          # <- by completion word
            assert rt.cudaMemcpy(got.ctypes.data, eng.device_out_slot(0), got.nbytes, 2) == 0
            for s in range(n_streams):
                port.mix(src[s, : block * cin], block, cin, cout, int(scales[s]), weights[s], m_in[s], m_out[s])
                want = port.mix(src[s, : block * cin], block, cin, cout, int(scales[s]), weights[s], m_in[s], m_out[s])
                assert np.array_equal(got[s, : block * cout], want), f"tick {it} stream {s}: output not complete when cmgpu_sync returned"
        # (every second sync of the loop can end by the word; a wait that outlasts the library's 60 us of polling falls
        #  back to the driver and is not counted, so leave slack for a busy box)
        assert eng.word_waits() >= 10, "cmgpu_sync never took the completion-word path for downmix launches"
        so = eng.snapshot()
        si = eng.input_snapshot()
        for s in range(n_streams):
            for st, want, ch in ((so[s], m_out[s], cout), (si[s], m_in[s], cin)):
                assert int(st.frames) == int(want.frames)
                for c in range(ch):
                    assert int(st.power[c]) == int(want.power[c])
                    assert int(st.channel_peak[c]) == int(want.channel_peak[c])
