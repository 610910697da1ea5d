"""The reference-named objects (coolmic_transform_*, coolmic_vumeter_*, coolmic_iohandle_*) exported
by the product, driven through the same call shapes as the reference's objects and compared with
them (oracle/_ref when present, the pinned port otherwise) and with the golden vectors. GPU only:
every read of these objects runs a tick on the device."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.test_oracle import same_result, unhex

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
KATS = json.loads((GOLD / "kat_appendix_b.json").read_text())
FUZZ = json.loads((GOLD / "fuzz_pipeline.json").read_text())
SINE = json.loads((GOLD / "sine.json").read_text())


@pytest.fixture(scope="module")
def shim(cm):
    from tests.shimlib import ShimLib
    return ShimLib()


@pytest.fixture(scope="module")
def checker(port):
    from oracle import pyoracle
    return pyoracle.ref() or port


@pytest.mark.parametrize("case", [k for k in KATS if "out" in k], ids=lambda k: k["name"])
def test_shim_objects_on_appendix_b(shim, case):
    pcm = np.array(case["in"], dtype=np.int16).view(np.uint8)[: case["in_bytes"]]
    gain = None if case["gain"] is None else tuple(case["gain"])
    out, rc = shim.transform(pcm, case["channels"], gain, src_chunk=case["src_chunk"])
    assert rc == case["gain_rc"]
    assert out.view(np.int16).tolist() == case["out"]
    got = shim.vumeter(out, case["channels"])
    assert len(got) == 1 and same_result(got[0], unhex(case["results"][0]))
    # and with the meter pulling straight from the transform's handle
    got = shim.chain(pcm, case["channels"], gain, src_chunk=case["src_chunk"])
    assert same_result(got[-1], unhex(case["results"][0]))


def test_shim_second_result_is_inval(shim):
    res = shim.vumeter(np.array(KATS[-1]["in"], dtype=np.int16), 1, result_every=1)
    want = [unhex(r) for r in KATS[-1]["results"]]
    assert len(res) == len(want) and all(same_result(a, b) for a, b in zip(res, want))


@pytest.mark.parametrize("idx", range(0, len(FUZZ), 2))
def test_shim_transform_on_fuzz_goldens(shim, idx):
    case = FUZZ[idx]
    pcm = np.array(case["in_bytes"], dtype=np.uint8)
    out, rc = shim.transform(pcm, case["channels"], tuple(case["gain"]), src_chunk=case["src_chunk"], pull=case["pull"])
    assert rc == case["gain_rc"]
    assert out.tolist() == case["out_bytes"]


@pytest.mark.parametrize("seed", range(10))
def test_shim_vs_reference_random_chunking(shim, checker, seed):
    rng = np.random.default_rng(500 + seed)
    ch = int(rng.integers(1, 17))
    nbytes = int(rng.integers(0, 30000))
    x = rng.integers(-32768, 32768, size=nbytes // 2 + 1).astype(np.int16)
    pcm = x.view(np.uint8)[:nbytes]
    scale = int(rng.integers(0, 65536)) if seed % 5 else 0
    gn = int(rng.choice([ch, 1, 2, 0]))
    gains = [int(v) for v in rng.integers(0, 65536, size=max(gn, 1))]
    gain = (gn, scale, gains[:gn] if gn else None)
    chunk = int(rng.choice([0, 1, 3, 7, 100, 4096]))
    pull = int(rng.choice([1024, 512, 33, 8192, 20000]))
    out_s, rc_s = shim.transform(pcm, ch, gain, src_chunk=chunk, pull=pull)
    out_r, rc_r = checker.transform(pcm, ch, gain, src_chunk=chunk, pull=pull)
    assert rc_s == rc_r and np.array_equal(out_s, out_r)
    for maxlen, every in [(-1, 0), (100, 3), (7, 5)]:
        res_s = shim.vumeter(out_s, ch, src_chunk=chunk, maxlen=maxlen, result_every=every)
        res_r = checker.vumeter(out_r, ch, src_chunk=chunk, maxlen=maxlen, result_every=every)
        assert len(res_s) == len(res_r)
        assert all(same_result(a, b) for a, b in zip(res_s, res_r)), (maxlen, every)


def test_shim_sine_golden_16k(shim, port):
    case = [s for s in SINE if s.get("rate") == 16000 and s.get("gain")][0]
    src = np.resize(np.array(case["period"], dtype=np.int16), case["bytes"] // 2)
    out, rc = shim.transform(src, 1, tuple(case["gain"]), rate=16000)
    assert f"{port.fnv1a64(out):016x}" == case["out_fnv1a64"]
    assert same_result(shim.vumeter(out, 1, rate=16000)[-1], unhex(case["result"]))


def test_shim_reads_ran_on_the_gpu(shim, cm):
    before = cm.lib().coolmic_b200_shim_launches()
    shim.chain(np.arange(4000, dtype=np.int16), 2, (1, 3, [2]))
    assert cm.lib().coolmic_b200_shim_launches() > before


@pytest.mark.parametrize("channels,block_frames,every,chunk", [(2, 256, 3, 0), (1, 1000, 1, 7), (8, 64, 0, 100),
                                                             (3, 333, 2, 0), (16, 50, 5, 3)])
def test_batch_mode_objects(shim, port, channels, block_frames, every, chunk):
    """member transforms + fused meters of one batch, two readers per stream.
    PCM must equal the reference transform's output; every result must equal the reference meter
    run over exactly the frames of its window of ticks."""
    rng = np.random.default_rng(channels * 1000 + block_frames)
    n = 13
    nbytes = 2 * channels * 1500 + 6            # ends inside a frame: the tail is withheld (Appendix B row 16)
    pcm = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    scale = rng.integers(0, 65536, size=n).astype(np.uint16)
    scale[0] = 0
    gain = rng.integers(0, 65536, size=(n, channels)).astype(np.uint16)
    gain[1] = scale[1]
    ticks, outs, results, flags = shim.batch(pcm, channels, scale, gain, src_chunk=chunk, block_frames=block_frames,
                                             result_every_ticks=every, pull=1024)
    assert flags == 0, "second reader saw different bytes, or EOF was not reported"
    whole = (nbytes // (2 * channels)) * 2 * channels
    frames_total = whole // (2 * channels)
    assert ticks == -(-frames_total // block_frames)
    for s in range(n):
        want, rc = port.transform(pcm[s], channels, (channels, int(scale[s]), gain[s].tolist()))
        assert rc == 0 and np.array_equal(outs[s], want), f"stream {s}"
        # windows: `every` ticks of block_frames frames each (the last one shorter)
        step = (every or ticks) * block_frames * 2 * channels
        want_res = []
        for lo in range(0, whole, step):
            want_res.append(port.vumeter(want[lo: lo + step], channels)[-1])
        got = [r for r in results[s] if r.get("rc", 0) == 0]
        assert len(got) == len(want_res), (len(got), len(want_res))
        for a, b in zip(got, want_res):
            assert same_result(a, b), f"stream {s}: {a} != {b}"


@pytest.mark.parametrize("channels,block_frames,slots,threads,chunk", [(2, 256, 3, 1, 0), (1, 100, 4, 3, 7), (8, 64, 2, 2, 100),
                                                                      (5, 333, 3, 4, 0), (2, 1000, 1, 2, 0)])
def test_ring_batch_producer_runs_ahead(shim, port, channels, block_frames, slots, threads, chunk):
    """The pipelined form of batch mode (coolmic_b200_batch_new_ring): ticks are only queued, the producer
    gets `slots` ticks ahead of the readers and is refused (COOLMIC_ERROR_BUSY) beyond that; readers
    follow tick by tick, each read waiting for its own slot's download only. PCM and the final results
    (one device round trip for all members) must equal the reference's, as in the synchronous form."""
    rng = np.random.default_rng(channels * 77 + block_frames + slots)
    n = 29
    nbytes = 2 * channels * 2500 + 2 * channels - 1     # ends inside a frame
    pcm = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    scale = rng.integers(0, 65536, size=n).astype(np.uint16)
    scale[0] = 0
    gain = rng.integers(0, 65536, size=(n, channels)).astype(np.uint16)
    ticks, outs, results, flags = shim.batch_ring(pcm, channels, scale, gain, src_chunk=chunk, block_frames=block_frames,
                                                  slots=slots, threads=threads, pull=777)
    assert flags == 0, f"ring semantics violated: flags {flags}"
    frames_total = nbytes // (2 * channels)
    assert ticks == -(-frames_total // block_frames)
    for s in range(n):
        want, rc = port.transform(pcm[s], channels, (channels, int(scale[s]), gain[s].tolist()))
        assert rc == 0 and np.array_equal(outs[s], want), f"stream {s}"
        assert same_result(results[s], port.vumeter(want, channels)[-1]), f"stream {s}"
