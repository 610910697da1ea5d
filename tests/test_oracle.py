"""The oracle is only trusted once it is pinned: the plain-C port (oracle/coolmic_oracle.c)
must reproduce every golden vector that tests/golden/make_golden.py captured from the
reference's own object code, and -- where that object code is available -- agree with it on
fresh random inputs too. No GPU needed."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


def unhex(res):
    r = dict(res)
    if r.get("rc", 0) == 0 and "global_power" in r:
        r["global_power"] = float.fromhex(r["global_power"])
        r["channel_power"] = [float.fromhex(x) for x in r["channel_power"]]
    return r


def same_result(a, b):
    """Bit-exact comparison of two result dicts (doubles compared as bit patterns; -inf == -inf)."""
    if a.get("rc", 0) != b.get("rc", 0):
        return False
    if a.get("rc", 0) != 0:
        return True
    for k in ("rate", "channels", "frames", "global_peak", "channel_peak"):
        if a[k] != b[k]:
            return False
    fa = [a["global_power"]] + list(a["channel_power"])
    fb = [b["global_power"]] + list(b["channel_power"])
    return all(np.float64(x).tobytes() == np.float64(y).tobytes() for x, y in zip(fa, fb))


def port_pipeline(port, pcm, ch, gain, src_chunk=0, pull=1024, result_every=0, rate=48000):
    """The reference wiring restated with the port's objects: the tee delivers the same byte
    stream to both readers, so the meter sees transform's output in 1024-byte reads."""
    out, rc = port.transform(pcm, ch, gain, src_chunk=src_chunk, pull=pull)
    return out, rc


KATS = json.loads((GOLD / "kat_appendix_b.json").read_text())
SINE = json.loads((GOLD / "sine.json").read_text())
FUZZ = json.loads((GOLD / "fuzz_pipeline.json").read_text())


@pytest.mark.parametrize("case", [k for k in KATS if "out" in k], ids=lambda k: k["name"])
def test_port_matches_appendix_b(port, case):
    pcm = np.array(case["in"], dtype=np.int16).view(np.uint8)[: case["in_bytes"]]
    gain = None if case["gain"] is None else tuple(case["gain"])
    out, rc = port.transform(pcm, case["channels"], gain, src_chunk=case["src_chunk"])
    assert rc == case["gain_rc"]
    assert out.view(np.int16).tolist() == case["out"]
    got = port.vumeter(out, case["channels"])
    want = [unhex(r) for r in case["results"]]
    assert len(got) == len(want) == 1
    assert same_result(got[0], want[0]), (got, want)


def test_appendix_b_values_are_the_surveys():
    """Spot-check that the regenerated vectors are the ones SURVEY.md Appendix B records."""
    by = {k["name"]: k for k in KATS}
    assert by["b06_mono_third"]["results"][0]["global_peak"] == 10922
    assert float.fromhex(by["b02_mono_half"]["results"][0]["global_power"]) == -12.205422581725447
    assert by["b11_tie_pos_first"]["results"][0]["global_peak"] == 5
    assert by["b12_tie_neg_first"]["results"][0]["global_peak"] == -5
    assert by["b13_global_interleaved_first"]["results"][0]["global_peak"] == -7
    assert by["b13_global_interleaved_first"]["results"][0]["channel_peak"] == [7, -7]
    assert by["b10_three_ch_inval"]["gain_rc"] == -10
    assert by["b16_eof_mid_frame"]["results"][0]["frames"] == 5
    assert float.fromhex(by["b17_eight_ch_partial"]["results"][0]["channel_power"][6]) == 0.0
    assert math.isinf(float.fromhex(by["b14_silence"]["results"][0]["global_power"]))
    assert by["b18_second_result_inval"]["results"][1]["rc"] == -10


def test_port_second_result_is_inval(port):
    res = port.vumeter(np.array(KATS[-1]["in"], dtype=np.int16), 1, result_every=1)
    want = [unhex(r) for r in KATS[-1]["results"]]
    assert len(res) == len(want)
    assert all(same_result(a, b) for a, b in zip(res, want))


@pytest.mark.parametrize("case", [s for s in SINE if "period" in s],
                         ids=lambda s: f"{s['rate']}Hz-{s['gain']}")
def test_port_matches_sine_goldens(port, case):
    period = np.array(case["period"], dtype=np.int16)
    src = np.resize(period, case["bytes"] // 2)      # cyclic read from phase 0
    assert f"{port.fnv1a64(src):016x}" == case["src_fnv1a64"]
    gain = None if case["gain"] is None else tuple(case["gain"])
    out, rc = port.transform(src, 1, gain, rate=case["rate"])
    assert rc == 0
    assert f"{port.fnv1a64(out):016x}" == case["out_fnv1a64"]
    got = port.vumeter(out, 1, rate=case["rate"])
    assert same_result(got[-1], unhex(case["result"]))


def test_sine_driver_is_mono_only():
    assert SINE[-1]["stereo_open"] is False


@pytest.mark.parametrize("idx", range(len(FUZZ)))
def test_port_matches_fuzz_goldens(port, idx):
    case = FUZZ[idx]
    pcm = np.array(case["in_bytes"], dtype=np.uint8)
    gain = tuple(case["gain"])
    out, rc = port.transform(pcm, case["channels"], gain, src_chunk=case["src_chunk"], pull=case["pull"])
    assert rc == case["gain_rc"]
    assert out.tolist() == case["out_bytes"]
    # the meter behind the tee reads the transform's output in <=1024-byte pieces, once per
    # consumer pull: reproduce that cadence (pull <= 1024 keeps both tee readers in lockstep)
    got = port.vumeter(out, case["channels"], maxlen=-1, result_every=0)
    want = [unhex(r) for r in case["results"]]
    if case["result_every"] == 0:
        assert len(want) == 1 and same_result(got[-1], want[0])


def test_fnv_matches_python(port):
    from oracle import pyoracle
    data = np.arange(1000, dtype=np.uint8)
    assert port.fnv1a64(data) == pyoracle.fnv1a64(data)


# ---- port vs the reference's own object code on fresh inputs (skipped where _ref is absent) ----

def test_ref_result_struct_is_192_bytes(ref):
    assert ref.sizeof_result() == 192


@pytest.mark.parametrize("seed", range(24))
def test_port_vs_reference_random(port, ref, seed):
    rng = np.random.default_rng(1000 + seed)
    ch = int(rng.integers(1, 17))
    nbytes = int(rng.integers(0, 20000))
    x = rng.integers(-32768, 32768, size=nbytes // 2 + 1).astype(np.int16)
    if seed % 3 == 0:
        x = (x // 64).astype(np.int16)
    pcm = x.view(np.uint8)[:nbytes]
    scale = int(rng.integers(0, 65536)) if seed % 7 else 0
    gn = int(rng.choice([ch, 1, 2, 0, 3]))
    gains = [int(v) for v in rng.integers(0, 65536, size=max(gn, 1))]
    gain = (gn, scale, gains[:gn] if gn else None)
    chunk = int(rng.choice([0, 1, 2, 3, 7, 100, 4096]))
    pull = int(rng.choice([1024, 512, 33, 8192]))
    out_r, rc_r = ref.transform(pcm, ch, gain, src_chunk=chunk, pull=pull)
    out_p, rc_p = port.transform(pcm, ch, gain, src_chunk=chunk, pull=pull)
    assert rc_r == rc_p
    assert np.array_equal(out_r, out_p)
    for maxlen, every in [(-1, 0), (100, 3), (1, 0), (7, 5)]:
        if maxlen == 1 and nbytes > 4000:
            continue
        res_r = ref.vumeter(out_r, ch, src_chunk=chunk, maxlen=maxlen, result_every=every)
        res_p = port.vumeter(out_p, ch, src_chunk=chunk, maxlen=maxlen, result_every=every)
        assert len(res_r) == len(res_p)
        assert all(same_result(a, b) for a, b in zip(res_r, res_p))


def test_reference_reproduces_goldens(ref):
    """The committed fixtures are regenerable: re-run two of them through oracle/_ref."""
    case = KATS[6]
    pcm = np.array(case["in"], dtype=np.int16).view(np.uint8)[: case["in_bytes"]]
    out, results, rc = ref.pipeline(pcm, case["channels"], tuple(case["gain"]), result_every=0)
    assert out.view(np.int16).tolist() == case["out"]
    assert same_result(results[0], unhex(case["results"][0]))
    s = SINE[4]
    src = ref.sine(s["rate"], s["bytes"])
    out, results, rc = ref.pipeline(src, 1, tuple(s["gain"]), rate=s["rate"], result_every=0)
    assert same_result(results[-1], unhex(s["result"]))


def test_gain_adapt_matches_reference_cases(port):
    # SURVEY.md Appendix A.2 / reference src/transform.c:195-222
    assert port.gain_adapt(2, 2, 10, [3, 4]) == (0, 10, [3, 4])
    assert port.gain_adapt(3, 1, 10, [7]) == (0, 10, [7, 7, 7])
    assert port.gain_adapt(1, 2, 2, [1, 3]) == (0, 2, [2])
    assert port.gain_adapt(1, 2, 2, [65535, 65535]) == (0, 2, [65535])
    rc, scale, g = port.gain_adapt(3, 2, 2, [1, 3], state=(5, [9, 9, 9]))
    assert (rc, scale, g) == (-10, 5, [9, 9, 9])
    assert port.gain_adapt(2, 0, 10, [1, 2], state=(5, [9, 9]))[:2] == (0, 0)
    assert port.gain_adapt(2, 2, 0, [1, 2], state=(5, [9, 9]))[:2] == (0, 0)
    assert port.gain_adapt(2, 2, 9, None, state=(5, [9, 9]))[:2] == (0, 0)


# ---- the encoder-side sample-format stage (enc_vorbis.c:108-117) ---------------------
PLANAR = json.loads((GOLD / "planar.json").read_text()) if (GOLD / "planar.json").exists() else []


@pytest.mark.parametrize("case", PLANAR, ids=lambda k: f"{k['channels']}ch_{k['frames']}f")
def test_planar_port_matches_reference_enc_vorbis_fixture(port, case):
    """oracle_planar (the port of enc_vorbis.c:108-117) against planes the reference's OWN enc_vorbis.c wrote
    (tests/golden/planar.json, generated by make_golden.py through oracle/_ref/libcoolmic_refenc.so)."""
    pcm = np.array(case["pcm"], dtype=np.int16)
    got = port.planar(pcm, case["channels"])
    want = np.array(case["planes_u32"], dtype=np.uint32)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want)


def test_planar_port_matches_reference_enc_vorbis_live(port):
    """... and against that object code run here on fresh inputs, every channel count, odd lengths."""
    from oracle import pyoracle
    enc = pyoracle.refenc()
    if enc is None:
        pytest.skip("oracle/_ref/libcoolmic_refenc.so is not available here")
    rng = np.random.default_rng(99)
    for ch in range(1, 17):
        frames = int(rng.integers(1, 3000))
        pcm = rng.integers(-32768, 32768, size=ch * frames).astype(np.int16)
        assert np.array_equal(port.planar(pcm, ch).view(np.uint32), enc.planes(pcm, ch).view(np.uint32)), ch
    # all 65,536 sample values
    allx = np.arange(-32768, 32768, dtype=np.int16)
    assert np.array_equal(port.planar(allx, 1).view(np.uint32), enc.planes(allx, 1).view(np.uint32))
