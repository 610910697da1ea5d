"""Drop-in proof. oracle/_ref/libcoolmic_dropin.so is the reference's own pull chain -- its
unmodified iohandle.c, tee.c, snddev.c and snddev_sine.c, wired by oracle/ref_harness.c in the order
of src/simple.c:212-229 -- with src/transform.c and src/vumeter.c REPLACED by the product's host shim
(libcoolmic-dsp_b200/csrc/host/*.c compiled with -DCOOLMIC_B200_WITH_IGLOO against the reference's
own headers). Every case below therefore runs
    mem/sine source -> [product transform, GPU] -> reference tee -> {consumer, [product vumeter, GPU]}
and must reproduce, bit for bit, what the all-reference build of the same harness produced (the
committed golden vectors, and oracle/_ref/libcoolmic_ref.so on fresh random inputs)."""
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
def dropin(cm):
    from oracle import pyoracle
    d = pyoracle.dropin()
    if d is None:
        pytest.skip("oracle/_ref/libcoolmic_dropin.so is not available here")
    return d


def launches(dropin):
    import ctypes
    fn = dropin.lib.coolmic_b200_shim_launches
    fn.restype = ctypes.c_uint64
    return int(fn())


def test_dropin_result_struct_is_the_references(dropin):
    assert dropin.sizeof_result() == 192


@pytest.mark.parametrize("case", [k for k in KATS if "out" in k], ids=lambda k: k["name"])
def test_dropin_pipeline_on_appendix_b(dropin, case):
    pcm = np.array(case["in"], dtype=np.int16).view(np.uint8)[: case["in_bytes"]]
    gain = None if case["gain"] is None else tuple(case["gain"])
    out, results, rc = dropin.pipeline(pcm, case["channels"], gain, src_chunk=case["src_chunk"], result_every=0)
    assert rc == case["gain_rc"]
    assert out.view(np.int16).tolist() == case["out"]
    assert same_result(results[-1], unhex(case["results"][0]))


@pytest.mark.parametrize("idx", range(0, len(FUZZ), 3))
def test_dropin_pipeline_on_fuzz_goldens(dropin, idx):
    case = FUZZ[idx]
    pcm = np.array(case["in_bytes"], dtype=np.uint8)
    out, results, rc = dropin.pipeline(pcm, case["channels"], tuple(case["gain"]), src_chunk=case["src_chunk"],
                                       pull=case["pull"], result_every=case["result_every"])
    assert rc == case["gain_rc"]
    assert out.tolist() == case["out_bytes"]
    want = [unhex(r) for r in case["results"]]
    assert len(results) == len(want)
    assert all(same_result(a, b) for a, b in zip(results, want))


@pytest.mark.parametrize("case", [s for s in SINE if "period" in s][:4],
                         ids=lambda s: f"{s['rate']}Hz-{s['gain']}")
def test_dropin_sine_driver_chain(dropin, case):
    """BASELINE config 1 in miniature: the reference's snddev_sine driver feeds the chain."""
    n = min(case["bytes"], 40960)
    src = dropin.sine(case["rate"], n)
    period = np.array(case["period"], dtype=np.int16)
    assert np.array_equal(src.view(np.int16), np.resize(period, n // 2))
    gain = None if case["gain"] is None else tuple(case["gain"])
    out, results, rc = dropin.pipeline(src, 1, gain, rate=case["rate"], result_every=20)
    assert rc == 0 and len(results) >= 2
    assert results[0]["frames"] == 20 * 512


@pytest.mark.parametrize("seed", range(8))
def test_dropin_vs_all_reference_build_random(dropin, ref, seed):
    rng = np.random.default_rng(7000 + seed)
    ch = int(rng.integers(1, 17))
    nbytes = int(rng.integers(1, 30000))
    x = rng.integers(-32768, 32768, size=nbytes // 2 + 1).astype(np.int16)
    pcm = x.view(np.uint8)[:nbytes]
    scale = int(rng.integers(1, 65536)) if seed % 5 else 0
    gn = int(rng.choice([ch, 1, ch]))
    gains = [int(v) for v in rng.integers(0, 65536, size=gn)]
    gain = (gn, scale, gains)
    chunk = int(rng.choice([0, 3, 100, 4096]))
    pull = int(rng.choice([1024, 512, 1000]))
    every = int(rng.choice([0, 3, 20]))
    out_r, res_r, rc_r = ref.pipeline(pcm, ch, gain, src_chunk=chunk, pull=pull, result_every=every)
    out_d, res_d, rc_d = dropin.pipeline(pcm, ch, gain, src_chunk=chunk, pull=pull, result_every=every)
    assert rc_r == rc_d
    assert np.array_equal(out_r, out_d)
    assert len(res_r) == len(res_d)
    assert all(same_result(a, b) for a, b in zip(res_r, res_d))


def test_dropin_reads_run_on_the_gpu(dropin):
    """The drop-in has no arithmetic of its own: a pipeline run shows up as kernel launches of the
    product library it links."""
    before = launches(dropin)
    pcm = (np.arange(4096) % 251 * 100).astype(np.int16)
    dropin.pipeline(pcm, 2, (2, 4, [3, 5]), result_every=0)
    assert launches(dropin) > before
