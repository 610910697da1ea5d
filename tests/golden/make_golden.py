#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the REFERENCE's own object code (oracle/_ref,
compiled unmodified from /root/reference/src by oracle/Makefile).

Run here, in the authoring container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no tests, fixtures or known-answer vectors for this path (SURVEY.md
section 4), so these files are what pins the oracle port and the CUDA path: every number in
them was produced by reference code, none by ours. Doubles are stored as C99 hex floats so
that comparisons can be bit-exact.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402

OUT = Path(__file__).resolve().parent
A12 = [-3, -1, 0, 1, 3, 32767, -32768, 21845, -21845, 100, -100, 7]


def hexify(res: dict) -> dict:
    r = dict(res)
    if r.get("rc", 0) != 0:
        return r
    r["global_power"] = float(r["global_power"]).hex()
    r["channel_power"] = [float(x).hex() for x in r["channel_power"]]
    return r


def kat_cases():
    """(name, channels, samples, gain, src_chunk, truncate_bytes)."""
    return [
        ("b01_mono_disabled", 1, A12, (0, 0, None), 0, None),
        ("b02_mono_half", 1, A12, (1, 2, [1]), 0, None),
        ("b03_mono_x1p5", 1, A12, (1, 2, [3]), 0, None),
        ("b04_mono_x65535", 1, A12, (1, 1, [65535]), 0, None),
        ("b05_mono_mute", 1, A12, (1, 1, [0]), 0, None),
        ("b06_mono_third", 1, A12, (1, 3, [1]), 0, None),
        ("b07_stereo_1_3", 2, A12, (2, 2, [1, 3]), 0, None),
        ("b08_stereo_broadcast", 2, A12, (1, 2, [3]), 0, None),
        ("b09_mono_from_stereo_setting", 1, A12, (2, 2, [1, 3]), 0, None),
        ("b10_three_ch_inval", 3, A12, (2, 2, [1, 3]), 0, None),
        ("b11_tie_pos_first", 1, [5, -5, -5, 5], None, 0, None),
        ("b12_tie_neg_first", 1, [-5, 5, 5, -5], None, 0, None),
        ("b13_global_interleaved_first", 2, [1, -7, 7, 3, -7, 7], None, 0, None),
        ("b14_silence", 2, [0, 0, 0, 0], None, 0, None),
        ("b15_three_byte_chunks", 2, A12, (1, 2, [3]), 3, None),
        ("b16_eof_mid_frame", 2, A12, (1, 2, [3]), 0, 22),
        ("b17_eight_ch_partial", 8, A12, (1, 2, [3]), 0, None),
    ]


def main():
    ref = po.ref()
    if ref is None:
        raise SystemExit("oracle/_ref is not built and /root/reference is absent")
    port = po.port()

    kats = []
    for name, ch, samples, gain, chunk, trunc in kat_cases():
        pcm = np.array(samples, dtype=np.int16).view(np.uint8)
        if trunc is not None:
            pcm = pcm[:trunc]
        out, results, rc = ref.pipeline(pcm, ch, gain, src_chunk=chunk, result_every=0)
        kats.append({
            "name": name, "channels": ch, "in": [int(x) for x in samples], "in_bytes": int(pcm.size),
            "gain": None if gain is None else [gain[0], gain[1], gain[2]],
            "src_chunk": chunk, "gain_rc": rc,
            "out": [int(x) for x in out.view(np.int16)],
            "results": [hexify(r) for r in results],
        })
    # row 18: a second result() right after a successful one
    res = ref.vumeter(np.array(A12, dtype=np.int16), 1, result_every=1)
    kats.append({"name": "b18_second_result_inval", "channels": 1, "in": A12,
                 "results": [hexify(r) for r in res]})
    (OUT / "kat_appendix_b.json").write_text(json.dumps(kats, indent=1) + "\n")

    # sine-driver goldens (Appendix B second table): real snddev_sine -> transform -> tee -> vumeter
    sine = []
    for rate, secs, gain in [(44100, 60, None), (44100, 60, (1, 2, [1])), (44100, 60, (1, 4, [3])),
                             (48000, 10, None), (48000, 10, (1, 4, [3])), (48000, 10, (1, 2, [3])),
                             (16000, 1, None), (16000, 1, (1, 1, [65535])),
                             (8000, 1, (1, 7, [5])), (96000, 1, (1, 1000, [999]))]:
        nbytes = rate * secs * 2
        src = ref.sine(rate, nbytes)
        out, results, rc = ref.pipeline(src, 1, gain, rate=rate, result_every=0)
        sine.append({
            "rate": rate, "secs": secs, "bytes": nbytes, "gain": None if gain is None else list(gain),
            "src_fnv1a64": f"{port.fnv1a64(src):016x}", "out_fnv1a64": f"{port.fnv1a64(out):016x}",
            "period": [int(x) for x in src.view(np.int16)[: rate // 1000]],
            "result": hexify(results[-1]),
        })
    # the driver refuses anything but mono (reference src/snddev_sine.c:172-173)
    sine.append({"rate": 48000, "stereo_open": ref.sine(48000, 64, channels=2) is not None})
    (OUT / "sine.json").write_text(json.dumps(sine, indent=1) + "\n")

    # seeded differential cases: small random inputs with every awkward shape, outputs stored whole
    rng = np.random.default_rng(0xC001)
    fuzz = []
    for i in range(48):
        ch = int(rng.integers(1, 17))
        nbytes = int(rng.integers(0, 700))
        kind = i % 4
        if kind == 0:
            x = rng.integers(-32768, 32768, size=(nbytes + 1) // 2, dtype=np.int64)
        elif kind == 1:
            x = rng.choice(np.array([-32768, -32767, -1, 0, 1, 32766, 32767]), size=(nbytes + 1) // 2)
        elif kind == 2:
            x = (rng.normal(0, 9000, size=(nbytes + 1) // 2)).clip(-32768, 32767).astype(np.int64)
        else:
            x = rng.integers(-40, 41, size=(nbytes + 1) // 2, dtype=np.int64)
        pcm = x.astype(np.int16).view(np.uint8)[:nbytes]
        scale = int(rng.choice([1, 2, 3, 7, 255, 256, 1000, 32767, 32768, 65535, int(rng.integers(1, 65536))]))
        gn = int(rng.choice([ch, 1, 2, 0]))
        gains = [int(v) for v in rng.integers(0, 65536, size=max(gn, 1))]
        if i % 5 == 0:
            gains = [int(min(65535, scale * k // 3)) for k in range(1, max(gn, 1) + 1)]
        gain = (gn, scale, gains[:gn] if gn else None)
        chunk = int(rng.choice([0, 1, 3, 5, 64, 1000]))
        pull = int(rng.choice([1024, 1024, 100, 4096, 62]))
        every = int(rng.choice([0, 1, 3]))
        out, results, rc = ref.pipeline(pcm, ch, gain, src_chunk=chunk, pull=pull, result_every=every)
        fuzz.append({
            "channels": ch, "in_bytes": [int(b) for b in pcm], "gain": [gain[0], gain[1], gain[2]],
            "src_chunk": chunk, "pull": pull, "result_every": every, "gain_rc": rc,
            "out_bytes": [int(b) for b in out], "results": [hexify(r) for r in results],
        })
    (OUT / "fuzz_pipeline.json").write_text(json.dumps(fuzz) + "\n")

    # The encoder-side sample-format stage: the reference's own src/enc_vorbis.c,
    # compiled unmodified against the libvorbis stand-in (oracle/Makefile target refenc), fed interleaved
    # S16 and asked for its float planes. Bit patterns, so that comparisons are exact.
    planar_file = OUT / "planar.json"
    enc = po.refenc()
    if enc is not None:
        rng = np.random.default_rng(20260109)
        planar = []
        edge = np.array([-32768, -32767, -1, 0, 1, 2, 3, 255, 256, 12345, -12345, 32766, 32767, 21845, -21845, 16384], np.int16)
        for ch, frames in [(1, 16), (2, 8), (1, 700), (2, 513), (3, 77), (4, 333), (6, 500), (8, 257), (16, 100), (5, 1)]:
            pcm = rng.integers(-32768, 32768, size=ch * frames).astype(np.int16)
            pcm[: min(edge.size, pcm.size)] = edge[: min(edge.size, pcm.size)]
            planes = enc.planes(pcm, ch)
            planar.append({"channels": ch, "frames": frames, "pcm": [int(v) for v in pcm],
                           "planes_u32": [[int(v) for v in planes[c].view(np.uint32)] for c in range(ch)]})
        planar_file.write_text(json.dumps(planar) + "\n")
    print("wrote", [p.name for p in sorted(OUT.glob("*.json"))])


if __name__ == "__main__":
    main()
