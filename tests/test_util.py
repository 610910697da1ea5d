"""coolmic_util_power2hue / peak2hue / ahsv2argb (reference src/util.c:59-139), the
colours the app paints its meter with. Host-side double arithmetic: the product's functions must
return the same doubles (bit patterns) and ARGB words as the reference's own object code
(oracle/_ref, when present) and as the committed fixtures generated from it. No GPU needed."""
import ctypes as C
import json
import math
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "util.json"


def bind(lib):
    lib.coolmic_util_ahsv2argb.restype = C.c_uint32
    lib.coolmic_util_ahsv2argb.argtypes = [C.c_double] * 4
    lib.coolmic_util_power2hue.restype = C.c_double
    lib.coolmic_util_power2hue.argtypes = [C.c_double, C.c_char_p]
    lib.coolmic_util_peak2hue.restype = C.c_double
    lib.coolmic_util_peak2hue.argtypes = [C.c_int16, C.c_char_p]
    return lib


def cases():
    rng = np.random.default_rng(4)
    powers = [-1e9, -96.3, -20.000001, -20.0, -19.999, -10.0, -6.0206, -3.0107752652057771, -1e-9, 0.0, 1.5,
              float("-inf")] + [float(x) for x in rng.uniform(-25, 1, 40)]
    peaks = [-32768, -32767, -30001, -30000, -28001, -28000, -1, 0, 1, 27999, 28000, 28001, 30000, 30001, 32766, 32767]
    hsv = [(1.0, h, 1.0, 1.0) for h in np.linspace(0, 2 * math.pi, 25)] + \
          [(0.5, 0.43, 0.8, 0.9), (1.0, 1.0, 1.0, 1.0), (2.0, math.pi * 2 / 3, 1.0, 0.5), (-1.0, 0.0, 0.3, 2.0),
           (1.0, 7.0, 1.0, 1.0), (1.0, -0.5, 1.0, 1.0)] + [tuple(float(x) for x in rng.uniform(0, 1, 4) * (1, 6.2, 1, 1)) for _ in range(40)]
    return powers, peaks, hsv


def run(lib):
    powers, peaks, hsv = cases()
    out = {"power2hue": [], "peak2hue": [], "ahsv2argb": [], "other_profile": []}
    for p in powers:
        out["power2hue"].append(float(lib.coolmic_util_power2hue(p, b"default")).hex())
    for k in peaks:
        out["peak2hue"].append(float(lib.coolmic_util_peak2hue(k, b"default")).hex())
    for a, h, s, v in hsv:
        out["ahsv2argb"].append(int(lib.coolmic_util_ahsv2argb(a, h, s, v)))
    out["other_profile"] = [float(lib.coolmic_util_power2hue(-5.0, b"other")).hex(),
                            float(lib.coolmic_util_peak2hue(100, b"other")).hex()]
    # the whole chain as the app uses it: result -> hue -> colour
    out["chain"] = [int(lib.coolmic_util_ahsv2argb(1.0, lib.coolmic_util_power2hue(p, b"default"), 1.0, 1.0)) for p in powers] + \
                   [int(lib.coolmic_util_ahsv2argb(1.0, lib.coolmic_util_peak2hue(k, b"default"), 1.0, 1.0)) for k in peaks]
    return out


def test_util_matches_committed_fixtures(cm):
    got = run(bind(cm.lib()))
    want = json.loads(GOLD.read_text())
    assert got == want


def test_util_matches_the_references_object_code(cm, ref):
    assert run(bind(cm.lib())) == run(bind(ref.lib))


def test_known_colours(cm):
    lib = bind(cm.lib())
    assert lib.coolmic_util_ahsv2argb(1.0, 0.0, 1.0, 1.0) == 0xFFFF0000                      # red at 0 dB / full scale
    assert lib.coolmic_util_ahsv2argb(1.0, lib.coolmic_util_power2hue(-30.0, b"default"), 1.0, 1.0) == 0xFF00FF18    # the reference takes the fraction of hue itself, hence the blue tinge
    assert lib.coolmic_util_peak2hue(32767, b"default") == 0.0 and lib.coolmic_util_peak2hue(-32768, b"default") == 0.0


if __name__ == "__main__":          # regenerate the fixtures from the reference's object code
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle import pyoracle
    r = pyoracle.ref()
    assert r is not None, "oracle/_ref is not available"
    GOLD.write_text(json.dumps(run(bind(r.lib)), indent=1) + "\n")
    print("wrote", GOLD)
