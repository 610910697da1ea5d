"""The kernel never divides: it multiplies by a per-(stream, channel) reciprocal (GainRow in
csrc/cmgpu_kernels.cuh). cmgpu_recipe_table() evaluates that exact integer recipe on the host
for all 65,536 inputs, so its equality with the reference arithmetic
(transform.c:110-123: trunc((int64)x*gain/scale), saturated) can be proven exhaustively here,
on CPU, before any GPU time is spent. No device work is involved."""
import ctypes as C

import numpy as np
import pytest

X = np.arange(-32768, 32768, dtype=np.int64)


def truth(g: int, d: int) -> np.ndarray:
    n = X * g
    q = np.abs(n) // d * np.sign(n)          # C division truncates toward zero
    return np.clip(q, -32768, 32767).astype(np.int16)


def table(cm, g: int, d: int) -> np.ndarray:
    out = np.empty(65536, dtype=np.int16)
    rc = cm.lib().cmgpu_recipe_table(g, d, out.ctypes.data_as(C.POINTER(C.c_int16)))
    assert rc == 0
    return out


def pairs():
    edge = [0, 1, 2, 3, 5, 7, 255, 256, 257, 999, 1000, 1001, 4095, 4096, 21845, 32767, 32768, 32769,
            43690, 65534, 65535]
    ps = {(g, d) for g in edge for d in edge if d}
    # the bench's gain table: scale = 1000 + s % 9000, gain = scale*3/4 + 37*((s+c) % 64)
    for s in range(0, 9000, 131):
        scale = 1000 + s
        for m in (0, 1, 31, 63):
            ps.add((scale * 3 // 4 + 37 * m, scale))
    rng = np.random.default_rng(7)
    for _ in range(400):
        ps.add((int(rng.integers(0, 65536)), int(rng.integers(1, 65536))))
    # ratios just below / at / above powers of two, where the reciprocal changes its pre-shift
    for d in (1, 3, 7, 1000, 4681, 65535):
        for t in range(0, 17):
            for dg in (-1, 0, 1):
                g = d * (1 << t) + dg
                if 0 <= g <= 65535:
                    ps.add((g, d))
    return sorted(ps)


def test_recipe_is_exact_for_all_inputs(cm):
    bad = []
    for g, d in pairs():
        if not np.array_equal(table(cm, g, d), truth(g, d)):
            bad.append((g, d))
    assert not bad, f"reciprocal recipe differs from trunc(x*g/d) for {bad[:10]} ({len(bad)} pairs)"


def test_recipe_matches_oracle_port(cm, port):
    x16 = X.astype(np.int16)
    for g, d in [(3, 4), (3, 2), (65535, 1), (1, 65535), (0, 9), (12345, 54321), (54321, 12345), (1, 3), (7, 7)]:
        out, rc = port.transform(x16, 1, (1, d, [g]))
        assert rc == 0
        assert np.array_equal(out.view(np.int16), table(cm, g, d))


def test_scale_zero_is_identity(cm):
    assert np.array_equal(table(cm, 123, 0), X.astype(np.int16))


@pytest.mark.parametrize("x,g,d,want", [(-3, 1, 2, -1), (-3, 3, 2, -4), (32767, 3, 2, 32767),
                                        (-32768, 3, 2, -32768), (-21845, 3, 2, -32767), (-1, 65535, 1, -32768),
                                        (1, 65535, 1, 32767), (-32768, 1, 1, -32768), (7, 1, 3, 2), (-7, 1, 3, -2)])
def test_recipe_known_answers(cm, x, g, d, want):
    # SURVEY.md Appendix B rows 2-6
    assert cm.lib().cmgpu_recipe_eval(g, d, x) == want
