"""Random schedules through the ring batch of the object shim, on CPU (csrc/host/*.c over tests/stub/cmgpu_stub.c).

tests/test_shim_host_logic.py drives the batch with two fixed patterns (tick-read-tick-read, and a producer that runs
ahead until it is refused). Here every step is drawn at random -- a tick, a read of a random length through a random
handle, a second handle obtained or dropped in mid-stream, a meter result -- over sources that deliver short reads and
"nothing now" (iohandle.h:41-53), and after every step the invariants of the contract are checked against the oracle:

  * every handle reads exactly the transform's output from the tick it was obtained at, in order, whatever the
    interleaving (the per-reader offsets of tee.c:167-206), and ends with eof() == 1;
  * tick() returns the frames it processed, and those are the bytes the fused vumeters report (coolmic_vumeter_read);
  * COOLMIC_ERROR_BUSY only ever comes with unread output pending (tee.c:145-151's back-pressure), never after every
    reader has caught up;
  * a result covers exactly the ticks since the previous one (vumeter.c:189-218), bit-identical doubles.
"""
import ctypes as C

import numpy as np
import pytest

from tests.shimlib import ShimLib
from tests.test_gpu_parity import same_result

BUSY = -12
READ_CB = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_size_t)
EOF_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)


class NativeResult(C.Structure):
    """coolmic_vumeter_result_t (include/coolmic_b200_shim.h, vumeter.h:48-83)."""
    _fields_ = [("rate", C.c_uint32), ("channels", C.c_uint), ("frames", C.c_size_t), ("global_peak", C.c_int16),
                ("global_power", C.c_double), ("channel_peak", C.c_int16 * 16), ("channel_power", C.c_double * 16)]

    def as_dict(self, rc):
        if rc != 0:
            return {"rc": rc}
        n = self.channels
        return {"rc": 0, "rate": int(self.rate), "channels": int(n), "frames": int(self.frames),
                "global_peak": int(self.global_peak), "global_power": float(self.global_power),
                "channel_peak": [int(self.channel_peak[c]) for c in range(n)],
                "channel_power": [float(self.channel_power[c]) for c in range(n)]}


@pytest.fixture(scope="module")
def api():
    L = ShimLib(stub=True).lib
    P = C.c_void_p
    L.coolmic_b200_batch_new_ring.restype = P
    L.coolmic_b200_batch_new_ring.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    L.coolmic_b200_batch_transform_new.restype = P
    L.coolmic_b200_batch_transform_new.argtypes = [P, C.c_char_p, P, C.c_uint32]
    L.coolmic_b200_batch_vumeter_new.restype = P
    L.coolmic_b200_batch_vumeter_new.argtypes = [P, P, C.c_char_p, P]
    L.coolmic_b200_batch_tick.argtypes = [P]
    L.coolmic_b200_batch_pending.restype = C.c_size_t
    L.coolmic_b200_batch_pending.argtypes = [P]
    L.coolmic_iohandle_new.restype = P
    L.coolmic_iohandle_new.argtypes = [C.c_char_p, P, P, P, READ_CB, EOF_CB]
    L.coolmic_iohandle_read.restype = C.c_ssize_t
    L.coolmic_iohandle_read.argtypes = [P, P, C.c_size_t]
    L.coolmic_iohandle_eof.argtypes = [P]
    L.coolmic_transform_attach_iohandle.argtypes = [P, P]
    L.coolmic_transform_get_iohandle.restype = P
    L.coolmic_transform_get_iohandle.argtypes = [P]
    L.coolmic_transform_set_master_gain.argtypes = [P, C.c_uint, C.c_uint16, C.POINTER(C.c_uint16)]
    L.coolmic_vumeter_read.restype = C.c_ssize_t
    L.coolmic_vumeter_read.argtypes = [P, C.c_ssize_t]
    L.coolmic_vumeter_result.argtypes = [P, C.POINTER(NativeResult)]
    L.coolmic_b200_unref.argtypes = [P]
    assert C.sizeof(NativeResult) == L.shimh_sizeof_result() == 192
    return L


class Source:
    """A memory iohandle that behaves like a capture device on a bad day: short reads, and now and then nothing."""

    def __init__(self, data: np.ndarray, seed: int, moody: bool):
        self.data, self.pos, self.rng, self.moody = data, 0, np.random.default_rng(seed), moody
        self.read_cb = READ_CB(self.read)
        self.eof_cb = EOF_CB(self.eof)

    def read(self, _ud, buf, length):
        left = self.data.size - self.pos
        if left == 0 or (self.moody and self.rng.random() < 0.15):
            return 0
        n = min(int(length), left, int(self.rng.integers(1, 600)) if self.moody else left)
        C.memmove(buf, self.data.ctypes.data + self.pos, n)
        self.pos += n
        return n

    def eof(self, _ud):
        return 1 if self.pos == self.data.size else 0


class Reader:
    def __init__(self, handle, stream, start):
        self.h, self.stream, self.pos = handle, stream, start


@pytest.mark.parametrize("seed", range(24))
def test_random_schedules_through_the_ring_batch(api, port, seed):
    L = api
    rng = np.random.default_rng(1000 + seed)
    channels = int(rng.integers(1, 5))
    fs = 2 * channels
    block = int(rng.integers(8, 160))
    slots = int(rng.integers(1, 5))
    threads = int(rng.integers(1, 4))
    n = int(rng.integers(2, 7))
    batch = L.coolmic_b200_batch_new_ring(0, channels, n, block, slots, threads)
    assert batch
    srcs, trs, vus, want, readers = [], [], [], [], []
    for s in range(n):
        nbytes = fs * int(rng.integers(block, 25 * block)) + int(rng.integers(0, fs))
        data = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
        scale = int(rng.integers(0, 65536)) if s else 0              # stream 0 in the reference's default state
        gain = rng.integers(0, 65536, size=channels).astype(np.uint16)
        out, rc = port.transform(data, channels, (channels, scale, gain.tolist()))
        assert rc == 0 and out.size == nbytes // fs * fs
        want.append(out)
        src = Source(data, 77 * seed + s, moody=bool(rng.integers(0, 2)))
        tr = L.coolmic_b200_batch_transform_new(batch, b"tr", None, 48000)
        assert tr
        assert L.coolmic_transform_set_master_gain(tr, channels, scale, gain.ctypes.data_as(C.POINTER(C.c_uint16))) == 0
        h = L.coolmic_iohandle_new(b"src", None, None, None, src.read_cb, src.eof_cb)
        assert L.coolmic_transform_attach_iohandle(tr, h) == 0
        L.coolmic_b200_unref(h)
        srcs.append(src); trs.append(tr)
        vus.append(L.coolmic_b200_batch_vumeter_new(batch, tr, b"vu", None))
        readers.append(Reader(L.coolmic_transform_get_iohandle(tr), s, 0))
    produced = [0] * n          # bytes of output the ticks so far have made, per stream
    window = [0] * n            # ... of which the last result() has covered this many
    scratch = np.zeros(4 * block * fs + 64, dtype=np.uint8)
    res = NativeResult()
    idle_ticks = 0

    def caught_up():
        return all(r.pos == produced[r.stream] for r in readers)

    def do_read(r, length):
        got = L.coolmic_iohandle_read(r.h, scratch.ctypes.data, length)
        assert got >= 0 and got % fs == 0 and got <= length
        assert r.pos + got <= produced[r.stream], "a reader got bytes no tick has produced"
        assert np.array_equal(scratch[:got], want[r.stream][r.pos: r.pos + got]), f"stream {r.stream} at {r.pos}"
        r.pos += got
        return got

    for step in range(20000):
        op = rng.random()
        if op < 0.30:
            rc = L.coolmic_b200_batch_tick(batch)
            if rc == BUSY:
                assert L.coolmic_b200_batch_pending(batch) > 0, "BUSY with nothing pending"
                assert not caught_up(), "BUSY although every reader has caught up"
                continue
            assert rc >= 0
            total = 0
            for s in range(n):
                d = L.coolmic_vumeter_read(vus[s], -1)
                assert d >= 0 and d % fs == 0
                produced[s] += d
                total += d // fs
            assert total == rc, "tick() and the fused vumeters disagree about the frames processed"
            done = all(src.pos == src.data.size for src in srcs)
            idle_ticks = idle_ticks + 1 if (rc == 0 and done) else 0
            if idle_ticks >= 2:
                break
        elif op < 0.80:
            r = readers[int(rng.integers(0, len(readers)))]
            do_read(r, int(rng.integers(1, scratch.size)))
        elif op < 0.86:
            s = int(rng.integers(0, n))
            if sum(1 for r in readers if r.stream == s) < 3:
                readers.append(Reader(L.coolmic_transform_get_iohandle(trs[s]), s, produced[s]))     # starts at the NEXT tick's output
        elif op < 0.90:
            extra = [i for i, r in enumerate(readers) if i >= n]
            if extra:
                r = readers.pop(extra[int(rng.integers(0, len(extra)))])
                L.coolmic_b200_unref(r.h)               # a lagging reader that goes away no longer holds the ring back
        else:
            s = int(rng.integers(0, n))
            rc = L.coolmic_vumeter_result(vus[s], C.byref(res))
            got = res.as_dict(rc)
            if produced[s] == window[s]:
                assert rc == -10                    # COOLMIC_ERROR_INVAL: nothing metered since the last result
            else:
                exp = port.vumeter(want[s][window[s]: produced[s]], channels)[-1]
                assert same_result(got, exp), f"stream {s} window [{window[s]}, {produced[s]})"
                window[s] = produced[s]
    else:
        pytest.fail("the schedule never drained its sources")
    # drain: every handle ends exactly at the end of its stream's output, then reports EOF
    for s in range(n):
        assert produced[s] == want[s].size, f"stream {s}: {produced[s]} of {want[s].size} bytes produced"
    for r in readers:
        while do_read(r, scratch.size):
            pass
        assert r.pos == want[r.stream].size
        assert L.coolmic_iohandle_eof(r.h) == 1
    assert L.coolmic_b200_batch_pending(batch) == 0
    assert L.coolmic_b200_batch_tick(batch) == 0
    for r in readers:
        L.coolmic_b200_unref(r.h)
    for s in range(n):
        L.coolmic_b200_unref(vus[s])
        L.coolmic_b200_unref(trs[s])
    L.coolmic_b200_unref(batch)
