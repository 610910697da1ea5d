"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, not product code.

ctypes bindings for the two CPU checkers built by oracle/Makefile:

* ``ref``  -- oracle/_ref/libcoolmic_ref.so: the reference's own transform.c / vumeter.c /
  tee.c / iohandle.c / snddev*.c object code behind oracle/ref_harness.c. ``None`` when the
  prebuilt file is absent and /root/reference is not there to build it from.
* ``dropin`` -- oracle/_ref/libcoolmic_dropin.so: the same harness and the reference's own iohandle.c /
  tee.c / snddev*.c, but with the PRODUCT's transform + vumeter host shim compiled in where
  src/transform.c and src/vumeter.c were (oracle/Makefile). Same ``RefLib`` interface; every
  transform / vumeter read in it runs on the GPU through the product library. It is the thing
  being tested, not a checker. GPU tests only.
* ``refenc`` -- oracle/_ref/libcoolmic_refenc.so: the reference's own src/enc_vorbis.c (unmodified)
  behind oracle/ref_enc_harness.c and a header-only libvorbis stand-in: its S16 -> planar float
  stage (enc_vorbis.c:76-122) as the oracle of the product's float planes. ``None`` when absent.
* ``port`` -- oracle/_build/libcoolmic_port.so: our plain-C restatement (coolmic_oracle.c).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libcoolmic_ref.so"
DROPIN_SO = HERE / "_ref" / "libcoolmic_dropin.so"
REFENC_SO = HERE / "_ref" / "libcoolmic_refenc.so"
PORT_SO = HERE / "_build" / "libcoolmic_port.so"
MAX_CH = 16


class Result(C.Structure):
    """Flat mirror of coolmic_vumeter_result_t (reference include/coolmic-dsp/vumeter.h:48-83)."""

    _fields_ = [
        ("rc", C.c_int32),
        ("rate", C.c_uint32),
        ("channels", C.c_uint32),
        ("global_peak", C.c_int32),
        ("frames", C.c_uint64),
        ("global_power", C.c_double),
        ("channel_peak", C.c_int32 * MAX_CH),
        ("channel_power", C.c_double * MAX_CH),
    ]

    def as_dict(self) -> dict:
        n = self.channels
        return {
            "rc": int(self.rc),
            "rate": int(self.rate),
            "channels": int(n),
            "frames": int(self.frames),
            "global_peak": int(self.global_peak),
            "global_power": float(self.global_power),
            "channel_peak": [int(self.channel_peak[c]) for c in range(n)],
            "channel_power": [float(self.channel_power[c]) for c in range(n)],
        }


class Meter(C.Structure):
    """oracle_meter_t: integer meter state between reset and result."""

    _fields_ = [
        ("power", C.c_int64 * MAX_CH),
        ("channel_peak", C.c_int16 * MAX_CH),
        ("global_peak", C.c_int16),
        ("frames", C.c_uint64),
    ]


def build(verbose: bool = False) -> None:
    """Compile the checkers (idempotent). Building the checker is not using it."""
    proc = subprocess.run(["make", "-C", str(HERE), "all"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout, proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("oracle build failed")


def _u16(a):
    if a is None:
        return None, None
    arr = np.ascontiguousarray(a, dtype=np.uint16)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_uint16))


def _bytes(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


class _Lib:
    def __init__(self, path: Path, prefix: str):
        self.path = path
        self.lib = C.CDLL(str(path))
        self.prefix = prefix

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)


class RefLib(_Lib):
    """The reference's own object code."""

    kind = "reference"

    def __init__(self, path: Path = REF_SO):
        super().__init__(path, "refh_")
        L = self.lib
        L.refh_sizeof_result.restype = C.c_uint
        L.refh_sine.restype = C.c_long
        L.refh_sine.argtypes = [C.c_uint, C.c_uint, C.c_size_t, C.c_size_t, C.c_void_p]
        L.refh_transform.restype = C.c_long
        L.refh_transform.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint,
                                     C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_int)]
        L.refh_vumeter.restype = C.c_long
        L.refh_vumeter.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_long, C.c_uint,
                                   C.POINTER(Result), C.c_size_t]
        L.refh_pipeline.restype = C.c_long
        L.refh_pipeline.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint,
                                    C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_uint,
                                    C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                    C.POINTER(Result), C.c_size_t, C.POINTER(C.c_int)]
        L.refh_bench.restype = C.c_double
        L.refh_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint,
                                 C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.c_size_t, C.c_uint, C.c_uint,
                                 C.POINTER(Result)]

    def sizeof_result(self) -> int:
        return int(self.lib.refh_sizeof_result())

    def sine(self, rate: int, nbytes: int, channels: int = 1, pull: int = 1024):
        """snddev_sine through the real driver; None if the driver refuses (e.g. stereo)."""
        out = np.zeros(nbytes, dtype=np.uint8)
        n = self.lib.refh_sine(rate, channels, nbytes, pull, out.ctypes.data)
        if n < 0:
            return None
        return out[:n].copy()

    def transform(self, pcm, channels, gain=None, rate=48000, src_chunk=0, pull=1024):
        """gain = None (setter never called) or (n, scale, [gains]) / (0, 0, None)."""
        src = _bytes(pcm)
        out = np.zeros(src.size + 64, dtype=np.uint8)
        rc = C.c_int(0)
        if gain is None:
            n = self.lib.refh_transform(src.ctypes.data, src.size, rate, channels, 0, 0, 0, None,
                                        src_chunk, pull, out.ctypes.data, out.size, C.byref(rc))
        else:
            gn, scale, gains = gain
            keep, gp = _u16(gains)
            n = self.lib.refh_transform(src.ctypes.data, src.size, rate, channels, 1, gn, scale, gp,
                                        src_chunk, pull, out.ctypes.data, out.size, C.byref(rc))
        if n < 0:
            raise RuntimeError("reference transform could not be constructed")
        return out[:n].copy(), int(rc.value)

    def vumeter(self, pcm, channels, rate=48000, src_chunk=0, maxlen=-1, result_every=0, cap=4096):
        src = _bytes(pcm)
        res = (Result * cap)()
        n = self.lib.refh_vumeter(src.ctypes.data, src.size, rate, channels, src_chunk, maxlen, result_every,
                                  res, cap)
        if n < 0:
            raise RuntimeError("reference vumeter could not be constructed")
        return [_res_dict(res[i]) for i in range(n)]

    def pipeline(self, pcm, channels, gain=None, rate=48000, src_chunk=0, pull=1024, result_every=20,
                 cap=65536):
        """source -> transform -> tee -> {consumer, vumeter}, the reference wiring."""
        src = _bytes(pcm)
        out = np.zeros(src.size + 64, dtype=np.uint8)
        res = (Result * cap)()
        rc = C.c_int(0)
        nout = C.c_size_t(0)
        if gain is None:
            args = (0, 0, 0, None)
        else:
            keep, gp = _u16(gain[2])
            args = (1, gain[0], gain[1], gp)
        n = self.lib.refh_pipeline(src.ctypes.data, src.size, rate, channels, *args, src_chunk, pull,
                                   result_every, out.ctypes.data, out.size, C.byref(nout), res, cap,
                                   C.byref(rc))
        if n < 0:
            raise RuntimeError("reference pipeline could not be constructed")
        return out[: nout.value].copy(), [_res_dict(res[i]) for i in range(min(n, cap))], int(rc.value)

    def bench(self, pcm2d: np.ndarray, channels, scale, gain, rate=48000, pull=1024, result_every=20,
              threads=1, want_out=False):
        """Many streams over `threads` pthreads through the reference wiring; returns
        (seconds, out or None, [last result per stream])."""
        assert pcm2d.dtype == np.int16 and pcm2d.ndim == 2 and pcm2d.flags.c_contiguous
        n_streams, samples = pcm2d.shape
        out = np.zeros_like(pcm2d) if want_out else None
        keep_s, sp = _u16(scale)
        keep_g, gp = _u16(gain)
        last = (Result * n_streams)()
        sec = self.lib.refh_bench(pcm2d.ctypes.data, out.ctypes.data if want_out else None, n_streams,
                                  samples * 2, rate, channels, sp, gp, pull, result_every, threads, last)
        if sec < 0:
            raise RuntimeError("reference bench failed")
        return sec, out, [_res_dict(last[i]) for i in range(n_streams)]


class PortLib(_Lib):
    """Our plain-C restatement."""

    kind = "port"

    def __init__(self, path: Path = PORT_SO):
        super().__init__(path, "oracle_")
        L = self.lib
        L.oracle_run_transform.restype = C.c_long
        L.oracle_run_transform.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_uint, C.c_uint,
                                           C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_void_p,
                                           C.c_size_t, C.POINTER(C.c_int)]
        L.oracle_run_vumeter.restype = C.c_long
        L.oracle_run_vumeter.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint, C.c_size_t, C.c_long,
                                         C.c_uint, C.POINTER(Result), C.c_size_t]
        L.oracle_batch.restype = None
        L.oracle_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), C.c_uint,
                                   C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.POINTER(Meter)]
        L.oracle_batch_threads.restype = C.c_double
        L.oracle_batch_threads.argtypes = L.oracle_batch.argtypes + [C.c_uint]
        L.oracle_meter_finalise.restype = C.c_int
        L.oracle_meter_finalise.argtypes = [C.POINTER(Meter), C.c_uint32, C.c_uint, C.POINTER(Result)]
        L.oracle_gain_adapt.restype = C.c_int
        L.oracle_gain_adapt.argtypes = [C.c_uint, C.c_uint, C.c_uint16, C.POINTER(C.c_uint16),
                                        C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
        L.oracle_mix_process.restype = None
        L.oracle_mix_process.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_uint16, C.POINTER(C.c_uint16),
                                         C.c_void_p, C.POINTER(Meter), C.POINTER(Meter)]
        L.oracle_fnv1a64.restype = C.c_uint64
        L.oracle_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]

    def transform(self, pcm, channels, gain=None, rate=48000, src_chunk=0, pull=1024):
        src = _bytes(pcm)
        out = np.zeros(src.size + 64, dtype=np.uint8)
        rc = C.c_int(0)
        if gain is None:
            args = (0, 0, 0, None)
        else:
            keep, gp = _u16(gain[2])
            args = (1, gain[0], gain[1], gp)
        n = self.lib.oracle_run_transform(src.ctypes.data, src.size, channels, *args, src_chunk, pull,
                                          out.ctypes.data, out.size, C.byref(rc))
        if n < 0:
            raise RuntimeError("oracle transform could not be constructed")
        return out[:n].copy(), int(rc.value)

    def vumeter(self, pcm, channels, rate=48000, src_chunk=0, maxlen=-1, result_every=0, cap=4096):
        src = _bytes(pcm)
        res = (Result * cap)()
        n = self.lib.oracle_run_vumeter(src.ctypes.data, src.size, rate, channels, src_chunk, maxlen,
                                        result_every, res, cap)
        if n < 0:
            raise RuntimeError("oracle vumeter could not be constructed")
        return [_res_dict(res[i]) for i in range(n)]

    def gain_adapt(self, stream_channels, n, scale, gains, state=None):
        """Returns (rc, scale, [gain per stream channel]) starting from `state` (scale, gains)."""
        st_scale = C.c_uint16(state[0] if state else 0)
        st_gain = (C.c_uint16 * MAX_CH)(*(state[1] if state else []))
        keep, gp = _u16(gains)
        rc = self.lib.oracle_gain_adapt(stream_channels, n, scale, gp, C.byref(st_scale), st_gain)
        return int(rc), int(st_scale.value), [int(st_gain[c]) for c in range(stream_channels)]

    def batch(self, pcm2d: np.ndarray, frames, channels, scale, gain, meters=None, threads=0):
        """In-place gain + metering of [n_streams][stride] S16. Returns (meters, seconds|None)."""
        assert pcm2d.dtype == np.int16 and pcm2d.ndim == 2 and pcm2d.flags.c_contiguous
        n_streams, stride = pcm2d.shape
        fr = np.ascontiguousarray(frames, dtype=np.uint32)
        keep_s, sp = _u16(scale)
        keep_g, gp = _u16(gain)
        if meters is None:
            meters = (Meter * n_streams)()
        frp = fr.ctypes.data_as(C.POINTER(C.c_uint32))
        if threads:
            sec = self.lib.oracle_batch_threads(pcm2d.ctypes.data, n_streams, stride, frp, channels, sp, gp,
                                                meters, threads)
            return meters, sec
        self.lib.oracle_batch(pcm2d.ctypes.data, n_streams, stride, frp, channels, sp, gp, meters)
        return meters, None

    def mix(self, pcm: np.ndarray, frames: int, cin: int, cout: int, scale: int, weights, meter_in=None, meter_out=None):
        """EXTENSION checker (parity unpinned): N -> M downmix of one stream; returns int16 [frames*cout]."""
        src = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.zeros(frames * cout, dtype=np.int16)
        keep, wp = _u16(np.asarray(weights).reshape(-1))
        self.lib.oracle_mix_process(src.ctypes.data, frames, cin, cout, scale, wp, out.ctypes.data,
                                    C.byref(meter_in) if meter_in is not None else None,
                                    C.byref(meter_out) if meter_out is not None else None)
        return out

    def planar(self, pcm, channels: int) -> np.ndarray:
        """enc_vorbis.c:108-117: interleaved int16 -> float32 [channels][frames] = sample / 32768.f."""
        src = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        frames = src.size // channels
        out = np.zeros((channels, max(frames, 1)), dtype=np.float32)
        self.lib.oracle_planar.restype = None
        self.lib.oracle_planar.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t]
        self.lib.oracle_planar(src.ctypes.data, frames, channels, out.ctypes.data, out.shape[1])
        return out[:, :frames]

    def finalise(self, meter: Meter, rate: int, channels: int) -> dict:
        res = Result()
        rc = self.lib.oracle_meter_finalise(C.byref(meter), rate, channels, C.byref(res))
        d = _res_dict(res)
        d["rc"] = int(rc)
        return d

    def fnv1a64(self, data) -> int:
        b = _bytes(data)
        return int(self.lib.oracle_fnv1a64(b.ctypes.data, b.size))


def _res_dict(r: Result) -> dict:
    d = r.as_dict()
    if d["rc"] != 0:
        return {"rc": d["rc"]}
    return d


def fnv1a64(data) -> int:
    """FNV-1a 64 in numpy-free Python for small inputs / as a cross-check of the C one."""
    h = 1469598103934665603
    for b in bytes(_bytes(data)):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


_ref = None
_port = None
_dropin = None


def dropin():
    """The reference pipeline with the product's transform + vumeter dropped in, or None when
    the prebuilt library is absent and cannot be built here."""
    global _dropin
    if _dropin is None:
        if not DROPIN_SO.exists() and os.path.isdir("/root/reference/src"):
            build()
        if DROPIN_SO.exists():
            _dropin = RefLib(DROPIN_SO)
            _dropin.kind = "dropin"
    return _dropin


def port() -> PortLib:
    global _port
    if _port is None:
        if not PORT_SO.exists():
            build()
        _port = PortLib()
    return _port


def ref():
    """The reference's own code, or None when neither the prebuilt library nor
    /root/reference is available (e.g. on a GPU box without the prebuilt file)."""
    global _ref
    if _ref is None:
        if not REF_SO.exists() and os.path.isdir("/root/reference/src"):
            build()
        if REF_SO.exists():
            _ref = RefLib()
    return _ref


class RefEncLib:
    """The reference's own enc_vorbis.c sample-format stage (interleaved S16 -> planar float / 32768.f)."""

    kind = "reference enc_vorbis.c"

    def __init__(self, path: Path = REFENC_SO):
        self.lib = C.CDLL(str(path))
        self.lib.refenc_vorbis_planes.restype = C.c_long
        self.lib.refenc_vorbis_planes.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t]

    def planes(self, pcm, channels: int) -> np.ndarray:
        """int16 [frames * channels] interleaved -> float32 [channels][frames], as enc_vorbis.c writes them."""
        src = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        frames = src.size // channels
        out = np.full((channels, max(frames, 1)), np.nan, dtype=np.float32)
        n = self.lib.refenc_vorbis_planes(src.ctypes.data, src.size * 2, channels, out.ctypes.data, out.shape[1])
        if n != frames:
            raise RuntimeError(f"refenc_vorbis_planes: {n} frames, expected {frames}")
        return out[:, :frames]


_refenc = None


def refenc():
    """RefEncLib or None (no prebuilt library and no /root/reference to build it from)."""
    global _refenc
    if _refenc is None:
        if not REFENC_SO.exists():
            try:
                build()
            except Exception:
                pass
        if REFENC_SO.exists():
            _refenc = RefEncLib()
    return _refenc
