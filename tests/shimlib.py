"""ctypes driver for tests/shim_harness.c: the product's coolmic_* objects behind the same
transform()/vumeter() call shape as oracle.pyoracle.RefLib, so tests read alike for both."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle.pyoracle import Result, _bytes, _res_dict, _u16

ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "libcoolmic-dsp_b200" / "lib"
SO = Path(__file__).resolve().parent / "_build" / "libshimh.so"


def build() -> Path:
    src = Path(__file__).resolve().parent / "shim_harness.c"
    if SO.exists() and SO.stat().st_mtime > max(src.stat().st_mtime, (LIBDIR / "libcoolmic_b200.so").stat().st_mtime):
        return SO
    SO.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-g", "-fPIC", "-Wall", "-Wextra", "-shared", "-o", str(SO), str(src),
                    "-L", str(LIBDIR), "-lcoolmic_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return SO


STUB_SRC = [ROOT / "tests" / "stub" / "cmgpu_stub.c", ROOT / "oracle" / "coolmic_oracle.c",
            Path(__file__).resolve().parent / "shim_harness.c"]
HOST_SRC = sorted((ROOT / "libcoolmic-dsp_b200" / "csrc" / "host").glob("*.c"))


def build_stub(sanitize: str = "", main: bool = False) -> Path:
    """The product's host shim (csrc/host/*.c) linked against tests/stub/cmgpu_stub.c -- a CPU stand-in
    for the cmgpu_* engine built on the oracle port -- so that the shim's HOST LOGIC runs without a GPU.
    `sanitize`: "", "address,undefined" or "thread"; `main`: the stand-alone scenario runner instead of
    a shared library (sanitizer runtimes do not like being loaded into python)."""
    tag = sanitize.replace(",", "_") or "plain"
    out = SO.parent / (f"shim_stub_{tag}" + ("" if main else ".so"))
    srcs = HOST_SRC + STUB_SRC + ([ROOT / "tests" / "stub" / "san_main.c"] if main else [])
    if out.exists() and out.stat().st_mtime > max(s.stat().st_mtime for s in srcs):
        return out
    SO.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-std=gnu11", "-O1", "-g", "-fPIC", "-Wall", "-Wextra", "-pthread", f"-I{ROOT / 'include'}"]
    if sanitize:
        cmd += [f"-fsanitize={sanitize}", "-fno-omit-frame-pointer"]
    if main:
        cmd += ["-DSHIM_BENCH_SELFTEST"]
    else:
        cmd += ["-shared"]
    subprocess.run(cmd + ["-o", str(out)] + [str(s) for s in srcs] + ["-lm"], check=True)
    return out


REFERENCE = Path("/root/reference")
DROPIN_REF_FILES = ["iohandle", "tee", "snddev", "snddev_sine", "snddev_null", "snddev_stdio", "logging", "coolmic-dsp"]


def build_dropin_stub():
    """The drop-in build of oracle/Makefile (`dropin`: the reference's own iohandle.c, tee.c, snddev*.c, logging.c and
    coolmic-dsp.c, unmodified and compiled from where they lie, src/transform.c and src/vumeter.c left out, the product's
    csrc/host/*.c with -DCOOLMIC_B200_WITH_IGLOO in their place, oracle/ref_harness.c on top) -- linked against the CPU
    stand-in of the engine instead of the CUDA library, so that the INTEGRATION build's host logic (libigloo object
    declarations, the reference's tee and iohandle driving the product's callbacks) runs without a GPU.
    Returns None where the reference's sources are not present (e.g. on the GPU box)."""
    if not (REFERENCE / "src" / "tee.c").exists():
        return None
    out = SO.parent / "libdropin_stub.so"
    oracle = ROOT / "oracle"
    srcs = ([REFERENCE / "src" / f"{f}.c" for f in DROPIN_REF_FILES] + HOST_SRC +
            [oracle / "ref_harness.c", ROOT / "tests" / "stub" / "cmgpu_stub.c", oracle / "coolmic_oracle.c"])
    if out.exists() and out.stat().st_mtime > max(s.stat().st_mtime for s in srcs):
        return out
    SO.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Wno-pointer-arith", "-O2", "-g", "-fPIC", "-D_GNU_SOURCE",
           "-DHAVE_SNDDRV_DRIVER_STDIO", "-DCOOLMIC_B200_WITH_IGLOO", f"-I{oracle / 'igloo_shim'}",
           f"-I{REFERENCE / 'include'}", f"-I{REFERENCE / 'src'}", f"-I{ROOT / 'include'}", "-shared", "-pthread",
           "-Wl,-Bsymbolic", "-o", str(out)] + [str(s) for s in srcs] + ["-lm"]
    subprocess.run(cmd, check=True)
    return out


class ShimLib:
    kind = "b200 shim"

    def __init__(self, stub: bool = False):
        self.lib = L = C.CDLL(str(build_stub() if stub else build()))
        L.shimh_sizeof_result.restype = C.c_uint
        L.shimh_null_checks.restype = C.c_int
        L.shimh_transform.restype = C.c_long
        L.shimh_transform.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint,
                                      C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_int)]
        L.shimh_vumeter.restype = C.c_long
        L.shimh_vumeter.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_long, C.c_uint,
                                    C.POINTER(Result), C.c_size_t]
        L.shimh_chain.restype = C.c_long
        L.shimh_chain.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint,
                                  C.POINTER(C.c_uint16), C.c_size_t, C.c_long, C.c_uint, C.POINTER(Result), C.c_size_t]

    def batch(self, pcm2d, channels, scale, gain, rate=48000, src_chunk=0, block_frames=256, result_every_ticks=0,
              pull=1024, cap=256):
        """n member transforms + fused meters on one batch engine. Returns (ticks, [out bytes per
        stream], [[results] per stream], fanout_flags)."""
        assert pcm2d.dtype == np.uint8 and pcm2d.ndim == 2 and pcm2d.flags.c_contiguous
        n, nbytes = pcm2d.shape
        out = np.zeros((n, nbytes), dtype=np.uint8)
        out_bytes = (C.c_size_t * n)()
        n_results = (C.c_size_t * n)()
        res = (Result * (n * cap))()
        mism = C.c_int(0)
        keep_s, sp = _u16(scale)
        keep_g, gp = _u16(gain)
        fn = self.lib.shimh_batch
        fn.restype = C.c_long
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.POINTER(C.c_uint16),
                       C.POINTER(C.c_uint16), C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_void_p,
                       C.POINTER(C.c_size_t), C.POINTER(Result), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        ticks = fn(pcm2d.ctypes.data, n, nbytes, rate, channels, sp, gp, src_chunk, block_frames, result_every_ticks,
                   pull, out.ctypes.data, out_bytes, res, cap, n_results, C.byref(mism))
        if ticks < 0:
            raise RuntimeError(f"shim batch harness failed: {ticks}")
        outs = [out[s, : out_bytes[s]].copy() for s in range(n)]
        results = [[_res_dict(res[s * cap + i]) for i in range(min(n_results[s], cap))] for s in range(n)]
        return int(ticks), outs, results, int(mism.value)

    def batch_ring(self, pcm2d, channels, scale, gain, rate=48000, src_chunk=0, block_frames=256, slots=3, threads=1,
                   pull=1024):
        """The ring batch with the producer running ahead of its readers. Returns (ticks, [out bytes per
        stream], [final result per stream], flags)."""
        assert pcm2d.dtype == np.uint8 and pcm2d.ndim == 2 and pcm2d.flags.c_contiguous
        n, nbytes = pcm2d.shape
        out = np.zeros((n, nbytes), dtype=np.uint8)
        out_bytes = (C.c_size_t * n)()
        res = (Result * n)()
        flags = C.c_int(0)
        keep_s, sp = _u16(scale)
        keep_g, gp = _u16(gain)
        fn = self.lib.shimh_batch_ring
        fn.restype = C.c_long
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.POINTER(C.c_uint16),
                       C.POINTER(C.c_uint16), C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_size_t, C.c_void_p,
                       C.POINTER(C.c_size_t), C.POINTER(Result), C.POINTER(C.c_int)]
        ticks = fn(pcm2d.ctypes.data, n, nbytes, rate, channels, sp, gp, src_chunk, block_frames, slots, threads, pull,
                   out.ctypes.data, out_bytes, res, C.byref(flags))
        if ticks < 0:
            raise RuntimeError(f"shim ring batch harness failed: {ticks}")
        outs = [out[s, : out_bytes[s]].copy() for s in range(n)]
        return int(ticks), outs, [_res_dict(res[s]) for s in range(n)], int(flags.value)

    def transform(self, pcm, channels, gain=None, rate=48000, src_chunk=0, pull=1024):
        src = _bytes(pcm)
        out = np.zeros(src.size + 64, dtype=np.uint8)
        rc = C.c_int(0)
        if gain is None:
            args = (0, 0, 0, None)
        else:
            keep, gp = _u16(gain[2])
            args = (1, gain[0], gain[1], gp)
        n = self.lib.shimh_transform(src.ctypes.data, src.size, rate, channels, *args, src_chunk, pull,
                                     out.ctypes.data, out.size, C.byref(rc))
        if n < 0:
            raise RuntimeError("shim transform could not be constructed")
        return out[:n].copy(), int(rc.value)

    def vumeter(self, pcm, channels, rate=48000, src_chunk=0, maxlen=-1, result_every=0, cap=4096):
        src = _bytes(pcm)
        res = (Result * cap)()
        n = self.lib.shimh_vumeter(src.ctypes.data, src.size, rate, channels, src_chunk, maxlen, result_every, res, cap)
        if n < 0:
            raise RuntimeError("shim vumeter could not be constructed")
        return [_res_dict(res[i]) for i in range(n)]

    def chain(self, pcm, channels, gain=None, rate=48000, src_chunk=0, maxlen=-1, result_every=0, cap=4096):
        src = _bytes(pcm)
        res = (Result * cap)()
        if gain is None:
            args = (0, 0, 0, None)
        else:
            keep, gp = _u16(gain[2])
            args = (1, gain[0], gain[1], gp)
        n = self.lib.shimh_chain(src.ctypes.data, src.size, rate, channels, *args, src_chunk, maxlen, result_every,
                                 res, cap)
        if n < 0:
            raise RuntimeError("shim chain could not be constructed")
        return [_res_dict(res[i]) for i in range(n)]
