"""ctypes binding for include/cmgpu.h. No arithmetic here: every call goes to the CUDA library,
and a missing library or a missing GPU is an error, never a fallback."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
import os
LIB_PATH = Path(os.environ.get("CMGPU_LIB", PKG / "lib" / "libcoolmic_b200.so"))
MAX_CH = 16

SEPARATE_OUT, NO_PINNED, FORCE_GENERIC, PLANAR_F32, MIX_OUTPUT_METER_ONLY = 0x1, 0x2, 0x4, 0x8, 0x10
TRANSFORM, METER, PLANAR = 0x1, 0x2, 0x4
FUSED = TRANSFORM | METER


class CmgpuError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: error {code}: {lib().cmgpu_last_error().decode(errors='replace')}")
        self.code = code


class MeterState(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64),
        ("power", C.c_int64 * MAX_CH),
        ("channel_peak", C.c_int16 * MAX_CH),
        ("global_peak", C.c_int16),
        ("reserved", C.c_int16 * 3),
    ]


class Result(C.Structure):
    _fields_ = [
        ("rate", C.c_uint32),
        ("channels", C.c_uint32),
        ("frames", C.c_uint64),
        ("global_peak", C.c_int16),
        ("global_power", C.c_double),
        ("channel_peak", C.c_int16 * MAX_CH),
        ("channel_power", C.c_double * MAX_CH),
    ]

    def as_dict(self) -> dict:
        n = self.channels
        return {
            "rc": 0, "rate": int(self.rate), "channels": int(n), "frames": int(self.frames),
            "global_peak": int(self.global_peak), "global_power": float(self.global_power),
            "channel_peak": [int(self.channel_peak[c]) for c in range(n)],
            "channel_power": [float(self.channel_power[c]) for c in range(n)],
        }


class Colors(C.Structure):
    _fields_ = [
        ("global_power_argb", C.c_uint32),
        ("global_peak_argb", C.c_uint32),
        ("channel_power_argb", C.c_uint32 * MAX_CH),
        ("channel_peak_argb", C.c_uint32 * MAX_CH),
        ("global_power_hue", C.c_double),
        ("channel_power_hue", C.c_double * MAX_CH),
    ]


COMM_ID_BYTES = 128
RESULTS_DEVICE_DB = 0x1

# every symbol include/cmgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U16P = C.POINTER(C.c_uint16)
SYMBOLS = {
    "cmgpu_version": (C.c_char_p, []),
    "cmgpu_device_count": (C.c_int, []),
    "cmgpu_last_error": (C.c_char_p, []),
    "cmgpu_host_alloc": (_P, [C.c_size_t]),
    "cmgpu_host_alloc_wc": (_P, [C.c_size_t]),
    "cmgpu_host_free": (None, [_P]),
    "cmgpu_ctx_create": (_P, [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]),
    "cmgpu_ctx_destroy": (None, [_P]),
    "cmgpu_channels": (C.c_uint, [_P]),
    "cmgpu_max_streams": (C.c_uint, [_P]),
    "cmgpu_ring_slots": (C.c_uint, [_P]),
    "cmgpu_block_frames": (C.c_uint, [_P]),
    "cmgpu_block_stride": (C.c_size_t, [_P]),
    "cmgpu_slot_bytes": (C.c_size_t, [_P]),
    "cmgpu_set_active_streams": (C.c_int, [_P, C.c_uint]),
    "cmgpu_stream_set_gain": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint16, _U16P]),
    "cmgpu_set_gain_table": (C.c_int, [_P, C.c_uint, C.c_uint, _U16P, _U16P]),
    "cmgpu_stream_get_gain": (C.c_int, [_P, C.c_uint, _U16P, _U16P]),
    "cmgpu_host_slot": (_P, [_P, C.c_uint]),
    "cmgpu_device_slot": (_P, [_P, C.c_uint]),
    "cmgpu_device_out_slot": (_P, [_P, C.c_uint]),
    "cmgpu_slot_set_frames": (C.c_int, [_P, C.c_uint, C.POINTER(C.c_uint32)]),
    "cmgpu_submit": (C.c_int, [_P, C.c_uint, _P]),
    "cmgpu_process": (C.c_int, [_P, C.c_uint, C.c_uint]),
    "cmgpu_process_cycle": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint]),
    "cmgpu_fetch": (C.c_int, [_P, C.c_uint, _P]),
    "cmgpu_device_planar_slot": (_P, [_P, C.c_uint]),
    "cmgpu_plane_stride": (C.c_size_t, [_P]),
    "cmgpu_fetch_planar": (C.c_int, [_P, C.c_uint, C.POINTER(C.c_float)]),
    "cmgpu_transfer_bytes": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cmgpu_sync": (C.c_int, [_P]),
    "cmgpu_slot_wait": (C.c_int, [_P, C.c_uint]),
    "cmgpu_meter_snapshot": (C.c_int, [_P, C.c_uint, C.c_uint, C.POINTER(MeterState), C.c_int]),
    "cmgpu_meter_reset": (C.c_int, [_P, C.c_uint, C.c_uint]),
    "cmgpu_meter_result": (C.c_int, [_P, C.c_uint, C.c_uint32, C.POINTER(Result)]),
    "cmgpu_finalise": (C.c_int, [C.POINTER(MeterState), C.c_uint32, C.c_uint, C.POINTER(Result)]),
    "cmgpu_device_meters": (_P, [_P]),
    "cmgpu_meter_row_u64": (C.c_uint, [_P]),
    "cmgpu_meter_decode": (C.c_int, [C.POINTER(C.c_uint64), C.c_uint, C.c_uint, C.POINTER(MeterState)]),
    "cmgpu_mix_ctx_create": (_P, [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]),
    "cmgpu_stream_set_mix": (C.c_int, [_P, C.c_uint, C.c_uint16, _U16P]),
    "cmgpu_mix_input_snapshot": (C.c_int, [_P, C.c_uint, C.c_uint, C.POINTER(MeterState), C.c_int]),
    "cmgpu_out_channels": (C.c_uint, [_P]),
    "cmgpu_out_block_stride": (C.c_size_t, [_P]),
    "cmgpu_host_out_slot": (_P, [_P, C.c_uint]),
    "cmgpu_time_process": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_float)]),
    "cmgpu_time_cycles": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_float)]),
    "cmgpu_time_single_tick": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cmgpu_link_probe": (C.c_int, [C.c_int, C.c_size_t, C.c_uint, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_float)]),
    "cmgpu_debug_violations": (C.c_int, []),
    "cmgpu_launch_count": (C.c_uint64, [_P]),
    "cmgpu_word_waits": (C.c_uint64, [_P]),
    "cmgpu_kernel_name": (C.c_char_p, [_P]),
    "cmgpu_meter_results": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint32, C.c_int, C.c_uint, C.POINTER(Result),
                                      C.POINTER(MeterState), C.POINTER(C.c_int)]),
    "cmgpu_comm_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "cmgpu_comm_create": (_P, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
    "cmgpu_comm_create_file": (_P, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "cmgpu_comm_adopt": (_P, [_P, C.c_int]),
    "cmgpu_comm_destroy": (None, [_P]),
    "cmgpu_comm_rank": (C.c_int, [_P]),
    "cmgpu_comm_size": (C.c_int, [_P]),
    "cmgpu_comm_nccl_version": (C.c_int, []),
    "cmgpu_comm_barrier": (C.c_int, [_P]),
    "cmgpu_comm_max": (C.c_int, [_P, C.POINTER(C.c_double), C.c_uint]),
    "cmgpu_comm_sum": (C.c_int, [_P, C.POINTER(C.c_double), C.c_uint]),
    "cmgpu_gather_results": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_int, C.POINTER(Result), C.POINTER(MeterState),
                                       C.POINTER(C.c_int), C.POINTER(C.c_uint)]),
    "cmgpu_comm_last_gather_ms": (C.c_float, [_P]),
    "cmgpu_meter_colors": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.POINTER(Colors)]),
    "cmgpu_tone_set_table": (C.c_int, [_P, C.POINTER(C.c_int16), C.c_uint]),
    "cmgpu_tone_fill": (C.c_int, [_P, C.c_uint, C.c_uint64, C.c_uint, C.c_uint, C.c_uint]),
    "cmgpu_noise_fill": (C.c_int, [_P, C.c_uint, C.c_uint64, C.c_uint, C.c_uint64, C.c_uint, C.c_uint]),
    "cmgpu_recipe_eval": (C.c_int, [C.c_uint16, C.c_uint16, C.c_int16]),
    "cmgpu_recipe_table": (C.c_int, [C.c_uint16, C.c_uint16, C.POINTER(C.c_int16)]),
}

_lib = None


def build_library(verbose: bool = False) -> Path:
    """Compile csrc/ for sm_100a into lib/libcoolmic_b200.so (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", str(PKG / "csrc")], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout, proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libcoolmic_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise CmgpuError(rc, what)


class Engine:
    """One cmgpu context: a device ring of [slots][streams][block_frames*channels] S16."""

    def __init__(self, channels: int, max_streams: int, block_frames: int, ring_slots: int = 1,
                 device: int = 0, flags: int = 0, out_channels: int = 0):
        self.L = lib()
        if out_channels:      # EXTENSION: downmix context (parity unpinned)
            self.ctx = self.L.cmgpu_mix_ctx_create(device, channels, out_channels, max_streams, ring_slots,
                                                   block_frames, flags)
        else:
            self.ctx = self.L.cmgpu_ctx_create(device, channels, max_streams, ring_slots, block_frames, flags)
        if not self.ctx:
            raise CmgpuError(-1, "cmgpu_ctx_create")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.out_stride = int(self.L.cmgpu_out_block_stride(self.ctx))
        self.max_streams = max_streams
        self.block_frames = block_frames
        self.ring_slots = ring_slots
        self.stride = int(self.L.cmgpu_block_stride(self.ctx))
        self.slot_bytes = int(self.L.cmgpu_slot_bytes(self.ctx))
        self.active = max_streams

    def close(self):
        if self.ctx:
            self.L.cmgpu_ctx_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def set_active(self, n: int):
        _check(self.L.cmgpu_set_active_streams(self.ctx, n), "cmgpu_set_active_streams")
        self.active = n

    def set_gain(self, stream: int, n: int, scale: int, gains) -> int:
        """Reference semantics (transform.c:195-222); returns the code instead of raising."""
        if gains is None:
            return self.L.cmgpu_stream_set_gain(self.ctx, stream, n, scale, None)
        arr = np.ascontiguousarray(gains, dtype=np.uint16)
        return self.L.cmgpu_stream_set_gain(self.ctx, stream, n, scale, arr.ctypes.data_as(_U16P))

    def set_gain_table(self, scale, gain, first: int = 0):
        s = np.ascontiguousarray(scale, dtype=np.uint16)
        g = np.ascontiguousarray(gain, dtype=np.uint16).reshape(s.size, self.channels)
        _check(self.L.cmgpu_set_gain_table(self.ctx, first, s.size, s.ctypes.data_as(_U16P),
                                           g.ctypes.data_as(_U16P)), "cmgpu_set_gain_table")

    def get_gain(self, stream: int):
        scale = C.c_uint16(0)
        g = (C.c_uint16 * MAX_CH)()
        _check(self.L.cmgpu_stream_get_gain(self.ctx, stream, C.byref(scale), g), "cmgpu_stream_get_gain")
        return int(scale.value), [int(g[c]) for c in range(self.channels)]

    def set_frames(self, slot: int, frames):
        if frames is None:
            _check(self.L.cmgpu_slot_set_frames(self.ctx, slot, None), "cmgpu_slot_set_frames")
            return
        f = np.ascontiguousarray(frames, dtype=np.uint32)
        assert f.size >= self.active
        _check(self.L.cmgpu_slot_set_frames(self.ctx, slot, f.ctypes.data_as(C.POINTER(C.c_uint32))),
               "cmgpu_slot_set_frames")

    # -- data path
    def host_slot(self, slot: int) -> np.ndarray:
        """The pinned staging slot as int16 [max_streams][stride/2]."""
        p = self.L.cmgpu_host_slot(self.ctx, slot)
        if not p:
            raise CmgpuError(-9, "cmgpu_host_slot")
        buf = (C.c_int16 * (self.slot_bytes // 2)).from_address(p)
        return np.frombuffer(buf, dtype=np.int16).reshape(self.max_streams, self.stride // 2)

    def host_out_slot(self, slot: int) -> np.ndarray:
        p = self.L.cmgpu_host_out_slot(self.ctx, slot)
        if not p:
            raise CmgpuError(-9, "cmgpu_host_out_slot")
        n = self.max_streams * self.out_stride // 2
        buf = (C.c_int16 * n).from_address(p)
        return np.frombuffer(buf, dtype=np.int16).reshape(self.max_streams, self.out_stride // 2)

    def set_mix(self, stream: int, scale: int, weights) -> int:
        w = np.ascontiguousarray(weights, dtype=np.uint16).reshape(self.out_channels, self.channels)
        return self.L.cmgpu_stream_set_mix(self.ctx, stream, scale, w.ctypes.data_as(_U16P))

    def input_snapshot(self, first: int = 0, count: int | None = None, reset: bool = False):
        count = self.active - first if count is None else count
        arr = (MeterState * count)()
        _check(self.L.cmgpu_mix_input_snapshot(self.ctx, first, count, arr, int(reset)), "cmgpu_mix_input_snapshot")
        return arr

    def submit(self, slot: int, host: np.ndarray | None = None):
        ptr = None if host is None else host.ctypes.data
        if host is not None:
            assert host.nbytes >= self.stride * self.active and host.flags.c_contiguous
        _check(self.L.cmgpu_submit(self.ctx, slot, ptr), "cmgpu_submit")

    def process(self, slot: int, flags: int = FUSED):
        _check(self.L.cmgpu_process(self.ctx, slot, flags), "cmgpu_process")

    def process_cycle(self, first_slot: int, n_slots: int, flags: int = FUSED):
        _check(self.L.cmgpu_process_cycle(self.ctx, first_slot, n_slots, flags), "cmgpu_process_cycle")

    def fetch(self, slot: int, host: np.ndarray | None = None):
        ptr = None if host is None else host.ctypes.data
        if host is not None:
            assert host.nbytes >= self.out_stride * self.active and host.flags.c_contiguous
        _check(self.L.cmgpu_fetch(self.ctx, slot, ptr), "cmgpu_fetch")

    def fetch_planar(self, slot: int) -> np.ndarray:
        """float32 [max_streams][channels][plane_stride] of the slot's last CMGPU_PLANAR tick (after sync())."""
        ps = int(self.L.cmgpu_plane_stride(self.ctx))
        out = np.zeros((self.max_streams, self.channels, ps), dtype=np.float32)
        _check(self.L.cmgpu_fetch_planar(self.ctx, slot, out.ctypes.data_as(C.POINTER(C.c_float))), "cmgpu_fetch_planar")
        return out

    def sync(self):
        _check(self.L.cmgpu_sync(self.ctx), "cmgpu_sync")

    def slot_wait(self, slot: int):
        _check(self.L.cmgpu_slot_wait(self.ctx, slot), "cmgpu_slot_wait")

    def device_slot(self, slot: int) -> int:
        return int(self.L.cmgpu_device_slot(self.ctx, slot))

    def device_out_slot(self, slot: int) -> int:
        return int(self.L.cmgpu_device_out_slot(self.ctx, slot))

    # -- meters
    def snapshot(self, first: int = 0, count: int | None = None, reset: bool = False):
        count = self.active - first if count is None else count
        arr = (MeterState * count)()
        _check(self.L.cmgpu_meter_snapshot(self.ctx, first, count, arr, int(reset)), "cmgpu_meter_snapshot")
        return arr

    def reset_meters(self, first: int = 0, count: int | None = None):
        count = self.active - first if count is None else count
        _check(self.L.cmgpu_meter_reset(self.ctx, first, count), "cmgpu_meter_reset")

    def result(self, stream: int, rate: int) -> dict:
        """coolmic_vumeter_result semantics: {'rc': -10} when nothing was metered."""
        res = Result()
        rc = self.L.cmgpu_meter_result(self.ctx, stream, rate, C.byref(res))
        if rc != 0:
            return {"rc": rc}
        return res.as_dict()

    @staticmethod
    def alloc_results(count: int):
        return (Result * count)(), (MeterState * count)(), (C.c_int * count)()

    def results(self, rate: int, first: int = 0, count: int | None = None, reset: bool = True, flags: int = 0, out=None):
        """coolmic_vumeter_result for many streams in one device round trip -> (Result[], MeterState[], rc[]).
        `out`: arrays from alloc_results() to fill instead of new ones."""
        count = self.active - first if count is None else count
        res, st, rcs = out if out is not None else self.alloc_results(count)
        _check(self.L.cmgpu_meter_results(self.ctx, first, count, rate, int(reset), flags, res, st, rcs),
               "cmgpu_meter_results")
        return res, st, rcs

    def colors(self, first: int = 0, count: int | None = None, alpha=1.0, saturation=1.0, value=1.0):
        count = self.active - first if count is None else count
        out = (Colors * count)()
        _check(self.L.cmgpu_meter_colors(self.ctx, first, count, alpha, saturation, value, out), "cmgpu_meter_colors")
        return out

    def tone_table(self, period):
        t = np.ascontiguousarray(period, dtype=np.int16)
        _check(self.L.cmgpu_tone_set_table(self.ctx, t.ctypes.data_as(C.POINTER(C.c_int16)), t.size), "cmgpu_tone_set_table")

    def tone_fill(self, slot: int, first_frame: int = 0, first_stream: int = 0, stream_step: int = 7, channel_step: int = 3):
        _check(self.L.cmgpu_tone_fill(self.ctx, slot, first_frame, first_stream, stream_step, channel_step), "cmgpu_tone_fill")

    def noise_fill(self, slot: int, first_frame: int = 0, first_stream: int = 0, seed: int = 0xC0011DC5,
                   every: int = 1, phase: int = 0):
        _check(self.L.cmgpu_noise_fill(self.ctx, slot, first_frame, first_stream, seed, every, phase), "cmgpu_noise_fill")

    def finalise(self, state: MeterState, rate: int, channels: int = 0) -> dict:
        res = Result()
        rc = self.L.cmgpu_finalise(C.byref(state), rate, channels or self.out_channels, C.byref(res))
        if rc != 0:
            return {"rc": rc}
        return res.as_dict()

    def device_meters(self) -> int:
        return int(self.L.cmgpu_device_meters(self.ctx))

    def meter_row_u64(self) -> int:
        return int(self.L.cmgpu_meter_row_u64(self.ctx))

    # -- measurement
    def time_process(self, reps: int, first_slot: int = 0, n_slots: int = 1, flags: int = FUSED) -> float:
        ms = C.c_float(0)
        _check(self.L.cmgpu_time_process(self.ctx, first_slot, n_slots, reps, flags, C.byref(ms)),
               "cmgpu_time_process")
        return float(ms.value)

    def time_cycles(self, cycles: int, first_slot: int = 0, n_slots: int = 1, flags: int = FUSED) -> float:
        ms = C.c_float(0)
        _check(self.L.cmgpu_time_cycles(self.ctx, first_slot, n_slots, cycles, flags, C.byref(ms)),
               "cmgpu_time_cycles")
        return float(ms.value)

    def transfer_bytes(self):
        """(host -> device, device -> host) PCM bytes cmgpu_submit / cmgpu_fetch have copied so far."""
        up, down = C.c_uint64(0), C.c_uint64(0)
        _check(self.L.cmgpu_transfer_bytes(self.ctx, C.byref(up), C.byref(down)), "cmgpu_transfer_bytes")
        return int(up.value), int(down.value)

    def time_single_tick(self, slot: int = 0, flags: int = FUSED, reps: int = 200):
        """(median, minimum) microseconds of cmgpu_process + cmgpu_sync on an idle context, timed in C."""
        med, mn = C.c_float(0), C.c_float(0)
        _check(self.L.cmgpu_time_single_tick(self.ctx, slot, flags, reps, C.byref(med), C.byref(mn)), "cmgpu_time_single_tick")
        return float(med.value), float(mn.value)

    def launch_count(self) -> int:
        return int(self.L.cmgpu_launch_count(self.ctx))

    def word_waits(self) -> int:
        return int(self.L.cmgpu_word_waits(self.ctx))

    def kernel_name(self) -> str:
        return self.L.cmgpu_kernel_name(self.ctx).decode()


class Comm:
    """NCCL communicator owned by the C library (cmgpu_comm_*): no torch, no MPI."""

    def __init__(self, device: int, rank: int, nranks: int, path: str | None = None, uid: bytes | None = None,
                 timeout_ms: int = 120000):
        self.L = lib()
        if uid is not None:
            buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(uid)
            self.comm = self.L.cmgpu_comm_create(device, rank, nranks, buf)
        else:
            self.comm = self.L.cmgpu_comm_create_file(device, rank, nranks, str(path).encode(), timeout_ms)
        if not self.comm:
            raise CmgpuError(-1, "cmgpu_comm_create")
        self.rank, self.size = rank, nranks

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * COMM_ID_BYTES)()
        _check(lib().cmgpu_comm_unique_id(buf), "cmgpu_comm_unique_id")
        return bytes(buf)

    def close(self):
        if self.comm:
            self.L.cmgpu_comm_destroy(self.comm)
            self.comm = None

    def barrier(self):
        _check(self.L.cmgpu_comm_barrier(self.comm), "cmgpu_comm_barrier")

    def max(self, *values: float):
        arr = (C.c_double * len(values))(*values)
        _check(self.L.cmgpu_comm_max(self.comm, arr, len(values)), "cmgpu_comm_max")
        return list(arr) if len(values) > 1 else arr[0]

    def sum(self, *values: float):
        arr = (C.c_double * len(values))(*values)
        _check(self.L.cmgpu_comm_sum(self.comm, arr, len(values)), "cmgpu_comm_sum")
        return list(arr) if len(values) > 1 else arr[0]

    @staticmethod
    def alloc_results(total_streams: int, nranks: int):
        return ((Result * total_streams)(), (MeterState * total_streams)(), (C.c_int * total_streams)(),
                (C.c_uint * nranks)())

    def gather_results(self, eng: "Engine", rate: int, total_streams: int, root: int = 0, reset: bool = True, out=None):
        """Collective. On the root: (Result[total], MeterState[total], rc[total], counts[nranks]); else None.
        `out`: arrays from alloc_results() to fill instead of new ones (root only)."""
        is_root = self.rank == root
        if is_root:
            res, st, rcs, counts = out if out is not None else self.alloc_results(total_streams, self.size)
        else:
            res = st = rcs = counts = None
        _check(self.L.cmgpu_gather_results(eng.ctx, self.comm, root, rate, int(reset), res, st, rcs, counts),
               "cmgpu_gather_results")
        return (res, st, rcs, list(counts)) if is_root else None

    def last_gather_ms(self) -> float:
        return float(self.L.cmgpu_comm_last_gather_ms(self.comm))


class PinnedArray:
    """int16 numpy view over page-locked host memory from cmgpu_host_alloc()."""

    def __init__(self, shape, wc: bool = False):
        n = int(np.prod(shape))
        self.ptr = (lib().cmgpu_host_alloc_wc if wc else lib().cmgpu_host_alloc)(n * 2)
        if not self.ptr:
            raise CmgpuError(-11, "cmgpu_host_alloc")
        buf = (C.c_int16 * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.int16).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().cmgpu_host_free(self.ptr)
            self.ptr = None


def link_probe(device: int = 0, nbytes: int = 256 << 20, reps: int = 8, both_only: bool = False, wc: bool = False) -> dict:
    """Host-link rates for pinned buffers (GB/s per direction): upload, download, both at once."""
    up, down, both = C.c_float(0), C.c_float(0), C.c_float(0)
    _check(lib().cmgpu_link_probe(device, nbytes, reps, int(wc), None if both_only else C.byref(up),
                                  None if both_only else C.byref(down), C.byref(both)), "cmgpu_link_probe")
    out = {"both_each_way_gbs": float(both.value)}
    if not both_only:
        out.update(h2d_gbs=float(up.value), d2h_gbs=float(down.value))
    return out


def state_dict(st: MeterState, channels: int) -> dict:
    return {
        "frames": int(st.frames),
        "power": [int(st.power[c]) for c in range(channels)],
        "channel_peak": [int(st.channel_peak[c]) for c in range(channels)],
        "global_peak": int(st.global_peak),
    }
