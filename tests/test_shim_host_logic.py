"""The host shim's logic on CPU: csrc/host/*.c linked against tests/stub/cmgpu_stub.c (a stand-in for
the cmgpu_* engine built on the oracle port, with imitated download asynchrony) instead of the CUDA
library. Covers what a GPU is not needed for -- framing and carry, reader cursors, the ring batch's slot
reuse, back-pressure and threads -- against the oracle, and runs a fixed scenario under ASan + UBSan
and under TSan. The GPU versions of the same tests are in tests/test_gpu_shim.py.
"""
import subprocess

import numpy as np
import pytest

from tests.shimlib import ShimLib, build_stub
from tests.test_gpu_parity import same_result


@pytest.fixture(scope="module")
def stub():
    return ShimLib(stub=True)


@pytest.mark.parametrize("channels,block_frames,every,chunk", [(2, 256, 3, 0), (1, 1000, 1, 7), (8, 64, 0, 100), (3, 333, 2, 0)])
def test_batch_mode_objects_host_logic(stub, port, channels, block_frames, every, chunk):
    rng = np.random.default_rng(channels * 1000 + block_frames)
    n = 13
    nbytes = 2 * channels * 1500 + 6
    pcm = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    scale = rng.integers(0, 65536, size=n).astype(np.uint16)
    scale[0] = 0
    gain = rng.integers(0, 65536, size=(n, channels)).astype(np.uint16)
    ticks, outs, results, flags = stub.batch(pcm, channels, scale, gain, src_chunk=chunk, block_frames=block_frames,
                                             result_every_ticks=every, pull=1024)
    assert flags == 0
    whole = (nbytes // (2 * channels)) * 2 * channels
    assert ticks == -(-(whole // (2 * channels)) // block_frames)
    for s in range(n):
        want, rc = port.transform(pcm[s], channels, (channels, int(scale[s]), gain[s].tolist()))
        assert rc == 0 and np.array_equal(outs[s], want), f"stream {s}"
        step = (every or ticks) * block_frames * 2 * channels
        want_res = [port.vumeter(want[lo: lo + step], channels)[-1] for lo in range(0, whole, step)]
        got = [r for r in results[s] if r.get("rc", 0) == 0]
        assert len(got) == len(want_res)
        assert all(same_result(a, b) for a, b in zip(got, want_res))


@pytest.mark.parametrize("channels,block_frames,slots,threads,chunk", [(2, 256, 3, 1, 0), (1, 100, 4, 3, 7), (8, 64, 2, 2, 100),
                                                                      (5, 333, 3, 4, 0), (2, 1000, 1, 2, 0), (2, 50, 8, 5, 3)])
def test_ring_batch_host_logic(stub, port, channels, block_frames, slots, threads, chunk):
    rng = np.random.default_rng(channels * 77 + block_frames + slots)
    n = 29
    nbytes = 2 * channels * 2500 + 2 * channels - 1
    pcm = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    scale = rng.integers(0, 65536, size=n).astype(np.uint16)
    scale[0] = 0
    gain = rng.integers(0, 65536, size=(n, channels)).astype(np.uint16)
    ticks, outs, results, flags = stub.batch_ring(pcm, channels, scale, gain, src_chunk=chunk, block_frames=block_frames,
                                                  slots=slots, threads=threads, pull=777)
    assert flags == 0, f"ring semantics violated: flags {flags}"
    assert ticks == -(-(nbytes // (2 * channels)) // block_frames)
    for s in range(n):
        want, rc = port.transform(pcm[s], channels, (channels, int(scale[s]), gain[s].tolist()))
        assert rc == 0 and np.array_equal(outs[s], want), f"stream {s}"
        assert same_result(results[s], port.vumeter(want, channels)[-1]), f"stream {s}"


@pytest.mark.parametrize("sanitize", ["address,undefined", "thread"])
def test_host_shim_under_sanitizers(sanitize):
    exe = build_stub(sanitize, main=True)
    env = {"ASAN_OPTIONS": "detect_leaks=1:abort_on_error=0", "UBSAN_OPTIONS": "halt_on_error=1:print_stacktrace=1",
           "TSAN_OPTIONS": "halt_on_error=1"}
    proc = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert proc.returncode == 0 and "san_main: ok" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-4000:]
