"""N>1 host logic on CPU: world_size-2 gloo. Each rank meters its own stream range (here: the
oracle stands in for the kernel and the rows are encoded exactly as the kernel encodes them),
the raw rows are gathered to rank 0 with the same helper bench.py uses over NCCL, decoded through
the C ABI and compared with a single-process run over all streams."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_package


def test_stream_range_partitions_exactly():
    cm = load_package()
    for total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            ranges = [cm.sharding.stream_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cm.sharding.stream_range(8, 2, 2)


def test_link_proportional_shards_of_the_end_to_end_leg():
    """bench.py's end-to-end leg on N > 1 GPUs gives each rank a contiguous stream range whose size follows
    the rank's share of the host links: the counts must partition the total, stay multiples of the quantum
    (the last rank takes the remainder), never be empty, and follow the weights."""
    import bench
    rng = np.random.default_rng(11)
    for total in (65536, 4096, 1000, 64, 7):
        for world in (1, 2, 4, 8):
            for _ in range(20):
                w = rng.uniform(5.0, 50.0, size=world).tolist()
                counts = bench.proportional_shards(w, total)
                assert len(counts) == world and sum(counts) == total and min(counts) >= (1 if total >= world else 0)
                if total // 64 >= world:
                    assert all(c % 64 == 0 for c in counts[:-1]) and min(counts) >= 64
                    ideal = [x / sum(w) * total for x in w]
                    assert all(abs(c - i) <= 128 + total % 64 for c, i in zip(counts, ideal)), (w, counts)
    assert bench.proportional_shards([34.0, 44.2], 65536) == [28480, 37056]
    assert bench.proportional_shards([1.0] * 8, 65536) == [8192] * 8
    assert bench.proportional_shards([0.0, 10.0], 1024) == [64, 960]      # a dead link still gets a quantum, not a division by zero


def encode_row(meter, positions, channels):
    """What fused_tick leaves in a meter row: key = mag<<47 | ~pos<<1 | neg, then power, frames."""
    row = np.zeros(2 * channels + 2, dtype=np.uint64)
    for c in range(channels):
        v = int(meter.channel_peak[c])
        if v:
            row[c] = (abs(v) << 47) | (((~positions[c]) & ((1 << 46) - 1)) << 1) | (1 if v < 0 else 0)
        row[channels + c] = int(meter.power[c])
    row[2 * channels] = int(meter.frames)
    return row


def first_peak_frames(pcm, channels):
    """Frame index of the first sample with the largest magnitude, per channel (the position the
    kernel would encode)."""
    x = pcm.astype(np.int64).reshape(-1, channels)
    return [int(np.argmax(np.abs(x[:, c]))) for c in range(channels)]


def make_world(total, channels, frames, seed=3):
    rng = np.random.default_rng(seed)
    pcm = rng.choice(np.array([-32768, -9, -1, 0, 1, 9, 32767], dtype=np.int16), size=(total, frames * channels))
    scale = rng.integers(1, 65536, size=total).astype(np.uint16)
    gain = rng.integers(0, 65536, size=(total, channels)).astype(np.uint16)
    return pcm, scale, gain


def shard_ranges(cm, total, world, weights):
    """Equal contiguous ranges (sharding.stream_range), or the link-proportional ones of bench.py's
    end-to-end leg for per-rank `weights`."""
    if weights is None:
        return [cm.sharding.stream_range(total, world, r) for r in range(world)]
    import bench
    counts = bench.proportional_shards(weights, total, quantum=1)
    firsts = [sum(counts[:r]) for r in range(world)]
    return [(f, f + n) for f, n in zip(firsts, counts)]


def worker(rank, world, port_no, total, channels, frames, result_q, weights=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle
        cm = load_package()
        port = pyoracle.port()
        pcm, scale, gain = make_world(total, channels, frames)
        ranges = shard_ranges(cm, total, world, weights)
        lo, hi = ranges[rank]
        mine = pcm[lo:hi].copy()
        meters, _ = port.batch(mine, np.full(hi - lo, frames, np.uint32), channels, scale[lo:hi], gain[lo:hi])
        rows = np.concatenate([encode_row(meters[i], first_peak_frames(mine[i], channels), channels)
                               for i in range(hi - lo)]) if hi > lo else np.zeros(0, np.uint64)
        counts = [(b - a) * (2 * channels + 2) for a, b in ranges]
        got = cm.sharding.gather_rows(torch.from_numpy(rows.view(np.int64)), dist, rank, world, counts)
        if rank == 0:
            allrows = np.concatenate([t.numpy() for t in got])
            states = cm.sharding.decode_rows(cm.lib(), allrows, channels)
            out = []
            for s in range(total):
                res = cm.Result()
                import ctypes as C
                rc = cm.lib().cmgpu_finalise(C.byref(states[s]), 48000, channels, C.byref(res))
                out.append((rc, res.as_dict() if rc == 0 else None))
            result_q.put(out)
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total,channels,weights", [(9, 2, None), (16, 1, None), (5, 8, None), (23, 2, [34.0, 9.5])])
def test_two_rank_gather_equals_single_process(total, channels, weights, port):
    """weights: the ranks' stream ranges follow their (here: made-up) host-link shares, as in bench.py's
    end-to-end leg -- what arrives at rank 0 must not depend on how the streams were cut."""
    frames = 257
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port_no, total, channels, frames, q, weights)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pcm, scale, gain = make_world(total, channels, frames)
    ref = pcm.copy()
    meters, _ = port.batch(ref, np.full(total, frames, np.uint32), channels, scale, gain)
    for s in range(total):
        want = port.finalise(meters[s], 48000, channels)
        rc, got = out[s]
        assert rc == want["rc"]
        if rc == 0:
            assert got["frames"] == want["frames"]
            assert got["channel_peak"] == want["channel_peak"]
            assert got["global_peak"] == want["global_peak"]
            assert [np.float64(v).tobytes() for v in got["channel_power"]] == \
                   [np.float64(v).tobytes() for v in want["channel_power"]]
            assert np.float64(got["global_power"]).tobytes() == np.float64(want["global_power"]).tobytes()
