"""One rank of tests/test_gpu_post.py::test_nccl_gather_in_c (one process per GPU, like a launcher would
start them; no torch). Rank r owns the contiguous stream range sharding.stream_range(total, n, r), runs
two ticks over it -- the second with only part of its streams active, so ranks bring different counts
-- and takes part in cmgpu_gather_results; rank 0 compares everything that arrived with the oracle.

    python tests/gather_worker.py RANK NRANKS ID_FILE TOTAL_STREAMS CHANNELS OUT_JSON
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tests.conftest import load_package  # noqa: E402


def world_inputs(total, channels, block, seed):
    rng = np.random.default_rng(seed)
    pcm = rng.choice(np.array([-32768, -32767, -9, -1, 0, 1, 9, 32767], dtype=np.int16), size=(total, block * channels))
    scale = rng.integers(0, 65536, size=total).astype(np.uint16)
    gain = rng.integers(0, 65536, size=(total, channels)).astype(np.uint16)
    return pcm, scale, gain


def main():
    rank, nranks, path, total, channels, out_path = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]),
                                                     int(sys.argv[5]), sys.argv[6])
    cm = load_package()
    block, rate = 480, 44100
    lo, hi = cm.sharding.stream_range(total, nranks, rank)
    n = hi - lo
    comm = cm.Comm(rank, rank, nranks, path=path)
    eng = cm.Engine(channels, n, block, device=rank)
    ok, detail = True, ""
    try:
        for rnd in range(2):
            pcm, scale, gain = world_inputs(total, channels, block, 50 + rnd)
            eng.set_gain_table(scale[lo:hi], gain[lo:hi])
            # round 1: every rank keeps a different number of streams active
            active = [cm.sharding.stream_range(total, nranks, r)[1] - cm.sharding.stream_range(total, nranks, r)[0]
                      for r in range(nranks)]
            if rnd == 1:
                active = [a - (r + 1) * 3 for r, a in enumerate(active)]
            eng.set_active(active[rank])
            host = eng.host_slot(0)
            host[:, : block * channels] = pcm[lo:hi]
            for _ in range(1 + rnd):
                eng.submit(0)
                eng.process(0)
            comm.barrier()
            out = comm.gather_results(eng, rate, sum(active), root=0)
            if rank == 0:
                from oracle import pyoracle
                port = pyoracle.port()
                res, st, rcs, counts = out
                if counts != active:
                    ok, detail = False, f"round {rnd}: counts {counts} != {active}"
                i = 0
                for r in range(nranks):
                    rlo = cm.sharding.stream_range(total, nranks, r)[0]
                    for s in range(active[r]):
                        g = rlo + s
                        work = pcm[g:g + 1].copy()
                        meters = None
                        for _ in range(1 + rnd):
                            work = pcm[g:g + 1].copy()
                            meters, _ = port.batch(work, np.array([block], np.uint32), channels, scale[g:g + 1],
                                                   gain[g:g + 1], meters=meters)
                        # (the integer state first: the oracle's finaliser resets its meter, like vumeter.c:214-215)
                        same_state = all(int(st[i].power[c]) == int(meters[0].power[c]) for c in range(channels))
                        want = port.finalise(meters[0], rate, channels)
                        got = res[i].as_dict()
                        same = same_state and rcs[i] == 0 and got["frames"] == want["frames"] \
                            and got["global_peak"] == want["global_peak"] and got["channel_peak"] == want["channel_peak"] \
                            and np.float64(got["global_power"]).tobytes() == np.float64(want["global_power"]).tobytes() \
                            and all(np.float64(a).tobytes() == np.float64(b).tobytes()
                                    for a, b in zip(got["channel_power"], want["channel_power"]))
                        if not same and ok:
                            ok, detail = False, f"round {rnd}: rank {r} stream {s} (global {g}): {got} != {want}"
                        i += 1
            # the gather reset what it took: a second one finds nothing
            out = comm.gather_results(eng, rate, sum(active), root=0)
            if rank == 0 and not all(rc == -10 for rc in out[2]):
                ok, detail = False, f"round {rnd}: meters were not reset by the gather"
        mx = comm.max(float(rank))
        sm = comm.sum(1.0)
        if mx != nranks - 1 or sm != nranks:
            ok, detail = False, f"max/sum over ranks: {mx}, {sm}"
    finally:
        eng.close()
        comm.close()
    if not ok:
        print(f"gather_worker rank {rank}: {detail}", flush=True)
    if rank == 0:
        Path(out_path).write_text(json.dumps({"ok": ok, "detail": detail, "ranks": nranks, "total_streams": total,
                                              "streams_checked": total if ok else 0,
                                              "nccl": int(cm.lib().cmgpu_comm_nccl_version())}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
