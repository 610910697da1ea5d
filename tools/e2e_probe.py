#!/usr/bin/env python3
"""e2e_probe.py -- the end-to-end loop of bench.py (pinned host PCM up, tick, transformed PCM down, meter
snapshot per step) on ONE device and without torch / NCCL, to be started once per GPU:
    CUDA_VISIBLE_DEVICES=k python tools/e2e_probe.py [--ticks 40] [--frames 12000] [--distinct 40] [--steps 5]
--distinct = how many different pinned tick buffers the step cycles through (pinned footprint)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--frames", type=int, default=12000)
    ap.add_argument("--ticks", type=int, default=40)
    ap.add_argument("--distinct", type=int, default=40)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-snapshot", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--torch", action="store_true", help="import torch and create its CUDA context on the device first")
    ap.add_argument("--start-at", type=float, default=0.0, help="wall-clock time (time.time()) to start the timed steps at")
    a = ap.parse_args()
    if a.torch:
        import torch
        torch.cuda.set_device(a.device)
        torch.zeros(1, device="cuda")
    cm = load_package()
    eng = cm.Engine(a.channels, a.streams, a.frames, ring_slots=4, device=a.device, flags=cm.NO_PINNED)
    s = np.arange(a.streams)
    scale = (1000 + s % 9000).astype(np.uint16)
    gain = (scale[:, None].astype(np.int64) * 3 // 4 + 37 * ((s[:, None] + np.arange(a.channels)) % 64)).astype(np.uint16)
    eng.set_gain_table(scale, gain)
    pin_in = cm.PinnedArray((a.distinct, a.streams, eng.stride // 2))
    pin_out = cm.PinnedArray((a.distinct, a.streams, eng.stride // 2))
    rng = np.random.default_rng(1)
    pin_in.array[:] = rng.integers(-20000, 20000, size=pin_in.array.shape, dtype=np.int16)

    def step():
        for t in range(a.ticks):
            slot = t % 4
            eng.submit(slot, pin_in.array[t % a.distinct])
            eng.process(slot)
            eng.fetch(slot, pin_out.array[t % a.distinct])
        if not a.no_snapshot:
            eng.snapshot(reset=True)

    step(); eng.sync()
    while time.time() < a.start_at:
        time.sleep(0.001)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    eng.sync()
    dt = time.perf_counter() - t0
    nbytes = a.steps * a.ticks * a.streams * eng.stride
    print(f"ticks={a.ticks} frames={a.frames} distinct={a.distinct}: {dt / a.steps * 1e3:.2f} ms/step, "
          f"{nbytes / dt / 1e9:.1f} GB/s each way, {nbytes / 2 / dt / 1e6:.0f} Msamples/s", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
