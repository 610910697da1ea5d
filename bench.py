#!/usr/bin/env python3
"""bench.py -- aggregate Msamples/s through the fused transform+vumeter tick (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (default `cfg5`, BASELINE.json configs[4], the configuration north_star states both targets
on): 65,536 independent 48 kHz stereo S16 streams x 1 s IN TOTAL, sharded by contiguous stream range
across the N GPUs (strong scaling; all of it fits one GPU: 12.6 GB in + 12.6 GB out), resident in one
device ring per GPU; one step = one fused tick (ONE kernel launch per GPU) over the rank's share =
6.29e9 samples and 25.2 GB of algorithmic traffic in total (2 B read + 2 B written per sample), far
larger than the 126 MB L2. No data-path collective; the per-stream meter results are gathered to rank
0 over NCCL by the C library (cmgpu_gather_results) outside the kernel-timed region and inside the
end-to-end one. No PyTorch anywhere: under torchrun this script only reads RANK / LOCAL_RANK /
WORLD_SIZE; the communicator, barriers and max-over-ranks are the library's own NCCL calls.

`value`  device-timed (CUDA events on the engine's compute stream, max over ranks), inputs in HBM.
`e2e`    the same work through the C ABI with pinned HOST buffers: per step every tick is uploaded,
         processed and downloaded through a 4-slot ring on three CUDA streams, then the meter results
         are gathered (NCCL at N>1) and finalised; wall clock, max over ranks. A second leg runs the
         reference-named object API (coolmic_b200_batch_tick over member transforms fed by memory
         iohandles) on the cfg2 shape; `link` is what the host link gives when every rank drives it.
`roofline` algorithmic bytes / measured kernel time against MEASURED_PEAKS.json's HBM copy rate: the
         burst figure of the timed steps, a >= 2 s sustained run, and one slot in place without overlap.
`cpu_baseline` the reference's own transform.c/tee.c/vumeter.c object code (oracle/_ref) on this
         box's host cores over a bounded sample of the same workload.

`--impl reference` times only that CPU reference (all host threads) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # device-timed run: `ticks` ticks of `frames` frames over a ring of `ring` slots per step
    # (graph=True: the ring's ticks issued as one cycle); e2e: `e2e_ticks` ticks of `e2e_frames`;
    # tone: period[(f + sstep*s + cstep*c) mod n] of the reference's snddev_sine period at `rate`
    "cfg2": dict(channels=2, streams=1024, rate=48000, frames=480000, ticks=1, ring=1, graph=False,
                 e2e_frames=12000, e2e_ticks=40, sstep=7, cstep=3,
                 desc="1,024 x 48 kHz stereo S16 streams x 10 s per GPU, one device ring, one fused tick per step"),
    "cfg3": dict(channels=1, streams=16384, rate=16000, frames=320, ticks=50, ring=50, graph=True,
                 e2e_frames=320, e2e_ticks=50, sstep=5, cstep=0,
                 desc="16,384 x 16 kHz mono streams, 20 ms (320-frame, 640-byte) stream-blocks; step = one cycle of 50 "
                      "ticks over a 50-slot ring (1.05 GB in+out) issued as ONE span launch (cmgpu_process_cycle)"),
    "cfg4a": dict(channels=8, streams=4096, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                  e2e_frames=2400, e2e_ticks=40, sstep=7, cstep=5,
                  desc="4,096 x 48 kHz 8-channel S16 streams x 2 s per GPU, per-channel gain + 8-channel meter (parity mode)"),
    "cfg5x": dict(channels=2, streams=65536, rate=48000, frames=49152, ticks=1, ring=1, graph=False,
                  e2e_frames=4800, e2e_ticks=10, strong=True, sstep=7, cstep=3,
                  desc="DIAGNOSTIC: cfg5 with 49,152 frames per stream (stream-blocks that divide into equal work items)"),
    "cfg6ch": dict(channels=6, streams=4096, rate=48000, frames=48000, ticks=1, ring=1, graph=False,
                   e2e_frames=4800, e2e_ticks=10, sstep=7, cstep=3,
                   desc="DIAGNOSTIC: 4,096 x 48 kHz 6-channel (5.1) streams x 1 s per GPU -- a channel count that does not tile a "
                        "16-byte vector (any-channel kernel)"),
    "cfg4b": dict(channels=8, streams=4096, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                  e2e_frames=9600, e2e_ticks=10, mix_out=2, bytes_per_sample=2.5, sstep=7, cstep=5,
                  desc="EXTENSION, PARITY UNPINNED (the reference has no downmix): 4,096 x 48 kHz 8-channel streams x 2 s "
                       "per GPU, 8->2 integer downmix + metering of the 8 input and 2 output channels; 16 B read + 4 B "
                       "written per frame; checked against our own CPU restatement only"),
    "cfg4c": dict(channels=8, streams=4096, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                  e2e_frames=9600, e2e_ticks=10, mix_out=2, bytes_per_sample=2.5, sstep=7, cstep=5, out_meter_only=True,
                  desc="EXTENSION, PARITY UNPINNED: cfg4b with only the 2 OUTPUT channels metered "
                       "(CMGPU_MIX_OUTPUT_METER_ONLY), not the 8 inputs"),
    "cfg1ch": dict(channels=1, streams=8192, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                   e2e_frames=4800, e2e_ticks=10, sstep=7, cstep=0,
                   desc="DIAGNOSTIC: 8,192 x 48 kHz mono streams x 2 s per GPU"),
    "cfg4ch": dict(channels=4, streams=4096, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                   e2e_frames=4800, e2e_ticks=10, sstep=7, cstep=3,
                   desc="DIAGNOSTIC: 4,096 x 48 kHz 4-channel streams x 2 s per GPU"),
    "cfg16ch": dict(channels=16, streams=2048, rate=48000, frames=96000, ticks=1, ring=1, graph=False,
                    e2e_frames=2400, e2e_ticks=10, sstep=7, cstep=3,
                    desc="DIAGNOSTIC: 2,048 x 48 kHz 16-channel streams x 2 s per GPU (the maximum, transform.h:35)"),
    "cfg5pt": dict(channels=2, streams=65536, rate=48000, frames=48000, ticks=1, ring=1, graph=False, passthrough=True,
                   bytes_per_sample=2.0, e2e_frames=4800, e2e_ticks=10, strong=True, sstep=7, cstep=3,
                   desc="DIAGNOSTIC: cfg5 with every stream in the reference's default state (no master gain, "
                        "transform.c:107-108): in place, metered only, nothing written -- 2 B read per sample"),
    "cfg2p": dict(channels=2, streams=1024, rate=48000, frames=240000, ticks=1, ring=1, graph=False, planar=True,
                  bytes_per_sample=8.0, e2e_frames=12000, e2e_ticks=20, sstep=7, cstep=3,
                  desc="DIAGNOSTIC (SURVEY 8f N2): cfg2 x 5 s with the float-plane second output (S16 -> planar float "
                       "/32768.f, enc_vorbis.c:108-117): 2 B read + 2 B + 4 B written per sample"),
    "cfg5": dict(channels=2, streams=65536, rate=48000, frames=48000, ticks=1, ring=1, graph=False,
                 e2e_frames=4800, e2e_ticks=10, strong=True, sstep=7, cstep=3,
                 desc="65,536 x 48 kHz stereo S16 streams x 1 s in total, sharded by stream across the GPUs"),
}
NOISE_EVERY, NOISE_PHASE = 16, 5


def gain_table(first_stream: int, n: int, channels: int):
    """scale = 1000 + s % 9000 (general division), gains around 3/4 with some
    streams above 1.0 (clipping). Every stream has an active, non-unity gain."""
    s = np.arange(first_stream, first_stream + n)
    scale = (1000 + s % 9000).astype(np.uint16)
    gain = (scale[:, None].astype(np.int64) * 3 // 4 + 37 * ((s[:, None] + np.arange(channels)) % 64)).astype(np.uint16)
    return scale, gain


def mix_table(first_stream: int, n: int, cin: int, cout: int):
    """Extension workload: scale as in gain_table, weights around scale/cin (a roughly unity-sum mix
    with some streams summing above 1.0, i.e. clipping)."""
    s = np.arange(first_stream, first_stream + n)
    scale = (1000 + s % 9000).astype(np.uint16)
    m = np.arange(cout)[None, :, None]
    c = np.arange(cin)[None, None, :]
    w = (scale[:, None, None].astype(np.int64) // cin + 11 * ((s[:, None, None] + c + 3 * m) % 32)).astype(np.uint16)
    return scale, w


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []          # (arrival time, fields)
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        inside = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.03)]
        # a timed region shorter than the sampling period: fall back to the samples nearest to it
        if not inside and self.rows and self.t0 is not None:
            inside = [r for t, r in sorted(self.rows, key=lambda tr: abs(tr[0] - self.t0))[:2]]
        for r in inside:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def cpu_reference_run(wl, budget_s=12.0, threads=None, streams=None):
    """The reference's own CPU path (oracle/_ref: mem -> transform -> tee -> {consumer, vumeter},
    1,024-byte pulls, result every 20 reads) over `streams` streams of the workload on `threads`
    pthreads, repeated until ~budget_s seconds have been spent. Falls back to the oracle port."""
    from oracle import pyoracle
    from libcoolmic_dsp_b200 import synth
    channels, rate, frames = wl["channels"], wl["rate"], wl["frames"] * wl["ticks"]
    threads = threads or os.cpu_count() or 1
    ref = pyoracle.ref()
    streams = streams or max(threads, min(4 * threads, 256))
    # keep the sample's memory bounded (~1 GB of input)
    while streams * frames * channels * 2 > (1 << 30) and frames > rate:
        frames //= 2
    while streams * frames * channels < 4_000_000:      # tiny blocks: more streams, same shape
        streams *= 2
    period = synth.load_period(rate)
    pcm = np.ascontiguousarray(synth.synth_rows(period, 0, streams, channels, frames, 0, wl["sstep"], wl["cstep"],
                                                NOISE_EVERY, NOISE_PHASE))
    scale, gain = gain_table(0, streams, channels)
    samples_per_pass = streams * frames * channels
    passes, spent = 0, 0.0
    if ref is not None:
        kind = "reference"
        while spent < budget_s and passes < 4096:
            sec, _, _ = ref.bench(pcm, channels, scale, gain, rate=rate, pull=1024, result_every=20, threads=threads)
            spent += sec
            passes += 1
    else:
        kind = "port"
        port = pyoracle.port()
        fr = np.full(streams, frames, dtype=np.uint32)
        while spent < budget_s and passes < 4096:
            work = pcm.copy()
            _, sec = port.batch(work, fr, channels, scale, gain, threads=threads)
            spent += sec
            passes += 1
    msps = samples_per_pass * passes / spent / 1e6
    sample = (f"{streams} streams x {frames / rate:g} s x {channels} ch ({samples_per_pass / 1e6:.1f} Msamples) x "
              f"{passes} passes, {'reference objects, 1024-byte pulls through tee, result every 20 reads' if kind == 'reference' else 'oracle port, whole-buffer calls'}")
    return {"value": msps, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": sample,
            "seconds": spent}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def rendezvous_path() -> str:
    """Where rank 0 publishes the NCCL unique id: unique per launcher (its pid and start time) and per
    MASTER_PORT, so that neither a concurrent job nor a stale file of a dead one can be picked up."""
    ppid = os.getppid()
    try:
        start = Path(f"/proc/{ppid}/stat").read_text().rsplit(")", 1)[1].split()[19]
    except Exception:
        start = "0"
    port = os.environ.get("MASTER_PORT", "0")
    return f"/tmp/cmgpu_nccl_{ppid}_{start}_{port}.id"


def emit(line: dict):
    """The ONE JSON line, on the real stdout (libraries such as NCCL print to fd 1 too)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    t_start = time.perf_counter()
    # keep stdout clean for the JSON line: everything else written to fd 1 goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-equal-shards", action="store_true",
                    help="end-to-end leg on N > 1 GPUs: equal stream counts per rank instead of counts in proportion to each rank's host-link rate")
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="fused", choices=["fused", "transform", "meter", "copy"],
                    help="diagnostic: which parts of the tick run in the device-timed loop (default: fused = the product)")
    ap.add_argument("--no-verify", action="store_true",
                    help="skip the parity spot check (rank 0 re-derives a few streams of every rank with the oracle after "
                         "the timed end-to-end steps -- checker only, outside every timed region)")
    ap.add_argument("--verify", action="store_true", help="(default; kept for older command lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sustained / in-place roofline figures, the link probe and the object-API end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--e2e-no-gather", action="store_true", help="diagnostic: skip the NCCL result gather in the end-to-end steps")
    ap.add_argument("--e2e-wc", action="store_true", help="diagnostic: write-combined pinned upload buffers")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank, world, local = dist_env()
    wl = WORKLOADS[args.workload]
    channels, streams_per_gpu, rate, desc = wl["channels"], wl["streams"], wl["rate"], wl["desc"]
    frames, ticks, ring = wl["frames"], wl["ticks"], wl["ring"]
    scaling = "weak"
    if wl.get("strong"):
        streams_per_gpu //= max(world, 1)
        scaling = "strong"
    samples_per_step_rank = streams_per_gpu * frames * ticks * channels
    ring_bytes = 2 * ring * streams_per_gpu * frames * channels * 2
    config = {"workload": f"{args.workload}: {desc}", "streams_per_gpu": streams_per_gpu, "channels": channels,
              "rate_hz": rate, "frames_per_tick": frames, "ticks_per_step": ticks, "ring_slots": ring,
              "cycle_api": bool(wl["graph"]), "gains": "every stream active, scale 1000+s%9000, gain ~0.75..3.1",
              "input": f"one period of the reference's snddev_sine driver at {rate} Hz (tests/golden/sine.json), "
                       f"sample = period[(f + {wl['sstep']}s + {wl['cstep']}c) mod n]; every {NOISE_EVERY}th stream full-range "
                       "splitmix64 noise; written by the on-device generator (cmgpu_tone_fill / cmgpu_noise_fill)",
              "l2": f"in+out rings of {ring_bytes / 1e9:.2f} GB per GPU cycled every step: far larger than the 126 MB L2, no flush needed",
              "sharding": "by stream, one process per GPU, no data-path collective; NCCL from the C library, no PyTorch"}
    if args.mode != "fused":
        config["diagnostic_mode"] = args.mode

    # ---------------------------------------------------------------- reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from __graft_entry__ import load_package
        load_package()
        runs = []
        cpu_reference_run(wl, budget_s=1.0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            runs.append(cpu_reference_run(wl, budget_s=max(0.25, 45.0 / max(args.steps, 1))))
        wall = time.perf_counter() - t0
        value = float(np.mean([r["value"] for r in runs]))
        base = dict(runs[-1]); base["value"] = value
        line = {"impl": "reference", "metric": "aggregate Msamples/s through transform+vumeter", "value": value,
                "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "int32 (S16 in/out, int64 power)", "data": "synthetic",
                "config": config, "cpu_baseline": base,
                "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ---------------------------------------------------------------- our arm
    from __graft_entry__ import load_package
    cm = load_package()
    from libcoolmic_dsp_b200 import synth
    if cm.lib().cmgpu_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the hot path has no CPU fallback")

    comm = None
    if world > 1:
        comm = cm.Comm(local, rank, world, path=rendezvous_path())

    def barrier():
        if comm is not None:
            comm.barrier()

    def max_over_ranks(x: float) -> float:
        return x if comm is None else float(comm.max(x))

    def sum_over_ranks(x: float) -> float:
        return x if comm is None else float(comm.sum(x))

    first_stream = rank * streams_per_gpu      # == sharding.stream_range(world * streams_per_gpu, world, rank)[0]
    scale, gain = gain_table(first_stream, streams_per_gpu, channels)
    period = synth.load_period(rate)
    fill = dict(stream_step=wl["sstep"], channel_step=wl["cstep"], noise_every=NOISE_EVERY, noise_phase=NOISE_PHASE)

    # ---- device-resident run: the step's ticks live in a ring of `ring` slots, out of place so
    #      that the input stays pristine across steps
    mix_out = wl.get("mix_out", 0)
    mix_flags = cm.MIX_OUTPUT_METER_ONLY if wl.get("out_meter_only") else 0
    planar = bool(wl.get("planar"))
    bytes_per_sample = wl.get("bytes_per_sample", 4.0)

    passthrough = bool(wl.get("passthrough"))

    def configure(e, first=None, n=None):
        if passthrough:
            return                      # no gain set: transform.c:107-108
        first = first_stream if first is None else first
        n = streams_per_gpu if n is None else n
        if mix_out:
            mscale, mw = mix_table(first, n, channels, mix_out)
            for i in range(n):
                assert e.set_mix(i, int(mscale[i]), mw[i]) == 0
        elif first == first_stream and n == streams_per_gpu:
            e.set_gain_table(scale, gain)
        else:
            e.set_gain_table(*gain_table(first, n, channels))

    eng = cm.Engine(channels, streams_per_gpu, frames, ring_slots=ring, device=local,
                    flags=cm.NO_PINNED | (0 if (mix_out or passthrough) else cm.SEPARATE_OUT) | (cm.PLANAR_F32 if planar else 0) | mix_flags,
                    out_channels=mix_out)
    configure(eng)
    eng.tone_table(period)
    for slot in range(ring):
        synth.device_fill(eng, slot, None, first_stream, slot * frames, **fill)
    eng.sync()

    pflags = {"fused": cm.FUSED, "transform": cm.TRANSFORM, "meter": cm.METER, "copy": 0}[args.mode]
    if planar:
        pflags |= cm.PLANAR

    def timed(e, steps):
        if wl["graph"]:
            return e.time_cycles(steps * (ticks // ring), 0, ring, flags=pflags)
        return e.time_process(steps * ticks, 0, ring, flags=pflags)

    # A run that saw a hardware or thermal slowdown is re-measured once (sw_power_cap is kept and noted).
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    remeasured = False
    for attempt in range(2):
        clocks = ClockSampler(local)
        clocks.start()
        timed(eng, args.warmup)
        eng.reset_meters()
        launches0 = eng.launch_count()
        barrier()
        clocks.mark_begin()
        ms_total = timed(eng, args.steps)
        clocks.mark_end()
        barrier()
        clk = clocks.stop()
        launches = eng.launch_count() - launches0
        ms_total = max_over_ranks(ms_total)
        throttled = max_over_ranks(1.0 if bad & set(clk.get("reasons") or []) else 0.0) > 0
        if not throttled or attempt == 1:
            break
        remeasured = True
        time.sleep(2.0)
    if remeasured:
        clk["remeasured_after_slowdown"] = True
    ms_step = ms_total / args.steps
    value = samples_per_step_rank * world / (ms_step * 1e-3) / 1e6

    kernel = eng.kernel_name()
    # meter sanity on what was just measured: K identical ticks -> K * frames frames per stream
    snap = eng.snapshot(0, min(4, streams_per_gpu))
    assert int(snap[0].frames) == (args.steps * ticks * frames if pflags & cm.METER else 0), "meter did not see every timed tick"
    # small-buffer regime: one tick alone, device time and launch-to-complete
    if wl["graph"]:
        dev_us = min(eng.time_process(1, 0, 1, flags=pflags) for _ in range(20)) * 1e3
        med_us, min_us = eng.time_single_tick(0, pflags, reps=200)
        tick_ms = 1e3 * frames / rate
        config["single_tick"] = {"device_us": dev_us, "launch_to_complete_us": med_us, "launch_to_complete_us_min": min_us,
                                 "cadence": f"real time needs one tick per {tick_ms:g} ms = {1e3 / tick_ms:g} ticks/s; the span "
                                            f"figure sustains {1e3 * ticks / ms_step:.0f} ticks/s",
                                 "note": "one tick of the same shape issued alone, outside the timed steps: device = between CUDA "
                                         "events; launch-to-complete = cmgpu_process + cmgpu_sync timed in C "
                                         "(cmgpu_time_single_tick, median / minimum of 200)"}
    peak, peak_src = measured_peak()
    launches_per_step = max(1, launches // max(args.steps, 1))
    alg_bytes = bytes_per_sample * samples_per_step_rank / launches_per_step        # per kernel launch
    ms_launch = ms_step / launches_per_step
    achieved = alg_bytes / (ms_launch * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kernel, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_launch,
                "launches_per_step": launches_per_step,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "how": f"{args.steps} back-to-back launches over the resident ring, separate output ring; consecutive launches "
                       "overlap at their edges (programmatic dependent launch)"}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            entry = json.loads(prof.read_text()).get(args.workload, {})
            if entry.get("dram_bytes_per_launch"):
                # the capture is of a ONE-GPU launch of this workload; a rank of a strong-scaling run launches over
                # its share of the streams, so the capture's bytes are scaled by that share (and said so)
                share = 1.0 / world if (wl.get("strong") and world > 1) else 1.0
                roofline["traffic"] = entry["dram_bytes_per_launch"] * share
                roofline["traffic_source"] = (f"ncu --set full capture profiles/{entry.get('from', '?')} (dram__bytes_read.sum + "
                                              "dram__bytes_write.sum of one launch on one GPU"
                                              + (f", x 1/{world}: this rank's share of the streams)" if share != 1.0 else ")"))
        except Exception:
            pass

    if not args.no_extras and not wl["graph"] and args.mode == "fused":
        # (a) sustained: >= 2 s of back-to-back ticks (the part runs into its power cap; clocks recorded)
        n_sus = max(args.steps, int(2200.0 / max(ms_step, 1e-3)) + 1)
        clocks = ClockSampler(local)
        clocks.start()
        barrier()
        clocks.mark_begin()
        ms_sus = timed(eng, n_sus)
        clocks.mark_end()
        clk_sus = clocks.stop()
        ms_sus = max_over_ranks(ms_sus) / n_sus
        sus = bytes_per_sample * samples_per_step_rank / (ms_sus * 1e-3) / 1e9
        roofline["sustained"] = {"steps": n_sus, "seconds": ms_sus * n_sus * 1e-3, "ms_per_step": ms_sus, "achieved": sus,
                                 "frac": sus / peak, "clocks": clk_sus}
    eng.close()

    if not args.no_extras and not wl["graph"] and args.mode == "fused" and not mix_out and not planar and not passthrough:
        # (b) the in-place figure: ONE slot transformed in place (what the reference does, transform.c:120),
        #     where the overlap rule forbids consecutive launches to overlap: each tick is a full dependency
        e2 = cm.Engine(channels, streams_per_gpu, frames, ring_slots=1, device=local, flags=cm.NO_PINNED)
        configure(e2)
        synth.device_fill(e2, 0, period, first_stream, 0, **fill)
        n_ip = max(20, min(args.steps, 50))
        e2.time_process(args.warmup, 0, 1, flags=pflags)
        barrier()
        ms_ip = max_over_ranks(e2.time_process(n_ip, 0, 1, flags=pflags)) / n_ip
        ip = bytes_per_sample * samples_per_step_rank / (ms_ip * 1e-3) / 1e9
        roofline["in_place_no_overlap"] = {"steps": n_ip, "ms_per_step": ms_ip, "achieved": ip, "frac": ip / peak,
                                           "note": "one ring slot, in place, every launch a full dependency of the previous one"}
        e2.close()

    # ---- end to end through the C ABI with host buffers: ticks through a 4-slot ring
    e2e = None
    if not args.no_e2e:
        tick_frames = wl["e2e_frames"]
        n_ticks = wl["e2e_ticks"]
        e2e_steps = args.e2e_steps or min(args.steps, 5)
        # Shards of the end-to-end leg. Every sample crosses the host link twice, and the GPUs of a box do not
        # get equal shares of it (links behind a common bridge halve each other: 9.6 to over 20 GB/s per rank on
        # an 8-GPU box): with equal stream counts the step ends when the rank on the slowest link does and the
        # others idle. So the stream counts follow what cmgpu_link_probe gives each rank while all ranks drive
        # their links at once -- still contiguous ranges, still no data-path exchange.
        total_streams = streams_per_gpu * world
        shard_counts = [streams_per_gpu] * world
        shard_gbs = None
        if comm is not None and world > 1 and not args.e2e_equal_shards:
            probe_bytes = min(256 << 20, max(16 << 20, streams_per_gpu * tick_frames * channels * 2))
            barrier()
            try:
                mine = cm.link_probe(local, probe_bytes, reps=max(4, (1 << 30) // probe_bytes), both_only=True)["both_each_way_gbs"]
            except Exception:            # the probe is a hint, not the metric: every rank then sees a 0 and all stay equal
                mine = 0.0
            shard_gbs = [float(x) for x in comm.sum(*[mine if r == rank else 0.0 for r in range(world)])]
            if min(shard_gbs) > 0.0:
                shard_counts = proportional_shards(shard_gbs, total_streams)
            else:
                shard_gbs = None
        e2e_n = shard_counts[rank]
        e2e_first = sum(shard_counts[:rank])
        # host inputs: written by the same device generator into a scratch (identity, in-place) context
        # and brought down once -- setup, not measured
        gen = cm.Engine(channels, e2e_n, tick_frames, ring_slots=2, device=local, flags=cm.NO_PINNED)
        gen.tone_table(period)
        shape = (n_ticks, e2e_n, gen.stride // 2)
        pin_in = cm.PinnedArray(shape, wc=args.e2e_wc)
        stage = cm.PinnedArray(shape[1:]) if args.e2e_wc else None
        for t in range(n_ticks):
            synth.device_fill(gen, t % 2, None, e2e_first, t * tick_frames, **fill)
            if stage is None:
                gen.fetch(t % 2, pin_in.array[t])
            else:
                gen.fetch(t % 2, stage.array)
                gen.sync()
                pin_in.array[t] = stage.array
        gen.sync()
        gen.close()
        if stage is not None:
            stage.free()

        eng = cm.Engine(channels, e2e_n, tick_frames, ring_slots=4, device=local, flags=cm.NO_PINNED | mix_flags,
                        out_channels=mix_out)
        configure(eng, e2e_first, e2e_n)
        pin_out = cm.PinnedArray((n_ticks, e2e_n, eng.out_stride // 2))
        gathered = [None]
        gather = not args.e2e_no_gather
        # result arrays allocated once (65,536 x 192 B of ctypes structures take milliseconds to create)
        gather_out = cm.Comm.alloc_results(total_streams, world) if (comm is not None and rank == 0) else None
        results_out = cm.Engine.alloc_results(e2e_n)

        presubmitted = [0]
        ahead = min(3, n_ticks)

        def e2e_step():
            # Upload, tick and download of every tick are queued on the engine's three streams; the
            # step's meter results are taken as soon as its last tick has run (stream order on the
            # compute stream), sent to rank 0 over NCCL and finalised there, while the downloads of this
            # step's last ticks overlap the uploads of the next step's first ones: those are queued
            # BEFORE the host waits for the results, so the upload stream never idles at a step boundary.
            # Everything is drained (eng.sync) before the clock stops. No barrier between steps: ranks
            # other than the root never wait on the host for the gather.
            for t in range(n_ticks):
                slot = t % 4
                if t >= presubmitted[0]:
                    eng.submit(slot, pin_in.array[t])
                eng.process(slot)
                eng.fetch(slot, pin_out.array[t])
            for t in range(ahead):                       # the next step's first uploads
                eng.submit(t % 4, pin_in.array[t])
            presubmitted[0] = ahead
            if comm is not None and gather:
                out = comm.gather_results(eng, rate, total_streams, out=gather_out)
            else:
                res, st, rcs = eng.results(rate, out=results_out)
                out = (res, st, rcs, [e2e_n])
            if out is not None:
                gathered[0] = out

        for _ in range(2):
            e2e_step()
        eng.sync()
        barrier()
        gather_ms = 0.0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
            if comm is not None:
                gather_ms += comm.last_gather_ms()
        eng.sync()
        wall_mine = time.perf_counter() - t0
        barrier()
        wall = max_over_ranks(wall_mine)
        spot = "skipped (--no-verify)"
        if rank == 0:
            res, st, rcs, counts = gathered[0]
            gathered_all = comm is not None and gather
            assert int(st[0].frames) == tick_frames * n_ticks and list(counts) == (shard_counts if gathered_all else [e2e_n])
            if not args.no_verify:
                spot = spot_check(cm, synth, wl, period, shard_counts if gathered_all else [e2e_n],
                                  channels, tick_frames, n_ticks, rate, pin_out.array, res, st)
        slot_bytes = e2e_n * eng.stride
        out_slot_bytes = e2e_n * eng.out_stride
        meter_bytes = e2e_n * eng.meter_row_u64() * 8
        samples_e2e_all = total_streams * tick_frames * n_ticks * channels
        e2e = {"parity_spot_check": spot, "through": "cmgpu_submit / cmgpu_process / cmgpu_fetch / cmgpu_gather_results (C ABI)",
               "value": samples_e2e_all * e2e_steps / wall / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": n_ticks * slot_bytes, "d2h_bytes_per_step": n_ticks * out_slot_bytes + meter_bytes,
               "steps": e2e_steps, "ms_per_step": 1e3 * wall / e2e_steps,
               "gbs_each_way_all_ranks": sum_over_ranks(n_ticks * slot_bytes * e2e_steps / wall_mine / 1e9),
               "nccl_gather_ms_per_step": (gather_ms / e2e_steps) if comm is not None else None,
               "how": f"{n_ticks} ticks of {tick_frames} frames per step through a 4-slot ring, pinned host buffers, "
                      "upload/compute/download on three CUDA streams (the next step's first uploads are queued before the "
                      "host waits for this step's results); per step the meter results of all ranks are gathered to "
                      "rank 0 by cmgpu_gather_results (NCCL send/recv of the raw rows on the compute stream, decode + dB "
                      "finalise on rank 0); steps are queued back to back and drained before the clock stops"}
        if world > 1:
            e2e["shards"] = {"streams_per_rank": shard_counts, "link_gbs_per_rank": shard_gbs,
                             "rule": "equal" if shard_gbs is None else "in proportion to each rank's share of the host links "
                                     "(cmgpu_link_probe, all ranks at once, both directions), contiguous stream ranges",
                             "note": "h2d / d2h bytes per step are rank 0's"}
            e2e["h2d_bytes_per_step_all_ranks"] = int(sum_over_ranks(float(n_ticks * slot_bytes)))
            e2e["d2h_bytes_per_step_all_ranks"] = int(sum_over_ranks(float(n_ticks * out_slot_bytes + meter_bytes)))
        eng.close()
        if not args.no_extras:
            # what the host link gives when every rank drives it at once, both directions (the e2e ceiling)
            probe_bytes = min(256 << 20, max(16 << 20, slot_bytes))
            barrier()
            link = cm.link_probe(local, probe_bytes, reps=max(4, (2 << 30) // probe_bytes), both_only=world > 1)
            link["both_each_way_gbs_all_ranks"] = sum_over_ranks(link["both_each_way_gbs"])
            link["chunk_bytes"] = probe_bytes
            e2e["link"] = link
            e2e["link_ceiling_gbs"] = link["both_each_way_gbs_all_ranks"]
            e2e["frac_of_link"] = e2e["gbs_each_way_all_ranks"] / max(link["both_each_way_gbs_all_ranks"], 1e-9)
        if not args.no_extras and not mix_out:
            # Pass-through streams, the reference's default state (no master gain: transform.c:107-108
            # leaves the buffer alone): uploaded from the pinned staging ring, metered, nothing downloaded.
            pt_ticks = min(n_ticks, 8)
            eng = cm.Engine(channels, e2e_n, tick_frames, ring_slots=4, device=local)
            for slot in range(4):
                eng.host_slot(slot)[:] = pin_in.array[slot % n_ticks]
            for _ in range(2):
                for t in range(pt_ticks):
                    eng.submit(t % 4); eng.process(t % 4); eng.fetch(t % 4)
            eng.results(rate, out=results_out)
            eng.sync()
            up0, down0 = eng.transfer_bytes()
            barrier()
            t0 = time.perf_counter()
            pt_steps = 3
            for _ in range(pt_steps):
                for t in range(pt_ticks):
                    eng.submit(t % 4); eng.process(t % 4); eng.fetch(t % 4)
                eng.results(rate, out=results_out)
            eng.sync()
            pt_wall = max_over_ranks(time.perf_counter() - t0)
            up1, down1 = eng.transfer_bytes()
            e2e["passthrough"] = {"value": total_streams * tick_frames * pt_ticks * channels * pt_steps / pt_wall / 1e6,
                                  "unit": "Msamples/s", "h2d_bytes_per_step": (up1 - up0) // pt_steps,
                                  "d2h_bytes_per_step": (down1 - down0) // pt_steps + meter_bytes,
                                  "note": "every stream in the reference's default state (no gain set): uploaded from the "
                                          "pinned staging ring, metered on the device, no PCM download -- the staging slot "
                                          "already holds the result"}
            eng.close()
        pin_in.free(); pin_out.free()

        if not args.no_extras and rank == 0 and not mix_out:
            try:
                e2e["object_api"] = object_api_leg(cm, synth, local)
            except Exception as exc:      # the leg is a report, not the metric
                e2e["object_api"] = {"error": repr(exc)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and mix_out:
        # no reference implementation exists for the extension: time our own restatement, one thread
        from oracle import pyoracle
        port = pyoracle.port()
        n = min(frames, 96000)
        pcm = np.ascontiguousarray(synth.synth_rows(period, 0, 1, channels, n, 0, wl["sstep"], wl["cstep"], 0))
        mscale, mw = mix_table(0, 1, channels, mix_out)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 5.0:
            port.mix(pcm[0], n, channels, mix_out, int(mscale[0]), mw[0], pyoracle.Meter(), pyoracle.Meter())
            reps += 1
        cpu = {"value": reps * n * channels / (time.perf_counter() - t0) / 1e6, "unit": "Msamples/s", "cores": 1,
               "kind": "port", "sample": f"1 stream x {n} frames x {reps} passes, our own restatement (no reference exists)"}
    elif rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(wl, budget_s=12.0)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {"metric": "aggregate Msamples/s through transform+vumeter", "value": value, "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "int32 (S16 in/out, int64 power)", "data": "synthetic", "config": config,
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "bench_wall_s": time.perf_counter() - t_start,
                "nccl": ({"version": int(cm.lib().cmgpu_comm_nccl_version()), "ranks": world,
                          "from": "libcoolmic_b200.so (cmgpu_comm_*), id exchanged through a file; no torch.distributed"}
                         if comm is not None else None)}
        emit(line)
    if comm is not None:
        comm.barrier()
        comm.close()
    return 0


def proportional_shards(weights, total, quantum=64):
    """Stream counts per rank in proportion to `weights`, multiples of `quantum` (the last rank takes the
    remainder), at least one quantum each; contiguous ranges follow from the running sum."""
    world = len(weights)
    w = [max(float(x), 1e-6) for x in weights]
    units = total // quantum
    if units < world:
        return [total // world + (1 if r < total % world else 0) for r in range(world)]
    raw = [x / sum(w) * units for x in w]
    counts = [max(1, int(x)) for x in raw]
    # hand the units rounding left over to the ranks with the largest remainders (or take them back)
    order = sorted(range(world), key=lambda r: raw[r] - int(raw[r]), reverse=True)
    i = 0
    while sum(counts) < units:
        counts[order[i % world]] += 1
        i += 1
    while sum(counts) > units:
        r = max(range(world), key=lambda k: counts[k])
        counts[r] -= 1
    counts = [c * quantum for c in counts]
    counts[-1] += total - sum(counts)
    return counts


def spot_check(cm, synth, wl, period, counts, channels, tick_frames, n_ticks, rate, out0, results, states):
    """Rank 0 re-derives, with the oracle port, what a few streams of EVERY rank must have produced
    in the last end-to-end step: transformed PCM (rank 0's own streams), the integer meter state and the
    finalised dB values (all ranks, from the rows that came over NCCL). The oracle is the checker here,
    nothing it computes is measured or shipped."""
    from oracle import pyoracle
    port = pyoracle.port()
    frames = tick_frames * n_ticks
    checked = 0
    world = len(counts)
    firsts = [sum(counts[:r]) for r in range(world)]

    def rows(g):
        return np.ascontiguousarray(synth.synth_rows(period, g, 1, channels, frames, 0, wl["sstep"], wl["cstep"],
                                                     NOISE_EVERY, NOISE_PHASE))

    if wl.get("mix_out"):
        # extension: our own restatement is the only checker there is (parity unpinned)
        cout = wl["mix_out"]
        for s in sorted({0, 5, counts[0] // 2, counts[0] - 1}):
            pcm = rows(s)
            mscale, mw = mix_table(s, 1, channels, cout)
            m_out = pyoracle.Meter()
            want = port.mix(pcm[0], frames, channels, cout, int(mscale[0]), mw[0], None, m_out)
            got = np.concatenate([out0[t, s, : tick_frames * cout] for t in range(n_ticks)])
            st = states[s]
            ok = np.array_equal(got, want) and int(st.frames) == frames
            for c in range(cout):
                ok = ok and int(st.power[c]) == int(m_out.power[c]) and int(st.channel_peak[c]) == int(m_out.channel_peak[c])
            if not ok:
                return f"MISMATCH (extension) stream {s}"
            checked += 1
        return f"ok: {checked} streams bit-exact vs our own CPU restatement (extension, parity unpinned)"
    for r in range(world):
        for s in sorted({0, min(5, counts[r] - 1), counts[r] // 2, counts[r] - 1}):
            g = firsts[r] + s
            pcm = rows(g)
            scale, gain = gain_table(g, 1, channels)
            meters, _ = port.batch(pcm, np.array([frames], np.uint32), channels, scale, gain)
            st = states[g]
            ok = int(st.frames) == frames and int(st.global_peak) == int(meters[0].global_peak)
            for c in range(channels):
                ok = ok and int(st.power[c]) == int(meters[0].power[c])
                ok = ok and int(st.channel_peak[c]) == int(meters[0].channel_peak[c])
            want = port.finalise(meters[0], rate, channels)
            got = results[g].as_dict()
            ok = ok and np.float64(got["global_power"]).tobytes() == np.float64(want["global_power"]).tobytes()
            for c in range(channels):
                ok = ok and np.float64(got["channel_power"][c]).tobytes() == np.float64(want["channel_power"][c]).tobytes()
            if r == 0:
                got_pcm = np.concatenate([out0[t, s, : tick_frames * channels] for t in range(n_ticks)])
                ok = ok and np.array_equal(got_pcm, pcm[0])
            if not ok:
                return f"MISMATCH rank {r} stream {s}"
            checked += 1
    return (f"ok: {checked} streams over {world} rank(s) bit-exact vs oracle (PCM on rank 0; integer meter state and dB "
            f"doubles on all, as gathered by cmgpu_gather_results)")


def object_api_leg(cm, synth, device, threads=None, slots=4, block=12000, n_ticks=40):
    """The path a libcoolmic-dsp maintainer would call (src/simple.c:212-229,445-505 with
    coolmic_b200_batch_tick in place of the per-stream pull loop): 1,024 member transforms fed by memory
    iohandles, fused meters, every transform's output read back through its iohandle -- the reference-
    named object API on a pipelined ring, host buffers on both sides."""
    lib = cm.lib()
    if not hasattr(lib, "coolmic_b200_bench_objects"):
        return {"error": "library built without the object-API bench driver"}
    fn = lib.coolmic_b200_bench_objects
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p,
                   C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    streams, channels = 1024, 2
    threads = threads or max(1, min(16, os.cpu_count() or 1))
    period = synth.load_period(48000)
    pcm = np.ascontiguousarray(synth.synth_rows(period, 0, streams, channels, block * 4, 0, 7, 3, NOISE_EVERY, NOISE_PHASE))
    secs = C.c_double(0)
    check = C.c_uint64(0)
    best = None
    for _ in range(3):
        rc = fn(device, channels, streams, block, n_ticks, slots, threads, pcm.shape[1] * 2, pcm.ctypes.data,
                C.byref(secs), C.byref(check))
        if rc != 0:
            return {"error": f"coolmic_b200_bench_objects: {rc}"}
        best = secs.value if best is None else min(best, secs.value)
    samples = streams * channels * block * n_ticks
    out = {"through": "coolmic_b200_batch_tick + coolmic_iohandle_read on 1,024 member transforms (memory iohandles in, "
                      "per-transform handles out), fused vumeters, coolmic_vumeter_result per stream at the end",
           "value": samples / best / 1e6, "unit": "Msamples/s", "ms_per_step": best * 1e3,
           "shape": f"{streams} x stereo x {block} frames x {n_ticks} ticks, {slots}-slot ring, {threads} host threads",
           "frames_metered_checksum": int(check.value)}
    alone = host_loop_alone(threads, slots, streams, block, n_ticks)
    if alone:
        out["host_loop_alone"] = alone
    return out


def host_loop_alone(threads, slots, streams, block, n_ticks):
    """The ceiling of the object-API leg on this host: the same driver and the same object code linked against
    an engine whose calls return at once (tools/hostprobe), i.e. only the two host copies per sample the
    iohandle contract costs (source -> pinned slot, pinned slot -> consumer). Diagnostic, no GPU involved."""
    exe = os.path.join(ROOT, "tools", "hostprobe", "hostloop")
    try:
        if not os.path.exists(exe):
            subprocess.run(["make", "-C", os.path.dirname(exe), "hostloop"], capture_output=True, timeout=120)
        r = subprocess.run([exe, str(threads), str(slots), str(streams), str(block), str(n_ticks)], capture_output=True,
                           text=True, timeout=120)
        f = r.stdout.split()
        secs = float(f[f.index("seconds") + 1])
        return {"value": streams * 2 * block * n_ticks / secs / 1e6, "unit": "Msamples/s", "ms_per_step": secs * 1e3,
                "note": "the object-API driver on an engine that returns at once: the host copies alone (no DMA competing "
                        "for host memory); no GPU run through the objects can be faster on this host"}
    except Exception:
        return None


if __name__ == "__main__":
    sys.exit(main())
