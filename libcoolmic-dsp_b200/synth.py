"""Synthetic inputs of the bench / test workloads, on the host (numpy) -- the same formulas the device
tone source (cmgpu_tone_fill / cmgpu_noise_fill, csrc/cmgpu_post.cu) evaluates, so that checkers can
re-derive any stream of a device-generated ring.

SURVEY.md 8d: stream s, channel c, frame f -> period[(f + a*s + b*c) mod n], where `period` is ONE
PERIOD OF THE REFERENCE'S OWN snddev_sine DRIVER at that rate (reference src/snddev_sine.c:118-150;
read through the real driver by tests/golden/make_golden.py and committed in tests/golden/sine.json --
the tables follow no closed formula, so they are taken from the driver's output, never re-typed), plus
a second, full-range data set x = (int16) splitmix64(seed ^ s<<40 ^ f<<4 ^ c) on every `noise_every`-th
stream to exercise the clamp and the peak tie-breaks.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
NOISE_SEED = 0xC0011DC5
_M64 = (1 << 64) - 1


def load_period(rate: int) -> np.ndarray:
    """One period of the reference's snddev_sine driver at `rate` (tests/golden/sine.json)."""
    for row in json.loads((ROOT / "tests" / "golden" / "sine.json").read_text()):
        if int(row["rate"]) == int(rate) and row.get("period"):
            return np.asarray(row["period"], dtype=np.int16)
    raise KeyError(f"no snddev_sine period recorded for {rate} Hz")


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def tone_rows(period: np.ndarray, first_stream: int, n: int, channels: int, frames: int, first_frame: int = 0,
              stream_step: int = 7, channel_step: int = 3) -> np.ndarray:
    """int16 [n][frames*channels]: period[(first_frame + f + stream_step*s + channel_step*c) mod len]."""
    plen = period.size
    s = np.arange(first_stream, first_stream + n, dtype=np.int64)
    # only `plen` distinct rows exist: build them once and gather
    f = np.arange(frames, dtype=np.int64)[:, None]
    c = np.arange(channels, dtype=np.int64)[None, :]
    base = (f + channel_step * c).reshape(-1)                       # [frames*channels]
    rows = np.stack([period[(p + base) % plen] for p in range(plen)])
    phase = (first_frame + stream_step * s) % plen
    return rows[phase]


def noise_rows(first_stream: int, n: int, channels: int, frames: int, first_frame: int = 0,
               seed: int = NOISE_SEED) -> np.ndarray:
    with np.errstate(over="ignore"):
        s = np.arange(first_stream, first_stream + n, dtype=np.uint64)[:, None, None]
        f = (np.uint64(first_frame) + np.arange(frames, dtype=np.uint64))[None, :, None]
        c = np.arange(channels, dtype=np.uint64)[None, None, :]
        x = np.uint64(seed) ^ (s << np.uint64(40)) ^ (f << np.uint64(4)) ^ c
        return splitmix64(x).astype(np.uint16).view(np.int16).reshape(n, frames * channels)


def synth_rows(period: np.ndarray, first_stream: int, n: int, channels: int, frames: int, first_frame: int = 0,
               stream_step: int = 7, channel_step: int = 3, noise_every: int = 16, noise_phase: int = 5,
               seed: int = NOISE_SEED) -> np.ndarray:
    """The bench data set: tone everywhere, noise on streams with s % noise_every == noise_phase."""
    out = tone_rows(period, first_stream, n, channels, frames, first_frame, stream_step, channel_step)
    if noise_every:
        s = np.arange(first_stream, first_stream + n)
        idx = np.nonzero(s % noise_every == noise_phase)[0]
        for i in idx:
            out[i] = noise_rows(first_stream + int(i), 1, channels, frames, first_frame, seed)[0]
    return out


def device_fill(eng, slot: int, period: np.ndarray | None, first_stream: int, first_frame: int = 0,
                stream_step: int = 7, channel_step: int = 3, noise_every: int = 16, noise_phase: int = 5,
                seed: int = NOISE_SEED) -> None:
    """The same data set written into a ring slot by the device generator (no host upload)."""
    if period is not None:
        eng.tone_table(period)
    eng.tone_fill(slot, first_frame, first_stream, stream_step, channel_step)
    if noise_every:
        eng.noise_fill(slot, first_frame, first_stream, seed, noise_every, noise_phase)
