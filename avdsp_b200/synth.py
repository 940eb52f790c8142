"""Deterministic synthetic PCM: identical on host (numpy) and device (torch).

Per stream s an LCG  u <- u*1664525 + 1013904223 (mod 2^32), u0 = 0x9E3779B9*(s+1), one draw per
(frame, input channel) in channel-minor order.  Stimuli:
  "noise"   sample = (int32)u >> 2      white, -12 dBFS peak  (throughput runs)
  "full"    sample = (int32)u           full-scale white      (exercises biquad / SAT saturation)
  "impulse" 0x7FFFFFFF at frame 0 on every channel, else 0
  "sine"    997 Hz, -1 dBFS, rounded to int32, channel c phase-shifted by c radians
"""
from __future__ import annotations

import math

import numpy as np

A, C = 1664525, 1013904223
GOLD = 0x9E3779B9
MASK = 0xFFFFFFFF


def _jump_table(n: int):
    """(a_k, c_k) with u_k = a_k*u0 + c_k  for k = 1..n  (mod 2^32)."""
    a = np.empty(n, dtype=np.uint64)
    c = np.empty(n, dtype=np.uint64)
    ak, ck = 1, 0
    for k in range(n):
        ak = (ak * A) & MASK
        ck = (ck * A + C) & MASK
        a[k], c[k] = ak, ck
    return a, c


def lcg_u32(n_streams: int, n_frames: int, n_ch: int, first_stream: int = 0, first_frame: int = 0) -> np.ndarray:
    """uint32 draws, shape [S, T, C]."""
    n = (first_frame + n_frames) * n_ch
    a, c = _jump_table(n)
    a, c = a[first_frame * n_ch:], c[first_frame * n_ch:]
    u0 = (GOLD * (np.arange(first_stream, first_stream + n_streams, dtype=np.uint64) + 1)) & MASK
    u = (u0[:, None] * a[None, :] + c[None, :]) & MASK
    return u.astype(np.uint32).reshape(n_streams, n_frames, n_ch)


def pcm(kind: str, n_streams: int, n_frames: int, n_ch: int, fs: int = 48000,
        first_stream: int = 0, first_frame: int = 0) -> np.ndarray:
    """int32 PCM [S, T, C] for one of the stimuli above."""
    if kind in ("noise", "full"):
        x = lcg_u32(n_streams, n_frames, n_ch, first_stream, first_frame).view(np.int32)
        return x >> 2 if kind == "noise" else x.copy()
    if kind == "impulse":
        x = np.zeros((n_streams, n_frames, n_ch), dtype=np.int32)
        if first_frame == 0 and n_frames:
            x[:, 0, :] = 0x7FFFFFFF
        return x
    if kind == "sine":
        t = np.arange(first_frame, first_frame + n_frames, dtype=np.float64)[None, :, None]
        ph = np.arange(n_ch, dtype=np.float64)[None, None, :] + 0.37 * np.arange(
            first_stream, first_stream + n_streams, dtype=np.float64)[:, None, None]
        amp = 10 ** (-1 / 20) * (2 ** 31 - 1)
        return np.round(amp * np.sin(2 * math.pi * 997.0 * t / fs + ph)).astype(np.int32)
    raise ValueError(kind)


def pcm_float(kind, *a, **k) -> np.ndarray:
    """float32 PCM in [-1,1) carried as int32 bit patterns (sample formats 5/6)."""
    return (pcm(kind, *a, **k).astype(np.float64) / 2 ** 31).astype(np.float32).view(np.int32)


def pcm_torch(kind: str, n_streams: int, n_frames: int, n_ch: int, device, first_stream: int = 0):
    """Same "noise"/"full" values generated directly on `device` as an int32 tensor [S, T, C]."""
    import torch
    assert kind in ("noise", "full")
    a, c = _jump_table(n_frames * n_ch)
    ta = torch.from_numpy(a.astype(np.int64)).to(device)
    tc = torch.from_numpy(c.astype(np.int64)).to(device)
    u0 = (GOLD * (torch.arange(first_stream, first_stream + n_streams, dtype=torch.int64, device=device) + 1)) & MASK
    out = torch.empty((n_streams, n_frames * n_ch), dtype=torch.int32, device=device)
    step = max(1, (1 << 24) // max(1, n_frames * n_ch))   # bound the int64 temporary
    for s0 in range(0, n_streams, step):
        u = (u0[s0:s0 + step, None] * ta[None, :] + tc[None, :]) & MASK
        u = torch.where(u >= (1 << 31), u - (1 << 32), u).to(torch.int32)
        out[s0:s0 + step] = (u >> 2) if kind == "noise" else u
    return out.view(n_streams, n_frames, n_ch)
