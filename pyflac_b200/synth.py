"""Deterministic synthetic PCM: "music-like" streams and the parity corpus.

stream ``s`` of a batch uses seed ``base_seed + s``; everything is numpy so the same samples can be
fed to the GPU engine and to libFLAC on the host in the same run.
"""
import numpy as np


def music_like(n, channels=2, sample_rate=48000, bps=16, seed=0):
    """Sum of 6 slowly amplitude-modulated sinusoids shared across channels (per-channel gain)
    plus per-channel AR(1) noise; scaled to half of full scale, rounded, clipped.
    Returns int16 (bps<=16) or int32 array of shape (n, channels)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sample_rate
    f = rng.uniform(50.0, 4000.0, 6)
    a = rng.uniform(0.02, 0.15, 6)
    fm = rng.uniform(0.1, 2.0, 6)
    ph = rng.uniform(0, 2 * np.pi, 6)
    base = np.zeros(n)
    for k in range(6):
        base += a[k] * (0.6 + 0.4 * np.sin(2 * np.pi * fm[k] * t)) * np.sin(2 * np.pi * f[k] * t + ph[k])
    out = np.empty((n, channels), np.float64)
    for c in range(channels):
        e = rng.standard_normal(n) * 0.01
        # AR(1), pole 0.95 (scipy is optional: fall back to a direct recursion in blocks)
        try:
            from scipy.signal import lfilter
            noise = lfilter([1.0], [1.0, -0.95], e)
        except Exception:  # pragma: no cover
            noise = np.empty(n)
            acc = 0.0
            for i in range(n):
                acc = 0.95 * acc + e[i]
                noise[i] = acc
        out[:, c] = (1.0 - 0.2 * c) * base + noise
    full = float(1 << (bps - 1))
    q = np.rint(out * 0.5 * full)
    q = np.clip(q, -full, full - 1)
    return q.astype(np.int16 if bps <= 16 else np.int32)


def music_like_batch(n_streams, n, channels=2, sample_rate=48000, bps=16, base_seed=0):
    """(n_streams, n, channels) array of independent music-like streams."""
    first = music_like(n, channels, sample_rate, bps, base_seed)
    out = np.empty((n_streams,) + first.shape, first.dtype)
    out[0] = first
    for s in range(1, n_streams):
        out[s] = music_like(n, channels, sample_rate, bps, base_seed + s)
    return out


def corpus_signal(kind, n, channels, bps, seed=0, sample_rate=48000):
    """Parity-corpus signals exercising each subframe type / branch."""
    rng = np.random.default_rng(seed)
    full = 1 << (bps - 1)
    dt = np.int16 if bps <= 16 else np.int32
    if kind == "music":
        return music_like(n, channels, sample_rate, bps, seed)
    if kind == "silence":
        return np.zeros((n, channels), dt)
    if kind == "dc":
        return np.full((n, channels), min(1234, full - 1), dt)
    if kind == "noise":      # full-scale white noise -> VERBATIM
        return rng.integers(-full, full, (n, channels)).astype(dt)
    if kind == "lownoise":   # small noise -> FIXED order 0 / small k
        return rng.integers(-3, 4, (n, channels)).astype(dt)
    if kind == "ramp":       # FIXED order 1/2
        r = (np.arange(n)[:, None] * (np.arange(channels)[None, :] + 1) * 3) % (full // 2)
        return r.astype(dt)
    if kind == "wasted":     # low bits zero -> wasted-bits path (side channel too)
        m = music_like(n, channels, sample_rate, bps, seed).astype(np.int64)
        return ((m >> 4) << 4).astype(dt)
    if kind == "sine":       # pure tone, very predictable -> high LPC gain, small residuals
        t = np.arange(n)[:, None]
        x = np.sin(2 * np.pi * 440.0 * t / sample_rate + np.arange(channels)[None, :]) * 0.8 * full
        return np.rint(x).clip(-full, full - 1).astype(dt)
    if kind == "square":     # full-scale alternating -> large fixed-predictor errors (accumulator wrap paths)
        x = np.where((np.arange(n)[:, None] + np.arange(channels)[None, :]) % 2 == 0, full - 1, -full)
        return x.astype(dt)
    if kind == "mixed":      # half music, then silence, then noise: forces every subframe type in one stream
        m = music_like(n, channels, sample_rate, bps, seed)
        a, b = n // 3, 2 * n // 3
        m[a:b] = 0
        m[b:] = rng.integers(-full, full, (n - b, channels)).astype(dt)
        return m
    if kind == "lr_uncorr":  # uncorrelated channels -> independent channel assignment
        return np.stack([music_like(n, 1, sample_rate, bps, seed + 17 * c)[:, 0] for c in range(channels)], 1)
    raise ValueError(kind)


CORPUS_KINDS = ["music", "silence", "dc", "noise", "lownoise", "ramp", "wasted", "sine", "square", "mixed",
                "lr_uncorr"]
