"""Synthetic PCM of the measurement plan: two sines + Gaussian noise, deterministic per seed.
Used by bench.py and the tools to generate inputs (the oracle keeps its own copy for the tests)."""
import numpy as np


def synth_pcm(seconds: float, srate: int, channels: int, seed: int) -> np.ndarray:
    """float32 [n, C]: x[n, c] = 0.25 g_c sin(2 pi 440 n / sr) + 0.10 sin(2 pi 3000 n / sr + 1) + 0.02 N(0, 1), clipped to
    [-1, 1], g = linspace(1.0, 0.3, C)."""
    n = int(round(seconds * srate))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)[:, None]
    g = np.linspace(1.0, 0.3, channels)[None, :]
    x = 0.25 * g * np.sin(2 * np.pi * 440 * t / srate) + 0.10 * np.sin(2 * np.pi * 3000 * t / srate + 1) \
        + 0.02 * rng.standard_normal((n, channels))
    return np.clip(x, -1, 1).astype(np.float32)
