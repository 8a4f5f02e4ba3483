"""Deterministic edge-case images shared by the golden generator and the parity tests.

Mirrors the survey's edge list: noise, checkerboard, constant 0/255, stripes,
gradients (including a wrapping one that overflows int64 in AVP at high near), widths 1-5,
heights 1-3, 1x1.
"""
from __future__ import annotations

import numpy as np

# (effort, near) pairs exercised on every edge image
SETTINGS_EDGE = [(0, 0), (1, 0), (1, 3), (1, 9), (2, 0), (2, 3), (2, 9), (3, 0), (3, 3), (3, 9)]


def _noise(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w), dtype=np.uint8)


def _smooth(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    v = 128 + 60 * np.sin(x / 7.0 + rng.random()) + 50 * np.cos(y / 5.0) + rng.normal(0, 3, size=(h, w))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def edge_cases():
    out = []
    for h, w in [(1, 1), (1, 2), (2, 1), (1, 5), (5, 1), (2, 2), (3, 3), (2, 7), (3, 4), (4, 3), (3, 17), (17, 3), (7, 64)]:
        out.append((f"noise_{h}x{w}", _noise(h, w, 100 + 31 * h + w)))
        out.append((f"smooth_{h}x{w}", _smooth(h, w, 200 + 31 * h + w)))
    out.append(("zeros_9x13", np.zeros((9, 13), np.uint8)))
    out.append(("full_9x13", np.full((9, 13), 255, np.uint8)))
    out.append(("zeros_40x40", np.zeros((40, 40), np.uint8)))
    out.append(("mid_33x70", np.full((33, 70), 128, np.uint8)))
    y, x = np.mgrid[0:48, 0:80]
    out.append(("checker_48x80", (((x + y) & 1) * 255).astype(np.uint8)))
    out.append(("checker4_48x80", ((((x >> 2) + (y >> 2)) & 1) * 255).astype(np.uint8)))
    out.append(("hstripes_48x80", ((y & 1) * 255).astype(np.uint8)))
    out.append(("vstripes_48x80", ((x & 1) * 255).astype(np.uint8)))
    out.append(("hgrad_48x80", (x * 255 // 79).astype(np.uint8)))
    out.append(("vgrad_48x80", (y * 255 // 47).astype(np.uint8)))
    out.append(("wrapgrad_64x96", ((np.mgrid[0:64, 0:96][1] * 13 + np.mgrid[0:64, 0:96][0] * 5) & 255).astype(np.uint8)))
    out.append(("noise_64x96", _noise(64, 96, 5)))
    out.append(("smooth_61x127", _smooth(61, 127, 6)))
    out.append(("smooth_128x130", _smooth(128, 130, 8)))
    out.append(("dark_50x50", np.clip(_smooth(50, 50, 9).astype(int) - 150, 0, 255).astype(np.uint8)))
    out.append(("bright_50x50", np.clip(_smooth(50, 50, 10).astype(int) + 150, 0, 255).astype(np.uint8)))
    return out
