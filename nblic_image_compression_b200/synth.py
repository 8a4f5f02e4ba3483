"""Deterministic synthetic "photographic-like" 8-bit gray images.

Integer-only value-noise generator used by the parity tests and by bench.py (the
reference ships no synthetic inputs; BASELINE.json's configs 3-5 name these shapes).
`gen(h, w, seed)` is the numpy definition; the CUDA library carries a bit-identical
device generator (`nblic_b200_synth_gray`) so that large batches never cross PCIe.
`occluders(h, w, seed)` is the only part that needs numpy's PCG64 stream, so the device
generator takes its 12x5 integer table as an argument.
"""
from __future__ import annotations

import numpy as np

_SHIFTS = (8, 7, 6, 5, 4, 3, 2)
_AMPS = (64, 48, 32, 20, 12, 7, 4)
N_OCCLUDERS = 12


def _h32(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & 0xFFFFFFFF
    v ^= v >> 16
    v = (v * 0x7FEB352D) & 0xFFFFFFFF
    v ^= v >> 15
    v = (v * 0x846CA68B) & 0xFFFFFFFF
    v ^= v >> 16
    return v


def _lattice(o: int, ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    k = (seed * 131 + o * 7919) & 0xFFFFFFFF
    inner = _h32(((iy.astype(np.uint64) * 0x85EBCA77) & 0xFFFFFFFF) ^ k)
    return (_h32(((ix.astype(np.uint64) * 0x9E3779B1) & 0xFFFFFFFF) ^ inner) & 255).astype(np.int64)


def occluders(h: int, w: int, seed: int) -> np.ndarray:
    """(12, 5) int64 table: cx, cy, rad, off, kind(0 disk / 1 box), in PCG64 call order."""
    rng = np.random.default_rng(seed)
    lo, hi = min(h, w) // 32 + 2, min(h, w) // 5 + 3
    tab = np.zeros((N_OCCLUDERS, 5), dtype=np.int64)
    for k in range(N_OCCLUDERS):
        cx = int(rng.integers(0, w))
        cy = int(rng.integers(0, h))
        rad = int(rng.integers(lo, hi))
        off = int(rng.integers(-60, 61))
        tab[k] = (cx, cy, rad, off, k & 1)
    return tab


def gen(h: int, w: int, seed: int) -> np.ndarray:
    """uint8 (h, w) raster, top-down."""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    acc = np.zeros((h, w), dtype=np.int64)
    for o, (sh, amp) in enumerate(zip(_SHIFTS, _AMPS)):
        S = 1 << sh
        ix, iy = x >> sh, y >> sh
        fx, fy = x & (S - 1), y & (S - 1)
        v00 = _lattice(o, ix, iy, seed)
        v10 = _lattice(o, ix + 1, iy, seed)
        v01 = _lattice(o, ix, iy + 1, seed)
        v11 = _lattice(o, ix + 1, iy + 1, seed)
        v = ((v00 * (S - fx) + v10 * fx) * (S - fy) + (v01 * (S - fx) + v11 * fx) * fy) >> (2 * sh)
        acc += amp * v
    img = acc // sum(_AMPS)
    img = 128 + ((img - 128) * 3) // 2
    for cx, cy, rad, off, kind in occluders(h, w, seed):
        if kind == 0:
            m = (x - cx) ** 2 + (y - cy) ** 2 < rad * rad
        else:
            m = (np.abs(x - cx) < rad) & (np.abs(y - cy) < rad // 2 + 1)
        img = img + off * m
    k = (seed * 977 + 12345) & 0xFFFFFFFF
    hn = _h32(((x.astype(np.uint64) * 0x27D4EB2D) & 0xFFFFFFFF) ^ _h32(y.astype(np.uint64) ^ k))
    nz = np.zeros((h, w), dtype=np.int64)
    for i in range(4):
        nz += ((hn >> (4 * i)) & 3).astype(np.int64)
    nz -= 6
    return np.clip(img + nz, 0, 255).astype(np.uint8)
