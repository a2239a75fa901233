"""Deterministic synthetic particle frames (the reference ships no datasets).

SURVEY.md section 8(d): a "dam-break" = jittered cubic lattice (spacing dx, uniform jitter
+-0.1 dx per axis) filling  {x in [-Lx/2, Lx/2], z in [-Lz/2, Lz/2], -1 <= y <= -1 + H(x, t)}:
a tall reservoir on the -x side and a parabolic tongue running toward +x; the floor is at
y = -1 to match FLOOR_HEIGHT in assets/shaders/advanced/composition.frag:39.
Defaults follow the reference's config: h = particleRadius = 0.1 (assets/config.yml:19),
dx = h/2 (SPlisHSPlasH convention: support radius = 4 x particle radius).
"""
from __future__ import annotations

import hashlib

import numpy as np

SEED = 20240229


def _profile(x: np.ndarray, L: float, t: float) -> np.ndarray:
    """Column height above the floor at abscissa x for scale L and time t in [0, 1]."""
    x_left = -1.2 * L
    x_gate = -0.4 * L                                      # reservoir: [x_left, x_gate]
    x_front = x_gate + 1.6 * L * (0.1 + 0.9 * t)           # tongue front advances with t
    area = L * (0.8 * L + 0.16 * L / 3.0)                  # cross-section conserved over t
    hr = area / ((x_gate - x_left) + (x_front - x_gate) / 3.0)
    xi = np.clip((x - x_gate) / (x_front - x_gate), 0.0, 1.0)
    hgt = np.where(x < x_gate, hr, hr * (1.0 - xi) ** 2)
    return np.where((x < x_left) | (x > x_front), 0.0, hgt)


def dam_break(n_target: int, h: float = 0.1, dx: float | None = None, t: float = 0.6,
              seed: int = SEED) -> np.ndarray:
    """Returns an (N, 3) float32 array with N within a few percent of n_target."""
    dx = h / 2.0 if dx is None else dx
    volume = n_target * dx ** 3
    L = (volume / (0.8 + 0.16 / 3.0)) ** (1.0 / 3.0)      # shape volume = (0.8 + 0.16/3) L^3
    xs = np.arange(-1.2 * L, 1.2 * L + dx, dx)
    zs = np.arange(-0.5 * L, 0.5 * L, dx)
    hx = _profile(xs, L, t)
    ny = np.floor(hx / dx).astype(np.int64)
    pts = []
    for xi, n in zip(xs, ny):
        if n <= 0:
            continue
        ys = -1.0 + 0.5 * dx + dx * np.arange(n)
        yy, zz = np.meshgrid(ys, zs, indexing="ij")
        col = np.empty((yy.size, 3), dtype=np.float64)
        col[:, 0] = xi
        col[:, 1] = yy.ravel()
        col[:, 2] = zz.ravel()
        pts.append(col)
    p = np.concatenate(pts, axis=0)
    rng = np.random.Generator(np.random.PCG64(seed))
    p += rng.uniform(-0.1 * dx, 0.1 * dx, size=p.shape)
    return np.ascontiguousarray(p.astype(np.float32))


def random_block(n: int, extent: float = 0.6, seed: int = SEED) -> np.ndarray:
    """Uniform random block, the pattern of Dataset::makeCube (src/app/Dataset.cpp:229-263)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    return np.ascontiguousarray(p)


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# BASELINE.json configs -> (particles, width, height)
CONFIGS = {
    "C1": dict(n_target=64_000, width=1280, height=720),
    "C2": dict(n_target=1_000_000, width=1920, height=1080),
    "C3": dict(n_target=4_000_000, width=3840, height=2160),
}
