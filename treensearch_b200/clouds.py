"""Synthetic point clouds for the configurations named in BASELINE.json / SURVEY.md §8d.

All generators are deterministic functions of their seed and return float32 arrays of shape (N, 3).
`uniform_cloud` reproduces ``std::mt19937 gen(seed); std::uniform_real_distribution<float>(0,1)`` filling
``P[3N]`` in xyzxyz order bit-for-bit (numpy's legacy RandomState uses the same MT19937 seeding, and libstdc++'s
generate_canonical<float,24> is one 32-bit draw scaled by 2^-32 in float).
"""
from __future__ import annotations

import math

import numpy as np


def radius_for_mean_neighbors(n_points: int, k_mean: float = 30.0, volume: float = 1.0) -> np.float32:
    """r such that a sphere of radius r holds k_mean points on average: r = cbrt(k / (n/V * 4pi/3))."""
    return np.float32(np.cbrt(k_mean * volume / (n_points * 4.18879020479)))


def _mt_uniform01(rs: np.random.RandomState, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.float32)
    step = 1 << 24
    for a in range(0, count, step):
        b = min(count, a + step)
        u = rs.randint(0, 2 ** 32, b - a, dtype=np.uint64).astype(np.uint32)
        f = u.astype(np.float32) / np.float32(4294967296.0)
        f[f >= 1.0] = np.nextafter(np.float32(1.0), np.float32(0.0))   # libstdc++ clamps the rounding-up case
        out[a:b] = f
    return out


def uniform_cloud(n_points: int, seed: int = 42, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """C1 / C2 / C5: uniform random points in [lo, hi)^3."""
    rs = np.random.RandomState(seed)
    u = _mt_uniform01(rs, 3 * n_points)
    if lo != 0.0 or hi != 1.0:
        u = (u * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
    return u.reshape(n_points, 3)


def lattice_cloud(bottom, top, spacing) -> tuple[np.ndarray, np.float32]:
    """The reference tests' fixture generator (tests/tests.cpp:16-32): regular lattice with float accumulation
    of the coordinate, x outermost / z innermost, and search_radius = 1.99 * spacing."""
    spacing = np.float32(spacing)

    def axis(b, t):
        vals = []
        x = np.float32(b)
        t = np.float32(t)
        while x <= t:
            vals.append(x)
            x = np.float32(x + spacing)
        return np.array(vals, dtype=np.float32)

    ax, ay, az = axis(bottom[0], top[0]), axis(bottom[1], top[1]), axis(bottom[2], top[2])
    X, Y, Z = np.meshgrid(ax, ay, az, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    return np.ascontiguousarray(pts), np.float32(np.float32(1.99) * spacing)


def sph_lattice(n_points: int, scale: float = 1.0):
    """Lattice in [-1,1]^3 with spacing scale * 2/cbrt(n) exactly as the reference tests build it
    (tests/tests.cpp:95-97, :119-121)."""
    particle_radius = np.float32(2.0 / math.pow(float(n_points), 1.0 / 3.0))
    return lattice_cloud((-1, -1, -1), (1, 1, 1), np.float32(np.float32(scale) * particle_radius))


def dam_break_cloud(n_points: int, seed: int = 1234):
    """C3: SPH dam-break-like clustered cloud: jittered lattice of spacing d filling a fluid column
    x in [0, 0.25 L], y in [0, H], z in [0, 0.5 W] of a 4:2:1 (L:H:W) tank plus a 2-layer pool on the floor.
    Returns (points, d, search_radius = 2.43 d)."""
    L, H, W = 4.0, 2.0, 1.0
    # solve for d: column cells + pool cells ~= n_points
    col_vol = (0.25 * L) * H * (0.5 * W)

    def count(d):
        nx, ny, nz = int(0.25 * L / d), int(H / d), int(0.5 * W / d)
        px, pz = int(L / d), int(W / d)
        pool = (px - nx) * 2 * pz + nx * 2 * max(pz - nz, 0)
        return nx * ny * nz + pool, (nx, ny, nz, px, pz)

    d = (col_vol / n_points) ** (1.0 / 3.0)
    for _ in range(60):
        c, _dims = count(d)
        d *= (c / n_points) ** (1.0 / 3.0)
    c, (nx, ny, nz, px, pz) = count(d)
    d32 = np.float32(d)
    gx, gy, gz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    col = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    # pool: 2 layers (y = 0, 1) everywhere on the floor outside the column footprint
    fx, fy, fz = np.meshgrid(np.arange(px), np.arange(2), np.arange(pz), indexing="ij")
    pool = np.stack([fx.ravel(), fy.ravel(), fz.ravel()], axis=1)
    outside = ~((pool[:, 0] < nx) & (pool[:, 2] < nz))
    pool = pool[outside]
    cells = np.concatenate([col, pool], axis=0)
    rs = np.random.RandomState(seed)
    jitter = (rs.random_sample((cells.shape[0], 3)) * 0.4 - 0.2)
    pts = ((cells + 0.5 + jitter) * d).astype(np.float32)
    if pts.shape[0] > n_points:
        pts = pts[:n_points]
    elif pts.shape[0] < n_points:          # top up with extra column points (keeps N exact)
        extra = n_points - pts.shape[0]
        e = rs.random_sample((extra, 3)) * np.array([0.25 * L, H, 0.5 * W])
        pts = np.concatenate([pts, e.astype(np.float32)], axis=0)
    return np.ascontiguousarray(pts), d32, np.float32(np.float32(2.43) * d32)


def advect(points: np.ndarray, d: float, step: int) -> np.ndarray:
    """One pseudo time step for C3: displace every point by a smooth divergence-free field with |u| <= 0.1 d."""
    p = points.astype(np.float64)
    a = 0.1 * float(d) / math.sqrt(2.0)
    ph = 0.37 * step
    ux = a * np.sin(2.0 * p[:, 1] + ph) * np.cos(3.0 * p[:, 2])
    uy = a * np.sin(2.0 * p[:, 2] + ph) * np.cos(3.0 * p[:, 0])
    uz = a * np.sin(2.0 * p[:, 0] + ph) * np.cos(3.0 * p[:, 1])
    return (p + np.stack([ux, uy, uz], axis=1)).astype(np.float32)


def two_set_cloud(n0: int = 2_000_000, n1: int = 500_000, seed: int = 7, k_mean: float = 30.0):
    """C4: set0 uniform in the unit cube, set1 on a thin slab z in [0, 0.05]; per-point radii r*U[1,1.5] (set0) and
    r*U[0.8,1.2] (set1) with r from k_mean at n0.  Returns (p0, r0, p1, r1, r)."""
    rs = np.random.RandomState(seed)
    r = radius_for_mean_neighbors(n0, k_mean)
    p0 = _mt_uniform01(rs, 3 * n0).reshape(n0, 3)
    p1 = _mt_uniform01(rs, 3 * n1).reshape(n1, 3).copy()
    p1[:, 2] *= np.float32(0.05)
    r0 = (r * (np.float32(1.0) + np.float32(0.5) * _mt_uniform01(rs, n0))).astype(np.float32)
    r1 = (r * (np.float32(0.8) + np.float32(0.4) * _mt_uniform01(rs, n1))).astype(np.float32)
    return np.ascontiguousarray(p0), r0, np.ascontiguousarray(p1), r1, r
