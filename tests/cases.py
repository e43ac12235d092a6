"""Seeded parity cases shared by the golden-fixture generator, the CPU tests and the GPU tests.

Every case is a dict:  sets = [(points float32 (n,3), radii float32 (n,) | None), ...], radius (fixed mode) | None,
pairs = [(set_i, set_j), ...], symmetric = bool.  All inputs are regenerated from seeds; only reference OUTPUTS are stored
under tests/golden/.
"""
from __future__ import annotations

import numpy as np

from treensearch_b200 import clouds


def _lattice_fixed(n):
    pts, r = clouds.sph_lattice(n)
    return dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True)


def _lattice_two_sets(n, scale=1.31):
    # tests/tests.cpp:114-145: second lattice 1.31x coarser, radii constant per set, 1->1 inactive
    p0, r0 = clouds.sph_lattice(n)
    p1, r1 = clouds.sph_lattice(n, scale)
    return dict(sets=[(p0, np.full(p0.shape[0], r0, np.float32)), (p1, np.full(p1.shape[0], r1, np.float32))],
                radius=None, pairs=[(0, 0), (0, 1), (1, 0)], symmetric=True)


def _uniform_fixed(n, seed, k=30.0):
    pts = clouds.uniform_cloud(n, seed)
    return dict(sets=[(pts, None)], radius=float(clouds.radius_for_mean_neighbors(n, k)), pairs=[(0, 0)], symmetric=True)


def _variable_random(n0, n1, seed, symmetric):
    rs = np.random.RandomState(seed)
    p0 = clouds.uniform_cloud(n0, seed)
    p1 = clouds.uniform_cloud(n1, seed + 1).copy()
    p1[:, 2] *= np.float32(0.1)
    r = clouds.radius_for_mean_neighbors(n0, 30.0)
    r0 = (r * (1.0 + 0.5 * rs.random_sample(n0))).astype(np.float32)
    r1 = (r * (0.8 + 0.4 * rs.random_sample(n1))).astype(np.float32)
    return dict(sets=[(p0, r0), (p1, r1)], radius=None, pairs=[(0, 0), (0, 1), (1, 0)], symmetric=symmetric)


def _duplicates(seed):
    # coincident points ARE neighbours; only the identical (set, index) is excluded (TreeNSearch.cpp:2464-2466)
    base = clouds.uniform_cloud(600, seed)
    pts = np.concatenate([base, base[:200], base[:50]], axis=0)
    return dict(sets=[(np.ascontiguousarray(pts), None)], radius=0.12, pairs=[(0, 0)], symmetric=True)


def _three_sets_all(seed):
    # combinatorial_stress_test style (tests/tests.cpp:287-427): coords in [0,10), radii in [0.5,1.0], all searches active
    rs = np.random.RandomState(seed)
    sets = []
    for n in (700, 0, 333):
        p = (rs.random_sample((n, 3)) * 10.0).astype(np.float32)
        r = (0.5 + 0.5 * rs.random_sample(n)).astype(np.float32)
        sets.append((p, r))
    pairs = [(i, j) for i in range(3) for j in range(3)]
    return dict(sets=sets, radius=None, pairs=pairs, symmetric=True)


def _clustered(seed):
    # a dense blob (hundreds of points inside one search radius) next to sparse background: exercises long candidate lists
    rs = np.random.RandomState(seed)
    blob = (0.5 + 0.01 * rs.standard_normal((1500, 3))).astype(np.float32)
    bg = rs.random_sample((2500, 3)).astype(np.float32)
    pts = np.ascontiguousarray(np.concatenate([blob, bg], axis=0))
    return dict(sets=[(pts, None)], radius=0.05, pairs=[(0, 0)], symmetric=True)


GOLDEN_CASES = {
    "lattice_fixed_1": lambda: _lattice_fixed(1),
    "lattice_fixed_100": lambda: _lattice_fixed(100),
    "lattice_fixed_3000": lambda: _lattice_fixed(3000),
    "lattice_two_sets_variable_1000": lambda: _lattice_two_sets(1000),
    "uniform_fixed_5000": lambda: _uniform_fixed(5000, 11),
    "variable_random_sym": lambda: _variable_random(4000, 1000, 21, True),
    "variable_random_asym": lambda: _variable_random(4000, 1000, 21, False),
    "duplicates": lambda: _duplicates(5),
    "three_sets_all_pairs": lambda: _three_sets_all(9),
    "clustered_blob": lambda: _clustered(13),
}


def configure(engine, case, f64_sets=()):
    """Feed a case to anything exposing the reference API subset (oracle loaders or treensearch_b200.TreeNSearch)."""
    if case["radius"] is not None:
        engine.set_search_radius(case["radius"])
    for s, (p, r) in enumerate(case["sets"]):
        if s in f64_sets and hasattr(engine, "add_point_set_f64"):
            engine.add_point_set_f64(p.astype(np.float64), None if r is None else r.astype(np.float64))
        else:
            engine.add_point_set(p, r)
    for (i, j) in case["pairs"]:
        engine.set_active_search(i, j, True)
    engine.set_symmetric_search(case["symmetric"])
    return engine
