#!/usr/bin/env python
"""Differential fuzzing of the CUDA engine against the brute-force port: random set counts / sizes / clustering / radii / active pairs /
options, several run() calls per engine with points moved in place (speculative grid, graph replay, re-runs).  Not a test of the suite
(run time is open ended): `python tools/fuzz_parity.py [seconds] [seed]`, prints the first failing configuration and exits 1."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np  # noqa: E402

import cases  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from oracle import loader  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rs = np.random.RandomState(seed)
t_end = time.time() + budget
n_cases = n_runs = 0


def cloud(n, kind, scale):
    if kind == 0:
        p = rs.random_sample((n, 3))
    elif kind == 1:                                   # thin slab
        p = rs.random_sample((n, 3)); p[:, rs.randint(0, 3)] *= 0.03
    elif kind == 2:                                   # blobs
        c = rs.random_sample((max(1, n // 400 + 1), 3))
        p = c[rs.randint(0, c.shape[0], n)] + 0.01 * rs.standard_normal((n, 3))
    elif kind == 3:                                   # lattice with duplicates
        g = int(round(max(n, 1) ** (1.0 / 3.0))) + 1
        p = rs.randint(0, g, (n, 3)) / float(g)
    else:                                             # far away from the origin
        p = rs.random_sample((n, 3)) + 50.0
    return np.ascontiguousarray((p * scale).astype(np.float32))


def compare(eng, case, tag):
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    for pr in case["pairs"]:
        a, b = eng.neighbor_csr(*pr), port.csr(*pr)
        if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])):
            print("MISMATCH", tag, "pair", pr, flush=True)
            return False
    return True


while time.time() < t_end:
    n_sets = int(rs.choice([1, 1, 1, 2, 3]))
    variable = bool(rs.randint(0, 2)) or n_sets > 1 and bool(rs.randint(0, 2))
    scale = float(rs.choice([1.0, 1.0, 10.0, 0.01]))
    sizes = [int(rs.choice([0, 1, 5, 33, 300, 3000, 20000], p=[0.05, 0.05, 0.1, 0.1, 0.2, 0.3, 0.2])) for _ in range(n_sets)]
    kinds = [int(rs.randint(0, 5)) for _ in range(n_sets)]
    n_ref = max(max(sizes), 50)
    r = scale * float(np.cbrt(rs.choice([8.0, 30.0, 60.0, 200.0]) / (n_ref * 4.19)))
    sets = []
    for n, k in zip(sizes, kinds):
        p = cloud(n, k, scale)
        rad = (r * (0.7 + 0.6 * rs.random_sample(n))).astype(np.float32) if variable else None
        sets.append((p, rad))
    pairs = [(i, j) for i in range(n_sets) for j in range(n_sets) if rs.random_sample() < 0.7] or [(0, 0)]
    sym = bool(rs.randint(0, 2))
    opts = {}
    if rs.random_sample() < 0.2:
        opts[t.TNSB_OPT_QUERY_KERNEL] = 1
    if rs.random_sample() < 0.15:
        opts[t.TNSB_OPT_BUILD] = 1
    if rs.random_sample() < 0.2:
        opts[t.TNSB_OPT_LIST_CAPACITY] = int(rs.choice([1, 8]))
    if rs.random_sample() < 0.3:
        opts[t.TNSB_OPT_ZERO_COPY_RESULTS] = 0
    if rs.random_sample() < 0.3:
        opts[t.TNSB_OPT_SORT_LISTS] = int(rs.choice([0, 1]))
    case = dict(sets=sets, radius=None if variable else r, pairs=pairs, symmetric=sym)
    tag = dict(seed=seed, case=n_cases, sizes=sizes, kinds=kinds, variable=variable, scale=scale, r=r, pairs=pairs, sym=sym, opts=opts)
    eng = t.TreeNSearch()
    for k, v in opts.items():
        eng.set_option(k, v)
    if not variable:
        eng.set_search_radius(r)
    for (p, rad) in sets:
        eng.add_point_set(p, rad, variable_radius=variable)
    for (i, j) in pairs:
        eng.set_active_search(i, j, True)
    eng.set_symmetric_search(sym)
    ok = True
    for step in range(int(rs.choice([1, 2, 5]))):
        if step > 0:
            for (p, rad) in sets:                       # move in place (same pointers): small drift, sometimes a jump out of the grid
                p += (scale * (0.0005 if rs.random_sample() < 0.8 else 0.2) * rs.standard_normal(p.shape)).astype(np.float32)
        eng.run()
        n_runs += 1
        if sum(sizes) <= 26000 or step == 0:
            ok = compare(eng, case, dict(tag, step=step, stats={k: eng.stats()[k] for k in ("brick_query", "n_slow_queries", "n_reruns", "graph_replay", "speculative_grid")}))
        if not ok:
            break
    eng.close()
    n_cases += 1
    if not ok:
        print(tag, flush=True)
        sys.exit(1)
print(f"fuzz ok: {n_cases} configurations, {n_runs} runs, seed {seed}, {budget:.0f} s")
