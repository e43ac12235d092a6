#!/usr/bin/env python
"""Diagnostic parity check of the default (brick) query against the restated oracle on a ladder of cases; prints the first
mismatching lists with their geometry instead of just failing.  Not a test and not a bench: a debugging aid for GPU runs."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))

import numpy as np  # noqa: E402

import cases  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from oracle import loader  # noqa: E402
from treensearch_b200 import clouds  # noqa: E402


def run_case(name, case, options=None):
    eng = t.TreeNSearch()
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    if case["radius"] is not None:
        eng.set_search_radius(case["radius"])
    keep = []
    for (p, r) in case["sets"]:
        keep.append((p, r))
        eng.add_point_set(p, r, variable_radius=(case["radius"] is None))
    for (i, j) in case["pairs"]:
        eng.set_active_search(i, j, True)
    eng.set_symmetric_search(case["symmetric"])
    t0 = time.time()
    eng.run()
    st = eng.stats()
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    ok = True
    for pr in case["pairs"]:
        a = eng.neighbor_csr(*pr)
        b = port.csr(*pr)
        if np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]):
            continue
        ok = False
        ca, cb = np.diff(a[0]), np.diff(b[0])
        bad = np.nonzero(ca != cb)[0]
        print(f"  pair {pr}: {bad.size} lists differ in length (of {ca.size}); totals {a[0][-1]} vs {b[0][-1]}")
        if bad.size == 0:
            # same lengths, different ids
            for i in range(ca.size):
                if not np.array_equal(a[1][a[0][i]:a[0][i + 1]], b[1][b[0][i]:b[0][i + 1]]):
                    bad = np.array([i])
                    break
        pi = case["sets"][pr[0]][0]
        pj = case["sets"][pr[1]][0]
        for i in bad[:4]:
            mine = set(a[1][a[0][i]:a[0][i + 1]].tolist())
            ref = set(b[1][b[0][i]:b[0][i + 1]].tolist())
            miss, extra = sorted(ref - mine)[:6], sorted(mine - ref)[:6]
            print(f"    point {i} at {pi[i]}: mine {len(mine)} ref {len(ref)} missing {miss} extra {extra}")
            for j in miss[:3]:
                print(f"       missing {j} at {pj[j]} d = {np.linalg.norm(pi[i].astype(np.float64) - pj[j].astype(np.float64)):.6g}")
    print(f"{'OK  ' if ok else 'FAIL'} {name}: brick={st['brick_query']} slow={st['n_slow_queries']} nbrs={st['n_neighbors']} reruns={st['n_reruns']} "
          f"query_ms={st['ms_query']:.3f} total_ms={st['ms_total_device']:.3f} wall={time.time() - t0:.2f}s", flush=True)
    return ok


def main():
    all_ok = True
    for name, fn in cases.GOLDEN_CASES.items():
        all_ok &= run_case(name, fn())
    n = 100_000
    all_ok &= run_case("uniform_100k", dict(sets=[(clouds.uniform_cloud(n, 42), None)], radius=float(clouds.radius_for_mean_neighbors(n)), pairs=[(0, 0)], symmetric=True))
    pts, d, r = clouds.dam_break_cloud(300_000)
    all_ok &= run_case("dambreak_300k", dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True))
    p0, r0, p1, r1, _ = clouds.two_set_cloud(200_000, 50_000)
    for sym in (True, False):
        all_ok &= run_case(f"twoset_sym{int(sym)}", dict(sets=[(p0, r0), (p1, r1)], radius=None, pairs=[(0, 0), (0, 1), (1, 0)], symmetric=sym))
    all_ok &= run_case("uniform_5000_limit", cases.GOLDEN_CASES["uniform_fixed_5000"](), {t.TNSB_OPT_QUERY_LIMIT: 3000}) if False else all_ok
    rs = np.random.RandomState(8)
    blob = (0.5 + 0.002 * rs.standard_normal((2600, 3))).astype(np.float32)
    bg = rs.random_sample((3000, 3)).astype(np.float32)
    all_ok &= run_case("dense_blob", dict(sets=[(np.ascontiguousarray(np.concatenate([blob, bg])), None)], radius=0.04, pairs=[(0, 0)], symmetric=True))
    n = 2_000_000
    all_ok &= run_case("uniform_2m", dict(sets=[(clouds.uniform_cloud(n, 42), None)], radius=float(clouds.radius_for_mean_neighbors(n)), pairs=[(0, 0)], symmetric=True))
    print("ALL OK" if all_ok else "SOME FAILED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
