"""Generates tests/golden/golden.npz from the UNMODIFIED reference (oracle/_ref/libtns_ref.so, built by oracle/Makefile
from /root/reference).  Run in the build container only:   python tests/golden/make_golden.py

For every case of tests/cases.py::GOLDEN_CASES and every active pair the file stores the reference's run() result as CSR
(offsets int64, indices int32, each list ascending -- the reference's own comparator sorts, tests/BruteforceNSearch.cpp:135).
Before writing, run() is cross-checked against run_scalar() and BruteforceNSearch on the same inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.loader import Reference  # noqa: E402
import cases  # noqa: E402


def main():
    out = {}
    for name, make in cases.GOLDEN_CASES.items():
        case = make()
        results = []
        for mode in (0, 1, 2):
            ref = cases.configure(Reference(), case)
            ref.run(mode)
            results.append({p: ref.csr(*p) for p in case["pairs"]})
        for p in case["pairs"]:
            for other, label in ((1, "run_scalar"), (2, "BruteforceNSearch")):
                same = np.array_equal(results[0][p][0], results[other][p][0]) and np.array_equal(results[0][p][1], results[other][p][1])
                if not same:
                    raise SystemExit(f"{name} pair {p}: run() != {label}")
            off, idx = results[0][p]
            out[f"{name}/{p[0]}_{p[1]}/offsets"] = off
            out[f"{name}/{p[0]}_{p[1]}/indices"] = idx
            print(f"{name:34s} pair {p}: n={off.shape[0] - 1:6d} K={off[-1]:8d}  (run == run_scalar == BruteforceNSearch)")
    path = os.path.join(os.path.dirname(__file__), "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
