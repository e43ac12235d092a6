"""CPU tests (-m "not gpu"): the restated oracle against the golden fixtures generated from the unmodified reference,
and -- where oracle/_ref is present -- against the reference itself."""
import numpy as np
import pytest

import cases
from conftest import csr_equal
from oracle import loader
from treensearch_b200 import clouds


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
@pytest.mark.parametrize("mode", [0, 1])
def test_port_matches_golden(golden, name, mode):
    case = cases.GOLDEN_CASES[name]()
    if mode == 0 and sum(p.shape[0] for p, _ in case["sets"]) > 6000:
        pytest.skip("all-pairs mode only for small cases")
    port = cases.configure(loader.OraclePort(), case)
    port.run(mode)
    for (i, j) in case["pairs"]:
        off, idx = port.csr(i, j)
        assert np.array_equal(off, golden[f"{name}/{i}_{j}/offsets"]), (name, i, j)
        assert np.array_equal(idx, golden[f"{name}/{i}_{j}/indices"]), (name, i, j)


def test_golden_covers_every_case(golden):
    for name, make in cases.GOLDEN_CASES.items():
        for (i, j) in make()["pairs"]:
            assert f"{name}/{i}_{j}/offsets" in golden


@pytest.mark.skipif(not loader.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
class TestAgainstReference:
    def test_c1_100k_uniform(self):
        # BASELINE.json configs[0]: 100K uniform, fixed radius, run() == run_scalar() == restated grid oracle
        n = 100_000
        pts = clouds.uniform_cloud(n, 42)
        r = clouds.radius_for_mean_neighbors(n)
        case = dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True)
        ref = cases.configure(loader.Reference(), case)
        ref.run(0)
        a = ref.csr(0, 0)
        assert a[0][-1] == 2864970          # k_mean = 28.65
        assert ref.n_unsorted_lists == 0    # reference lists are ascending
        ref.run(1)
        assert csr_equal(a, ref.csr(0, 0))
        port = cases.configure(loader.OraclePort(), case)
        port.run(1)
        assert csr_equal(a, port.csr(0, 0))

    def test_mixed_float_double(self):
        # tests/tests.cpp:147-186: one float set + one double set
        case = cases.GOLDEN_CASES["lattice_two_sets_variable_1000"]()
        ref = cases.configure(loader.Reference(), case, f64_sets=(1,))
        ref.run(0)
        port = cases.configure(loader.OraclePort(), case)
        port.run(1)
        for p in case["pairs"]:
            assert csr_equal(ref.csr(*p), port.csr(*p))

    def test_reference_zsort_is_a_morton_order(self):
        pts = clouds.uniform_cloud(20000, 3).copy()
        case = dict(sets=[(pts, None)], radius=0.05, pairs=[(0, 0)], symmetric=True)
        ref = cases.configure(loader.Reference(), case)
        ref.prepare_zsort()
        order = ref.zsort_order(0)
        assert np.array_equal(np.sort(order), np.arange(20000))


def test_morton_bit_order():
    # libmorton: x -> bit 0, y -> bit 1, z -> bit 2 (morton_BMI.h:40-52)
    assert loader.morton3d_64(1, 0, 0) == 1
    assert loader.morton3d_64(0, 1, 0) == 2
    assert loader.morton3d_64(0, 0, 1) == 4
    assert loader.morton3d_64(2, 0, 0) == 8
    assert loader.morton3d_64(0x1fffff, 0x1fffff, 0x1fffff) == (1 << 63) - 1


def test_lattice_generator_matches_reference_fixture():
    # tests/tests.cpp:16-32 at n = 100: spacing 2/cbrt(100), 5 points per axis, r = 1.99 * spacing
    pts, r = clouds.sph_lattice(100)
    assert pts.shape == (125, 3)
    assert np.isclose(r, 1.99 * 2.0 / 100 ** (1 / 3), rtol=1e-6)
    assert np.all(pts[0] == -1.0)
    # x outermost, z innermost
    assert pts[1, 2] > pts[0, 2] and pts[1, 0] == pts[0, 0]


def test_digest_is_order_independent():
    ids = np.array([5, 1, 9, 9, 1, 5, 7], dtype=np.int32)
    pos = np.array([0, 3, 6], dtype=np.int64)
    cnt = np.array([3, 3, 1], dtype=np.int32)
    d = loader.list_digests(ids, pos, cnt)
    assert d[0] == d[1] and d[0] != d[2]
