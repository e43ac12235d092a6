"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (via the Python mirror of the reference class),
against the restated oracle, the golden fixtures generated from the unmodified reference and -- when oracle/_ref travelled
with the snapshot -- the reference itself.  Bar: bit-exact integer index sets per (set_i, set_j, i)."""
import numpy as np
import pytest

import cases
from conftest import csr_equal
from oracle import loader
from treensearch_b200 import clouds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tnsb(built_library):
    import treensearch_b200 as t
    return t


def run_engine(t, case, f64_sets=(), options=None):
    eng = t.TreeNSearch()
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    if case["radius"] is not None:
        eng.set_search_radius(case["radius"])
    for s, (p, r) in enumerate(case["sets"]):
        if s in f64_sets:
            p = p.astype(np.float64)
            r = None if r is None else r.astype(np.float64)
        eng.add_point_set(p, r, variable_radius=(case["radius"] is None))
        eng._test_keep = getattr(eng, "_test_keep", []) + [(p, r)]
    for (i, j) in case["pairs"]:
        eng.set_active_search(i, j, True)
    eng.set_symmetric_search(case["symmetric"])
    eng.run()
    return eng


def assert_matches_port(eng, case, pairs=None):
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    for p in (pairs or case["pairs"]):
        a = eng.neighbor_csr(*p)
        b = port.csr(*p)
        assert np.array_equal(a[0], b[0]), f"pair {p}: neighbour counts differ"
        assert np.array_equal(a[1], b[1]), f"pair {p}: neighbour ids differ"


# ---------------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_golden(tnsb, golden, name):
    case = cases.GOLDEN_CASES[name]()
    eng = run_engine(tnsb, case)
    for (i, j) in case["pairs"]:
        off, idx = eng.neighbor_csr(i, j)
        assert np.array_equal(off, golden[f"{name}/{i}_{j}/offsets"]), (name, i, j)
        assert np.array_equal(idx, golden[f"{name}/{i}_{j}/indices"]), (name, i, j)


def test_neighborlist_handle_layout(tnsb):
    # NeighborList is a view of [n, j0, j1 ...] (NeighborList.h:8-39): size() is the word in front of the ids
    case = cases.GOLDEN_CASES["lattice_fixed_100"]()
    eng = run_engine(tnsb, case)
    ragged, pos = eng.neighbor_lists(0, 0)
    for i in (0, 7, 124):
        nl = eng.get_neighborlist(0, 0, i)
        assert nl.size() == ragged[pos[i]] == len(list(nl))
        got = []
        eng.for_each_neighbor(0, 0, i, got.append)
        assert got == list(nl)
        assert i not in got
    assert eng.get_neighborlist_n_bytes() == 4 * ragged.shape[0]


# ---------------------------------------------------------------------------------------------------- reference test cases
@pytest.mark.parametrize("n", [1, 100, 10000])
def test_one_set_fixed_radius(tnsb, n):            # tests/tests.cpp:91-112
    pts, r = clouds.sph_lattice(n)
    case = dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True)
    assert_matches_port(run_engine(tnsb, case), case)


@pytest.mark.parametrize("n", [1, 100, 10000])
def test_two_dynamic_sets_variable_radius(tnsb, n):   # tests/tests.cpp:114-145
    case = cases._lattice_two_sets(n)
    eng = run_engine(tnsb, case)
    assert_matches_port(eng, case)
    assert not eng.is_search_active(1, 1)
    with pytest.raises(tnsb.TreeNSearchError):
        eng.neighbor_lists(1, 1)


@pytest.mark.parametrize("n", [100, 10000])
def test_mixed_float_double_point_sets(tnsb, n):      # tests/tests.cpp:147-186
    case = cases._lattice_two_sets(n, 1.33)
    assert_matches_port(run_engine(tnsb, case, f64_sets=(1,)), case)
    assert_matches_port(run_engine(tnsb, case, f64_sets=(0, 1)), case)


def test_resize_variable_radius(tnsb):                # tests/tests.cpp:188-237
    full = cases._lattice_two_sets(10000)
    (p0, r0), (p1, r1) = full["sets"]
    eng = tnsb.TreeNSearch()
    eng.add_point_set(p0, r0, n_points=p0.shape[0] // 2)
    eng.add_point_set(p1, r1, n_points=p1.shape[0] // 2)
    for p in full["pairs"]:
        eng.set_active_search(*p, True)
    for frac in (2, 1, 3):
        n0, n1 = p0.shape[0] // frac, p1.shape[0] // frac
        eng.resize_point_set(0, p0, r0, n_points=n0)
        eng.resize_point_set(1, p1, r1, n_points=n1)
        eng.run()
        assert eng.get_n_points_in_set(0) == n0 and eng.get_total_n_points() == n0 + n1
        sub = dict(full, sets=[(p0[:n0], r0[:n0]), (p1[:n1], r1[:n1])])
        assert_matches_port(eng, sub)


def test_combinatorial_counts_with_zsort(tnsb):
    """combinatorial_stress_test of the reference (tests/tests.cpp:287-427): every combination of "interesting" particle counts over
    1, 2 and 3 variable-radius sets (coordinates in [0, 10), radii in [0.5, 1.0], all searches active), run, prepare_zsort +
    apply_zsort of positions and radii, run again.  The reference only looks for crashes there (its comparison is compiled out);
    here both runs of every case are compared with the brute-force port.  The count lists are thinned for 2 and 3 sets (the full
    cross product is 175 000 engine constructions)."""
    rs = np.random.RandomState(42)
    counts1 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 23, 24, 25, 100, 1000, 10000, 10001, 10009]
    counts2 = [0, 1, 2, 8, 9, 17, 100, 1000, 10007]
    counts3 = [0, 1, 17, 1000]
    combos = [(a,) for a in counts1] + [(a, b) for a in counts2 for b in counts2] + [(a, b, c) for a in counts3 for b in counts3 for c in counts3]
    for counts in combos:
        sets = []
        for n in counts:
            p = (rs.random_sample((n, 3)) * 10.0).astype(np.float32)
            r = (0.5 + 0.5 * rs.random_sample(n)).astype(np.float32)
            sets.append((p, r))
        k = len(counts)
        pairs = [(i, j) for i in range(k) for j in range(k)]
        case = dict(sets=sets, radius=None, pairs=pairs, symmetric=True)
        eng = tnsb.TreeNSearch()
        for (p, r) in sets:
            eng.add_point_set(p, r, variable_radius=True)
        eng.set_all_searches(True)
        eng.run()
        check = sum(counts) <= 12000 or len(counts) == 1          # the brute-force port is quadratic: the largest pairs are run, not compared
        if check:
            assert_matches_port(eng, case)
        eng.prepare_zsort()
        for s, (p, r) in enumerate(sets):
            if counts[s] > 0:
                eng.apply_zsort(s, p, 3)
                eng.apply_zsort(s, r, 1)
        eng.run()
        if check:
            assert_matches_port(eng, case)
        assert eng.stats()["n_queries"] == sum(counts) * k
        eng.close()


def test_dynamic_emitter_style_sequence(tnsb):       # tests/tests.cpp:434-514 (shortened): add / remove / replace with empty sets
    rs = np.random.RandomState(123)
    eng = tnsb.TreeNSearch()
    empty = np.zeros((0, 3), np.float32)
    for _ in range(2):
        eng.add_point_set(empty, np.zeros(0, np.float32))
    eng.set_all_searches(True)
    store = [None, None]
    sizes = [0, 0]
    for it in range(40):
        s = int(rs.randint(0, 2))
        action = int(rs.randint(0, 3))
        amount = int(rs.randint(1, 21))
        sizes[s] = sizes[s] + amount if action == 0 else (max(0, sizes[s] - amount) if action == 1 else amount)
        p = (rs.random_sample((sizes[s], 3)) * 10.0).astype(np.float32)
        r = np.full(sizes[s], 0.5, np.float32)
        store[s] = (p, r)
        eng.resize_point_set(s, p, r)
        eng.run()
        sets = [st if st is not None else (empty, np.zeros(0, np.float32)) for st in store]
        case = dict(sets=sets, radius=None, pairs=[(0, 0), (0, 1), (1, 0), (1, 1)], symmetric=True)
        assert_matches_port(eng, case)


# ---------------------------------------------------------------------------------------------------- configs of BASELINE.json
def test_c1_100k_uniform(tnsb):
    n = 100_000
    pts = clouds.uniform_cloud(n, 42)
    case = dict(sets=[(pts, None)], radius=float(clouds.radius_for_mean_neighbors(n)), pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case)
    a = eng.neighbor_csr(0, 0)
    assert a[0][-1] == 2864970
    assert_matches_port(eng, case)
    if loader.reference_available():
        ref = cases.configure(loader.Reference(), case)
        ref.run(0)
        assert csr_equal(a, ref.csr(0, 0))
        ref.run(2)                                   # BruteforceNSearch, ~6 s on 8 cores
        assert csr_equal(a, ref.csr(0, 0))


def _digest_compare(eng, pair, ref_off, ref_idx):
    ragged, pos = eng.neighbor_lists(*pair)
    cnt = ragged[pos]
    assert np.array_equal(cnt.astype(np.int64), np.diff(ref_off)), "neighbour counts differ"
    mine = loader.list_digests(ragged, pos + 1, cnt)
    theirs = loader.csr_digests(ref_off, ref_idx)
    bad = np.nonzero(mine != theirs)[0]
    assert bad.size == 0, f"{bad.size} neighbour lists differ, first at point {bad[:5]}"


def test_c2_10m_uniform(tnsb):
    n = 10_000_000
    pts = clouds.uniform_cloud(n, 42)
    r = clouds.radius_for_mean_neighbors(n)
    case = dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case)
    st = eng.stats()
    assert st["n_neighbors"] == 296973820            # K at 10M, seed 42
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    _digest_compare(eng, (0, 0), *port.csr(0, 0))
    if loader.reference_available():
        ref = cases.configure(loader.Reference(), case)
        ref.run(0)
        _digest_compare(eng, (0, 0), *ref.csr(0, 0, sort_lists=False))
    # idempotence: a second run on the same data gives the same sets
    ragged, pos = eng.neighbor_lists(0, 0)
    d1 = loader.list_digests(ragged, pos + 1, ragged[pos])
    eng.run()
    ragged, pos = eng.neighbor_lists(0, 0)
    assert np.array_equal(d1, loader.list_digests(ragged, pos + 1, ragged[pos]))


def _c3_loop(tnsb, n, n_steps, zsort_every, check_steps, use_reference):
    """BASELINE config 3: dam-break cloud, prepare_zsort + apply_zsort before step 0 and every `zsort_every` steps, points advected
    by <= 0.1 d per step; digests of all lists against the reference / the port on `check_steps`, neighbour totals on
    every step."""
    pts, d, r = clouds.dam_break_cloud(n)
    pts = pts.copy()
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(float(r))
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    for step in range(n_steps):
        if step % zsort_every == 0:
            eng.prepare_zsort()
            eng.apply_zsort(0, pts, 3)
        eng.run()
        if step in check_steps:
            case = dict(sets=[(pts, None)], radius=float(r), pairs=[(0, 0)], symmetric=True)
            if use_reference:
                ref = cases.configure(loader.Reference(), case)
                ref.run(0)
                _digest_compare(eng, (0, 0), *ref.csr(0, 0, sort_lists=False))
                assert ref.pair_total(0, 0) == eng.stats()["n_neighbors"]
                ref.close()
            else:
                port = cases.configure(loader.OraclePort(), case)
                port.run(1)
                _digest_compare(eng, (0, 0), *port.csr(0, 0))
                port.close()
        pts[...] = clouds.advect(pts, d, step)
    assert eng.stats()["n_neighbors"] / n > 40
    return eng


def test_c3_dam_break_with_zsort(tnsb):
    _c3_loop(tnsb, 2_000_000, 3, 2, (0, 1, 2), use_reference=False)


def test_c3_10m_dam_break_zsort_every_10_vs_reference(tnsb):
    """Config 3 at its stated size: 10M points, 20 steps, zsort before steps 0 and 10, against the unmodified reference."""
    if not loader.reference_available():
        pytest.skip("oracle/_ref not present")
    eng = _c3_loop(tnsb, 10_000_000, 20, 10, (0, 9, 10, 19), use_reference=True)
    assert eng.stats()["brick_query"] == 1


def test_c4_two_sets_variable_radii(tnsb):
    p0, r0, p1, r1, _ = clouds.two_set_cloud(400_000, 100_000)
    for sym in (True, False):
        case = dict(sets=[(p0, r0), (p1, r1)], radius=None, pairs=[(0, 0), (0, 1), (1, 0)], symmetric=sym)
        eng = run_engine(tnsb, case)
        port = cases.configure(loader.OraclePort(), case)
        port.run(1)
        for p in case["pairs"]:
            _digest_compare(eng, p, *port.csr(*p))


@pytest.mark.parametrize("sym", [True, False])
def test_c4_2m_500k_vs_reference(tnsb, sym):
    """Config 4 at its stated size: 2M fluid + 500K boundary points, per-point radii, searches 0->0, 0->1, 1->0, against the reference."""
    if not loader.reference_available():
        pytest.skip("oracle/_ref not present")
    p0, r0, p1, r1, _ = clouds.two_set_cloud(2_000_000, 500_000)
    case = dict(sets=[(p0, r0), (p1, r1)], radius=None, pairs=[(0, 0), (0, 1), (1, 0)], symmetric=sym)
    eng = run_engine(tnsb, case)
    eng.run()                                       # second run: column height picked from the first run's longest list
    ref = cases.configure(loader.Reference(), case)
    ref.run(0)
    for p in case["pairs"]:
        _digest_compare(eng, p, *ref.csr(*p, sort_lists=False))
    assert not eng.is_search_active(1, 1)


# ---------------------------------------------------------------------------------------------------- engine behaviour
def test_empty_inputs(tnsb):
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(0.1)
    eng.add_point_set(np.zeros((0, 3), np.float32))
    eng.set_active_search(0, 0, True)
    eng.run()
    ragged, pos = eng.neighbor_lists(0, 0)
    assert ragged.shape[0] == 0 and pos.shape[0] == 0
    eng.prepare_zsort()
    assert eng.get_zsort_order(0).shape[0] == 0
    # no sets at all
    eng2 = tnsb.TreeNSearch()
    eng2.set_search_radius(0.1)
    eng2.run()
    assert eng2.get_n_sets() == 0 and eng2.get_total_n_points() == 0


def test_64bit_keys_sparse_domain(tnsb):
    # radius tiny against the extent: > 1024 cells per axis -> 63-bit Morton keys
    rs = np.random.RandomState(4)
    centers = (rs.random_sample((300, 3)) * 100.0).astype(np.float32)
    pts = np.ascontiguousarray((centers[:, None, :] + 0.02 * rs.standard_normal((300, 20, 3)).astype(np.float32)).reshape(-1, 3))
    case = dict(sets=[(pts, None)], radius=0.03, pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case)
    assert eng.stats()["key_bits"] > 30
    assert_matches_port(eng, case)


def test_long_lists_and_dense_cells(tnsb):
    # one list longer than the per-warp staging buffer (> 2047 ids) and candidate lists longer than the register path
    rs = np.random.RandomState(8)
    blob = (0.5 + 0.002 * rs.standard_normal((2600, 3))).astype(np.float32)
    bg = rs.random_sample((3000, 3)).astype(np.float32)
    pts = np.ascontiguousarray(np.concatenate([blob, bg]))
    case = dict(sets=[(pts, None)], radius=0.04, pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case)
    assert eng.stats()["n_neighbors"] > 2600 * 2500
    assert_matches_port(eng, case)


def test_list_buffer_overflow_rerun(tnsb):
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_LIST_CAPACITY: 1})
    assert eng.stats()["n_reruns"] >= 1
    assert_matches_port(eng, case)


def test_sorted_lists_option(tnsb):
    case = cases.GOLDEN_CASES["clustered_blob"]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_SORT_LISTS: 1})
    ragged, pos = eng.neighbor_lists(0, 0)
    off, idx = eng.neighbor_csr(0, 0, sort_lists=False)      # stored order
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    assert csr_equal((off, idx), port.csr(0, 0))            # ascending exactly like the reference's lists


def test_host_lists_are_ascending_by_default(tnsb, golden):
    """TNSB_OPT_SORT_LISTS = -1 (default): lists mirrored to the host come out ascending, exactly the reference's sequences
    (SURVEY.md §0.6); device-resident lists stay in traversal order unless the option is 1."""
    for name in ("uniform_fixed_5000", "variable_random_sym", "duplicates"):
        case = cases.GOLDEN_CASES[name]()
        eng = run_engine(tnsb, case)
        assert eng.stats()["brick_query"] == 1
        for (i, j) in case["pairs"]:
            off, idx = eng.neighbor_csr(i, j, sort_lists=False)          # stored order
            assert np.array_equal(off, golden[f"{name}/{i}_{j}/offsets"]) and np.array_equal(idx, golden[f"{name}/{i}_{j}/indices"]), (name, i, j)
    # long lists (two ranking registers per lane and more) on the dense blob
    case = cases.GOLDEN_CASES["clustered_blob"]()
    eng = run_engine(tnsb, case)
    eng.run()
    off, idx = eng.neighbor_csr(0, 0, sort_lists=False)
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    poff, pidx = port.csr(0, 0)
    cnt = np.diff(off)
    fast = np.nonzero(cnt <= 96)[0]                  # lists that went through the hit columns (the slow path writes traversal order)
    assert np.array_equal(off, poff)
    for i in fast[:: max(1, fast.size // 400)]:
        assert np.array_equal(idx[off[i]:off[i + 1]], pidx[poff[i]:poff[i + 1]])


def test_query_limit_halo_points(tnsb):
    # points with index >= limit are find-only (halo of a Z-slab shard): the first `limit` lists equal the full search
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_QUERY_LIMIT: 3000})
    off, idx = eng.neighbor_csr(0, 0)
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    poff, pidx = port.csr(0, 0)
    assert off.shape[0] == 3001
    assert np.array_equal(off, poff[:3001]) and np.array_equal(idx, pidx[: poff[3000]])


def test_device_resident_io(tnsb):
    import torch
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    pts = torch.from_numpy(case["sets"][0][0]).cuda()
    eng = tnsb.TreeNSearch()
    eng.set_option(tnsb.TNSB_OPT_HOST_RESULTS, 0)
    eng.set_search_radius(case["radius"])
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    st = eng.stats()
    assert st["h2d_bytes"] == 0 and st["d2h_bytes"] == 0
    with pytest.raises(tnsb.TreeNSearchError):
        eng.neighbor_lists(0, 0)
    d_ragged, d_pos, n_ints = eng.neighbor_lists_device(0, 0)
    assert n_ints == st["n_list_ints"] >= st["n_neighbors"] + 5000      # + alignment padding between flushes
    # pinned host input takes the same path as pageable input
    pinned = torch.from_numpy(case["sets"][0][0]).pin_memory()
    eng2 = tnsb.TreeNSearch()
    eng2.set_search_radius(case["radius"])
    eng2.add_point_set(pinned)
    eng2.set_active_search(0, 0, True)
    eng2.run()
    assert eng2.stats()["n_neighbors"] == st["n_neighbors"]
    assert_matches_port(eng2, case)


def test_zero_copy_results(tnsb):
    # lists written by the kernel straight into mapped pinned host memory: same sets, no separate D2H of the ids
    case = cases.GOLDEN_CASES["variable_random_sym"]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_ZERO_COPY_RESULTS: 1})
    assert_matches_port(eng, case)
    eng.run()
    assert_matches_port(eng, case)
    eng.set_option(tnsb.TNSB_OPT_ZERO_COPY_RESULTS, 0)      # switching back re-sizes the HBM buffer
    eng.run()
    assert_matches_port(eng, case)


def test_zsort_permutation_and_rerun(tnsb):
    pts = clouds.uniform_cloud(50_000, 17).copy()
    vel = np.arange(50_000, dtype=np.float32)
    r = float(clouds.radius_for_mean_neighbors(50_000))
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(r)
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.prepare_zsort()                                  # no run() yet: from scratch (TreeNSearch.cpp:2592-2595)
    order = eng.get_zsort_order(0).copy()
    assert np.array_equal(np.sort(order), np.arange(50_000))
    before = pts.copy()
    eng.apply_zsort(0, pts, 3)
    eng.apply_zsort(0, vel, 1)
    assert np.array_equal(pts, before[order]) and np.array_equal(vel, order.astype(np.float32))
    # the new order is a Z-order: Morton keys of the grid cells are non-decreasing
    st_cell = r * (1.0 + 1.0 / 8192.0)
    eng.run()
    st = eng.stats()
    cell = np.floor((pts.astype(np.float64) - np.array(st["domain_bottom"], np.float64)) / st_cell).astype(np.int64)
    keys = np.array([loader.morton3d_64(*c) for c in cell[::97]], dtype=np.uint64)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    case = dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True)
    assert_matches_port(eng, case)
    # after a run the order comes from the existing grid (TreeNSearch.cpp:2598-2661); sorted data -> identity-like order
    eng.prepare_zsort()
    order2 = eng.get_zsort_order(0)
    assert np.array_equal(np.sort(order2), np.arange(50_000))
    assert np.array_equal(order2, np.arange(50_000))     # stable sort of already Z-sorted data is the identity


def test_error_behaviour(tnsb):
    pts = clouds.uniform_cloud(100, 1)
    rad = np.full(100, 0.1, np.float32)
    eng = tnsb.TreeNSearch()
    eng.add_point_set(pts, rad)
    with pytest.raises(tnsb.TreeNSearchError, match="Cannot set a global search radius"):      # TreeNSearch.cpp:22-25
        eng.set_search_radius(0.1)
    eng = tnsb.TreeNSearch()
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    with pytest.raises(tnsb.TreeNSearchError, match="not all point sets have per-point search radius"):   # :388-391
        eng.run()
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(-1.0)
    eng.add_point_set(pts)
    with pytest.raises(tnsb.TreeNSearchError, match="global_search_radius <= 0"):               # :378-381
        eng.run()
    eng = tnsb.TreeNSearch()
    eng.set_cell_size(0.5)
    with pytest.raises(tnsb.TreeNSearchError, match="Cell size already set"):                   # :175-178
        eng.set_cell_size(0.6)
    with pytest.raises(tnsb.TreeNSearchError, match="Cannot resize a set that was not previously added"):   # :69-72
        eng.resize_point_set(3, pts)
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(0.1)
    eng.add_point_set(pts)
    with pytest.raises(tnsb.TreeNSearchError, match="previously didn't have one"):              # :73-76
        eng.resize_point_set(0, pts, rad)


def test_set_active_search_overloads(tnsb):
    pts = clouds.uniform_cloud(10, 1)
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(0.5)
    for _ in range(3):
        eng.add_point_set(pts)
    eng.set_active_search(1, True, False)        # set 1 searches in all, is found by none (TreeNSearch.cpp:225-235)
    table = [[eng.is_search_active(i, j) for j in range(3)] for i in range(3)]
    assert table == [[False, False, False], [True, True, True], [False, False, False]]
    eng.set_active_search(1, False, True)        # order matters: the search row overwrites (1,1)
    table = [[eng.is_search_active(i, j) for j in range(3)] for i in range(3)]
    assert table == [[False, True, False], [False, False, False], [False, True, False]]
    eng.set_all_searches(True)
    assert all(eng.is_search_active(i, j) for i in range(3) for j in range(3))
    assert eng.does_set_exist(2) and not eng.does_set_exist(3)


def test_device_apply_zsort_matches_host(tnsb):
    """tnsb_apply_zsort_device: fused gather of several resident arrays (float32 xyz, float64 velocity, int32 tag) in one launch,
    in place and out of place, against the host apply_zsort (TreeNSearch.h:443-481)."""
    import torch
    n = 60_000
    pts = clouds.uniform_cloud(n, 23).copy()
    vel = np.random.RandomState(1).standard_normal((n, 3))
    tag = np.arange(n, dtype=np.int32)
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(float(clouds.radius_for_mean_neighbors(n)))
    d_pts = torch.from_numpy(pts).cuda()
    eng.add_point_set(d_pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    eng.prepare_zsort()                     # grid of the last run is valid: order from the resident records, no upload
    assert eng.stats()["h2d_bytes"] == 0
    order = eng.get_zsort_order(0).copy()
    assert np.array_equal(np.sort(order), np.arange(n))
    d_vel, d_tag = torch.from_numpy(vel).cuda(), torch.from_numpy(tag).cuda()
    out_tag = torch.empty_like(d_tag)
    eng.apply_zsort_device(0, [d_pts, d_vel], [3, 3])            # in place, two arrays, one launch
    eng.apply_zsort_device(0, [d_tag], [1], out=[out_tag])       # out of place
    assert np.array_equal(d_pts.cpu().numpy(), pts[order])
    assert np.array_equal(d_vel.cpu().numpy(), vel[order])
    assert np.array_equal(out_tag.cpu().numpy(), order) and np.array_equal(d_tag.cpu().numpy(), tag)
    # the single-array float32 entry point and the host template agree
    d2 = torch.from_numpy(pts).cuda()
    eng.apply_zsort(0, d2, 3)
    host = pts.copy()
    eng.apply_zsort(0, host, 3)
    assert np.array_equal(d2.cpu().numpy(), host) and np.array_equal(host, pts[order])
    # searching the permuted points gives the same sets (renumbered)
    eng.run()
    case = dict(sets=[(host, None)], radius=float(clouds.radius_for_mean_neighbors(n)), pairs=[(0, 0)], symmetric=True)
    assert_matches_port(eng, case)


def test_zsort_reuses_resident_grid_same_order(tnsb):
    """prepare_zsort() after run() (resident records) hands out the same permutation as prepare_zsort() from scratch."""
    n = 80_000
    pts, d, r = clouds.dam_break_cloud(n)
    a = tnsb.TreeNSearch()
    a.set_search_radius(float(r)); a.add_point_set(pts); a.set_active_search(0, 0, True)
    a.prepare_zsort()
    scratch = a.get_zsort_order(0).copy()
    b = tnsb.TreeNSearch()
    b.set_search_radius(float(r)); b.add_point_set(pts); b.set_active_search(0, 0, True)
    b.run()
    b.prepare_zsort()
    assert np.array_equal(b.get_zsort_order(0), scratch)


def test_pin_user_memory_option(tnsb):
    """TNSB_OPT_PIN_USER_MEMORY: pageable user arrays are registered once and re-read on every run (the engine borrows them)."""
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    pts = case["sets"][0][0].copy()
    eng = tnsb.TreeNSearch()
    eng.set_option(tnsb.TNSB_OPT_PIN_USER_MEMORY, 1)
    eng.set_search_radius(case["radius"])
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    assert_matches_port(eng, case)
    pts[:, 0] = pts[::-1, 0].copy()                 # the user mutates the borrowed array in place
    eng.run()
    assert_matches_port(eng, dict(case, sets=[(pts, None)]))
    assert eng.stats()["h2d_bytes"] == pts.nbytes


def test_print_state_and_stats(tnsb, capsys):
    case = cases.GOLDEN_CASES["variable_random_sym"]()
    eng = run_engine(tnsb, case)
    eng.print_state()
    out = capsys.readouterr().out
    assert "NEIGHBORLISTS" in out and "set_0 -> set_1" in out and "n_neighbors set_1 -> set_0 [min, max, avg]" in out
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    off, _ = port.csr(0, 0)
    cnt = np.diff(off)
    line = [l for l in out.splitlines() if l.startswith("n_neighbors set_0 -> set_0")][0]
    assert f"[{cnt.min()}, {cnt.max()}," in line


def test_cell_size_semantics(tnsb):
    """set_cell_size (TreeNSearch.cpp:173-181, :300, :368): any positive value is accepted once and does not change the results; exactly 0
    is the reference's "cell_size is not set" error at run(); a negative value means "use the default"."""
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    for cs in (0.5 * case["radius"], 3.0 * case["radius"], -1.0):
        eng = tnsb.TreeNSearch()
        eng.set_cell_size(cs)
        eng.set_search_radius(case["radius"])
        eng.add_point_set(case["sets"][0][0])
        eng.set_active_search(0, 0, True)
        eng.run()
        assert_matches_port(eng, case)
    eng = tnsb.TreeNSearch()
    eng.set_cell_size(0.0)
    eng.set_search_radius(case["radius"])
    eng.add_point_set(case["sets"][0][0])
    with pytest.raises(tnsb.TreeNSearchError, match="cell_size is not set"):
        eng.run()


def test_speculative_grid_reuse_and_fallback(tnsb):
    """Steady state: run() reuses the previous grid without waiting for the world box (one host round trip per run); a device-side
    check notices when the cloud leaves that grid -- or the radii change -- and the run repeats itself with a fresh grid."""
    rs = np.random.RandomState(3)
    pts = (rs.random_sample((30_000, 3)) * 2.0).astype(np.float32)
    r = 0.08
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(r)
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    assert eng.stats()["speculative_grid"] == 0
    pts += np.float32(0.01)                                     # small motion: still inside the grid (4 cells of room)
    eng.run()
    st = eng.stats()
    assert st["speculative_grid"] == 1 and st["n_reruns"] == 0
    assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))
    pts[:1000] += np.float32(0.7)                               # a splash outside the old grid
    eng.run()
    st = eng.stats()
    assert st["n_reruns"] >= 1 and st["speculative_grid"] == 0
    assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))
    eng.run()
    assert eng.stats()["speculative_grid"] == 1
    eng.set_search_radius(0.05)                                 # a different radius never reuses the old grid
    eng.run()
    assert eng.stats()["speculative_grid"] == 0
    assert_matches_port(eng, dict(sets=[(pts, None)], radius=0.05, pairs=[(0, 0)], symmetric=True))
    # variable radii: the device check also watches the radius range
    rad = np.full(pts.shape[0], 0.06, np.float32)
    e2 = tnsb.TreeNSearch()
    e2.add_point_set(pts, rad)
    e2.set_active_search(0, 0, True)
    e2.run(); e2.run()
    assert e2.stats()["speculative_grid"] == 1
    rad[::7] = 0.09                                             # larger than the grid was built for
    e2.run()
    assert e2.stats()["n_reruns"] >= 1
    assert_matches_port(e2, dict(sets=[(pts, rad)], radius=None, pairs=[(0, 0)], symmetric=True))


def test_list_pos_32_bit_host_mirror(tnsb):
    """The host mirror of list_pos travels as uint32 (tnsb_get_neighborlists_u32, what the C++ drop-in header reads); the 64-bit getter
    widens it on the host on first use.  Same positions either way, run after run."""
    import ctypes as C
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    eng = run_engine(tnsb, case)
    n = case["sets"][0][0].shape[0]
    for _ in range(3):
        eng.run()
        rag = C.POINTER(C.c_int32)()
        p32 = C.POINTER(C.c_uint32)()
        n_ints = C.c_int64()
        eng._check(eng._lib.tnsb_get_neighborlists_u32(eng._h, 0, 0, C.byref(rag), C.byref(p32), C.byref(n_ints)))
        pos32 = np.ctypeslib.as_array(p32, shape=(n,)).astype(np.int64)
        eng._views.clear()
        ragged, pos64 = eng.neighbor_lists(0, 0)
        assert np.array_equal(pos32, np.asarray(pos64))
        assert eng.stats()["d2h_bytes"] <= 4 * int(n_ints.value) + 4 * n
    assert_matches_port(eng, case)


def test_graph_replay_small_steady_state(tnsb):
    """Small problems in steady state: the enqueue phase of run() becomes ONE CUDA graph launch (captured once two consecutive runs
    looked alike, replayed while configuration / pointers / sizes / buffers are unchanged).  Moving points, host and device arrays,
    two sets with variable radii; anything that changes the key -- or a cloud that leaves the reused grid -- falls back to a plain run."""
    import torch
    rs = np.random.RandomState(11)
    base = (rs.random_sample((20_000, 3)) * 2.0).astype(np.float32)
    r = 0.09
    for device_resident in (False, True):
        pts = base.copy()
        arr = torch.from_numpy(pts).cuda() if device_resident else pts
        eng = tnsb.TreeNSearch()
        eng.set_search_radius(r)
        eng.add_point_set(arr)
        eng.set_active_search(0, 0, True)
        replays = 0
        for step in range(7):
            pts += np.float32(0.002)                             # the host array moves in place; the device copy follows it
            if device_resident:
                arr.copy_(torch.from_numpy(pts))
            eng.run()
            st = eng.stats()
            replays += st["graph_replay"]
            if step >= 4:
                assert st["graph_replay"] == 1 and st["speculative_grid"] == 1, (device_resident, step, st)
            assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))
        assert replays >= 3
        pts[:500] += np.float32(0.9)                             # a splash outside the reused grid: the graph run is void and repeats itself
        if device_resident:
            arr.copy_(torch.from_numpy(pts))
        eng.run()
        st = eng.stats()
        assert st["n_reruns"] >= 1 and st["graph_replay"] == 0
        assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))
        for _ in range(4):
            eng.run()
        assert eng.stats()["graph_replay"] == 1                  # captured again for the new grid
        eng.set_active_search(0, 0, False)                       # a change of the configuration changes the key
        eng.run()
        assert eng.stats()["graph_replay"] == 0 and eng.stats()["n_neighbors"] == 0
        eng.set_active_search(0, 0, True)
        eng.run()
        assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))
    # two sets, variable radii, three pairs
    case = cases.GOLDEN_CASES["variable_random_sym"]()
    eng = run_engine(tnsb, case)
    for _ in range(5):
        eng.run()
    assert eng.stats()["graph_replay"] == 1
    assert_matches_port(eng, case)
    # switched off
    eng = run_engine(tnsb, case)
    import os
    os.environ["TNSB_GRAPH"] = "0"
    try:
        e2 = run_engine(tnsb, case)
        for _ in range(4):
            e2.run()
        assert e2.stats()["graph_replay"] == 0
        assert_matches_port(e2, case)
    finally:
        del os.environ["TNSB_GRAPH"]


def test_resize_fast_path_keeps_grid(tnsb):
    """resize_point_set with the same pointer and size is a no-op (TreeNSearch.cpp:77-79, :107-109): the grid stays valid for prepare_zsort."""
    import torch
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    pts = torch.from_numpy(case["sets"][0][0]).cuda()
    eng = tnsb.TreeNSearch()
    eng.set_search_radius(case["radius"])
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    launches = eng.stats()["n_kernel_launches"]
    eng.resize_point_set(0, pts)
    eng.prepare_zsort()                              # still the resident grid: far fewer launches than a build from scratch
    assert eng.stats()["n_kernel_launches"] == launches
    eng.run()
    assert_matches_port(eng, case)


# ---------------------------------------------------------------------------------------------------- the cell kernel (option 1)
# TNSB_OPT_QUERY_KERNEL = 1: cell = r grid sorted by Morton keys, a lane owns a candidate (csrc/query.cuh); the default (0) is the
# brick query on the half-radius grid (csrc/query_brick.cuh), which falls back to the cell kernel on huge sparse domains.  Same contract.
@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_cell_kernel_golden(tnsb, golden, name):
    case = cases.GOLDEN_CASES[name]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_QUERY_KERNEL: 1})
    for (i, j) in case["pairs"]:
        off, idx = eng.neighbor_csr(i, j)
        assert np.array_equal(off, golden[f"{name}/{i}_{j}/offsets"]), (name, i, j)
        assert np.array_equal(idx, golden[f"{name}/{i}_{j}/indices"]), (name, i, j)


def test_cell_kernel_slow_paths_and_limits(tnsb):
    opt = {tnsb.TNSB_OPT_QUERY_KERNEL: 1}
    # neighbourhoods larger than a tile, lists longer than the private lists and than the staging buffer
    rs = np.random.RandomState(8)
    blob = (0.5 + 0.002 * rs.standard_normal((2600, 3))).astype(np.float32)
    bg = rs.random_sample((3000, 3)).astype(np.float32)
    pts = np.ascontiguousarray(np.concatenate([blob, bg]))
    case = dict(sets=[(pts, None)], radius=0.04, pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case, options=opt)
    assert eng.stats()["n_neighbors"] > 2600 * 2500
    assert_matches_port(eng, case)
    # 64-bit keys on a sparse domain: hash of the occupied cells instead of the prefix table
    rs = np.random.RandomState(3)
    centers = rs.random_sample((40, 3)) * 1000.0
    pts = np.ascontiguousarray((centers[rs.randint(0, 40, 4000)] + 0.05 * rs.standard_normal((4000, 3))).astype(np.float32))
    case = dict(sets=[(pts, None)], radius=0.02, pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case, options=opt)
    assert eng.stats()["key_bits"] > 32
    assert_matches_port(eng, case)
    # halo points (find-only) + list buffer overflow re-run
    case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_QUERY_KERNEL: 1, tnsb.TNSB_OPT_LIST_CAPACITY: 1})
    assert eng.stats()["n_reruns"] >= 1
    assert_matches_port(eng, case)


def test_cell_kernel_c1_and_zsort(tnsb):
    n = 100_000
    pts = clouds.uniform_cloud(n, 42).copy()
    r = float(clouds.radius_for_mean_neighbors(n))
    case = dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True)
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_QUERY_KERNEL: 1})
    assert_matches_port(eng, case)
    # the order handed to the user is the libmorton Z-order
    eng.prepare_zsort()
    order = eng.get_zsort_order(0).copy()
    assert np.array_equal(np.sort(order), np.arange(n))
    eng.apply_zsort(0, pts, 3)
    eng.run()
    st = eng.stats()
    cell = np.floor((pts.astype(np.float64) - np.array(st["domain_bottom"], np.float64)) / (r * (1.0 + 1.0 / 8192.0))).astype(np.int64)
    keys = np.array([loader.morton3d_64(*c) for c in cell[::211]], dtype=np.uint64)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    assert_matches_port(eng, dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True))


# ---------------------------------------------------------------------------------------------------- build paths
@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("name", ["uniform_fixed_5000", "variable_random_sym", "clustered_blob", "duplicates", "three_sets_all_pairs"])
def test_radix_build_matches_golden(tnsb, golden, name, kernel):
    """TNSB_OPT_BUILD = 1 forces the LSD radix sort (the default picks the bucket build for these small dense grids)."""
    case = cases.GOLDEN_CASES[name]()
    eng = run_engine(tnsb, case, options={tnsb.TNSB_OPT_BUILD: 1, tnsb.TNSB_OPT_QUERY_KERNEL: kernel})
    assert eng.stats()["sort_passes"] >= 1
    for (i, j) in case["pairs"]:
        off, idx = eng.neighbor_csr(i, j)
        assert np.array_equal(off, golden[f"{name}/{i}_{j}/offsets"]), (name, i, j)
        assert np.array_equal(idx, golden[f"{name}/{i}_{j}/indices"]), (name, i, j)
