"""GPU test: the reference's OWN test program (tests/main.cpp + tests/tests.cpp + tests/BruteforceNSearch.cpp, unmodified),
compiled against the drop-in header include/TreeNSearch and linked to libtnsb.so (oracle/Makefile target `ref-tests`).
Covers one_set_fixed_radius, two_dynamic_sets_variable_radius, mixed_float_double_point_sets and resize_variable_radius at
n = 1, 100, 10000 with the thread / recursion-cap sweeps and the zsort round trip (tests/tests.cpp:34-237), each compared
against BruteforceNSearch by the reference's own comparator."""
import os
import subprocess

import pytest

from oracle import loader

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(loader.REF_TESTS_BIN), reason="oracle/_ref/ref_tests_on_b200 not built (needs /root/reference)")
def test_reference_test_program_passes_on_the_cuda_engine(built_library):
    res = subprocess.run([loader.REF_TESTS_BIN], capture_output=True, text=True, timeout=1200)
    out = res.stdout
    assert res.returncode == 0, out[-3000:] + res.stderr[-2000:]
    assert "FAILED" not in out, out[-3000:]
    # per size: 3 cases x 7 'passed!' lines (tests/tests.cpp:34-89) + 3 lines of resize_variable_radius (:188-237)
    assert out.count("passed!") == 3 * (3 * 7 + 3), out[-3000:]
    assert "Runtime parallel SIMD" in out


@pytest.mark.skipif(not os.path.exists(loader.REF_STRESS_BIN), reason="oracle/_ref/ref_stress_on_b200 not built (needs /root/reference)")
def test_reference_dynamic_emitter_stress_test_all_10000_steps(built_library):
    """tests/tests.cpp:434-514, unmodified and at its full length: 10 000 steps of adding / removing / replacing up to 20 points in one
    of two variable-radius sets (all searches active, sets start empty), every step compared with BruteforceNSearch by the reference's
    own comparator.  The reference keeps it behind `if (false)` in main.cpp; oracle/stress_main.cpp calls it."""
    res = subprocess.run([loader.REF_STRESS_BIN], capture_output=True, text=True, timeout=1200)
    out = res.stdout
    assert res.returncode == 0, out[-3000:] + res.stderr[-2000:]
    assert "FAILED" not in out, out[-3000:]
    assert "Dynamic Emitter Stress Test Passed!" in out, out[-3000:]
