"""The CSVs the reference's own (unmodified) drivers wrote on B200 -- once linked against this library, once against the
reference library (profiles/r01_reference_drivers/, produced by tools/ref_drivers.py run) -- checked offline: every
emulation cell of the accuracy tables is the same string in both, the DGEMM table equals the reference's published GH200
table (fixture tests/golden/published_d_accuracy_GH200.csv), and every timing row carries identical errors."""
import glob
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(ROOT, "profiles", "r01_reference_drivers")


def _tool():
    spec = importlib.util.spec_from_file_location("ref_drivers", os.path.join(ROOT, "tools", "ref_drivers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _csv(lib, kind, prec):
    c = glob.glob(os.path.join(DIR, f"{lib}_oz2_results_{prec}_{kind}_*.csv"))
    assert len(c) == 1, c
    return _tool().read_csv(c[0])


def test_accuracy_tables_identical_between_libraries():
    t = _tool()
    for prec, cells in (("d", 760), ("f", 448)):
        ours, moduli = t.accuracy_table(_csv("ours", "accuracy", prec))
        ref, _ = t.accuracy_table(_csv("ref", "accuracy", prec))
        n = 0
        for key, row in ours.items():
            if t.is_emulation(key[1]):
                assert row == ref[key], key
                n += len(row)
        assert n == cells


def test_dgemm_accuracy_table_equals_the_published_one():
    t = _tool()
    ours, moduli = t.accuracy_table(_csv("ours", "accuracy", "d"))
    pub, pm = t.accuracy_table(t.read_csv(os.path.join(ROOT, "tests", "golden", "published_d_accuracy_GH200.csv")))
    assert moduli == pm
    n = 0
    for key, row in pub.items():
        assert ours[key] == row, key
        n += len(row)
    assert n == 30 * 19        # 5 phi x 3 k (<= 4096 kept in the fixture) x fast / accurate, 2..20 moduli


def test_timing_rows_carry_identical_errors_and_are_faster():
    t = _tool()
    for prec, rows in (("d", 152), ("f", 112)):
        ours, ref = t.time_table(_csv("ours", "time", prec)), t.time_table(_csv("ref", "time", prec))
        n = 0
        for key, o in ours.items():
            if not t.is_emulation(key[1]):
                continue
            r = ref[key]
            assert (o["relerr_max"], o["relerr_med"]) == (r["relerr_max"], r["relerr_med"]), key
            assert float(o["TFLOPS"]) > float(r["TFLOPS"]), key
            n += 1
        assert n == rows


DIR2 = os.path.join(ROOT, "profiles", "r02_reference_drivers")


def test_round2_drivers_accuracy_tables_identical_between_libraries():
    """test_mixed_double (FP64 x FP32 -> FP64), test_mixed_float (-> FP32) and test_float_complex, the three drivers round 1 only
    built (GEMMul8/testing/test_mixed_double.cu:162, test_mixed_float.cu:200, test_float_complex.cu:24,201): every emulation
    cell of their accuracy tables -- 5 phi x 4 k x 2..20 moduli x fast / accurate -- is the same string with both libraries."""
    t = _tool()
    for prec, cells in (("dfd", 760), ("dff", 448), ("fC", 448)):
        def one(lib):
            c = glob.glob(os.path.join(DIR2, f"{lib}_oz2_results_{prec}_accuracy_*.csv"))
            assert len(c) == 1, c
            return t.accuracy_table(t.read_csv(c[0]))[0]
        ours, ref = one("ours"), one("ref")
        n = 0
        for key, row in ours.items():
            if t.is_emulation(key[1]):
                assert row == ref[key], (prec, key)
                n += len(row)
        assert n == cells


def test_round2_drivers_timing_rows_carry_identical_errors_and_are_faster():
    """flops_check of test_mixed_double, test_mixed_float and test_float_complex (sizes 1024 .. 8192, 2 .. 20 moduli, fast and accurate): every
    emulation row has the same relerr_max / relerr_med strings with both libraries, and this library is the faster one."""
    t = _tool()
    for prec, rows in (("dfd", 152), ("dff", 112), ("fC", 112)):
        def one(lib):
            c = glob.glob(os.path.join(DIR2, f"{lib}_oz2_results_{prec}_time_*.csv"))
            assert len(c) == 1, c
            return t.time_table(t.read_csv(c[0]))
        ours, ref = one("ours"), one("ref")
        n = 0
        for key, o in ours.items():
            if not t.is_emulation(key[1]):
                continue
            r = ref[key]
            assert (o["relerr_max"], o["relerr_med"]) == (r["relerr_max"], r["relerr_med"]), (prec, key)
            assert float(o["TFLOPS"]) > float(r["TFLOPS"]), (prec, key)
            n += 1
        assert n == rows
