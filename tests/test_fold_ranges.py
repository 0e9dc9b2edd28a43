"""Interval bookkeeping of the fold-arithmetic transforms (tools/fold_ranges.py): every multiplied word stays inside
the window sol_mul admits and every word inside the signed 64-bit range, for the worst prime of the fast path."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("fold_ranges", os.path.join(ROOT, "tools", "fold_ranges.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fold_ranges_hold():
    fr = _load()
    out, log = fr.check_all()      # asserts inside
    assert out and log
    assert {"forward_3n", "inverse_3n", "forward_std", "inverse_std", "rows_fwd_after_shoup", "rows_inv_to_shoup",
            "forward_3n_a10", "inverse_3n_a16", "forward_std_10", "inverse_std_17"} <= set(out)
    for name, w in out.items():
        assert w.mag() < 4 * fr.Q, name


def test_model_constants_match_the_kernels():
    """the schedule constants the interval model restates are the ones the kernels compile with"""
    import re
    fr = _load()
    src = open(os.path.join(ROOT, "matrix-fhe-lattigo_b200", "csrc", "ntt_fold.cuh")).read()
    assert int(re.search(r"constexpr int FWD_PERIOD = (\d+);", src).group(1)) == fr.FWD_PERIOD
    assert "fwd_ku(int r) { return (r & 1) == 0; }" in src and "fwd_kv(int r) { return (r & 1) == 1; }" in src   # recentre U on even, V on odd layers
    assert "if (l & 1) fbf_inv<1, true>" in src                                                                   # inverse: odd layers recentre and fold the sums
    ring = open(os.path.join(ROOT, "matrix-fhe-lattigo_b200", "csrc", "b200ring.cu")).read()
    assert "k >= 50 && k <= 60 && (c < 0 ? -c : c) <= (1LL << (k - 33))" in ring                                   # the admissibility the model's C_MAX, K stand for
    assert fr.K == 60 and fr.C_MAX == 2 ** 27
