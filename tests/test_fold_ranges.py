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
