"""Range bookkeeping of the fold-arithmetic transforms (csrc/solinas.cuh, csrc/ntt_fold.cuh).

The fold butterflies work on signed 64-bit words that are only congruent to the field element; nothing is reduced
unless the schedule says so.  This module restates every kernel's schedule (which layers recentre by +-q, where the
folds sit) on intervals, in exact rational arithmetic, for the worst prime the fast path admits, and checks the two
hardware limits:

  * a multiplied word must satisfy |y| <= Y_MAX (the four partial products of sol_mul must not overflow 64 bits),
  * every word must stay inside the signed 64-bit range.

tests/test_fold_ranges.py runs check_all(); the kernels' comments quote the windows printed by `python tools/fold_ranges.py`.
All quantities are in units of 2^k (k = bit position of the prime's leading term, q = 2^k + c).
"""
from fractions import Fraction as F

K = 60                      # worst case of the fast path: the largest k it admits
C_MAX = F(2) ** 27          # and the largest |c|
UNIT = F(2) ** K
Q = 1 + C_MAX / UNIT        # q in units of 2^k (c > 0 is the worse sign for the upper ends)
S64 = F(2) ** 63 / UNIT     # 8 units
Y_MAX = S64 - F(2) ** 59 / UNIT   # |y| limit of sol_mul: |y1| <= 2^31 - 2^27 keeps B = y0*w_hi + y1*v_hi inside 64 bits


class Iv:
    """closed interval [lo, hi]"""

    def __init__(self, lo, hi):
        self.lo, self.hi = F(lo), F(hi)
        assert self.lo <= self.hi

    def __add__(self, o):
        o = o if isinstance(o, Iv) else Iv(o, o)
        return Iv(self.lo + o.lo, self.hi + o.hi)

    def __sub__(self, o):
        o = o if isinstance(o, Iv) else Iv(o, o)
        return Iv(self.lo - o.hi, self.hi - o.lo)

    def __neg__(self):
        return Iv(-self.hi, -self.lo)

    def union(self, o):
        return Iv(min(self.lo, o.lo), max(self.hi, o.hi))

    def mag(self):
        return max(abs(self.lo), abs(self.hi))

    def __repr__(self):
        return "[%.3f, %.3f]" % (float(self.lo), float(self.hi))


def mul(y, what=""):
    """window of sol_mul(y, w) for any twiddle: A + 2^s*B.lo - c*B.hi   (solinas.cuh)"""
    Y = y.mag()
    assert Y <= Y_MAX, "multiplied word out of range %s: %s" % (what, y)
    s = K - 32
    y1 = Y * UNIT / 2 ** 32 + 1                                  # |y1| <= |y|/2^32 + 1
    A = (F(2) ** 31 * F(2) ** (s - 1) + y1 * F(2) ** (s - 1)) / UNIT
    Bhi = (F(2) ** 31 * F(2) ** 31 + y1 * F(2) ** 31) / 2 ** 32 + 1
    cB = C_MAX * Bhi / UNIT
    return Iv(-(A + cB), 1 + A + cB)


def fold(v):
    """window of sol_fold(v): (v mod 2^k) - (v >> k)*c"""
    h = int(v.mag()) + 1
    e = h * C_MAX / UNIT
    return Iv(-e, 1 + e)


def fold_u():
    """window of sol_fold_u of any unsigned 64-bit word"""
    e = F(2) ** (64 - K) * C_MAX / UNIT
    return Iv(-e, 1)


def check_word(v, what=""):
    assert v.mag() < S64, "64-bit overflow %s: %s" % (what, v)
    return v


# ---- forward radix-2 layers (Cooley-Tukey): U' = U + T - ku*q, V' = U - T + kv*q ------------------------------------
FWD_PERIOD = 5
FWD_PATTERN = [(1, 0), (0, 1), (1, 0), (0, 1), (1, 0)]   # (ku, kv) of the r-th layer after a fold (ntt_fold.cuh: fwd_ku / fwd_kv)


def fwd_layers(w, n, r0=0, log=None, name=""):
    for i in range(n):
        ku, kv = FWD_PATTERN[r0 + i]
        t = mul(w, "%s layer %d" % (name, r0 + i))
        o0 = w + t - ku * Q
        o1 = w - t + kv * Q
        w = check_word(o0.union(o1), name)
        if log is not None:
            log.append("%s r=%d: %s" % (name, r0 + i, w))
    return w


# ---- inverse radix-2 layers (Gentleman-Sande): U' = U + V - ks*q, V' = (U - V)*w -----------------------------------------
def inv_layer(ws, wp, ks, name=""):
    """ws / wp: windows of the words that were sums / products in the previous layer (a pair is two of one kind)"""
    outs, outp = None, None
    for w in (ws, wp):
        if w is None:
            continue
        d = w - w
        t = mul(d, name)
        s = check_word(w + w - ks * Q, name)
        outs = s if outs is None else outs.union(s)
        outp = t if outp is None else outp.union(t)
    return outs, outp


def rows_fwd(log=None):
    """k_rows_fwd_f: fold_u at the load, layers 0-3, fold, layers 4-7, canonical reduction"""
    w = fold_u()
    w = fwd_layers(w, 4, 0, log, "rows_fwd a")
    w = fold(w)
    w = fwd_layers(w, 4, 0, log, "rows_fwd b")
    assert w.mag() < S64 - Q, "canonical reduction needs |v| < 2^63 - q"
    return w


def inv_layers(ws, wp, n, l0=0, log=None, name=""):
    """n Gentleman-Sande layers of ntt_fold.cuh inv_regs_f: the sums of odd layers are recentred by -q and folded"""
    for l in range(l0, l0 + n):
        ws, wp = inv_layer(ws, wp, l & 1, "%s layer %d" % (name, l))
        if l & 1:
            ws = fold(ws)
        if log is not None:
            log.append("%s layer %d: sums %s products %s" % (name, l, ws, wp))
    return ws, wp


def rows_inv(log=None):
    """k_rows_inv_f: fold_u at the load, 8 layers"""
    ws, wp = inv_layers(fold_u(), None, 8, 0, log, "rows_inv")
    return ws.union(wp)


def check_all():
    log = []
    out = {"rows_fwd": rows_fwd(log), "rows_inv": rows_inv(log)}
    return out, log


if __name__ == "__main__":
    out, log = check_all()
    print("units of 2^%d, worst prime q = 2^%d + 2^27; multiplied words <= %.3f, all words < %.1f" % (K, K, float(Y_MAX), float(S64)))
    for line in log:
        print(" ", line)
    for k, v in out.items():
        print(k, "leaves", v)
