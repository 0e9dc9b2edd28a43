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

K = 60                      # worst case of the fast path: the largest k it admits (least headroom below 2^63)
C_MAX = F(2) ** 27          # and the largest |c| there; the fast path admits |c| <= 2^(k-33), which scales every window
                            # below to the same number of units of 2^k for smaller k (b200ring.cu fold_prime_ok)
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
FWD_PERIOD = 5   # ntt_fold.cuh: every FWD_PERIOD layers all words are folded; layer l recentres U (l % 5 even) or V (odd)


def fwd_layers(w, l0, n, log=None, name=""):
    """layers l0 .. l0+n-1 of fwd_regs_f / fwd_regs_fr (numbered across both passes)"""
    for l in range(l0, l0 + n):
        r = l % FWD_PERIOD
        if l > 0 and r == 0:
            w = fold(w)
        ku, kv = (1, 0) if r % 2 == 0 else (0, 1)
        t = mul(w, "%s layer %d" % (name, l))
        w = check_word((w + t - ku * Q).union(w - t + kv * Q), name)
        if log is not None:
            log.append("%s layer %d: %s" % (name, l, w))
    return w


def canon(w, what):
    assert w.mag() < S64 - Q, "canonical reduction needs |v| < 2^63 - q (%s: %s)" % (what, w)
    return Iv(0, Q)


# ---- inverse radix-2 layers (Gentleman-Sande): U' = U + V - ks*q (folded in odd layers), V' = (U - V)*w ------------------
def inv_layers(ws, wp, l0, n, log=None, name=""):
    """ws / wp: windows of the words that were sums / products in the previous layer (a pair is two of one kind)"""
    for l in range(l0, l0 + n):
        outs, outp = None, None
        for w in (ws, wp):
            if w is None:
                continue
            t = mul(w - w, "%s layer %d" % (name, l))
            sm = check_word(w + w - (l & 1) * Q, name)
            outs = sm if outs is None else outs.union(sm)
            outp = t if outp is None else outp.union(t)
        ws, wp = outs, outp
        if l & 1:
            ws = fold(ws)
        if log is not None:
            log.append("%s layer %d: sums %s products %s" % (name, l, ws, wp))
    return ws, wp


def forward_3n(log=None, radix2_layers=14):
    """k_cols_fwd_m3_15_{f,b} + k_rows_fwd_f: caller's words below 4 * 2^k (contract: below 2q); cyclotomic layer, radix-3 layer (outputs folded),
    radix-2 layers 0..13, canonical reduction"""
    w = Iv(0, 4)
    tt = mul(w, "cyclotomic")
    lo = check_word(w + tt - Q, "cyclotomic lo")
    hi = check_word(w + w - tt - 2 * Q, "cyclotomic hi")
    x = lo.union(hi)
    b0 = fold(x)
    t1 = mul(x, "radix-3 t1")
    t3 = mul(t1 - t1, "radix-3 t3")
    out = fold(check_word((b0 + t1 + t1).union(b0 - t1 + t3).union(b0 - t1 - t3), "radix-3"))
    if log is not None:
        log.append("forward 3N after the radix-3 layer: %s" % out)
    w = fwd_layers(out, 0, radix2_layers, log, "forward 3N")
    return canon(w, "forward 3N")


def forward_std(log=None, layers=16):
    """k_cols_fwd_std16_f + k_rows_fwd_f: caller's words below 2.5 * 2^k (the reference's transforms take words below
    2q, ring/ntt.go:155-171), layers 0..15 with the first fold before layer 5, canonical reduction"""
    w = fwd_layers(Iv(0, F(5, 2)), 0, layers, log, "forward 2^16")
    return canon(w, "forward 2^16")


def rows_fwd_after_shoup(log=None):
    """k_rows_fwd_f<.., -1>: the Shoup column pass left unsigned words below 8q: fold_u at the load, layers 0..7"""
    return canon(fwd_layers(fold_u(), 0, 8, log, "rows after Shoup columns"), "rows after Shoup columns")


def inverse_3n(log=None, radix2_layers=14):
    """k_rows_inv_f + k_cols_inv_m3_15_{f,b}: fold_u of any input word, radix-2 layers 0..13, inverse radix-3 layer,
    inverse cyclotomic layer in its two-product form (hh = k1 (lo - hi), ll = k4 (lo + hi) - hh / 2), canonical"""
    ws, wp = inv_layers(fold_u(), None, 0, radix2_layers, log, "inverse 3N")
    y = ws.union(wp)
    sg = mul(y - y, "inverse radix-3 s")
    y0 = y if (radix2_layers - 1) & 1 else fold(y)                   # last layer even: its sums arrive unfolded, one summand is folded first
    x0 = fold(check_word(y0 + y + y, "inverse radix-3 x0"))         # folded: lo + hi below must stay a valid multiplicand
    x12 = mul(check_word(y - y - sg, "inverse radix-3 operand"), "inverse radix-3 x1, x2")
    x = x0.union(x12)
    hh = mul(check_word(x - x, "inverse cyclotomic lo - hi"), "inverse cyclotomic hh")
    half = Iv((hh.lo - Q) / 2, (hh.hi + Q) / 2)                      # sol_half: (v + q) >> 1 for odd v
    ll = check_word(mul(check_word(x + x, "inverse cyclotomic lo + hi"), "inverse cyclotomic ll") - half, "inverse cyclotomic ll")
    if log is not None:
        log.append("inverse 3N: x0 %s, products %s, ll %s, hh %s" % (x0, x12, ll, hh))
    canon(ll, "inverse 3N ll")
    return canon(hh, "inverse 3N hh")


def inverse_std(log=None, layers=16):
    """k_rows_inv_f + k_cols_inv_std16_f: layers 0..14, then the last layer carries N^-1 on both outputs"""
    ws, wp = inv_layers(fold_u(), None, 0, layers - 1, log, "inverse 2^16")
    w = ws.union(wp)
    out = mul(check_word(w + w, "last layer sum"), "last layer").union(mul(w - w, "last layer difference"))
    return canon(out, "inverse 2^16")


def rows_inv_to_shoup(log=None):
    """k_rows_inv_f<.., false>: 8 layers, then + q so that the Shoup column pass sees unsigned words below 4q"""
    ws, wp = inv_layers(fold_u(), None, 0, 8, log, "rows before Shoup columns")
    w = ws.union(wp) + Q
    assert w.lo >= 0 and w.hi < 4 * Q, w
    return w


def check_all():
    """the headline shapes with their logs, then every shape the generic fold column kernels cover (same schedule, other depths)"""
    log = []
    out = {"forward_3n": forward_3n(log), "inverse_3n": inverse_3n(log), "forward_std": forward_std(log), "inverse_std": inverse_std(log),
           "rows_fwd_after_shoup": rows_fwd_after_shoup(log), "rows_inv_to_shoup": rows_inv_to_shoup(log)}
    for a in range(10, 17):          # N = 3 * 2^a: a - 1 radix-2 layers
        out["forward_3n_a%d" % a] = forward_3n(None, a - 1)
        out["inverse_3n_a%d" % a] = inverse_3n(None, a - 1)
    for logn in range(10, 18):       # N = 2^logn
        out["forward_std_%d" % logn] = forward_std(None, logn)
        out["inverse_std_%d" % logn] = inverse_std(None, logn)
    return out, log


if __name__ == "__main__":
    out, log = check_all()
    print("units of 2^%d, worst prime q = 2^%d + 2^27; multiplied words <= %.3f, all words < %.1f" % (K, K, float(Y_MAX), float(S64)))
    for line in log:
        print(" ", line)
    for k, v in out.items():
        print(k, "leaves", v)
