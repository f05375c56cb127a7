"""Strong-Wolfe line search of the device-resident L-BFGS driver (models._Base._bfgs_gpu).

Host logic only: `phi(alpha)` moves the device state to x0 + alpha d, runs the residual kernels and returns the three
scalars the search branches on.  Bracketing + zoom with safeguarded cubic interpolation (Nocedal & Wright, Numerical
Optimization, alg. 3.5 / 3.6); the constants the caller passes (1e-3, 0.9) are those of L-BFGS-B's line search, which
the reference reaches through ScipyOptimizerInterface (PlateHoleQuarter/train/train.py:240-247)."""
from __future__ import annotations

import math


def cubic_min(a, fa, da, b, fb, db):
    """Minimiser of the cubic through (a, fa, f'a), (b, fb, f'b); None when it has none."""
    if a == b:
        return None
    d1 = da + db - 3.0 * (fa - fb) / (a - b)
    rad = d1 * d1 - da * db
    if not rad >= 0.0:
        return None
    d2 = math.sqrt(rad) * (1.0 if b > a else -1.0)
    den = db - da + 2.0 * d2
    if den == 0.0:
        return None
    return b - (b - a) * (db + d2 - d1) / den


def strong_wolfe(phi, f0, dphi0, alpha, c1=1e-3, c2=0.9, maxls=20, budget=None, first=None):
    """phi(alpha) -> (f, dphi, aux).  Returns (ok, alpha, f, dphi, aux); when ok, the LAST phi call was made at the returned
    alpha (the device state is left there).  budget() -> evaluations still allowed (maxfun), or None.
    first = (f, dphi, aux): the result of phi(alpha) at the initial trial step when the caller has evaluated it already
    (the device-resident driver queues the first trial behind the direction kernel without a host round trip)."""
    a_lo, f_lo, d_lo = 0.0, f0, dphi0
    a_hi = f_hi = d_hi = None
    last = None
    for ls in range(maxls):
        if budget is not None and budget() <= 0 and last is not None:
            break
        if ls == 0 and first is not None:
            fa, da, aux = first
        else:
            fa, da, aux = phi(alpha)
        last = (alpha, fa, da, aux)
        if not math.isfinite(fa):                                      # stepped out of the representable region: pull back
            a_hi, f_hi, d_hi = alpha, math.inf, None
            alpha = 0.5 * (a_lo + alpha) if a_lo > 0 else 0.1 * alpha
            continue
        armijo = fa <= f0 + c1 * alpha * dphi0
        if a_hi is None:                                               # bracketing phase
            if not armijo or (a_lo > 0 and fa >= f_lo):
                a_hi, f_hi, d_hi = alpha, fa, da
            elif abs(da) <= -c2 * dphi0:
                return True, alpha, fa, da, aux
            elif da >= 0:
                a_hi, f_hi, d_hi = a_lo, f_lo, d_lo
                a_lo, f_lo, d_lo = alpha, fa, da
            else:                                                      # still descending: extrapolate
                new = cubic_min(a_lo, f_lo, d_lo, alpha, fa, da)
                a_lo, f_lo, d_lo = alpha, fa, da
                alpha = min(max(new if new is not None else 2.0 * alpha, 1.1 * alpha), 4.0 * alpha)
                continue
        else:                                                          # zoom phase
            if not armijo or fa >= f_lo:
                a_hi, f_hi, d_hi = alpha, fa, da
            else:
                if abs(da) <= -c2 * dphi0:
                    return True, alpha, fa, da, aux
                if da * (a_hi - a_lo) >= 0:
                    a_hi, f_hi, d_hi = a_lo, f_lo, d_lo
                a_lo, f_lo, d_lo = alpha, fa, da
        new = None
        if d_hi is not None and math.isfinite(f_hi):
            new = cubic_min(a_lo, f_lo, d_lo, a_hi, f_hi, d_hi)
        elif math.isfinite(f_hi):
            den = 2.0 * (f_hi - f_lo - d_lo * (a_hi - a_lo))
            new = a_lo - d_lo * (a_hi - a_lo) ** 2 / den if den != 0.0 else None
        lo, hi = min(a_lo, a_hi), max(a_lo, a_hi)
        if hi - lo <= 1e-9 * max(hi, 1e-30):
            break
        if new is None or not (lo + 0.1 * (hi - lo) <= new <= hi - 0.1 * (hi - lo)):
            new = 0.5 * (lo + hi)
        alpha = new
    # exhausted: accept the best sufficient-decrease point seen, if any
    if a_lo > 0.0 and f_lo < f0:
        if last is None or last[0] != a_lo:
            if budget is not None and budget() <= 0:                   # maxfun is a hard limit: no re-evaluation at a_lo, the caller restores x0
                return False, 0.0, f0, dphi0, None
            fa, da, aux = phi(a_lo)
            return True, a_lo, fa, da, aux
        return True, last[0], last[1], last[2], last[3]
    return False, 0.0, f0, dphi0, None
