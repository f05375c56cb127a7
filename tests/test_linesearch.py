"""Host logic of the device-resident L-BFGS driver: the strong-Wolfe search and a numpy model of the
two-loop recursion the CUDA kernel implements, run as a full L-BFGS on analytic functions and compared with SciPy's L-BFGS-B."""
import math

import numpy as np
import scipy.optimize

from pinn_elastodynamics_b200.linesearch import cubic_min, strong_wolfe


def _phi_of(f, g, x, d):
    def phi(a):
        xx = x + a * d
        return float(f(xx)), float(g(xx) @ d), xx
    return phi


def test_cubic_min_recovers_cubic():
    f = lambda t: (t - 1.0) ** 2 * (t + 2.0)          # local minimum at t = 1
    df = lambda t: 2 * (t - 1) * (t + 2) + (t - 1) ** 2
    assert abs(cubic_min(0.2, f(0.2), df(0.2), 1.7, f(1.7), df(1.7)) - 1.0) < 1e-12
    assert cubic_min(0.0, 0.0, -1.0, 1.0, -1.0, -1.0) is None           # a straight line has no minimiser


def test_strong_wolfe_conditions_hold():
    rng = np.random.default_rng(0)
    rosen, rosen_g = scipy.optimize.rosen, scipy.optimize.rosen_der
    for trial in range(50):
        x = rng.normal(size=6)
        d = -rosen_g(x) * rng.uniform(0.01, 3.0)       # badly scaled descent directions
        f0, dphi0 = rosen(x), rosen_g(x) @ d
        ok, a, fa, da, _ = strong_wolfe(_phi_of(rosen, rosen_g, x, d), f0, dphi0, 1.0, 1e-3, 0.9, 30)
        assert ok
        assert fa <= f0 + 1e-3 * a * dphi0 + 1e-12
        assert abs(da) <= 0.9 * abs(dphi0) + 1e-12


def test_strong_wolfe_backs_out_of_nan_region():
    f = lambda x: math.log(1.0 - x[0]) * -1.0 + x[0] ** 2 if x[0] < 1 else float('nan')    # barrier at x = 1
    g = lambda x: np.array([1.0 / (1.0 - x[0]) + 2 * x[0]]) if x[0] < 1 else np.array([float('nan')])
    x, d = np.array([-2.0]), np.array([10.0])
    ok, a, fa, da, _ = strong_wolfe(_phi_of(f, g, x, d), f(x), g(x) @ d, 1.0, 1e-3, 0.9, 40)
    assert ok and math.isfinite(fa) and fa < f(x)


def test_last_evaluation_is_the_returned_point():
    rosen, rosen_g = scipy.optimize.rosen, scipy.optimize.rosen_der
    calls = []
    x = np.array([-1.2, 1.0, 0.3]); d = -rosen_g(x)
    base = _phi_of(rosen, rosen_g, x, d)
    def phi(a):
        calls.append(a)
        return base(a)
    ok, a, *_ = strong_wolfe(phi, rosen(x), rosen_g(x) @ d, 1.0, 1e-3, 0.9, 25)
    assert ok and calls[-1] == a


def _two_loop(g, S, Y):
    """numpy statement of csrc/pe_lbfgs.cu:lbfgs_direction_kernel (pairs ordered oldest -> newest)"""
    q = g.copy(); al = []
    for s, y in zip(reversed(S), reversed(Y)):
        a = (s @ q) / (y @ s); al.append(a); q -= a * y
    if S:
        q *= (Y[-1] @ S[-1]) / (Y[-1] @ Y[-1])
    for (s, y), a in zip(zip(S, Y), reversed(al)):
        b = (y @ q) / (y @ s); q += (a - b) * s
    return -q


def test_lbfgs_with_this_search_reaches_scipy_minimum():
    rosen, rosen_g = scipy.optimize.rosen, scipy.optimize.rosen_der
    x = np.full(10, -1.0); m = 10
    S, Y = [], []
    f, g = rosen(x), rosen_g(x)
    nfev = 1
    for it in range(500):
        if np.abs(g).max() < 1e-8:
            break
        d = _two_loop(g, S, Y)
        calls = [0]
        def phi(a):
            calls[0] += 1
            xx = x + a * d
            return float(rosen(xx)), float(rosen_g(xx) @ d), xx
        ok, a, f, _, xn = strong_wolfe(phi, f, g @ d, 1.0 if S else min(1.0, 1.0 / np.linalg.norm(g)), 1e-3, 0.9, 30)
        assert ok
        nfev += calls[0]
        gn = rosen_g(xn)
        S.append(xn - x); Y.append(gn - g)
        S, Y = S[-m:], Y[-m:]
        x, g = xn, gn
    ref = scipy.optimize.minimize(rosen, np.full(10, -1.0), jac=rosen_g, method='L-BFGS-B',
                                  options=dict(maxcor=10, ftol=1e-15, gtol=1e-8, maxiter=500))
    assert f < 1e-12 and ref.fun < 1e-10
    assert nfev < 2.0 * ref.nfev + 20            # comparable evaluation count
