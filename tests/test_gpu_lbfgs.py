"""Device-resident L-BFGS (csrc/pe_lbfgs.cu, models._Base._bfgs_gpu) against the SciPy-driven path that
mirrors ScipyOptimizerInterface.minimize (PlateHoleQuarter/train/train.py:240-247, 522-525)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_torch as R
from tests.test_linesearch import _two_loop

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pe():
    import pinn_elastodynamics_b200 as pe_
    return pe_


def _P(t):
    return C.c_void_p(t.data_ptr())


def test_direction_and_pair_kernels_match_numpy(pe):
    """two-loop recursion, ring-buffer addressing, rho/gamma bookkeeping: fp64 numpy statement, tolerance 2e-5 relative
    (fp32 storage of S, Y, d; double accumulation of every dot product)"""
    from pinn_elastodynamics_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(3)
    n, m = 10655 + 5, 7
    dev = torch.device('cuda:0')
    S = torch.zeros(m, n, device=dev); Y = torch.zeros(m, n, device=dev); state = torch.zeros(m + 3, device=dev)
    d = torch.zeros(n, device=dev); res = torch.zeros(2, device=dev)
    A = rng.normal(size=(n,)) ** 2 + 0.5                      # SPD diagonal "Hessian": y = A s keeps y.s > 0
    pairs_s, pairs_y = [], []
    x = rng.normal(size=n).astype(np.float32); g = (A * x).astype(np.float32)
    head = -1
    for k in range(11):                                       # wraps the ring buffer (11 > m)
        xn = (x + 0.1 * rng.normal(size=n)).astype(np.float32); gn = (A * xn).astype(np.float32)
        head = (head + 1) % m
        tx, txp, tg, tgp = (torch.from_numpy(a).to(dev) for a in (xn, x, gn, g))
        L.check(lib.pe_lbfgs_store_pair(n, m, head, _P(tx), _P(txp), _P(tg), _P(tgp), _P(S), _P(Y), _P(state), None), 'store')
        pairs_s.append((xn - x).astype(np.float64)); pairs_y.append((gn - g).astype(np.float64))
        pairs_s, pairs_y = pairs_s[-m:], pairs_y[-m:]
        x, g = xn, gn
        count = min(k + 1, m)
        tg = torch.from_numpy(g).to(dev)
        L.check(lib.pe_lbfgs_direction(n, m, count, head, _P(tg), _P(S), _P(Y), _P(state), _P(d), None), 'direction')
        ref = _two_loop(g.astype(np.float64), pairs_s, pairs_y)
        got = d.cpu().numpy().astype(np.float64)
        assert np.linalg.norm(got - ref) <= 2e-5 * np.linalg.norm(ref), k
        L.check(lib.pe_vec_dot_max(n, _P(tg), _P(d), _P(res), None), 'dot')
        r = res.cpu().numpy()
        np.testing.assert_allclose(r[0], g.astype(np.float64) @ got, rtol=1e-5)
        assert r[1] == np.abs(g).max()
        st = state.cpu().numpy()
        np.testing.assert_allclose(st[m + 1], pairs_y[-1] @ pairs_s[-1], rtol=1e-5)
        np.testing.assert_allclose(st[0], (pairs_y[-1] @ pairs_s[-1]) / (pairs_y[-1] @ pairs_y[-1]), rtol=1e-5)
    out = torch.zeros(n, device=dev)
    L.check(lib.pe_vec_axpy(n, _P(out), _P(tg), 0.25, _P(d), None), 'axpy')
    exp = (np.float64(0.25) * d.cpu().numpy().astype(np.float64) + g.astype(np.float64)).astype(np.float32)   # fmaf = one rounding
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=2e-7, atol=0)


def _plate(pe, Collo, HOLE, layers, Ws, bs, engine='simt'):
    m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=engine)
    m.uv_net.set_weights(Ws, bs)
    return m


@pytest.mark.parametrize('engine', ['simt', 'tcf'])
def test_gpu_driver_reaches_scipy_loss(pe, golden, engine):
    """Same objective, same options: after the same evaluation budget the device-resident driver must be at least as low as
    1.5x the loss SciPy's L-BFGS-B reaches (different line search => different iterates; both monotone), and the loss it
    reports must equal a fresh evaluation at the parameters it leaves in the network."""
    g = golden('synthetic_5x50.npz')
    layers = [3, 50, 50, 50, 50, 50, 5] if engine == 'tcf' else [3, 20, 20, 5]
    Ws, bs = R.xavier_params(layers, seed=21)
    Collo, HOLE = g['f5_collo'][:2000], g['f5_hole'][:200]
    opts = dict(maxiter=150, maxfun=150, maxcor=50, maxls=50, ftol=1e-5 * np.finfo(float).eps)
    ms = _plate(pe, Collo, HOLE, layers, Ws, bs, engine)
    seq_s = []; ms.callback = lambda l: seq_s.append(l)
    rs = ms.train_bfgs(opts)
    mg = _plate(pe, Collo, HOLE, layers, Ws, bs, engine)
    seq_g = []; mg.callback = lambda l: seq_g.append(l)
    rg = mg.train_bfgs(dict(opts, driver='gpu'))
    assert len(seq_g) == rg.nfev and rg.nfev <= 150 + 1
    assert seq_g[0] == pytest.approx(seq_s[0], rel=1e-6)             # same starting point, same kernels
    assert rg.fun < 0.2 * seq_g[0]                                   # it optimises
    assert rg.fun <= 1.5 * rs.fun, (rg.fun, rs.fun)
    t = mg._evaluate_terms()
    assert mg._total(t) == pytest.approx(rg.fun, rel=1e-5)
    np.testing.assert_allclose(mg.uv_net.get_flat(), rg.x, rtol=0, atol=0)


def test_gpu_driver_limits_and_pretraining(pe, golden):
    """maxfun / maxiter are honoured; the pre-training entry points accept the driver switch (plate:527-559)"""
    g = golden('synthetic_5x50.npz')
    layers = [3, 20, 20, 5]
    Ws, bs = R.xavier_params(layers, seed=5)
    m = _plate(pe, g['f5_collo'][:500], g['f5_hole'][:50], layers, Ws, bs)
    n = []; m.callback = lambda l: n.append(l)
    r = m.train_bfgs(dict(maxiter=1000, maxfun=12, maxcor=5, maxls=50, driver='gpu'))
    assert r.nfev <= 13 and not r.success and 'EVALUATIONS' in r.message
    r = m.train_bfgs(dict(maxiter=3, maxfun=1000, maxcor=5, maxls=50, driver='gpu'))
    assert r.nit == 3 and 'ITERATIONS' in r.message
    m.bfgs_driver = 'gpu'                                             # attribute switch instead of the options key
    f0 = r.fun
    r = m.train_bfgs(dict(maxiter=10, maxfun=30, maxcor=5, maxls=50))
    assert r.fun < f0


@pytest.mark.parametrize('engine,layers', [('simt', [3, 18, 22, 5]), ('simt', [3, 20, 20, 5]), ('tcf', [3, 50, 50, 50, 50, 50, 5])])
def test_gradient_pads_are_zero(pe, golden, engine, layers):
    """The device-resident optimiser takes dot products over the PADDED parameter vector, so the pad entries of the reduced
    gradient (row strides rounded to 4, slack between matrices) must be exactly zero and stay zero through Adam steps."""
    g = golden('synthetic_5x50.npz')
    Ws, bs = R.xavier_params(layers, seed=9)
    m = _plate(pe, g['f5_collo'][:700], g['f5_hole'][:90], layers, Ws, bs, engine)
    net = m.uv_net
    real = net.pack(np.ones(net.P, np.float32)) != 0
    m.train(3, 1e-3)
    m.engine.evaluate()
    grad = m.engine.out[:net.Pp].cpu().numpy()
    assert np.all(grad[~real] == 0.0) and np.any(grad[real] != 0.0)
    assert np.all(net.params.cpu().numpy()[~real] == 0.0)
