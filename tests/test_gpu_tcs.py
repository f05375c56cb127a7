"""GPU parity tests of the tcgen05 engines -- the fp16-pair engine 'tcf' (csrc/pe_tcf.cu, what 'auto' selects) and its TF32x3 A/B partner
'tc3s' (csrc/pe_tcs.cu): F5 (K = 5 jet streams) and F7
(K = 4, 7 outputs) against the float64 oracle, same tolerances as tests/test_gpu_parity.py:
  loss terms rel 1e-5, gradient <= 2e-5 of each W_l / b_l block's max at random init (3e-5 for the 14-wide nets),
  Adam loss curve: weighted total 1e-5, individual terms 3e-5 (tensor-core accumulator truncation, see test_gpu_parity.py).
(The first two tcgen05 generations it was bit-identical to were removed in round 2; the fp16-pair engine 'tcf' superseded it as 'auto'.)
"""
import numpy as np
import pytest
import torch

from oracle import ref_torch as R
from tests.util import per_layer_grad_err, random_biases, rel_err, unpack_golden
from tests.test_gpu_parity import _check, _plate, _wave_sets

pytestmark = pytest.mark.gpu
ENGINES = [('tcf', 8), ('tc3s', 5)]          # (name, PE_ENGINE_* id): the fp16-pair engine ('auto') and its TF32x3 A/B partner


@pytest.fixture(params=ENGINES, ids=[e[0] for e in ENGINES])
def eng(request):
    return request.param


@pytest.fixture(scope='module')
def pe():
    assert torch.cuda.is_available()
    import pinn_elastodynamics_b200 as pe
    return pe


def test_f5_5x50_random_init(pe, golden, eng):
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    bs = random_biases(bs, 3)
    orc = R.Oracle('plate', Ws, bs)
    sets = {'Collo': g['f5_collo'], 'HOLE': g['f5_hole']}
    T, loss, gref = orc.loss_and_grad(sets)
    m = _plate(pe, sets['Collo'], sets['HOLE'], layers, Ws, bs, engine=eng[0])
    _check(m, T, ('loss_f_uv', 'loss_f_s', 'loss_HOLE'), gref, layers, 1e-5, 2e-5)
    assert m.engine.terms[0].engine == eng[1] and m.engine.terms[1].fused_into is m.engine.terms[0]


def test_run_to_run_deterministic(pe, golden, eng):
    """fixed-order accumulation everywhere: two evaluations give identical bits"""
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    outs = []
    for _ in range(2):
        m = _plate(pe, g['f5_collo'], g['f5_hole'], layers, Ws, bs, engine=eng[0])
        m.engine.evaluate()
        outs.append(m.engine.out.cpu().numpy().copy())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize('n', [1, 127, 128, 129, 128 * 150 + 5])
def test_f5_ragged_and_large_point_counts(pe, n, eng):
    """tail tiles (n % 128 != 0), fewer points than a tile, more tiles than SMs; narrow nets (K-steps < 7)"""
    rng = np.random.default_rng(n)
    layers = [3, 24, 40, 5]
    Ws, bs = R.xavier_params(layers, seed=9)
    bs = random_biases(bs, 4)
    Collo = rng.uniform([0, 0, 0], [.5, .5, 10], (n, 3))
    HOLE = rng.uniform([0, 0, 0], [.1, .1, 10], (max(1, n // 7), 3))
    orc = R.Oracle('plate', Ws, bs)
    T, loss, gref = orc.loss_and_grad({'Collo': Collo, 'HOLE': HOLE})
    m = _plate(pe, Collo, HOLE, layers, Ws, bs, engine=eng[0])
    _check(m, T, ('loss_f_uv', 'loss_f_s', 'loss_HOLE'), gref, layers, 2e-5, 5e-5)
    assert m.engine.terms[0].engine == eng[1]


def test_f5_composite_5x50(pe, golden, eng):
    g = golden('plate_ckpt.npz')
    di, pa = unpack_golden(g, 'dist'), unpack_golden(g, 'part')
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=77)
    bs = random_biases(bs, 8)
    Collo, HOLE = g['collo'][:1000], g['hole'][:100]
    orc = R.Oracle('plate', Ws, bs, dist=di, part=pa)
    T, loss, gref = orc.loss_and_grad({'Collo': Collo, 'HOLE': HOLE})
    m = _plate(pe, Collo, HOLE, layers, Ws, bs, dist=di, part=pa, engine=eng[0])
    _check(m, T, ('loss_f_uv', 'loss_f_s', 'loss_HOLE'), gref, layers, 2e-5, 5e-5)


def test_f5_adam_curve_matches_golden(pe, golden, eng):
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    m = _plate(pe, g['f5_collo'], g['f5_hole'], layers, Ws, bs, engine=eng[0])
    l_uv, l_s, l_h, loss = m.train(20, 5e-4)
    C = g['f5_curve']
    np.testing.assert_allclose(l_uv, C[:, 0], rtol=3e-5)
    np.testing.assert_allclose(l_s, C[:, 1], rtol=3e-5)
    np.testing.assert_allclose(l_h, C[:, 2], rtol=3e-5)
    np.testing.assert_allclose(loss, C[:, 3], rtol=1e-5)
    assert rel_err(m.uv_net.get_flat(), g['f5_params_after']) <= 1e-5


# ------------------------------------------------------------------------------ wave formulation on tensor cores (K = 4, 7 outputs)
def test_f7_semi_5x50(pe, golden, eng):
    """semi:228-272 residuals, loss 5,5,2,2,2 (semi:127): collocation term on tcgen05, IC term fused as extra tiles"""
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [7]
    Ws, bs = R.xavier_params(layers, seed=1111)
    Ws[0] = g['f7_W0']
    sets = _wave_sets(g, 'f7_')
    m = pe.DeepHPM(sets['Collo'], sets['SRC'], sets['IC'], sets['UP'], layers, None, None, verbose=False, engine=eng[0])
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    t = m.engine.terms_host()
    assert m.engine.terms[0].engine == eng[1], 'tensor-core engine was not selected for the F7 collocation term'
    assert m.engine.terms[1].fused_into is m.engine.terms[0]
    np.testing.assert_allclose(t[:5], g['f7_terms'][:5], rtol=1e-5)
    assert m._total(t) == pytest.approx(g['f7_terms'][5], rel=1e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), g['f7_grad'], layers)
    assert max(e for _, e in errs) <= 2e-5, errs


def test_f7_random_biases_and_ragged(pe, eng):
    rng = np.random.default_rng(5)
    layers = [3, 40, 56, 24, 7]
    Ws, bs = R.xavier_params(layers, seed=13)
    bs = random_biases(bs, 14)
    lb, ub = np.array([-15., -15., 0.]), np.array([15., 15., 16.])
    for n in (1, 130, 128 * 149 + 3):
        sets = {'Collo': rng.uniform(lb, ub, (n, 3)) * [0.1, 0.1, 0.2], 'IC': rng.uniform(lb, ub, (max(1, n // 9), 3)) * 0.1,
                'UP': rng.uniform(lb, ub, (max(1, n // 11), 3)) * 0.1,
                'SRC': np.concatenate([rng.uniform(lb, ub, (max(1, n // 5), 3)) * 0.1, rng.standard_normal((max(1, n // 5), 2)) * 0.1], 1)}
        orc = R.Oracle('semi', Ws, bs)
        T, loss, gref = orc.loss_and_grad(sets)
        m = pe.DeepHPM(sets['Collo'], sets['SRC'], sets['IC'], sets['UP'], layers, None, None, verbose=False, engine=eng[0])
        m.uv_net.set_weights(Ws, bs)
        t, gr = _check(m, T, ('loss_f_uv', 'loss_f_s', 'loss_IC', 'loss_SRC', 'loss_NB'), gref, layers, 2e-5, 5e-5)
        assert m.engine.terms[0].engine == eng[1]
        assert m._total(t) == pytest.approx(loss, rel=2e-5)


@pytest.mark.parametrize('variant', ['inf', 'conf'])
def test_f7_other_variants(pe, variant, eng):
    """inf: input normalisation 2(X-lb)/(ub-lb)-1 (inf:191) through the tensor-core first layer; conf: FIX term (conf:156)"""
    rng = np.random.default_rng(11)
    layers = [3, 14, 14, 14, 7]
    Ws, bs = R.xavier_params(layers, seed=5)
    bs = random_biases(bs, 6)
    lb, ub = np.array([0., 0., 0.]), np.array([30., 30., 20.])
    sets = {'Collo': rng.uniform(lb, ub, (333, 3)), 'IC': rng.uniform(lb, ub, (45, 3)), 'UP': rng.uniform(lb, ub, (37, 3)),
            'FIXED': rng.uniform(lb, ub, (65, 3)),
            'SRC': np.concatenate([rng.uniform(lb, ub, (50, 3)), rng.standard_normal((50, 2)) * 0.1], 1)}
    orc = R.Oracle(variant, Ws, bs, lb=lb, ub=ub)
    T, loss, gref = orc.loss_and_grad(sets)
    if variant == 'inf':
        m = pe.DeepHPM(sets['Collo'], sets['SRC'], sets['IC'], sets['UP'], layers, lb, ub, variant='inf', verbose=False, engine=eng[0])
        names = ('loss_f_uv', 'loss_f_s', 'loss_IC', 'loss_SRC')
    else:
        m = pe.DeepElasticWave(sets['Collo'], sets['SRC'], sets['IC'], sets['FIXED'], None, layers, None, None, lb, ub, verbose=False, engine=eng[0])
        names = ('loss_f_uv', 'loss_f_s', 'loss_SRC', 'loss_IC', 'loss_FIX')
    m.uv_net.set_weights(Ws, bs)
    t, g = _check(m, T, names, gref, layers, 1e-5, 3e-5)
    assert m.engine.terms[0].engine == eng[1]
    assert m._total(t) == pytest.approx(loss, rel=1e-5)


def test_f7_adam_curve_and_chunking(pe, golden, eng):
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [7]
    Ws, bs = R.xavier_params(layers, seed=1111)
    Ws[0] = g['f7_W0']
    sets = _wave_sets(g, 'f7_')
    m = pe.DeepHPM(sets['Collo'], sets['SRC'], sets['IC'], sets['UP'], layers, None, None, verbose=False, engine=eng[0])
    m.uv_net.set_weights(Ws, bs)
    out = m.train(20, 5e-4, 1)
    C = g['f7_curve']
    for i, col in enumerate((0, 1, 2, 3)):
        np.testing.assert_allclose(out[i], C[:, col], rtol=3e-5)
    np.testing.assert_allclose(out[4], C[:, 5], rtol=1e-5)
    # batch_num chunking (semi:299-326): 3 chunks x 2 iterations vs oracle
    m2 = pe.DeepHPM(sets['Collo'], sets['SRC'], sets['IC'], sets['UP'], layers, None, None, verbose=False, engine=eng[0])
    m2.uv_net.set_weights(Ws, bs)
    orc = R.Oracle('semi', Ws, bs)
    rec = orc.train(sets, 2, 5e-4, batch_num=3)
    out2 = m2.train(2, 5e-4, 3)
    assert len(out2[4]) == 6
    np.testing.assert_allclose(out2[4], rec['loss'], rtol=1e-5)


def test_engine_auto_and_env_default(pe, golden, monkeypatch, eng):
    """engine='auto' = the fp16-pair tcgen05 engine for the collocation term (wide nets fall back to the SIMT engine per term); a constructor
    call with the reference's own signature (no engine argument) reads $PE_ENGINE and defaults to 'auto'."""
    g = golden('synthetic_5x50.npz')
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    outs = {}
    if eng[0] != 'tcf':
        pytest.skip("'auto' is the fp16-pair engine")
    for name in ('tcf', 'auto'):
        m = _plate(pe, g['f5_collo'], g['f5_hole'], layers, Ws, bs, engine=name)
        m.engine.evaluate()
        assert m.engine.terms[0].engine == eng[1]
        outs[name] = m.engine.out.cpu().numpy().copy()
    assert np.array_equal(outs['tcf'], outs['auto'])
    monkeypatch.setenv('PE_ENGINE', 'auto')
    m = pe.PINN(g['f5_collo'], g['f5_hole'], None, None, None, None, None, None, layers, None, None, None, None, verbose=False)
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    assert m.engine.terms[0].engine == eng[1] and np.array_equal(m.engine.out.cpu().numpy(), outs['auto'])
    monkeypatch.setenv('PE_ENGINE', 'simt')          # the opt-out for a caller that keeps the reference's constructor signature
    m = pe.PINN(g['f5_collo'], g['f5_hole'], None, None, None, None, None, None, layers, None, None, None, None, verbose=False)
    m.engine.build()
    assert m.engine.terms[0].engine == 0
    monkeypatch.delenv('PE_ENGINE')                  # default: 'auto'
    m = pe.PINN(g['f5_collo'], g['f5_hole'], None, None, None, None, None, None, layers, None, None, None, None, verbose=False)
    m.engine.build()
    assert m.engine.terms[0].engine == eng[1]
    # hidden width 70 > 56: 'auto' keeps every term on the SIMT engine
    wide = [3] + 3 * [70] + [5]
    m = pe.PINN(g['f5_collo'][:256], g['f5_hole'][:32], None, None, None, None, None, None, wide, None, None, None, None, verbose=False, engine='auto')
    m.engine.evaluate()
    assert all(t.engine == 0 for t in m.engine.terms) and np.isfinite(m.engine.terms_host()).all()
