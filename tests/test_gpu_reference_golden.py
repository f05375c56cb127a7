"""GPU: the CUDA path, called through the drop-in classes exactly as the reference's classes are called, against golden vectors
produced by the reference's own class files (tests/golden/reference_tf1shim.npz; see tests/golden/make_reference_golden.py and
tests/test_reference_golden.py for how they were made and how the oracle is pinned to them).

fp32 tolerances (the reference computes plate / semi / conf in float64, inf in float32):
  loss terms 1e-5 (2e-5 composite), gradient <= 2e-5 of each W_l / b_l block's max (3e-5 wave nets, 5e-5 composite),
  Adam loss curves (the weighted total = the curve north_star names): 1e-5 on BOTH engines (measured: SIMT <= 7.5e-6, fp16-pair tcgen05 engine
  <= 8.3e-6, profiles/r2_refgold_report_tcf.jsonl); per-term curves 1e-5, except the smallest term of the deliberately violent plate trajectory
  on the tcgen05 engine (3e-5; see the comment there);
  predicted fields 2e-5 of the field's max.
"""
import numpy as np
import pytest
import torch

from tests.util import layers_of, per_layer_grad_err, rel_err, unpack_golden

pytestmark = pytest.mark.gpu
ENGINES = ['simt', 'tcf']


@pytest.fixture(scope='module')
def pe():
    assert torch.cuda.is_available()
    import pinn_elastodynamics_b200 as pe
    return pe


@pytest.fixture(scope='module')
def G(golden):
    return golden('reference_tf1shim.npz')


def _uv(G, kind):
    Ws, bs = unpack_golden(G, f'{kind}_uv')
    return [np.asarray(w, np.float64) for w in Ws], [np.asarray(b, np.float64) for b in bs]


def _plate_model(pe, G, engine, composite):
    S = {k: G['plate_' + k] for k in ('Collo', 'HOLE', 'IC', 'LF', 'RT', 'UP', 'LW', 'DIST', 'lb', 'ub')}
    Ws, bs = _uv(G, 'plate')
    layers = layers_of(Ws)
    if composite:
        di, pa = unpack_golden(G, 'plate_dist'), unpack_golden(G, 'plate_part')
        m = pe.PINN(S['Collo'], S['HOLE'], S['IC'], S['LF'], S['RT'], S['UP'], S['LW'], S['DIST'], layers, layers_of(di[0]), layers_of(pa[0]),
                    S['lb'], S['ub'], verbose=False, engine=engine)
        m.dist_net.set_weights(*di); m.part_net.set_weights(*pa); m.refresh_composite()
    else:
        m = pe.PINN(S['Collo'], S['HOLE'], None, None, None, None, None, None, layers, None, None, S['lb'], S['ub'], verbose=False, engine=engine)
    m.uv_net.set_weights(Ws, bs)
    return m, layers, S


def _fields_close(pred, ref, tol=2e-5):
    pred = np.concatenate(pred, 1)
    for c in range(ref.shape[1]):
        assert np.abs(pred[:, c] - ref[:, c]).max() <= tol * max(1.0, np.abs(ref[:, c]).max()), c


@pytest.mark.parametrize('engine', ENGINES)
def test_plate_plain_against_reference_source(pe, G, engine):
    m, layers, S = _plate_model(pe, G, engine, composite=False)
    m.engine.evaluate()
    t = m.engine.terms_host()
    ref = G['plate_plain_terms']
    np.testing.assert_allclose(t[:3], ref[:3], rtol=1e-5)
    assert m._total(t) == pytest.approx(ref[3], rel=1e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G['plate_plain_grad'], layers)
    assert max(e for _, e in errs) <= 2e-5, errs
    xs = S['Collo'][:40]
    _fields_close(m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), G['plate_plain_predict'])
    # PINN.train(iter, learning_rate) -> (loss_f_uv[], loss_f_s[], loss_HOLE[], loss[]), recorded after each update (plate:475-506)
    out = m.train(20, 5e-4)
    C = G['plate_plain_adam']
    # This trajectory is violent on purpose (loss 73 -> 1.5 -> 4 within 20 steps at lr 5e-4): it amplifies a per-evaluation error about
    # tenfold.  The weighted total -- the loss curve north_star asks to match to 1e-5 -- is held to 1e-5 on both engines (measured 7.5e-6 SIMT,
    # 8.2e-6 tcgen05).  Per term: SIMT 1e-5; tcgen05 1e-5 on loss_f_uv / loss_f_s and 3e-5 on loss_HOLE, the smallest term (0.3 % of the
    # total), measured 1.9e-5 at the step-14 minimum (profiles/r2_refgold_report_tcf.jsonl).
    np.testing.assert_allclose(out[3], C[:, 3], rtol=1e-5)
    for i in range(3):
        np.testing.assert_allclose(out[i], C[:, i], rtol=3e-5 if (engine != 'simt' and i == 2) else 1e-5)
    assert rel_err(m.uv_net.get_flat(), G['plate_plain_params_after_adam']) <= 1e-5      # measured 3e-7 (simt), 4e-7 (tcf)


@pytest.mark.parametrize('engine', ENGINES)
def test_plate_composite_against_reference_source(pe, G, engine):
    m, layers, S = _plate_model(pe, G, engine, composite=True)
    m.engine.evaluate()
    t = m.engine.terms_host()
    ref = G['plate_comp_terms']            # loss_f_uv, loss_f_s, loss_HOLE, loss, loss_PART, loss_DIST
    np.testing.assert_allclose(t[:3], ref[:3], rtol=2e-5)
    assert m._total(t) == pytest.approx(ref[3], rel=2e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G['plate_comp_grad'], layers)
    assert max(e for _, e in errs) <= 5e-5, errs
    xs = S['Collo'][:40]
    _fields_close(m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), G['plate_comp_predict'])
    _fields_close(m.predict_D(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), G['plate_comp_predict_D'])
    _fields_close(m.predict_P(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), G['plate_comp_predict_P'])
    if engine == 'simt':       # pre-training losses of the frozen nets (plate:194-215), evaluated by the SIMT engine in either case
        m._pre_engines()
        m.part_engine.evaluate(); m.dist_engine.evaluate()
        assert m.part_engine.terms_host()[0] == pytest.approx(ref[4], rel=1e-5)
        assert m.dist_engine.terms_host()[0] == pytest.approx(ref[5], rel=1e-5)
    out = m.train(8, 5e-4)
    np.testing.assert_allclose(out[3], G['plate_comp_adam'][:, 3], rtol=1e-5)


@pytest.mark.parametrize('engine', ENGINES)
@pytest.mark.parametrize('kind', ['semi', 'inf', 'conf'])
def test_waves_against_reference_source(pe, G, kind, engine):
    S = {k: G[f'{kind}_{k}'] for k in (('Collo', 'SRC', 'IC', 'FIXED', 'lb', 'ub') if kind == 'conf' else ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub'))}
    Ws, bs = _uv(G, kind)
    layers = layers_of(Ws)
    if kind == 'conf':
        m = pe.DeepElasticWave(S['Collo'], S['SRC'], S['IC'], S['FIXED'], None, layers, None, None, S['lb'], S['ub'], verbose=False, engine=engine)
    else:
        m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers, S['lb'], S['ub'], variant=kind, verbose=False, engine=engine)
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    t = m.engine.terms_host()
    ref = G[f'{kind}_terms']
    n = len(ref) - 1
    np.testing.assert_allclose(t[:n], ref[:n], rtol=2e-5 if kind == 'inf' else 1e-5)
    assert m._total(t) == pytest.approx(ref[n], rel=2e-5 if kind == 'inf' else 1e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G[f'{kind}_grad'], layers)
    assert max(e for _, e in errs) <= (1e-4 if kind == 'inf' else 3e-5), errs
    xs = S['Collo'][:40]
    _fields_close(m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), G[f'{kind}_predict'], 5e-5 if kind == 'inf' else 2e-5)
    # train(iter, learning_rate, batch_num = 2): the reference's chunked Adam loop (semi:289-326); the weighted total is the last column
    out = m.train(6, 1e-3, 2)
    C = G[f'{kind}_adam_b2']
    np.testing.assert_allclose(out[-1], C[:, -1], rtol=2e-3 if kind == 'inf' else 1e-5)      # inf: the reference script itself runs in float32
    np.testing.assert_allclose(out[0], C[:, 0], rtol=2e-3 if kind == 'inf' else 1e-5)


# ------------------------------------------------------------------------------ L-BFGS-B sequences of the reference's own train_bfgs (plate:508-525, semi:330-346)
def _bfgs_check(seen, ref, engine):
    """The reference's callback saw `ref` (one loss per function evaluation, float64).  In fp32 the first evaluations are the same points --
    the initial one and the first line-search trials -- and must agree to the evaluation tolerance; later iterates depend on line-search
    decisions that rounding may flip, so the rest of the trace is held to the reached loss level only."""
    assert len(seen) >= 3 and len(ref) >= 3
    np.testing.assert_allclose(seen[:3], ref[:3], rtol=1e-5 if engine == 'simt' else 2e-5)
    assert min(seen) <= 1.02 * min(ref) + 1e-12


@pytest.mark.parametrize('engine', ENGINES)
def test_plate_lbfgs_sequence_against_reference_source(pe, G, engine):
    m, layers, S = _plate_model(pe, G, engine, composite=False)
    m.train(20, 5e-4)                                  # the reference's trace continues from its Adam state (make_reference_golden.py)
    seen = []
    m.callback = lambda loss: seen.append(float(loss))
    res = m.train_bfgs(dict(maxiter=6, maxfun=24, maxcor=50, maxls=50, ftol=0.00001 * np.finfo(float).eps))      # plate:243-247, budget as in the generator
    _bfgs_check(seen, G['plate_plain_bfgs_losses'], engine)
    assert res.nfev == len(seen)
    assert rel_err(m.uv_net.get_flat(), G['plate_plain_params_after_bfgs']) <= 5e-3


@pytest.mark.parametrize('engine', ENGINES)
def test_semi_lbfgs_sequence_against_reference_source(pe, G, engine):
    S = {k: G[f'semi_{k}'] for k in ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub')}
    Ws, bs = _uv(G, 'semi')
    m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers_of(Ws), S['lb'], S['ub'], variant='semi', verbose=False, engine=engine)
    m.uv_net.set_weights(Ws, bs)
    m.train(6, 1e-3, 2)
    n0 = len(m.loss_rec)
    m.train_bfgs(1, dict(maxiter=5, maxfun=20, maxcor=50, maxls=50, ftol=0.001 * np.finfo(float).eps))           # semi:133-137
    seen = m.loss_rec[n0:]                             # semi:287: the callback appends every evaluation's loss
    _bfgs_check(seen, G['semi_bfgs_losses'], engine)
    assert rel_err(m.uv_net.get_flat(), G['semi_params_after_bfgs']) <= 5e-3
