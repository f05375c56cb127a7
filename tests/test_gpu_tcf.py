"""The fp16-pair tcgen05 engine ('tcf', csrc/pe_tcf.cu; what engine='auto' selects) against the golden vectors produced by the reference's own
class files (tests/golden/reference_tf1shim.npz) and against the SIMT engine.  Bars: loss terms 1e-5, gradient blocks 2e-5 (the float32 `inf`
script: 2e-5 / 1e-4), the 20-step reference Adam curve 1e-5 (north_star)."""
import os

import numpy as np
import pytest
import torch

from tests.util import layers_of, per_layer_grad_err, unpack_golden

pytestmark = pytest.mark.gpu
TCF = 8


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


def test_plate_plain_against_reference_source(pe, G):
    Ws, bs = _uv(G, 'plate')
    layers = layers_of(Ws)
    m = pe.PINN(G['plate_Collo'], G['plate_HOLE'], None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tcf')
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    assert m.engine.terms[0].engine == TCF and m.engine.terms[1].fused_into is m.engine.terms[0]
    t = m.engine.terms_host()
    ref = G['plate_plain_terms']
    np.testing.assert_allclose(t[:3], ref[:3], rtol=1e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G['plate_plain_grad'], layers)
    assert max(e for _, e in errs) <= 2e-5, errs
    out = m.train(20, 5e-4)
    np.testing.assert_allclose(out[3], G['plate_plain_adam'][:, 3], rtol=1e-5)


@pytest.mark.parametrize('kind', ['semi', 'inf', 'conf'])
def test_waves_against_reference_source(pe, G, kind):
    S = {k: G[f'{kind}_{k}'] for k in (('Collo', 'SRC', 'IC', 'FIXED', 'lb', 'ub') if kind == 'conf' else ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub'))}
    Ws, bs = _uv(G, kind)
    layers = layers_of(Ws)
    if kind == 'conf':
        m = pe.DeepElasticWave(S['Collo'], S['SRC'], S['IC'], S['FIXED'], None, layers, None, None, S['lb'], S['ub'], verbose=False, engine='tcf')
    else:
        m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers, S['lb'], S['ub'], variant=kind, verbose=False, engine='tcf')
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    assert m.engine.terms[0].engine == TCF
    t = m.engine.terms_host()
    ref = G[f'{kind}_terms']
    n = len(ref) - 1
    np.testing.assert_allclose(t[:n], ref[:n], rtol=2e-5 if kind == 'inf' else 1e-5)
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G[f'{kind}_grad'], layers)
    assert max(e for _, e in errs) <= (1e-4 if kind == 'inf' else 3e-5), errs


@pytest.mark.parametrize('n', [1, 127, 129, 128 * 150 + 5])
def test_ragged_point_counts_match_the_simt_engine(pe, n):
    from oracle import ref_torch as R
    rng = np.random.default_rng(n)
    layers = [3, 14, 30, 50, 5]
    Ws, bs = R.xavier_params(layers, seed=4)
    bs = [rng.standard_normal(b.shape) * 0.1 for b in bs]
    Collo = rng.uniform([0, 0, 0], [.5, .5, 10], (n, 3))
    th = rng.uniform(0, np.pi / 2, max(n // 10, 1))
    HOLE = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), rng.uniform(0, 10, th.size)], 1)
    outs = {}
    for eng in ('simt', 'tcf', 'tcf'):
        m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=eng)
        m.uv_net.set_weights(Ws, bs)
        m.engine.evaluate()
        outs.setdefault(eng, []).append((m.engine.terms_host()[:3].copy(), m.engine.grad_compact_host().astype(np.float64)))
    assert np.array_equal(outs['tcf'][0][1], outs['tcf'][1][1])           # run-to-run deterministic
    np.testing.assert_allclose(outs['tcf'][0][0], outs['simt'][0][0], rtol=1e-5)
    errs = per_layer_grad_err(outs['tcf'][0][1], outs['simt'][0][1], layers)
    assert max(e for _, e in errs) <= 3e-5, errs
