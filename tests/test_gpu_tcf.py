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


@pytest.mark.parametrize('case', ['plate', 'plate_composite', 'semi'])
def test_predict_on_the_tensor_core_forward_sweep(pe, G, case):
    """`predict` of a frame batch (>= 512 points) runs the forward sweep of the tcgen05 engine (pe_launch_fields_tcf); small batches and
    PE_FIELDS_ENGINE=simt run the fp32 SIMT fields kernel.  Same fields to 2e-5 of each field's max (the bar of predict against the
    reference's own predict output), ragged batch size, composite u = P + D N included (plate:382-395, 449-461)."""
    from pinn_elastodynamics_b200 import _lib as L
    from tests.test_gpu_reference_golden import _plate_model
    lib = L.load()
    rng = np.random.default_rng(5)
    n = 128 * 37 + 61
    if case == 'semi':
        S = {k: G[f'semi_{k}'] for k in ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub')}
        Ws, bs = _uv(G, 'semi')
        m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers_of(Ws), S['lb'], S['ub'], variant='semi', verbose=False, engine='tcf')
        m.uv_net.set_weights(Ws, bs)
        X = rng.uniform(S['lb'], S['ub'], (n, 3))
    else:
        m, _, _ = _plate_model(pe, G, 'tcf', composite=(case == 'plate_composite'))
        X = rng.uniform([0, 0, 0], [.5, .5, 10], (n, 3))
    out = {}
    try:
        for name, eng in (('simt', L.ENGINE_SIMT_FP32), ('tcf', L.ENGINE_TCF), ('tcf2', L.ENGINE_TCF)):
            lib.pe_debug_set_fields_engine(eng)
            out[name] = np.concatenate(m.predict(X[:, 0:1], X[:, 1:2], X[:, 2:3]), 1).astype(np.float64)
    finally:
        lib.pe_debug_set_fields_engine(-1)
    assert np.array_equal(out['tcf'], out['tcf2'])
    assert not np.array_equal(out['tcf'], out['simt'])                 # the two kernels really are different code paths
    for c in range(8):
        assert np.abs(out['tcf'][:, c] - out['simt'][:, c]).max() <= 2e-5 * max(1.0, np.abs(out['simt'][:, c]).max()), c


def test_16_bit_forward_mode_tolerance(pe, G):
    """BASELINE config 3 names a '16-bit forward / fp32 gradient' mode: engine 'tcf16' computes the forward layer GEMMs as single fp16 x fp16
    products (fp32 accumulation; operands rounded to 11 bits) and keeps the fp16-pair arithmetic for the adjoint and weight-gradient GEMMs.
    Its tolerance class against the reference's own values (half-space wave script, semi:99-127) is stated here: loss terms 5e-3, gradient
    blocks 5e-3 of each block's max (measured on B200: 2.2e-3 / 9.3e-4, profiles/r2_tcf16_check.jsonl) -- NOT the 1e-5 class of 'tcf'."""
    S = {k: G[f'semi_{k}'] for k in ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub')}
    Ws, bs = _uv(G, 'semi')
    layers = layers_of(Ws)
    m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers, S['lb'], S['ub'], variant='semi', verbose=False, engine='tcf16')
    m.uv_net.set_weights(Ws, bs)
    m.engine.evaluate()
    assert m.engine.terms[0].engine == 9
    t = m.engine.terms_host()
    ref = G['semi_terms']
    n = len(ref) - 1
    np.testing.assert_allclose(t[:n], ref[:n], rtol=5e-3)
    assert not np.allclose(t[:2], ref[:2], rtol=1e-6)                 # it really is the reduced-precision path
    errs = per_layer_grad_err(m.engine.grad_compact_host(), G['semi_grad'], layers)
    assert max(e for _, e in errs) <= 5e-3, errs
