"""CPU tests of the oracle itself: finite differences, golden regression, shipped-checkpoint pins.

The reference has no tests; these pin the restatement in oracle/ref_torch.py by
(i) finite differences, (ii) the survey's known-answer residual losses of the shipped plate
checkpoints, (iii) FEM rel-L2 bands, (iv) agreement of the two independent restatements
(nested autograd vs forward-jet + hand adjoint)."""
import numpy as np
import pytest
import torch

from oracle import jet_numpy as J
from oracle import ref_torch as R


def _unpack(g, prefix):
    Ws, bs, i = [], [], 0
    while f'{prefix}_W{i}' in g:
        Ws.append(g[f'{prefix}_W{i}']); bs.append(g[f'{prefix}_b{i}']); i += 1
    return Ws, bs


def test_finite_difference_f5_and_f7():
    rng = np.random.default_rng(7)
    for kind, O in (('plate', 5), ('semi', 7), ('conf', 7), ('inf', 7)):
        layers = [3, 10, 10, O]
        Ws, bs = R.xavier_params(layers, seed=2)
        bs = [rng.standard_normal(b.shape) * 0.1 for b in bs]
        sets = {'Collo': rng.uniform(0, 1, (40, 3)), 'HOLE': rng.uniform(0, .1, (9, 3)),
                'IC': rng.uniform(0, 1, (9, 3)), 'SRC': rng.uniform(0, 1, (9, 5)),
                'UP': rng.uniform(0, 1, (9, 3)), 'FIXED': rng.uniform(0, 1, (9, 3))}
        orc = R.Oracle(kind, Ws, bs, lb=np.zeros(3), ub=np.ones(3))
        _, loss, g = orc.loss_and_grad(sets)
        flat = orc.flat_params()
        for idx in rng.choice(flat.size, 12, replace=False):
            h = 1e-6
            p = flat.copy(); p[idx] += h; orc.set_flat_params(p); _, lp, _ = orc.loss_and_grad(sets)
            p = flat.copy(); p[idx] -= h; orc.set_flat_params(p); _, lm, _ = orc.loss_and_grad(sets)
            fd = (lp - lm) / (2 * h)
            assert abs(fd - g[idx]) <= 1e-6 * max(1.0, abs(g[idx])), (kind, idx, fd, g[idx])


def test_plate_checkpoint_known_answers(golden):
    g = golden('plate_ckpt.npz')
    uv, di, pa = _unpack(g, 'uv'), _unpack(g, 'dist'), _unpack(g, 'part')
    assert [w.shape for w in uv[0]] == [(3, 70)] + 7 * [(70, 70)] + [(70, 5)]
    orc = R.Oracle('plate', *uv, dist=di, part=pa)
    T, loss, grad = orc.loss_and_grad({'Collo': g['collo'], 'HOLE': g['hole']})
    # SURVEY.md section 8(c) known-answer values (float64) for the shipped checkpoints
    assert T['loss_f_uv'] == pytest.approx(3.8686e-05, rel=2e-5)
    assert T['loss_f_s'] == pytest.approx(2.4435e-05, rel=2e-5)
    np.testing.assert_allclose([T['loss_f_uv'], T['loss_f_s'], T['loss_HOLE'], loss], g['terms'], rtol=1e-10)
    np.testing.assert_allclose(grad, g['grad'], rtol=1e-8, atol=1e-12)


def test_plate_checkpoint_vs_fem(golden):
    """Loose physics pin: composite prediction vs FEM frames, bands from SURVEY.md section 4."""
    g = golden('plate_ckpt.npz')
    orc = R.Oracle('plate', *_unpack(g, 'uv'), dist=_unpack(g, 'dist'), part=_unpack(g, 'part'))
    bands = {2: 0.03, 3: 0.06, 4: 0.012, 5: 0.12, 6: 0.04}      # fem col -> max rel-L2 (u, v, s11, s22, s12)
    for k in (10, 20, 50):
        A = g[f'fem{k}']
        t = np.full((A.shape[0], 1), k * 0.125)
        pred = np.concatenate(orc.predict(A[:, 0:1], A[:, 1:2], t), 1)
        np.testing.assert_allclose(pred, g[f'pred{k}'], rtol=1e-9, atol=1e-12)
        for c, band in bands.items():
            rel = np.linalg.norm(pred[:, c - 2] - A[:, c]) / np.linalg.norm(A[:, c])
            assert rel < band, (k, c, rel)


def test_semi_checkpoint(golden):
    g = golden('semi_ckpt.npz')
    uv = _unpack(g, 'uv')
    assert uv[0][0].dtype == np.float32 and [w.shape for w in uv[0]][-1] == (100, 7)
    orc = R.Oracle('semi', *uv)
    sets = {'Collo': g['collo'], 'IC': g['ic'], 'UP': g['up'], 'SRC': g['src']}
    T, loss, grad = orc.loss_and_grad(sets)
    np.testing.assert_allclose([T['loss_f_uv'], T['loss_f_s'], T['loss_IC'], T['loss_SRC'], T['loss_NB'], loss],
                               g['terms'], rtol=1e-10)
    np.testing.assert_allclose(grad, g['grad'], rtol=1e-8, atol=1e-12)
    A = g['fem8']
    pred = np.concatenate(orc.predict(A[:, 0:1], A[:, 1:2], np.full((A.shape[0], 1), 2.0)), 1)
    rel = np.linalg.norm(pred[:, 0] - A[:, 2]) / np.linalg.norm(A[:, 2])
    assert rel < 0.03, rel                                         # 1.2 % at frame 8


def test_jet_restatement_matches_autograd(golden):
    g = golden('synthetic_5x50.npz')
    Ws, bs = R.xavier_params([3] + 5 * [50] + [5], seed=1111)
    orc = R.Oracle('plate', Ws, bs)
    T, _ = orc.loss_terms({'Collo': g['f5_collo'], 'HOLE': g['f5_hole']})
    gs = torch.autograd.grad(10 * (T['loss_f_uv'] + T['loss_f_s']), orc.params())
    gref = np.concatenate([x.numpy().ravel() for x in gs])
    luv, ls, dW, db = J.loss_grad_residual('f5', g['f5_collo'], Ws, bs, 10, 10, 20.0, .25, 1.0)
    gj = np.concatenate([x.ravel() for x in dW] + [x.ravel() for x in db])
    assert luv == pytest.approx(float(T['loss_f_uv']), rel=1e-12)
    assert ls == pytest.approx(float(T['loss_f_s']), rel=1e-12)
    assert np.abs(gj - gref).max() <= 1e-12 * np.abs(gref).max()
    # float32 evaluation of the same algebra: the error level the fp32 CUDA path is held to
    Ws32 = [w.astype(np.float32) for w in Ws]; bs32 = [b.astype(np.float32) for b in bs]
    luv32, ls32, dW32, _ = J.loss_grad_residual('f5', g['f5_collo'].astype(np.float32), Ws32, bs32, 10, 10, 20.0, .25, 1.0,
                                                dtype=np.float32)
    assert luv32 == pytest.approx(luv, rel=2e-5) and ls32 == pytest.approx(ls, rel=2e-5)


def test_golden_curve_regression(golden):
    g = golden('synthetic_5x50.npz')
    Ws, bs = R.xavier_params([3] + 5 * [50] + [5], seed=1111)
    orc = R.Oracle('plate', Ws, bs)
    rec = orc.train({'Collo': g['f5_collo'], 'HOLE': g['f5_hole']}, 5, 5e-4)
    np.testing.assert_allclose(rec['loss'], g['f5_curve'][:5, 3], rtol=1e-9)


def test_tf1_adam_differs_from_torch_adam():
    """epsilon is added to sqrt(v) un-bias-corrected; one step on a known gradient."""
    orc = R.Oracle('plate', [np.ones((3, 5))], [np.zeros((1, 5))])
    g = [torch.full((3, 5), 1e-6, dtype=torch.float64), torch.full((1, 5), 1e-6, dtype=torch.float64)]
    orc.adam_step(g, 1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = 1.0 - lr_t * (0.1 * 1e-6) / (np.sqrt(0.001 * 1e-12) + 1e-8)
    assert float(orc.W[0][0, 0]) == pytest.approx(expect, rel=1e-12)
