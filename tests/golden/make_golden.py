"""Generate the committed golden fixtures under tests/golden/ (run once, in the BUILD container).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Sources:
  * /root/reference shipped checkpoints (*.pickle) and FEM frames (ProbeData-k.mat) -- the only
    artefacts of the reference that pin numbers; the TF1 code itself cannot run
    here, so oracle outputs stored below are produced by oracle/ref_torch.py (float64 autograd
    restatement) and are "oracle-pinned", not "reference-pinned".
  * numpy default_rng seeds written next to each block.
Nothing at test / bench time reads /root/reference: everything needed is copied into the .npz files.
"""
import os
import pickle
import sys

import numpy as np
import scipy.io

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_torch as R  # noqa: E402

REF = '/root/reference'


def fem(path, n, seed, shift=0.0):
    d = scipy.io.loadmat(path)
    cols = [d[k].flatten() for k in ('x', 'y', 'u', 'v', 's11', 's22', 's12')]
    A = np.stack(cols, 1)
    idx = np.random.default_rng(seed).choice(A.shape[0], n, replace=False)
    A = A[np.sort(idx)]
    A[:, 0] += shift
    A[:, 1] += shift
    return A


def pack(Ws, bs, prefix):
    out = {}
    for i, (w, b) in enumerate(zip(Ws, bs)):
        out[f'{prefix}_W{i}'] = np.asarray(w)
        out[f'{prefix}_b{i}'] = np.asarray(b)
    return out


def plate():
    d = REF + '/PlateHoleQuarter/train/'
    uv = pickle.load(open(d + 'uvNN_float64.pickle', 'rb'))
    di = pickle.load(open(d + 'distNN_float64.pickle', 'rb'))
    pa = pickle.load(open(d + 'partNN_float64.pickle', 'rb'))
    out = {}
    out.update(pack(*uv, 'uv')); out.update(pack(*di, 'dist')); out.update(pack(*pa, 'part'))
    # survey known-answer set: 5,000 uniform points rng(0) minus the r<=0.1 hole -> 4,862 rows
    r = np.random.default_rng(0)
    P = r.uniform([0, 0, 0], [.5, .5, 10], (5000, 3))
    P = P[np.hypot(P[:, 0], P[:, 1]) > 0.1]
    th = np.random.default_rng(1).uniform(0, np.pi / 2, 400)
    HOLE = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), np.random.default_rng(2).uniform(0, 10, 400)], 1)
    orc = R.Oracle('plate', uv[0], uv[1], dist=di, part=pa)
    T, loss, g = orc.loss_and_grad({'Collo': P, 'HOLE': HOLE})
    out['collo'] = P; out['hole'] = HOLE
    out['terms'] = np.array([T['loss_f_uv'], T['loss_f_s'], T['loss_HOLE'], loss])
    out['grad'] = g
    # FEM frames (t = k*0.125 s, plate:980-993 with N_t=81, MAX_T=10)
    for k in (10, 20, 50):
        A = fem(REF + f'/PlateHoleQuarter/FEM_result/Quarter_plate_hole_dynamic/ProbeData-{k}.mat', 400, k)
        t = np.full((A.shape[0], 1), k * 0.125)
        pred = orc.predict(A[:, 0:1], A[:, 1:2], t)
        out[f'fem{k}'] = A
        out[f'pred{k}'] = np.concatenate(pred, 1)
    np.savez_compressed(os.path.join(HERE, 'plate_ckpt.npz'), **out)
    print('plate', out['terms'])


def semi():
    uv = pickle.load(open(REF + '/ElasticWaveSemiInfinite/uv_NN#16s.pickle', 'rb'))
    out = pack(*uv, 'uv')
    r = np.random.default_rng(5)
    lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
    P = r.uniform(lb, ub, (3000, 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 2.0]
    IC = r.uniform(lb, ub, (300, 3)); IC[:, 2] = 0
    UP = r.uniform(lb, ub, (300, 3)); UP[:, 1] = 15
    th = r.uniform(0, 2 * np.pi, 300); ts = r.uniform(0, 16, 300)
    amp = (1 - 2 * (np.pi * (ts - 3) / 3) ** 2) * np.exp(-(np.pi * (ts - 3) / 3) ** 2) * 0.1
    SRC = np.stack([2 * np.cos(th), 2 * np.sin(th), ts, amp * np.cos(th), amp * np.sin(th)], 1)
    orc = R.Oracle('semi', uv[0], uv[1])
    sets = {'Collo': P, 'IC': IC, 'UP': UP, 'SRC': SRC}
    T, loss, g = orc.loss_and_grad(sets)
    out.update(collo=P, ic=IC, up=UP, src=SRC, grad=g,
               terms=np.array([T['loss_f_uv'], T['loss_f_s'], T['loss_IC'], T['loss_SRC'], T['loss_NB'], loss]))
    for k in (8, 24):   # frame k <-> t = k/4 s, FEM coords shifted by -45 (semi:475-476)
        A = fem(REF + f'/ElasticWaveSemiInfinite/FEM_result/ProbeData-{k}.mat', 400, k, shift=-45.0)
        t = np.full((A.shape[0], 1), k / 4.0)
        pred = orc.predict(A[:, 0:1], A[:, 1:2], t)
        out[f'fem{k}'] = A
        out[f'pred{k}'] = np.concatenate(pred, 1)
    np.savez_compressed(os.path.join(HERE, 'semi_ckpt.npz'), **out)
    print('semi', out['terms'])


def synthetic():
    """BASELINE config-1 shaped case, small: 5x50 nets, Xavier seed 1111, 20 Adam steps (oracle f64)."""
    out = {}
    r = np.random.default_rng(1111)
    P = r.uniform([0, 0, 0], [.5, .5, 10], (700, 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 0.1][:512]
    th = r.uniform(0, np.pi / 2, 64)
    HOLE = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), r.uniform(0, 10, 64)], 1)
    layers = [3] + 5 * [50] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    orc = R.Oracle('plate', Ws, bs)
    sets = {'Collo': P, 'HOLE': HOLE}
    T, loss, g = orc.loss_and_grad(sets)
    out.update(f5_collo=P, f5_hole=HOLE, f5_grad=g, f5_terms=np.array([T['loss_f_uv'], T['loss_f_s'], T['loss_HOLE'], loss]))
    rec = orc.train(sets, 20, 5e-4)
    out['f5_curve'] = np.stack([rec['loss_f_uv'], rec['loss_f_s'], rec['loss_HOLE'], rec['loss']], 1)
    out['f5_params_after'] = orc.flat_params()
    # F7 half-space shaped
    lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
    P = r.uniform(lb, ub, (700, 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 2.0][:512]
    IC = r.uniform(lb, ub, (64, 3)); IC[:, 2] = 0
    UP = r.uniform(lb, ub, (64, 3)); UP[:, 1] = 15
    th = r.uniform(0, 2 * np.pi, 64); ts = r.uniform(0, 16, 64)
    SRC = np.stack([2 * np.cos(th), 2 * np.sin(th), ts, 0.1 * np.cos(th) * np.sin(ts), 0.1 * np.sin(th) * np.sin(ts)], 1)
    layers = [3] + 5 * [50] + [7]
    Ws, bs = R.xavier_params(layers, seed=1111)
    # scale first layer so tanh is not saturated on the [-15,15] domain
    Ws[0] = Ws[0] * 0.1
    orc = R.Oracle('semi', Ws, bs)
    sets = {'Collo': P, 'IC': IC, 'UP': UP, 'SRC': SRC}
    T, loss, g = orc.loss_and_grad(sets)
    out.update(f7_collo=P, f7_ic=IC, f7_up=UP, f7_src=SRC, f7_grad=g, f7_W0=Ws[0],
               f7_terms=np.array([T['loss_f_uv'], T['loss_f_s'], T['loss_IC'], T['loss_SRC'], T['loss_NB'], loss]))
    rec = orc.train(sets, 20, 5e-4)
    out['f7_curve'] = np.stack([rec['loss_f_uv'], rec['loss_f_s'], rec['loss_IC'], rec['loss_SRC'], rec['loss_NB'], rec['loss']], 1)
    np.savez_compressed(os.path.join(HERE, 'synthetic_5x50.npz'), **out)
    print('synthetic', out['f5_terms'], out['f7_terms'])


if __name__ == '__main__':
    plate(); semi(); synthetic()
