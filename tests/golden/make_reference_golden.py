"""Golden vectors produced by the REFERENCE'S OWN class files, executed unmodified from /root/reference in the build container.

    PYTHONPATH=/root/repo python tests/golden/make_reference_golden.py          -> tests/golden/reference_tf1shim.npz

TensorFlow 1.x cannot be imported here, so the files run on oracle/tf1_shim.py, a small graph-mode emulator of the TF
primitives they call (placeholder / Variable / matmul / tanh / gradients / reduce_mean / Session / AdamOptimizer /
ScipyOptimizerInterface) over torch CPU.  Everything above the primitives -- net_uv, net_e, net_f_sig, net_t, the composite
P + D*N, the loss assembly and weights, the Adam loop with its post-update bookkeeping, batch_num chunking, the L-BFGS-B call --
is the reference's code running as written (PINN: PlateHoleQuarter/train/train.py:26-612; DeepHPM: ElasticWaveSemiInfinite/
ElasticWave.py:23-392 and ElasticWaveInfinite/ElasticWave.py:21-376; DeepElasticWave: ElasticWaveConfined/ElasticWave.py:21-475).
Parity label of these fixtures: "reference source over a TF1 shim" (stronger than the oracle-generated synthetic_5x50.npz,
weaker than real TensorFlow: Adam / L-BFGS-B glue and the primitives follow TF's documentation).

Weights are loaded through the reference's own load_NN from pickles written here (numpy seeds below), inputs are small
random point sets.  tests/test_reference_golden.py checks the float64 oracle against this file on the CPU;
tests/test_gpu_reference_golden.py checks the CUDA path against it on the GPU.  Nothing at test time reads /root/reference.
"""
import contextlib
import io
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tf1_shim as tf  # noqa: E402

REF = '/root/reference'
TMP = tempfile.mkdtemp(prefix='pe_refgold_')
OUT = {}


def xavier(layers, rng, bias_scale=0.0, w_scale=1.0):
    Ws, bs = [], []
    for i in range(len(layers) - 1):
        fi, fo = layers[i], layers[i + 1]
        std = np.sqrt(2.0 / (fi + fo))
        w = np.clip(rng.standard_normal((fi, fo)), -2, 2) * std * w_scale
        Ws.append(w)
        bs.append(rng.standard_normal((1, fo)) * bias_scale)
    return Ws, bs


def dump(Ws, bs, name):
    path = os.path.join(TMP, name)
    with open(path, 'wb') as f:
        pickle.dump([Ws, bs], f)
    return path


def store(prefix, Ws, bs):
    for i, (w, b) in enumerate(zip(Ws, bs)):
        OUT[f'{prefix}_W{i}'] = np.asarray(w)
        OUT[f'{prefix}_b{i}'] = np.asarray(b)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def flat_grad(model, feed):
    """d loss / d (uv_weights + uv_biases), var_list order -- what ScipyOptimizerInterface packs (plate:241)."""
    var_list = model.uv_weights + model.uv_biases
    gs = model.sess.run(tf.gradients(model.loss, var_list), feed)
    return np.concatenate([np.asarray(g, np.float64).ravel() for g in gs])


def flat_params(model):
    return np.concatenate([np.asarray(a, np.float64).ravel() for a in model.sess.run(model.uv_weights + model.uv_biases)])


def bfgs_trace(model, maxiter, *args):
    """run the reference's train_bfgs with a short iteration budget; returns the losses its callback saw."""
    seen = []
    model.optimizer.optimizer_kwargs['options'] = dict(model.optimizer.optimizer_kwargs['options'], maxiter=maxiter, maxfun=4 * maxiter)
    model.callback = lambda loss: seen.append(float(loss))
    quiet(model.train_bfgs, *args)
    return np.array(seen)


# ----------------------------------------------------------------------------------------------------------- plate (F5)
def plate_sets(rng, n_c=256):
    lb, ub = np.array([0., 0, 0]), np.array([.5, .5, 10.])
    P = rng.uniform(lb, ub, (3 * n_c, 3))
    Collo = P[np.hypot(P[:, 0], P[:, 1]) > 0.1][:n_c]
    th = rng.uniform(0, np.pi / 2, 48)
    HOLE = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), rng.uniform(0, 10, 48)], 1)
    IC = rng.uniform(lb, ub, (40, 3)); IC[:, 2] = 0
    LF = rng.uniform(lb, ub, (40, 3)); LF[:, 0] = 0
    RT = rng.uniform(lb, ub, (40, 3)); RT[:, 0] = .5
    RT = np.concatenate([RT, 0.5 * np.sin(2 * np.pi * RT[:, 2:3] / 5 + 1.5 * np.pi) + 0.5], 1)     # plate:903
    UP = rng.uniform(lb, ub, (40, 3)); UP[:, 1] = .5
    LW = rng.uniform(lb, ub, (40, 3)); LW[:, 1] = 0
    DIST = np.concatenate([rng.uniform(lb, ub, (64, 3)), rng.uniform(0, 0.4, (64, 5))], 1)
    return dict(Collo=Collo, HOLE=HOLE, IC=IC, LF=LF, RT=RT, UP=UP, LW=LW, DIST=DIST, lb=lb, ub=ub)


def plate_feed(m):
    return {m.x_c_tf: m.x_c, m.y_c_tf: m.y_c, m.t_c_tf: m.t_c, m.x_IC_tf: m.x_IC, m.y_IC_tf: m.y_IC, m.t_IC_tf: m.t_IC,
            m.x_HOLE_tf: m.x_HOLE, m.y_HOLE_tf: m.y_HOLE, m.t_HOLE_tf: m.t_HOLE, m.x_LF_tf: m.x_LF, m.y_LF_tf: m.y_LF, m.t_LF_tf: m.t_LF,
            m.x_RT_tf: m.x_RT, m.y_RT_tf: m.y_RT, m.t_RT_tf: m.t_RT, m.s11_RT_tf: m.s11_RT, m.x_UP_tf: m.x_UP, m.y_UP_tf: m.y_UP, m.t_UP_tf: m.t_UP,
            m.x_LW_tf: m.x_LW, m.y_LW_tf: m.y_LW, m.t_LW_tf: m.t_LW, m.x_dist_tf: m.x_dist, m.y_dist_tf: m.y_dist, m.t_dist_tf: m.t_dist,
            m.u_dist_tf: m.u_dist, m.v_dist_tf: m.v_dist, m.s11_dist_tf: m.s11_dist, m.s22_dist_tf: m.s22_dist, m.s12_dist_tf: m.s12_dist}


def plate():
    mod = tf.load_reference_module(REF + '/PlateHoleQuarter/train/train.py', 'ref_plate')
    rng = np.random.default_rng(20261017)
    S = plate_sets(rng)
    for k, v in S.items():
        OUT['plate_' + k] = v
    uv_layers = [3] + 5 * [50] + [5]
    uvW, uvb = xavier(uv_layers, rng, bias_scale=0.1)
    store('plate_uv', uvW, uvb)
    uv_path = dump(uvW, uvb, 'uv.pickle')
    # (a) plain net: D == 1 and P == 0 through the reference's own composite (dist net: zero weights, unit last bias; part net: zeros)
    tiny = [3, 4, 5]
    dW = [np.zeros((3, 4)), np.zeros((4, 5))]; db = [np.zeros((1, 4)), np.ones((1, 5))]
    pW = [np.zeros((3, 4)), np.zeros((4, 5))]; pb = [np.zeros((1, 4)), np.zeros((1, 5))]
    args = [S[k] for k in ('Collo', 'HOLE', 'IC', 'LF', 'RT', 'UP', 'LW', 'DIST')]
    tf.reset_default_graph()
    m = quiet(mod.PINN, *args, uv_layers, tiny, tiny, S['lb'], S['ub'], partDir=dump(pW, pb, 'p0.pickle'), distDir=dump(dW, db, 'd1.pickle'), uvDir=uv_path)
    feed = plate_feed(m)
    OUT['plate_plain_terms'] = np.array(m.sess.run([m.loss_f_uv, m.loss_f_s, m.loss_HOLE, m.loss], feed), np.float64)
    OUT['plate_plain_grad'] = flat_grad(m, feed)
    xs = S['Collo'][:40]
    OUT['plate_plain_predict'] = np.concatenate(m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), 1)
    curves = quiet(m.train, 20, 5e-4)                                       # the reference's Adam loop (plate:475-506)
    OUT['plate_plain_adam'] = np.array(curves, np.float64).T                 # [20, 4]: loss_f_uv, loss_f_s, loss_HOLE, loss
    OUT['plate_plain_params_after_adam'] = flat_params(m)
    OUT['plate_plain_bfgs_losses'] = bfgs_trace(m, 6)                        # continues from the Adam state (plate:508-525)
    OUT['plate_plain_params_after_bfgs'] = flat_params(m)
    # (b) composite with non-trivial frozen dist / part nets
    small = [3, 10, 10, 5]
    dW, db = xavier(small, rng, bias_scale=0.3)
    pW, pb = xavier(small, rng, bias_scale=0.3)
    store('plate_dist', dW, db); store('plate_part', pW, pb)
    tf.reset_default_graph()
    m = quiet(mod.PINN, *args, uv_layers, small, small, S['lb'], S['ub'], partDir=dump(pW, pb, 'p.pickle'), distDir=dump(dW, db, 'd.pickle'), uvDir=uv_path)
    feed = plate_feed(m)
    OUT['plate_comp_terms'] = np.array(m.sess.run([m.loss_f_uv, m.loss_f_s, m.loss_HOLE, m.loss, m.loss_PART, m.loss_DIST], feed), np.float64)
    OUT['plate_comp_grad'] = flat_grad(m, feed)
    OUT['plate_comp_predict'] = np.concatenate(m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), 1)
    OUT['plate_comp_predict_D'] = np.concatenate(m.predict_D(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), 1)
    OUT['plate_comp_predict_P'] = np.concatenate(m.predict_P(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3]), 1)
    OUT['plate_comp_adam'] = np.array(quiet(m.train, 8, 5e-4), np.float64).T


# ----------------------------------------------------------------------------------------------------------- waves (F7)
def wave_sets(rng, lb, ub, n_c=256, fixed=False):
    Collo = rng.uniform(lb, ub, (n_c, 3))
    IC = rng.uniform(lb, ub, (48, 3)); IC[:, 2] = lb[2]
    UP = rng.uniform(lb, ub, (48, 3)); UP[:, 1] = ub[1]
    th = rng.uniform(0, 2 * np.pi, 64); ts = rng.uniform(lb[2], ub[2], 64)
    cx, cy = 0.5 * (lb[0] + ub[0]), 0.5 * (lb[1] + ub[1])
    amp = 0.1 * np.sin(ts)
    SRC = np.stack([cx + 2 * np.cos(th), cy + 2 * np.sin(th), ts, amp * np.cos(th), amp * np.sin(th)], 1)
    S = dict(Collo=Collo, SRC=SRC, IC=IC, UP=UP, lb=lb, ub=ub)
    if fixed:
        FIX = rng.uniform(lb, ub, (48, 3)); FIX[:24, 0] = lb[0]; FIX[24:, 1] = lb[1]
        S['FIXED'] = FIX
        S['DIST'] = np.concatenate([rng.uniform(lb, ub, (32, 3)), rng.uniform(0, 1, (32, 5))], 1)
    return S


def wave_feed(m, kind, a=None, b=None):
    sl = slice(a, b)
    d = {m.x_c_tf: m.x_c[sl], m.y_c_tf: m.y_c[sl], m.t_c_tf: m.t_c[sl], m.x_IC_tf: m.x_IC, m.y_IC_tf: m.y_IC, m.t_IC_tf: m.t_IC,
         m.x_SRC_tf: m.x_SRC, m.y_SRC_tf: m.y_SRC, m.t_SRC_tf: m.t_SRC, m.u_SRC_tf: m.u_SRC, m.v_SRC_tf: m.v_SRC}
    if kind in ('semi', 'inf'):
        d.update({m.x_UP_tf: m.x_UP, m.y_UP_tf: m.y_UP, m.t_UP_tf: m.t_UP})
    else:
        d.update({m.x_FIX_tf: m.x_FIX, m.y_FIX_tf: m.y_FIX, m.t_FIX_tf: m.t_FIX})
    return d


def wave(kind):
    path = {'semi': '/ElasticWaveSemiInfinite/ElasticWave.py', 'inf': '/ElasticWaveInfinite/ElasticWave.py', 'conf': '/ElasticWaveConfined/ElasticWave.py'}[kind]
    mod = tf.load_reference_module(REF + path, 'ref_' + kind)
    rng = np.random.default_rng({'semi': 11, 'inf': 12, 'conf': 13}[kind])
    lb, ub = {'semi': (np.array([-15., -15, 0]), np.array([15., 15, 16])), 'inf': (np.array([0., 0, 0]), np.array([30., 30, 20])),
              'conf': (np.array([-15., -15, 0]), np.array([15., 15, 14]))}[kind]
    S = wave_sets(rng, lb, ub, fixed=(kind == 'conf'))
    for k, v in S.items():
        OUT[f'{kind}_{k}'] = v
    layers = [3] + 5 * [50] + [7]
    # un-normalised inputs reach |x| = 15..30: scale the first layer so that tanh is not saturated (inf normalises itself, inf:191)
    W, b = xavier(layers, rng, bias_scale=0.1)
    if kind != 'inf':
        W[0] = W[0] * 0.1
    if kind == 'inf':      # the float32 script: load_NN takes the dtype of the pickled arrays (inf:181-182), as its own save_NN writes them
        W = [w.astype(np.float32) for w in W]; b = [x.astype(np.float32) for x in b]
    store(f'{kind}_uv', W, b)
    uv_path = dump(W, b, f'{kind}_uv.pickle')
    tf.reset_default_graph()
    if kind == 'conf':
        m = quiet(mod.DeepElasticWave, S['Collo'], S['SRC'], S['IC'], S['FIXED'], S['DIST'], layers, [3, 4, 5], [3, 4, 5], lb, ub, uvDir=uv_path)
    else:
        m = quiet(mod.DeepHPM, S['Collo'], S['SRC'], S['IC'], S['UP'], layers, lb, ub, ExistModel=1, modelDir=uv_path)
    feed = wave_feed(m, kind)
    names = {'semi': ['loss_f_uv', 'loss_f_s', 'loss_IC', 'loss_SRC', 'loss_NB', 'loss'], 'inf': ['loss_f_uv', 'loss_f_s', 'loss_IC', 'loss_SRC', 'loss'],
             'conf': ['loss_f_uv', 'loss_f_s', 'loss_SRC', 'loss_IC', 'loss_FIX', 'loss']}[kind]
    OUT[f'{kind}_term_names'] = np.array(names)
    OUT[f'{kind}_terms'] = np.array(m.sess.run([getattr(m, n) for n in names], feed), np.float64)
    OUT[f'{kind}_grad'] = flat_grad(m, feed)
    xs = S['Collo'][:40]
    OUT[f'{kind}_predict'] = np.concatenate([np.asarray(a, np.float64) for a in m.predict(xs[:, 0:1], xs[:, 1:2], xs[:, 2:3])], 1)
    # the reference's Adam loop with batch_num = 2 chunks (semi:289-319): iter steps on each chunk in sequence
    OUT[f'{kind}_adam_b2'] = np.array(quiet(m.train, 6, 1e-3, 2), np.float64).T
    OUT[f'{kind}_params_after_adam'] = flat_params(m)
    if kind == 'semi':
        OUT['semi_bfgs_losses'] = bfgs_trace(m, 5, 1)
        OUT['semi_params_after_bfgs'] = flat_params(m)



# ----------------------------------------------------------------------------------------------------------- host preprocessing helpers
def helpers():
    """the reference's own module-level point generators (pure numpy; plate:614-656,857-869, semi:397-409,633-657, conf:477-526,869-872)"""
    P = tf.load_reference_module(REF + '/PlateHoleQuarter/train/train.py', 'ref_plate_h')
    S = tf.load_reference_module(REF + '/ElasticWaveSemiInfinite/ElasticWave.py', 'ref_semi_h')
    C = tf.load_reference_module(REF + '/ElasticWaveConfined/ElasticWave.py', 'ref_conf_h')
    rng = np.random.default_rng(5)
    x, y, t = P.GenDistPt(xmin=0, xmax=0.5, ymin=0, ymax=0.5, tmin=0, tmax=10, xc=0, yc=0, r=0.1, num_surf_pt=11, num=9, num_t=5)   # plate:921
    XYT = np.concatenate((x, y, t), 1)
    OUT['h_plate_GenDistPt'] = XYT
    OUT['h_plate_GenDist'] = P.GenDist(XYT)
    pts = rng.uniform([0, 0, 0], [.5, .5, 10], (500, 3))
    OUT['h_pts_plate'] = pts
    OUT['h_plate_DelHolePT'] = P.DelHolePT(pts, xc=0, yc=0, r=0.1)
    OUT['h_plate_GenHoleSurfPT'] = np.concatenate(P.GenHoleSurfPT(xc=0, yc=0, r=0.1, N_PT=17), 1)
    OUT['h_semi_CartGrid'] = np.concatenate(S.CartGrid(xmin=-15, xmax=15, ymin=-15, ymax=15, tmin=0, tmax=16, num=6, num_t=4), 1)
    OUT['h_semi_GenCirclePT'] = np.concatenate(S.GenCirclePT(xc=0, yc=0, r=2.0, N_PT=23), 1)
    ptw = rng.uniform([-15, -15, 0], [15, 15, 16], (500, 3))
    ptw[:40, :2] = 2.0 * np.stack([np.cos(np.linspace(0, 6, 40)), np.sin(np.linspace(0, 6, 40))], 1)        # rows on / near the r = 2 circle
    OUT['h_pts_wave'] = ptw
    OUT['h_semi_DelSrcPT'] = S.DelSrcPT(ptw, xc=0, yc=0, r=2.0)          # keeps dst >= r (semi:657)
    OUT['h_conf_DelSrcPT'] = C.DelSrcPT(ptw, xc=0, yc=0, r=2.0)          # keeps dst >  r (conf:872)
    x, y, t = C.GenDistPt(xmin=-15, xmax=15, ymin=-15, ymax=15, tmin=0, tmax=14, xc=0, yc=0, r=2.0, num_surf_pt=13, num=8, num_t=4)
    XYT = np.concatenate((x, y, t), 1)
    OUT['h_conf_GenDistPt'] = XYT
    OUT['h_conf_GenDist'] = C.GenDist(XYT)


if __name__ == '__main__':
    helpers()
    plate()
    for k in ('semi', 'inf', 'conf'):
        wave(k)
    out = os.path.join(HERE, 'reference_tf1shim.npz')
    np.savez_compressed(out, **OUT)
    print('wrote', out, os.path.getsize(out), 'bytes,', len(OUT), 'arrays')
    for k in sorted(OUT):
        if OUT[k].size <= 8:
            print(' ', k, OUT[k])
