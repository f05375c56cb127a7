"""Host-side pre/post-processing helpers of the reference drivers, vectorised.

Same names, arguments and return shapes as the reference's free functions; the O(N) Python loops
(`GenDist` plate:649-654, `DelHolePT` plate:859, `DelSrcPT` semi:653-657) are replaced by numpy expressions that
produce identical arrays.  Nothing here touches the GPU.

    plate = PlateHoleQuarter/train/train.py      semi = ElasticWaveSemiInfinite/ElasticWave.py
    inf   = ElasticWaveInfinite/ElasticWave.py   conf = ElasticWaveConfined/ElasticWave.py
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- sampling
def lhs(n, samples=None, rng=None):
    """Latin-hypercube sample in [0,1]^n, `samples` rows -- the call signature the reference uses from pyDOE
    (`lhs(3, 70000)`, plate:903).  pyDOE (third-party, not vendored) is absent here; this follows its published
    `_lhsclassic` algorithm: one uniform draw per stratum and dimension, then an independent random permutation per
    dimension.  With rng=None the global numpy stream is used in the same call order as pyDOE (`rand` then one
    `permutation` per column), so `np.random.seed(1111)` (plate:22) gives pyDOE's point sets."""
    samples = n if samples is None else samples
    R = np.random if rng is None else rng
    u = R.rand(samples, n) if rng is None else rng.random((samples, n))
    cut = np.linspace(0, 1, samples + 1)
    a, b = cut[:samples], cut[1:samples + 1]
    pts = u * (b - a)[:, None] + a[:, None]
    H = np.empty_like(pts)
    for j in range(n):
        order = R.permutation(samples)
        H[:, j] = pts[order, j]
    return H


def DelHolePT(XYT_c, xc=0, yc=0, r=0.1):
    """Drop points inside the hole, keeps dst > r (plate:857-860)."""
    XYT_c = np.asarray(XYT_c)
    dst = ((XYT_c[:, 0] - xc) ** 2 + (XYT_c[:, 1] - yc) ** 2) ** 0.5      # same expression as the reference: identical rounding on the hole edge
    return XYT_c[dst > r, :]


def DelSrcPT(XYT_c, xc, yc, r, strict=False):
    """Drop points inside the source circle: keeps dst >= r (semi:653-657, inf) or dst > r with strict=True (conf:869-872)."""
    XYT_c = np.asarray(XYT_c)
    dst = ((XYT_c[:, 0] - xc) ** 2 + (XYT_c[:, 1] - yc) ** 2) ** 0.5
    return XYT_c[(dst > r) if strict else (dst >= r), :]


def GenHoleSurfPT(xc, yc, r, N_PT):
    """Quarter-circle hole surface points (plate:862-869)."""
    theta = np.linspace(0.0, np.pi / 2.0, N_PT)
    return (r * np.cos(theta) + xc)[:, None], (r * np.sin(theta) + yc)[:, None]


def GenCirclePT(xc, yc, r, N_PT):
    """Full-circle source points (semi:633-650, conf:849-866)."""
    theta = np.linspace(0.0, 2 * np.pi, N_PT)
    return (r * np.cos(theta) + xc)[:, None], (r * np.sin(theta) + yc)[:, None]


def CartGrid(xmin, xmax, ymin, ymax, tmin, tmax, num, num_t):
    """x-y-t Cartesian grid, flattened columns (semi:397-408)."""
    x = np.linspace(xmin, xmax, num=num)
    y = np.linspace(ymin, ymax, num=num)
    t = np.linspace(tmin, tmax, num=num_t)
    xxx, yyy, ttt = np.meshgrid(x, y, t)
    return xxx.flatten()[:, None], yyy.flatten()[:, None], ttt.flatten()[:, None]


def GenDistPt(xmin, xmax, ymin, ymax, tmin, tmax, xc, yc, r, num_surf_pt, num, num_t, arc=np.pi / 2.0):
    """Grid minus the hole + surface refinement points, times a time grid (plate:614-641; conf:477-508 uses arc=2*pi)."""
    x = np.linspace(xmin, xmax, num=num)
    y = np.linspace(ymin, ymax, num=num)
    x, y = np.meshgrid(x, y)
    keep = ((x - xc) ** 2 + (y - yc) ** 2) ** 0.5 >= r
    x, y = x[keep].flatten(), y[keep].flatten()
    theta = np.linspace(0.0, arc, num_surf_pt)
    x = np.concatenate((x, r * np.cos(theta) + xc))
    y = np.concatenate((y, r * np.sin(theta) + yc))
    t = np.linspace(tmin, tmax, num=num_t)
    xxx, ttt = np.meshgrid(x, t)
    yyy, _ = np.meshgrid(y, t)
    return xxx.flatten()[:, None], yyy.flatten()[:, None], ttt.flatten()[:, None]


def GenDist(XYT_dist):
    """Analytic distance-function targets of the plate (plate:643-656): [x, y, t, D_u, D_v, D_s11, D_s22, D_s12]."""
    X = np.asarray(XYT_dist, dtype=float)
    x, y, t = X[:, 0:1], X[:, 1:2], X[:, 2:3]
    d_u = np.minimum(t, x)
    d_v = np.minimum(t, y)
    d_s11 = np.minimum(t, 0.5 - x)
    d_s22 = np.minimum(t, 0.5 - y)
    d_s12 = np.minimum.reduce([t, y, 0.5 - y, x, 0.5 - x])
    return np.concatenate((X, d_u, d_v, d_s11, d_s22, d_s12), 1)


def GenDist_confined(XYT_dist):
    """Distance targets of the confined-wave script (conf:510-526; built there but unused by its net_uv)."""
    X = np.asarray(XYT_dist, dtype=float)
    x, y, t = X[:, 0:1], X[:, 1:2], X[:, 2:3]
    d = np.minimum.reduce([t, (x ** 2 + y ** 2) ** 0.5 - 2.0, 15 - x, x + 15, 15 - y, y + 15]) / 10.0
    one = np.ones_like(d)
    return np.concatenate((X, d, d, one, one, one), 1)


def shuffle(*arrays, rng=None):
    """In-place row shuffles, one independent permutation per array (semi:660-664)."""
    for a in arrays:
        (np.random if rng is None else rng).shuffle(a)


# ----------------------------------------------------------------------------- loads / sources
def plate_traction(t, period=5.0):
    """s11 on the loaded edge: 0.5 sin(2 pi t / period + 3 pi / 2) + 0.5 (plate:923-926)."""
    return 0.5 * np.sin((2 * np.pi / period) * t + 3 * np.pi / 2) + 0.5


def ricker(t, ts=3.0, tsh=3.0, Amp=1.0):
    """Ricker source amplitude (semi:726; inf:697)."""
    a = np.pi ** 2 * (t - ts) ** 2 / tsh ** 2
    return Amp * (2 * a - 1) * np.exp(-a)


def gauss_pulse(t, t0=2.0, width=0.5, Amp=0.5):
    """Gauss source amplitude 0.5 exp(-((t-2)/0.5)^2) (conf:960)."""
    return Amp * np.exp(-((t - t0) / width) ** 2)


def source_ring(xc, yc, r, N_PT, tt, amplitude):
    """SRC array [x, y, t, u, v]: radial displacement amplitude(t) on the source circle (semi:716-729)."""
    xx, yy = GenCirclePT(xc, yc, r, N_PT)
    x_S, t_S = np.meshgrid(xx, tt)
    y_S, _ = np.meshgrid(yy, tt)
    x_S, y_S, t_S = x_S.flatten()[:, None], y_S.flatten()[:, None], t_S.flatten()[:, None]
    amp = amplitude(t_S)
    return np.concatenate((x_S, y_S, t_S, amp * (x_S - xc) / r, amp * (y_S - yc) / r), 1)


# ----------------------------------------------------------------------------- default problem definitions
def plate_point_sets(rng=None, scale=1.0):
    """The point sets of the plate driver (plate:892-929); `scale` shrinks every LHS count (1.0 = the reference's sizes)."""
    n = lambda k: max(int(k * scale), 8)
    lb, ub = np.array([0, 0, 0.0]), np.array([0.5, 0.5, 10.0])
    x_d, y_d, t_d = GenDistPt(0, 0.5, 0, 0.5, 0, 10, 0, 0, 0.1, num_surf_pt=40, num=21, num_t=21)
    DIST = GenDist(np.concatenate((x_d, y_d, t_d), 1))
    IC = DelHolePT(lb + np.array([0.5, 0.5, 0.0]) * lhs(3, n(5000), rng))
    XYT_c = np.concatenate((lb + (ub - lb) * lhs(3, n(70000), rng), lb + np.array([0.15, 0.15, 10.0]) * lhs(3, n(40000), rng)), 0)
    XYT_c = DelHolePT(XYT_c)
    xx, yy = GenHoleSurfPT(0, 0, 0.1, 83)
    tt = np.linspace(0, 10, 121)[1:]
    x_ho, t_ho = np.meshgrid(xx, tt)
    y_ho, _ = np.meshgrid(yy, tt)
    HOLE = np.concatenate((x_ho.flatten()[:, None], y_ho.flatten()[:, None], t_ho.flatten()[:, None]), 1)
    LW = np.array([0.1, 0.0, 0.0]) + np.array([0.4, 0.0, 10]) * lhs(3, n(8000), rng)
    UP = np.array([0.0, 0.5, 0.0]) + np.array([0.5, 0.0, 10]) * lhs(3, n(8000), rng)
    LF = np.array([0.0, 0.1, 0.0]) + np.array([0.0, 0.4, 10]) * lhs(3, n(8000), rng)
    RT = np.array([0.5, 0.0, 0.0]) + np.array([0.0, 0.5, 10]) * lhs(3, n(13000), rng)
    RT = np.concatenate((RT, plate_traction(RT[:, 2:3])), 1)
    XYT_c = np.concatenate((XYT_c, HOLE[::4, :], LF[::5, :], RT[::5, 0:3], UP[::5, :], LW[::5, :]), 0)
    return dict(Collo=XYT_c, HOLE=HOLE, IC=IC, LF=LF, RT=RT, UP=UP, LW=LW, DIST=DIST, lb=lb, ub=ub)


def semi_point_sets(MAX_T=16.0, rng=None, scale=1.0, shuffled=True):
    """The point sets of the half-space wave driver (semi:675-729), row-shuffled like the driver does before training (semi:765)."""
    n = lambda k: max(int(k * scale), 8)
    lb, ub = np.array([-15, -15, 0.0]), np.array([15, 15, MAX_T])
    xc, yc, r = 0.0, 0.0, 2.0
    xy_IC = np.array([-15, -15]) + np.array([30, 30]) * lhs(2, n(12000), rng)
    IC = np.concatenate((xy_IC, 0 * xy_IC[:, 0:1]), 1)
    xt_up = np.array([-15, 0]) + np.array([30, MAX_T]) * lhs(2, n(15000), rng)
    UP = np.concatenate((xt_up[:, 0:1], np.full_like(xt_up[:, 0:1], 15.0), xt_up[:, 1:2]), 1)
    XYT_c = np.concatenate((lb + (ub - lb) * lhs(3, n(120000), rng),
                            np.array([xc - r - 2, yc - r - 2, 0.0]) + np.array([2 * (r + 2), 2 * (r + 2), MAX_T]) * lhs(3, n(15000), rng),
                            np.array([-15, 15 - 6, 0.0]) + np.array([30, 6, MAX_T]) * lhs(3, n(20000), rng)), 0)
    XYT_c = DelSrcPT(XYT_c, xc, yc, r)
    tt = np.concatenate((np.linspace(0, 6, 153), np.linspace(6, MAX_T, 63)))[1:]
    SRC = source_ring(xc, yc, r, 150, tt, ricker)
    if shuffled:
        shuffle(XYT_c, SRC, IC, UP, rng=rng)
    return dict(Collo=XYT_c, SRC=SRC, IC=IC, UP=UP, lb=lb, ub=ub)


def inf_point_sets(MAX_T=20.0, rng=None, scale=1.0, shuffled=True):
    """The point sets of the infinite-domain wave driver (inf:641-705): IC on a 101 x 101 grid, UP 150 x 201 (built, unused by the loss),
    120 k + 10 k LHS collocation points minus the source disc, Ricker source on 200 circle points x 352 times; row-shuffled like the
    driver does before training (inf:732)."""
    n = lambda k: max(int(k * scale), 8)
    lb, ub = np.array([0.0, 0.0, 0.0]), np.array([30, 30, MAX_T])
    xc, yc, r = 15.0, 15.0, 2.0
    IC = np.concatenate(CartGrid(xmin=0, xmax=30, ymin=0, ymax=30, tmin=0, tmax=0, num=101, num_t=1), 1)
    x_up, t_up = np.meshgrid(np.linspace(0, 30, 150), np.linspace(0, MAX_T, 201))
    x_up, t_up = x_up.flatten()[:, None], t_up.flatten()[:, None]
    UP = np.concatenate((x_up, np.full((x_up.size, 1), 30.0), t_up), 1)
    XYT_c = lb + (ub - lb) * lhs(3, n(120000), rng)
    XYT_c_ext = np.array([xc - r - 1, yc - r - 1, 0.0]) + np.array([2 * (r + 1), 2 * (r + 1), MAX_T]) * lhs(3, n(10000), rng)
    XYT_c = DelSrcPT(np.concatenate((XYT_c, XYT_c_ext), axis=0), xc, yc, r)
    tt = np.linspace(0, MAX_T, 353)[1:]
    SRC = source_ring(xc, yc, r, 200, tt, ricker)
    if shuffled:
        shuffle(XYT_c, SRC, IC, UP, rng=rng)
    return dict(Collo=XYT_c, SRC=SRC, IC=IC, UP=UP, lb=lb, ub=ub)


def conf_point_sets(MAX_T=14.0, rng=None, scale=1.0):
    """The point sets of the confined-wave driver (conf:883-968): DIST targets (built, unused by its net_uv), 6 k LHS IC points minus the
    source disc, four fixed edges of 7 k points each, 120 k + 15 k + an edge band of a 50 k LHS set as collocation points, Gauss-pulse
    source on 200 circle points x 281 times."""
    n = lambda k: max(int(k * scale), 8)
    lb, ub = np.array([-15.0, -15.0, 0.0]), np.array([15.0, 15.0, MAX_T])
    xc, yc, r = 0.0, 0.0, 2.0
    x_d, y_d, t_d = GenDistPt(xmin=-15.0, xmax=15.0, ymin=-15.0, ymax=15.0, tmin=0, tmax=MAX_T, xc=0.0, yc=0.0, r=2.0, num_surf_pt=120, num=21, num_t=21,
                              arc=2 * np.pi)
    DIST = GenDist_confined(np.concatenate((x_d, y_d, t_d), 1))
    IC = DelSrcPT(lb + np.array([30.0, 30.0, 0.0]) * lhs(3, n(6000), rng), 0, 0, 2.0, strict=True)
    LW = np.array([-15.0, -15.0, 0.0]) + np.array([30.0, 0.0, MAX_T]) * lhs(3, n(7000), rng)
    UP = np.array([-15.0, 15.0, 0.0]) + np.array([30.0, 0.0, MAX_T]) * lhs(3, n(7000), rng)
    LF = np.array([-15.0, -15.0, 0.0]) + np.array([0.0, 30.0, MAX_T]) * lhs(3, n(7000), rng)
    RT = np.array([15.0, -15.0, 0.0]) + np.array([0.0, 30.0, MAX_T]) * lhs(3, n(7000), rng)
    FIXED = np.concatenate((LF, RT, LW, UP), 0)
    XYT_c = lb + (ub - lb) * lhs(3, n(120000), rng)
    XYT_c_ext = np.array([xc - r - 1, yc - r - 1, 0.0]) + np.array([2 * (r + 1), 2 * (r + 1), MAX_T]) * lhs(3, n(15000), rng)
    XYT_c_ext2 = lb + (ub - lb) * lhs(3, n(50000), rng)
    XYT_c_ext2 = XYT_c_ext2[(np.abs(XYT_c_ext2[:, 0]) > 12) | (np.abs(XYT_c_ext2[:, 1]) > 12), :]
    XYT_c = DelSrcPT(np.concatenate((XYT_c, XYT_c_ext, XYT_c_ext2), axis=0), xc, yc, r, strict=True)
    tt = np.concatenate((np.linspace(0, 4, 141), np.linspace(4, MAX_T, 141)), 0)[1:]
    SRC = source_ring(xc, yc, r, 200, tt, gauss_pulse)
    return dict(Collo=XYT_c, SRC=SRC, IC=IC, FIXED=FIXED, DIST=DIST, lb=lb, ub=ub)


def time_march(make_model, point_sets, horizons, train, checkpoint=None):
    """The drivers' curriculum ("Need pretraining!! (i.e. train for 10s -> 15s -> 25s)", inf:636-638; "train 7s -> 14s", conf:884;
    semi:670-672): train on [0, T_1], save, rebuild the point sets for the longer horizon T_2, warm-start from the saved weights, ...
    The reference does this by editing MAX_T and the pickle name by hand between runs; here it is one loop.

    make_model(sets, warm_start_path_or_None) -> model;  point_sets(MAX_T) -> dict (e.g. semi_point_sets);
    train(model, MAX_T) runs the stage's optimiser calls;  checkpoint(MAX_T) -> path of the pickle written after the stage
    (default: a temporary file per stage).  Returns the last model and the list of (MAX_T, path)."""
    import os
    import tempfile
    saved, model, path = [], None, None
    tmp = tempfile.mkdtemp(prefix='pe_march_') if checkpoint is None else None
    for T in horizons:
        model = make_model(point_sets(T), path)
        train(model, T)
        path = checkpoint(T) if checkpoint is not None else os.path.join(tmp, 'uv_NN_%gs.pickle' % T)
        try:
            model.save_NN(path)
        except TypeError:                      # PINN / DeepElasticWave: save_NN(fileDir, TYPE)
            model.save_NN(path, 'UV')
        saved.append((T, path))
    return model, saved


# ----------------------------------------------------------------------------- FEM ground truth + metrics
def preprocess(dir):
    """FEM frame loader (plate:658-676; semi:411-435 returns the same columns plus amp / Mises): flattened (N,1) columns
    x, y, u, v, s11, s22, s12."""
    import scipy.io
    data = scipy.io.loadmat(dir)
    return tuple(np.asarray(data[k]).flatten()[:, None] for k in ('x', 'y', 'u', 'v', 's11', 's22', 's12'))


def rel_l2(pred, ref):
    """Relative L2 error of one field -- the number behind the reference's side-by-side PINN-vs-FEM panels (plate:678-855)."""
    pred, ref = np.asarray(pred, float).ravel(), np.asarray(ref, float).ravel()
    return float(np.linalg.norm(pred - ref) / np.linalg.norm(ref))


def fem_metrics(pred_fields, fem_fields, names=('u', 'v', 's11', 's22', 's12')):
    """dict name -> rel-L2 for the first len(names) fields of predict()'s tuple against the FEM columns."""
    return {n: rel_l2(p, f) for n, p, f in zip(names, pred_fields, fem_fields)}
