"""CPU tests of the vectorised host preprocessing against loop restatements of the reference helpers."""
import numpy as np

from oracle import preprocess_loops as O
from pinn_elastodynamics_b200 import preprocess as P


def test_lhs_is_a_latin_hypercube_and_follows_the_global_stream():
    np.random.seed(1111)
    H = P.lhs(3, 1000)
    assert H.shape == (1000, 3) and H.min() >= 0 and H.max() <= 1
    for j in range(3):      # exactly one point per stratum and dimension
        assert np.array_equal(np.sort(np.floor(H[:, j] * 1000).astype(int)), np.arange(1000))
    np.random.seed(1111)
    assert np.array_equal(P.lhs(3, 1000), H)                      # reproducible from np.random.seed like pyDOE
    # published pyDOE _lhsclassic call order: one rand(samples, n), then one permutation per column
    np.random.seed(7)
    u = np.random.rand(50, 2); cut = np.linspace(0, 1, 51)
    pts = u * (cut[1:] - cut[:-1])[:, None] + cut[:-1, None]
    exp = np.stack([pts[np.random.permutation(50), j] for j in range(2)], 1)
    np.random.seed(7)
    assert np.array_equal(P.lhs(2, 50), exp)
    G = P.lhs(3, 64, rng=np.random.default_rng(0))
    assert np.array_equal(np.sort(np.floor(G[:, 1] * 64).astype(int)), np.arange(64))


def test_vectorised_deletions_and_distance_targets_match_the_loops():
    rng = np.random.default_rng(3)
    X = rng.uniform([0, 0, 0], [.5, .5, 10], (2000, 3))
    X[:5, :2] = [[0.1, 0.0], [0.0, 0.1], [0.06, 0.08], [0.0, 0.0], [0.5, 0.5]]       # points on / inside the hole edge
    np.testing.assert_array_equal(P.DelHolePT(X), O.DelHolePT(X))
    np.testing.assert_array_equal(P.GenDist(X), O.GenDist(X))
    Y = rng.uniform([-15, -15, 0], [15, 15, 14], (2000, 3))
    Y[:3, :2] = [[2.0, 0.0], [0.0, -2.0], [1.2, 1.6]]
    np.testing.assert_array_equal(P.DelSrcPT(Y, 0, 0, 2.0), O.DelSrcPT(Y, 0, 0, 2.0))
    np.testing.assert_array_equal(P.DelSrcPT(Y, 0, 0, 2.0, strict=True), O.DelSrcPT(Y, 0, 0, 2.0, strict=True))
    assert len(O.DelSrcPT(Y, 0, 0, 2.0)) > len(O.DelSrcPT(Y, 0, 0, 2.0, strict=True))   # the two scripts differ on the circle itself
    np.testing.assert_allclose(P.GenDist_confined(Y), O.GenDist_confined(Y), rtol=1e-15, atol=0)


def test_grids_sources_and_default_point_sets():
    x, y, t = P.GenDistPt(0, 0.5, 0, 0.5, 0, 10, 0, 0, 0.1, num_surf_pt=40, num=21, num_t=21)
    nxy = (21 * 21 - int((np.hypot(*np.meshgrid(np.linspace(0, .5, 21), np.linspace(0, .5, 21))) < 0.1).sum())) + 40
    assert x.shape == (nxy * 21, 1) and t.min() == 0 and t.max() == 10
    xx, yy = P.GenHoleSurfPT(0, 0, 0.1, 83)
    np.testing.assert_allclose(np.hypot(xx, yy), 0.1, rtol=1e-14)
    assert xx[0, 0] == 0.1 and abs(xx[-1, 0]) < 1e-16
    cx, cy = P.GenCirclePT(15, 15, 2, 200)
    np.testing.assert_allclose(np.hypot(cx - 15, cy - 15), 2, rtol=1e-14)
    g = P.CartGrid(-15, 15, -15, 15, 0, 16, 5, 3)
    assert g[0].shape == (75, 1) and set(np.unique(g[2])) == {0.0, 8.0, 16.0}
    assert P.plate_traction(np.array([0.0, 2.5, 5.0])).round(12).tolist() == [0.0, 1.0, 0.0]      # plate:923-926
    assert abs(P.ricker(np.array([3.0]))[0] + 1.0) < 1e-15                                        # -Amp at t = ts (semi:726)
    sets = P.plate_point_sets(rng=np.random.default_rng(1), scale=0.02)
    assert sets['HOLE'].shape == (83 * 120, 3) and sets['RT'].shape[1] == 4 and sets['DIST'].shape[1] == 8
    assert (np.hypot(sets['IC'][:, 0], sets['IC'][:, 1]) > 0.1).all() and (sets['IC'][:, 2] == 0).all()
    w = P.semi_point_sets(rng=np.random.default_rng(2), scale=0.02)
    assert w['SRC'].shape == (150 * 215, 5) and (np.hypot(w['Collo'][:, 0], w['Collo'][:, 1]) >= 2.0).all()
    assert (w['UP'][:, 1] == 15.0).all() and (w['IC'][:, 2] == 0).all()
    a = np.arange(12.).reshape(6, 2); b = a.copy()
    P.shuffle(a, rng=np.random.default_rng(0))
    assert sorted(map(tuple, a)) == sorted(map(tuple, b)) and not np.array_equal(a, b)


def test_fem_metrics():
    ref = [np.array([[1.0], [2.0], [2.0]]), np.array([[0.0], [3.0], [4.0]])]
    pred = [r * 1.01 for r in ref]
    m = P.fem_metrics(pred, ref, names=('u', 'v'))
    assert abs(m['u'] - 0.01) < 1e-12 and abs(m['v'] - 0.01) < 1e-12


def test_helpers_match_the_reference_functions(golden):
    """arrays returned by the reference's own module-level helpers (called in place from /root/reference by
    tests/golden/make_reference_golden.py: plate:614-656,857-869; semi:397-409,633-657; conf:477-526,869-872)"""
    G = golden('reference_tf1shim.npz')
    cat = lambda cols: np.concatenate(cols, 1)
    XYT = cat(P.GenDistPt(xmin=0, xmax=0.5, ymin=0, ymax=0.5, tmin=0, tmax=10, xc=0, yc=0, r=0.1, num_surf_pt=11, num=9, num_t=5))
    np.testing.assert_array_equal(XYT, G['h_plate_GenDistPt'])
    np.testing.assert_array_equal(P.GenDist(XYT), G['h_plate_GenDist'])
    np.testing.assert_array_equal(P.DelHolePT(G['h_pts_plate'], xc=0, yc=0, r=0.1), G['h_plate_DelHolePT'])
    np.testing.assert_array_equal(cat(P.GenHoleSurfPT(xc=0, yc=0, r=0.1, N_PT=17)), G['h_plate_GenHoleSurfPT'])
    np.testing.assert_array_equal(cat(P.CartGrid(xmin=-15, xmax=15, ymin=-15, ymax=15, tmin=0, tmax=16, num=6, num_t=4)), G['h_semi_CartGrid'])
    np.testing.assert_array_equal(cat(P.GenCirclePT(xc=0, yc=0, r=2.0, N_PT=23)), G['h_semi_GenCirclePT'])
    np.testing.assert_array_equal(P.DelSrcPT(G['h_pts_wave'], 0, 0, 2.0), G['h_semi_DelSrcPT'])
    np.testing.assert_array_equal(P.DelSrcPT(G['h_pts_wave'], 0, 0, 2.0, strict=True), G['h_conf_DelSrcPT'])
    XYT = cat(P.GenDistPt(xmin=-15, xmax=15, ymin=-15, ymax=15, tmin=0, tmax=14, xc=0, yc=0, r=2.0, num_surf_pt=13, num=8, num_t=4, arc=2 * np.pi))
    np.testing.assert_array_equal(XYT, G['h_conf_GenDistPt'])
    np.testing.assert_allclose(P.GenDist_confined(XYT), G['h_conf_GenDist'], rtol=1e-15, atol=0)


def test_default_point_sets_equal_the_reference_drivers(golden):
    """plate / semi / inf / conf_point_sets() against fingerprints of the arrays the reference's own `__main__` bodies build (executed in
    place from /root/reference by tests/golden/make_driver_point_sets.py; plate:871-929, semi:667-765, inf:634-732, conf:881-968), with
    the numpy global stream seeded like the scripts (np.random.seed(1111), plate:22).  Equal means byte-equal float64 arrays."""
    import hashlib
    G = golden('driver_point_sets.npz')
    for kind, build in (('plate', P.plate_point_sets), ('semi', P.semi_point_sets), ('inf', P.inf_point_sets), ('conf', P.conf_point_sets)):
        np.random.seed(1111)
        sets = build()
        names = [k[len(kind) + 1:-6] for k in G.files if k.startswith(kind + '_') and k.endswith('_shape')]
        assert names
        for name in names:
            a = np.ascontiguousarray(np.asarray(sets[name], np.float64))
            assert tuple(G[f'{kind}_{name}_shape']) == a.shape, (kind, name, a.shape)
            np.testing.assert_allclose([a.sum(), (a * a).sum()], G[f'{kind}_{name}_sums'], rtol=1e-12, err_msg=f'{kind} {name}')
            assert np.array_equal(np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8), G[f'{kind}_{name}_sha256']), (kind, name)


def test_time_march_warm_starts_each_stage(tmp_path):
    """the drivers' hand-edited curriculum (semi:670-672, inf:636-638, conf:884) as one loop: every stage is built on the longer horizon's
    point sets and warm-started from the previous stage's pickle"""
    calls = []

    class Model:
        def __init__(self, sets, warm):
            self.sets, self.warm, self.w = sets, warm, (0 if warm is None else open(warm).read())

        def save_NN(self, path):
            open(path, 'w').write(str(int(self.w) + 1))

    def train(m, T):
        calls.append((T, m.sets['T'], m.warm))

    model, saved = P.time_march(lambda s, w: Model(s, w), lambda T: dict(T=T), [7.0, 14.0, 20.0], train, checkpoint=lambda T: str(tmp_path / f'uv_{T:g}.pickle'))
    assert [c[0] for c in calls] == [7.0, 14.0, 20.0] and [c[1] for c in calls] == [7.0, 14.0, 20.0]
    assert calls[0][2] is None and calls[1][2] == saved[0][1] and calls[2][2] == saved[1][1]
    assert open(saved[-1][1]).read() == '3' and model.sets['T'] == 20.0
