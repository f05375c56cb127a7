"""Loop-for-loop restatements of the reference's host preprocessing helpers -- TEST INFRASTRUCTURE ONLY (checker for
pinn_elastodynamics_b200/preprocess.py).  plate = PlateHoleQuarter/train/train.py, semi = ElasticWaveSemiInfinite/ElasticWave.py,
conf = ElasticWaveConfined/ElasticWave.py."""
import numpy as np


def GenDist(XYT_dist):                      # plate:643-656
    z = lambda: np.zeros_like(XYT_dist[:, 0:1])
    dist_u, dist_v, dist_s11, dist_s22, dist_s12 = z(), z(), z(), z(), z()
    for i in range(len(XYT_dist)):
        dist_u[i, 0] = min(XYT_dist[i][2], XYT_dist[i][0])
        dist_v[i, 0] = min(XYT_dist[i][2], XYT_dist[i][1])
        dist_s11[i, 0] = min(XYT_dist[i][2], 0.5 - XYT_dist[i][0])
        dist_s22[i, 0] = min(XYT_dist[i][2], 0.5 - XYT_dist[i][1])
        dist_s12[i, 0] = min(XYT_dist[i][2], XYT_dist[i][1], 0.5 - XYT_dist[i][1], XYT_dist[i][0], 0.5 - XYT_dist[i][0])
    return np.concatenate((XYT_dist, dist_u, dist_v, dist_s11, dist_s22, dist_s12), 1)


def GenDist_confined(XYT_dist):             # conf:510-526
    d = np.zeros_like(XYT_dist[:, 0:1])
    for i in range(len(XYT_dist)):
        d[i, 0] = min(XYT_dist[i][2], ((XYT_dist[i][0]) ** 2 + (XYT_dist[i][1]) ** 2) ** 0.5 - 2.0,
                      15 - XYT_dist[i][0], XYT_dist[i][0] + 15, 15 - XYT_dist[i][1], XYT_dist[i][1] + 15) / 10.0
    one = np.ones_like(d)
    return np.concatenate((XYT_dist, d, d, one, one, one), 1)


def DelHolePT(XYT_c, xc=0, yc=0, r=0.1):    # plate:857-860
    dst = np.array([((xyt[0] - xc) ** 2 + (xyt[1] - yc) ** 2) ** 0.5 for xyt in XYT_c])
    return XYT_c[dst > r, :]


def DelSrcPT(XYT_c, xc, yc, r, strict=False):   # semi:653-657 (>=), conf:869-872 (>)
    dst = np.array([((xyt[0] - xc) ** 2 + (xyt[1] - yc) ** 2) ** 0.5 for xyt in XYT_c])
    return XYT_c[dst > r, :] if strict else XYT_c[dst >= r, :]
