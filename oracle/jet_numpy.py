"""numpy restatement of the KERNEL ALGORITHM (forward-mode jets + hand-derived reverse sweep)
--  TEST INFRASTRUCTURE ONLY, never imported by the product package.

`oracle/ref_torch.py` restates the reference graph with nested autograd (the tf.gradients structure).
This file restates the algebra the CUDA kernels implement so that the
hand-derived adjoint formulas can be checked against autograd on the CPU, in float64 or float32,
before a GPU is involved.  Streams: 0 = value, 1 = d/dx, 2 = d/dy, 3 = d/dt, 4 = d2/dt2.

Reference equations: plate = PlateHoleQuarter/train/train.py:404-439 (F5), 452-461 (traction);
semi = ElasticWaveSemiInfinite/ElasticWave.py:228-272 (F7).
"""
from __future__ import annotations

import numpy as np


def forward_jets(X, Ws, bs, K, lb=None, ub=None, dtype=np.float64):
    """Returns (Y[K,N,O], stash) where stash holds per-layer inputs/activations for the reverse sweep."""
    X = np.asarray(X, dtype)
    N = X.shape[0]
    A = np.zeros((K, N, 3), dtype)
    scale = np.ones(3, dtype)
    if lb is not None:
        lb = np.asarray(lb, dtype); ub = np.asarray(ub, dtype)
        scale = (2.0 / (ub - lb)).astype(dtype)
        A[0] = (2.0 * (X - lb) / (ub - lb) - 1.0).astype(dtype)
    else:
        A[0] = X
    if K >= 4:
        A[1, :, 0] = scale[0]
        A[2, :, 1] = scale[1]
        A[3, :, 2] = scale[2]
    elif K == 2:                     # (value, d/dt) stream set used by the dt data terms
        A[1, :, 2] = scale[2]
    layers_in, acts = [], []
    for l, (W, b) in enumerate(zip(Ws, bs)):
        W = np.asarray(W, dtype); b = np.asarray(b, dtype).reshape(1, -1)
        layers_in.append(A)
        Z = A @ W                    # [K,N,dout]
        Z[0] += b
        if l == len(Ws) - 1:
            return Z, (layers_in, acts, scale)
        a = np.tanh(Z[0]); s = 1 - a * a
        An = np.empty_like(Z)
        An[0] = a
        for k in range(1, K):
            An[k] = s * Z[k]
        if K == 5:
            An[4] = s * Z[4] - 2 * a * s * Z[3] * Z[3]
        acts.append((a, s, Z))
        A = An
    raise AssertionError


def backward_jets(Ybar, Ws, stash, K):
    """Reverse sweep: Ybar[K,N,O] -> list of dW, list of db."""
    layers_in, acts, _ = stash
    L = len(Ws)
    dWs = [None] * L
    dbs = [None] * L
    Zbar = Ybar
    for l in range(L - 1, -1, -1):
        A = layers_in[l]
        W = np.asarray(Ws[l], Zbar.dtype)
        dWs[l] = np.einsum('kni,knj->ij', A, Zbar)
        dbs[l] = Zbar[0].sum(0, keepdims=True)
        if l == 0:
            break
        Abar = Zbar @ W.T            # adjoint of layer l-1 outputs
        a, s, Z = acts[l - 1]
        Zb = np.empty_like(Abar)
        for k in range(1, K):
            Zb[k] = s * Abar[k]
        acc = np.zeros_like(a)
        for k in range(1, K):
            acc += Z[k] * Abar[k]
        zv = Abar[0] - 2 * a * acc
        if K == 5:
            Zb[3] = s * Abar[3] - 4 * a * s * Z[3] * Abar[4]
            zv = zv - 2 * (1 - 3 * a * a) * Z[3] * Z[3] * Abar[4]
        Zb[0] = s * zv
        Zbar = Zb
    return dWs, dbs


def composite(Y, D, P):
    """u = P + D*N with jets (plate:382-387; product rule).  Y,D,P: [5,N,5]."""
    U = np.empty_like(Y)
    U[0] = P[0] + D[0] * Y[0]
    for k in (1, 2, 3):
        U[k] = P[k] + D[k] * Y[0] + D[0] * Y[k]
    U[4] = P[4] + D[4] * Y[0] + 2 * D[3] * Y[3] + D[0] * Y[4]
    return U


def composite_adjoint(Ubar, D):
    Yb = np.zeros_like(Ubar)
    Yb[0] = D[0] * Ubar[0] + D[1] * Ubar[1] + D[2] * Ubar[2] + D[3] * Ubar[3] + D[4] * Ubar[4]
    for k in (1, 2):
        Yb[k] = D[0] * Ubar[k]
    Yb[3] = D[0] * Ubar[3] + 2 * D[3] * Ubar[4]
    Yb[4] = D[0] * Ubar[4]
    return Yb


def residual_f5(U, E, mu, rho):
    """plate:416-437.  U[5,N,5] (u,v,s11,s22,s12).  Returns f[5,N]: f_u,f_v,f_s11,f_s22,f_s12."""
    c11 = E / (1 - mu * mu); c12 = E * mu / (1 - mu * mu); G = E / (2 * (1 + mu))
    e11 = U[1, :, 0]; e22 = U[2, :, 1]; e12 = U[2, :, 0] + U[1, :, 1]
    f_s11 = U[0, :, 2] - (c11 * e11 + c12 * e22)
    f_s22 = U[0, :, 3] - (c12 * e11 + c11 * e22)
    f_s12 = U[0, :, 4] - G * e12
    f_u = U[1, :, 2] + U[2, :, 4] - rho * U[4, :, 0]
    f_v = U[2, :, 3] + U[1, :, 4] - rho * U[4, :, 1]
    return np.stack([f_u, f_v, f_s11, f_s22, f_s12])


def residual_f5_adjoint(fbar, U, E, mu, rho):
    c11 = E / (1 - mu * mu); c12 = E * mu / (1 - mu * mu); G = E / (2 * (1 + mu))
    Ub = np.zeros_like(U)
    fu, fv, f11, f22, f12 = fbar
    Ub[0, :, 2] = f11; Ub[0, :, 3] = f22; Ub[0, :, 4] = f12
    Ub[1, :, 0] = -(c11 * f11 + c12 * f22)          # e11 = u_x
    Ub[2, :, 1] = -(c12 * f11 + c11 * f22)          # e22 = v_y
    Ub[2, :, 0] = -G * f12                          # u_y
    Ub[1, :, 1] = -G * f12                          # v_x
    Ub[1, :, 2] = fu; Ub[2, :, 4] += fu; Ub[4, :, 0] = -rho * fu
    Ub[2, :, 3] = fv; Ub[1, :, 4] += fv; Ub[4, :, 1] = -rho * fv
    return Ub


def residual_f7(Y, E, mu, rho):
    """semi:245-270.  Y[4,N,7] (u,v,ut,vt,s11,s22,s12). Returns f[7,N]: f_u,f_v,f_ut,f_vt,f_s11,f_s22,f_s12."""
    coef = E / ((1 + mu) * (1 - 2 * mu)); c11 = coef * (1 - mu); c12 = coef * mu; G = E / (2 * (1 + mu))
    e11 = Y[1, :, 0]; e22 = Y[2, :, 1]; e12 = Y[2, :, 0] + Y[1, :, 1]
    f_s11 = Y[0, :, 4] - (c11 * e11 + c12 * e22)
    f_s22 = Y[0, :, 5] - (c12 * e11 + c11 * e22)
    f_s12 = Y[0, :, 6] - G * e12
    f_ut = Y[3, :, 0] - Y[0, :, 2]
    f_vt = Y[3, :, 1] - Y[0, :, 3]
    f_u = Y[1, :, 4] + Y[2, :, 6] - rho * Y[3, :, 2]
    f_v = Y[2, :, 5] + Y[1, :, 6] - rho * Y[3, :, 3]
    return np.stack([f_u, f_v, f_ut, f_vt, f_s11, f_s22, f_s12])


def residual_f7_adjoint(fbar, Y, E, mu, rho):
    coef = E / ((1 + mu) * (1 - 2 * mu)); c11 = coef * (1 - mu); c12 = coef * mu; G = E / (2 * (1 + mu))
    Yb = np.zeros_like(Y)
    fu, fv, fut, fvt, f11, f22, f12 = fbar
    Yb[0, :, 4] = f11; Yb[0, :, 5] = f22; Yb[0, :, 6] = f12
    Yb[1, :, 0] = -(c11 * f11 + c12 * f22)
    Yb[2, :, 1] = -(c12 * f11 + c11 * f22)
    Yb[2, :, 0] = -G * f12
    Yb[1, :, 1] = -G * f12
    Yb[3, :, 0] = fut; Yb[0, :, 2] = -fut
    Yb[3, :, 1] = fvt; Yb[0, :, 3] = -fvt
    Yb[1, :, 4] = fu; Yb[2, :, 6] += fu; Yb[3, :, 2] = -rho * fu
    Yb[2, :, 5] = fv; Yb[1, :, 6] += fv; Yb[3, :, 3] = -rho * fv
    return Yb


def loss_grad_residual(kind, X, Ws, bs, w_uv, w_s, E, mu, rho, n_global=None, lb=None, ub=None,
                       dist=None, part=None, dtype=np.float64):
    """Collocation part of the loss: w_uv*loss_f_uv + w_s*loss_f_s and its flat gradient.

    Returns (loss_f_uv, loss_f_s, dWs, dbs)."""
    K = 5 if kind == 'f5' else 4
    N = X.shape[0]
    n_global = N if n_global is None else n_global
    Y, stash = forward_jets(X, Ws, bs, K, lb, ub, dtype)
    if kind == 'f5':
        if dist is not None:
            D, _ = forward_jets(X, dist[0], dist[1], 5, dtype=dtype)
            P, _ = forward_jets(X, part[0], part[1], 5, dtype=dtype)
            U = composite(Y, D, P)
        else:
            U = Y
        f = residual_f5(U, E, mu, rho)
        l_uv = (f[0] ** 2 + f[1] ** 2).sum() / n_global
        l_s = (f[2] ** 2 + f[3] ** 2 + f[4] ** 2).sum() / n_global
        fbar = 2 * f / n_global
        fbar[:2] *= w_uv; fbar[2:] *= w_s
        Ub = residual_f5_adjoint(fbar, U, E, mu, rho)
        Yb = composite_adjoint(Ub, D) if dist is not None else Ub
    else:
        f = residual_f7(Y, E, mu, rho)
        l_uv = (f[:4] ** 2).sum() / n_global
        l_s = (f[4:] ** 2).sum() / n_global
        fbar = 2 * f / n_global
        fbar[:4] *= w_uv; fbar[4:] *= w_s
        Yb = residual_f7_adjoint(fbar, Y, E, mu, rho)
    dWs, dbs = backward_jets(Yb.astype(dtype), Ws, stash, K)
    return float(l_uv), float(l_s), dWs, dbs
