"""CPU float64 ORACLE for the PINN-elastodynamics hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, line by line, the TensorFlow-1 graph of the reference scripts with
torch.autograd (nested `torch.autograd.grad(..., create_graph=True)` standing in for
`tf.gradients`).  It is the checker for the CUDA path; nothing under
`pinn_elastodynamics_b200/` may import it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it.

PARITY STATUS: pinned to the reference's own source.  TensorFlow 1.x cannot be imported here and the reference ships no
tests or golden loss/gradient vectors, so the four reference class files are executed unmodified on a
TF1-primitive shim (oracle/tf1_shim.py) by tests/golden/make_reference_golden.py, and this oracle must reproduce their outputs
(tests/golden/reference_tf1shim.npz; tests/test_reference_golden.py: terms / gradients 1e-10, Adam loops 1e-9, L-BFGS-B sequences
1e-7).  Not covered: TensorFlow's own kernels -- the primitives' semantics come from TF's documentation.  Further pins:
(i) finite differences of its own loss, (ii) the shipped checkpoints: residual losses of the trained plate nets
(loss_f_uv = 3.8686e-05, loss_f_s = 2.4435e-05 on the survey's 4,862-point set) and FEM rel-L2 bands, see
tests/test_oracle.py and tests/golden/make_golden.py.

Reference files (relative to /root/reference):
  plate = PlateHoleQuarter/train/train.py
  inf   = ElasticWaveInfinite/ElasticWave.py
  semi  = ElasticWaveSemiInfinite/ElasticWave.py
  conf  = ElasticWaveConfined/ElasticWave.py
"""
from __future__ import annotations

import numpy as np
import torch

DT = torch.float64


def _t(a, dtype=DT):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def _grad(y, x):
    """tf.gradients(y, x)[0]: d(sum y)/dx; rows are independent so this is the per-row partial."""
    g, = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True, allow_unused=True)
    if g is None:
        g = torch.zeros_like(x)
    return g


# ----------------------------------------------------------------------------- parameters
def xavier_params(layers, seed=1111, dtype=np.float64):
    """Xavier N(0, 2/(in+out)) truncated at 2 sigma, zero biases (plate:258-274).

    The TF1 truncated-normal stream (graph seed 1111, plate:23) cannot be reproduced without TF, so
    parity runs inject these explicit arrays into both the oracle and the CUDA path."""
    rng = np.random.default_rng(seed)
    Ws, bs = [], []
    for l in range(len(layers) - 1):
        fi, fo = layers[l], layers[l + 1]
        std = np.sqrt(2.0 / (fi + fo))
        w = rng.standard_normal((fi, fo))
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        Ws.append((w * std).astype(dtype))
        bs.append(np.zeros((1, fo), dtype=dtype))
    return Ws, bs


def flatten_params(Ws, bs):
    """ScipyOptimizerInterface packing order: var_list = weights + biases (plate:241)."""
    return np.concatenate([np.asarray(w).ravel() for w in Ws] + [np.asarray(b).ravel() for b in bs])


def unflatten_params(flat, layers):
    Ws, bs, o = [], [], 0
    for l in range(len(layers) - 1):
        n = layers[l] * layers[l + 1]
        Ws.append(np.asarray(flat[o:o + n]).reshape(layers[l], layers[l + 1]))
        o += n
    for l in range(len(layers) - 1):
        n = layers[l + 1]
        bs.append(np.asarray(flat[o:o + n]).reshape(1, n))
        o += n
    return Ws, bs


# ----------------------------------------------------------------------------- network
def neural_net(X, weights, biases, lb=None, ub=None):
    """plate:308-320 (semi:195-206, conf:232-243); with lb/ub given: inf:188-199 normalisation."""
    H = X
    if lb is not None:
        H = 2.0 * (X - lb) / (ub - lb) - 1.0          # inf:191
    for W, b in zip(weights[:-1], biases[:-1]):
        H = torch.tanh(torch.add(torch.matmul(H, W), b))
    return torch.add(torch.matmul(H, weights[-1]), biases[-1])


class Oracle:
    """Loss graph of one reference class.

    kind: 'plate' (PINN, F5 plane stress, optional dist/part composite), 'inf', 'semi' (DeepHPM, F7
    plane strain), 'conf' (DeepElasticWave, F7).
    """

    def __init__(self, kind, uv_weights, uv_biases, dist=None, part=None, lb=None, ub=None,
                 E=None, mu=0.25, rho=1.0, hole_r=0.1, dtype=DT):
        self.kind = kind
        self.dtype = dtype
        self.W = [_t(w, dtype).clone().requires_grad_(True) for w in uv_weights]
        self.b = [_t(b, dtype).reshape(1, -1).clone().requires_grad_(True) for b in uv_biases]
        self.dist = None if dist is None else ([_t(w, dtype) for w in dist[0]], [_t(b, dtype).reshape(1, -1) for b in dist[1]])
        self.part = None if part is None else ([_t(w, dtype) for w in part[0]], [_t(b, dtype).reshape(1, -1) for b in part[1]])
        self.E = (20.0 if kind == 'plate' else 2.5) if E is None else E   # plate:39, semi:35
        self.mu, self.rho, self.hole_r = mu, rho, hole_r
        self.lb = None if lb is None else _t(lb, dtype)
        self.ub = None if ub is None else _t(ub, dtype)
        self.normalize = (kind == 'inf')                                  # inf:191 only
        # Adam slots (persist across train() calls like TF graph-level slots)
        self.m = [torch.zeros_like(p) for p in self.params()]
        self.v = [torch.zeros_like(p) for p in self.params()]
        self.step = 0

    def params(self):
        return self.W + self.b

    def _nn(self, X):
        if self.normalize:
            return neural_net(X, self.W, self.b, self.lb, self.ub)
        return neural_net(X, self.W, self.b)

    # ---- net_uv
    def net_uv(self, x, y, t):
        X = torch.cat([x, y, t], 1)
        out = self._nn(X)
        if self.kind == 'plate':                                           # plate:358-388
            cols = [out[:, i:i + 1] for i in range(5)]
            if self.dist is not None:
                D = neural_net(X, *self.dist)
                P = neural_net(X, *self.part)
                cols = [P[:, i:i + 1] + D[:, i:i + 1] * cols[i] for i in range(5)]
            return tuple(cols)                                             # u, v, s11, s22, s12
        return tuple(out[:, i:i + 1] for i in range(7))                    # semi:208-218  u,v,ut,vt,s11,s22,s12

    def net_e(self, x, y, t):                                              # plate:390-396, semi:220-226
        o = self.net_uv(x, y, t)
        u, v = o[0], o[1]
        e11 = _grad(u, x)
        e22 = _grad(v, y)
        e12 = _grad(u, y) + _grad(v, x)
        return e11, e22, e12

    def net_f_sig(self, x, y, t):
        E, mu, rho = self.E, self.mu, self.rho
        if self.kind == 'plate':                                           # plate:404-439
            u, v, s11, s22, s12 = self.net_uv(x, y, t)
            e11, e22, e12 = self.net_e(x, y, t)
            sp11 = E / (1 - mu * mu) * e11 + E * mu / (1 - mu * mu) * e22
            sp22 = E * mu / (1 - mu * mu) * e11 + E / (1 - mu * mu) * e22
            sp12 = E / (2 * (1 + mu)) * e12
            f_s11, f_s12, f_s22 = s11 - sp11, s12 - sp12, s22 - sp22
            s11_1 = _grad(s11, x)
            s12_2 = _grad(s12, y)
            u_t = _grad(u, t)
            u_tt = _grad(u_t, t)
            s22_2 = _grad(s22, y)
            s12_1 = _grad(s12, x)
            v_t = _grad(v, t)
            v_tt = _grad(v_t, t)
            f_u = s11_1 + s12_2 - rho * u_tt
            f_v = s22_2 + s12_1 - rho * v_tt
            return f_u, f_v, f_s11, f_s22, f_s12
        # F7 plane strain: semi:228-272, inf:221-265, conf:304-348
        u, v, ut, vt, s11, s22, s12 = self.net_uv(x, y, t)
        e11, e22, e12 = self.net_e(x, y, t)
        coef = E / ((1 + mu) * (1 - 2 * mu))
        sp11 = coef * (1 - mu) * e11 + coef * mu * e22
        sp22 = coef * mu * e11 + coef * (1 - mu) * e22
        sp12 = E / (2 * (1 + mu)) * e12
        f_s11, f_s12, f_s22 = s11 - sp11, s12 - sp12, s22 - sp22
        f_ut = _grad(u, t) - ut
        f_vt = _grad(v, t) - vt
        s11_1 = _grad(s11, x)
        s12_2 = _grad(s12, y)
        u_tt = _grad(ut, t)
        s22_2 = _grad(s22, y)
        s12_1 = _grad(s12, x)
        v_tt = _grad(vt, t)
        f_u = s11_1 + s12_2 - rho * u_tt
        f_v = s22_2 + s12_1 - rho * v_tt
        return f_u, f_v, f_ut, f_vt, f_s11, f_s22, f_s12

    def net_t(self, x, y, t):                                              # plate:452-461
        r = self.hole_r
        u, v, s11, s22, s12 = self.net_uv(x, y, t)
        nx, ny = -x / r, -y / r
        return s11 * nx + s12 * ny, s12 * nx + s22 * ny

    # ---- losses
    @staticmethod
    def _cols(A, req=True):
        A = _t(A)
        return [A[:, i:i + 1].clone().requires_grad_(req and i < 3) for i in range(A.shape[1])]

    @staticmethod
    def _ms(a):
        return torch.mean(torch.square(a))

    def loss_terms(self, sets):
        """sets: dict of numpy arrays.  plate: Collo[N,3], HOLE[N,3].  inf/semi: Collo, SRC[N,5], IC, UP.
        conf: Collo, SRC, IC, FIXED.  Returns (dict of scalar tensors, total loss)."""
        ms = self._ms
        k = self.kind
        x, y, t = self._cols(sets['Collo'])[:3]
        f = self.net_f_sig(x, y, t)
        T = {}
        if k == 'plate':                                                   # plate:187-193,217
            T['loss_f_uv'] = ms(f[0]) + ms(f[1])
            T['loss_f_s'] = ms(f[2]) + ms(f[3]) + ms(f[4])
            xh, yh, th = self._cols(sets['HOLE'], req=False)[:3]
            tx, ty = self.net_t(xh, yh, th)
            T['loss_HOLE'] = ms(tx) + ms(ty)
            loss = 10 * (T['loss_f_uv'] + T['loss_f_s'] + T['loss_HOLE'])
            return T, loss
        T['loss_f_uv'] = ms(f[0]) + ms(f[1]) + ms(f[2]) + ms(f[3])          # semi:112-118
        T['loss_f_s'] = ms(f[4]) + ms(f[5]) + ms(f[6])
        ic = self._cols(sets['IC'], req=False)
        o = self.net_uv(*ic[:3])
        T['loss_IC'] = ms(o[0]) + ms(o[1]) + ms(o[2]) + ms(o[3])            # semi:119-122
        src = self._cols(sets['SRC'], req=False)
        o = self.net_uv(*src[:3])
        T['loss_SRC'] = ms(o[0] - src[3]) + ms(o[1] - src[4])               # semi:123-124
        if k in ('inf', 'semi'):
            up = self._cols(sets['UP'], req=False)
            o = self.net_uv(*up[:3])
            T['loss_NB'] = ms(o[5]) + ms(o[6])                              # semi:125-126
        if k == 'conf':
            fx = self._cols(sets['FIXED'], req=False)
            o = self.net_uv(*fx[:3])
            T['loss_FIX'] = ms(o[0]) + ms(o[1])                             # conf:147-148
        if k == 'inf':                                                      # inf:119 (NB commented out)
            loss = T['loss_f_uv'] + T['loss_f_s'] + T['loss_IC'] + T['loss_SRC']
        elif k == 'semi':                                                   # semi:127
            loss = 5 * T['loss_f_uv'] + 5 * T['loss_f_s'] + 2 * T['loss_IC'] + 2 * T['loss_SRC'] + 2 * T['loss_NB']
        else:                                                               # conf:156
            loss = 5 * T['loss_f_uv'] + 5 * T['loss_f_s'] + T['loss_SRC'] + T['loss_IC'] + T['loss_FIX']
        return T, loss

    def loss_and_grad(self, sets):
        """-> (terms dict of floats, loss float, flat gradient in var_list order weights+biases)."""
        T, loss = self.loss_terms(sets)
        gs = torch.autograd.grad(loss, self.params())
        flat = np.concatenate([g.detach().numpy().ravel() for g in gs])
        return {k: float(v.detach()) for k, v in T.items()}, float(loss.detach()), flat

    # ---- TF1 Adam: eps added to sqrt(v) un-bias-corrected
    def adam_step(self, grads, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.step += 1
        tt = self.step
        lr_t = lr * np.sqrt(1.0 - beta2 ** tt) / (1.0 - beta1 ** tt)
        with torch.no_grad():
            for p, g, m, v in zip(self.params(), grads, self.m, self.v):
                m.mul_(beta1).add_(g, alpha=1 - beta1)
                v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
                p.sub_(lr_t * m / (v.sqrt() + eps))

    def train(self, sets, iters, learning_rate, batch_num=None):
        """Adam loop with the reference's bookkeeping: losses recorded AFTER each update
        (plate:496-506; chunking semi:299-326).  Returns dict name -> list."""
        rec = {}
        N = sets['Collo'].shape[0]
        chunks = [(0, N)] if batch_num is None else \
            [(int(i * N / batch_num), int((i + 1) * N / batch_num)) for i in range(batch_num)]
        for (a, b) in chunks:
            cur = dict(sets)
            cur['Collo'] = sets['Collo'][a:b]
            for _ in range(iters):
                T, loss = self.loss_terms(cur)
                gs = torch.autograd.grad(loss, self.params())
                self.adam_step(gs, learning_rate)
                with torch.enable_grad():
                    T, loss = self.loss_terms(cur)
                for kk, vv in T.items():
                    rec.setdefault(kk, []).append(float(vv.detach()))
                rec.setdefault("loss", []).append(float(loss.detach()))
        return rec

    def flat_params(self):
        return np.concatenate([p.detach().numpy().ravel() for p in self.params()])

    def set_flat_params(self, flat):
        o = 0
        with torch.no_grad():
            for p in self.params():
                n = p.numel()
                p.copy_(_t(flat[o:o + n], self.dtype).reshape(p.shape))
                o += n

    def predict(self, x_star, y_star, t_star):
        """plate:561-570 / semi:348-358: (u, v, s11, s22, s12, e11, e22, e12), each (N,1) numpy."""
        x = _t(x_star, self.dtype).reshape(-1, 1).clone().requires_grad_(True)
        y = _t(y_star, self.dtype).reshape(-1, 1).clone().requires_grad_(True)
        t = _t(t_star, self.dtype).reshape(-1, 1).clone().requires_grad_(True)
        o = self.net_uv(x, y, t)
        e = self.net_e(x, y, t)
        if self.kind == 'plate':
            outs = (o[0], o[1], o[2], o[3], o[4]) + e
        else:
            outs = (o[0], o[1], o[4], o[5], o[6]) + e
        return tuple(a.detach().numpy() for a in outs)


# ----------------------------------------------------------------------------- plate pre-training losses
def _net_cols(W, b, A, need_t=False):
    A = _t(A)
    x, y = A[:, 0:1], A[:, 1:2]
    t = A[:, 2:3].clone().requires_grad_(need_t)
    out = neural_net(torch.cat([x, y, t], 1), W, b)
    return [out[:, i:i + 1] for i in range(out.shape[1])], t


def loss_dist(dist_W, dist_b, DIST, IC):
    """plate:194-200: fit of the distance-function net D to its targets + zero d/dt at the IC points
    (net_dist plate:322-329, net_dist_dt plate:331-345).  Returns (loss, flat grad) with W, b lists of numpy arrays."""
    W = [_t(w).clone().requires_grad_(True) for w in dist_W]
    b = [_t(x).reshape(1, -1).clone().requires_grad_(True) for x in dist_b]
    ms = lambda a: torch.mean(torch.square(a))
    D, _ = _net_cols(W, b, DIST)
    tg = _t(DIST)
    loss = sum(ms(D[c] - tg[:, 3 + c:4 + c]) for c in range(5))
    Di, t = _net_cols(W, b, IC, need_t=True)
    loss = loss + ms(_grad(Di[0], t)) + ms(_grad(Di[1], t))
    gs = torch.autograd.grad(loss, W + b)
    return float(loss.detach()), np.concatenate([g.numpy().ravel() for g in gs])


def loss_part(part_W, part_b, IC, LF, RT, UP, LW):
    """plate:201-215 with net_part plate:347-356."""
    W = [_t(w).clone().requires_grad_(True) for w in part_W]
    b = [_t(x).reshape(1, -1).clone().requires_grad_(True) for x in part_b]
    ms = lambda a: torch.mean(torch.square(a))
    P, t = _net_cols(W, b, IC, need_t=True)
    loss = sum(ms(P[c]) for c in range(5)) + ms(_grad(P[0], t)) + ms(_grad(P[1], t))
    P, _ = _net_cols(W, b, LF)
    loss = loss + ms(P[0]) + ms(P[4])
    P, _ = _net_cols(W, b, RT)
    loss = loss + ms(P[2] - _t(RT)[:, 3:4]) + ms(P[4])
    P, _ = _net_cols(W, b, LW)
    loss = loss + ms(P[1]) + ms(P[4])
    P, _ = _net_cols(W, b, UP)
    loss = loss + ms(P[3]) + ms(P[4])
    gs = torch.autograd.grad(loss, W + b)
    return float(loss.detach()), np.concatenate([g.numpy().ravel() for g in gs])
