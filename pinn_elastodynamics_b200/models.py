"""Drop-in model classes mirroring the reference's Python class surface.

    PINN             PlateHoleQuarter/train/train.py:26-612            (alias PhysicsInformedNN, BASELINE.json)
    DeepHPM          ElasticWaveInfinite/ElasticWave.py:21-376 (variant='inf')
                     ElasticWaveSemiInfinite/ElasticWave.py:23-394 (variant='semi', default)
    DeepElasticWave  ElasticWaveConfined/ElasticWave.py:21-474

Same constructor arguments, method names, return tuples, print formats and error behaviour
(`load_NN` asserts the layer count, plate:299).  The TF1 graph + session is replaced by
`LossEngine` (hand-written sm_100a kernels behind include/pinn_elasto.h); parameters live on the GPU
in fp32.  Semantics kept: losses are recorded AFTER each update (plate:497,502-505) -- obtained for
free as the pre-update terms of the next step plus one final evaluation; Adam slots persist across
train() calls (graph-level slots); each train_bfgs call restarts L-BFGS memory; `batch_num` chunk i
= rows [int(i*N/B), int((i+1)*N/B)) (semi:299-302); only the uv network is trained by train / train_bfgs.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import scipy.optimize          # L-BFGS-B driver of train_bfgs (imported here: ~0.4 s the first time, not inside the optimisation call)
import torch

from . import _lib as L
from .engine import LossEngine, Network
from .linesearch import strong_wolfe


def xavier_init_lists(layers, rng):
    """Xavier N(0, 2/(in+out)) truncated at 2 sigma + zero biases (plate:258-274).  The TF1 RNG stream
    cannot be reproduced; a numpy Generator seeded with the reference's seed 1111 (plate:22) is used."""
    Ws, bs = [], []
    for l in range(len(layers) - 1):
        fi, fo = layers[l], layers[l + 1]
        w = rng.standard_normal((fi, fo))
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        Ws.append(w * np.sqrt(2.0 / (fi + fo)))
        bs.append(np.zeros((1, fo)))
    return Ws, bs


def _col(a):
    return np.asarray(a, dtype=np.float64).reshape(-1, 1)


class _Base:
    """Shared machinery: networks, engine, Adam loop, SciPy L-BFGS-B driver, predict, pickle I/O."""

    formulation = L.RES_F7
    term_names = ()          # names of loss-term slots, in slot order
    term_weights = ()        # weight of each slot in the total loss
    adam_returns = ()        # which histories train() returns (names; 'loss' = weighted total)
    print_fmt = 'It: %d, Loss: %.3e'
    bfgs_options = dict(maxiter=1000, maxfun=1000, maxcor=50, maxls=50, ftol=1e-3 * np.finfo(float).eps)

    def _init_common(self, uv_layers, lb, ub, engine, verbose, dtype):
        self.count = 0
        self.lb = lb
        self.ub = ub
        self.uv_layers = list(uv_layers)
        self.verbose = verbose
        self.dtype = dtype
        # engine: 'auto' (the tcgen05 engine for the collocation term where the net fits it -- hidden width <= 56, F5 / F7 -- and the fp32
        # SIMT engine for every other term and every other net), 'simt' (fp32 FFMA everywhere: the parity anchor), or an explicit name
        # from _lib.ENGINES.  None (a caller that keeps the reference's constructor signature) reads $PE_ENGINE, default 'auto'.
        if engine is None:
            engine = os.environ.get('PE_ENGINE', 'auto')
        self._engine_name = engine
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.uv_net = Network(self.uv_layers, self.device)
        self.engine = LossEngine(self.uv_net, engine)

    # ---- parameter I/O (plate:258-306)
    def initialize_NN(self, layers):
        if not hasattr(self, '_rng'):
            self._rng = np.random.default_rng(1111)
        return xavier_init_lists(layers, self._rng)

    def load_NN(self, fileDir, layers):
        with open(fileDir, 'rb') as f:
            uv_weights, uv_biases = pickle.load(f)
        # Stored model must have the same # of layers (plate:299)
        assert len(layers) == (len(uv_weights) + 1)
        if self.verbose:
            for _ in uv_weights:
                print("Load NN parameters successfully...")
        return [np.asarray(w) for w in uv_weights], [np.asarray(b) for b in uv_biases]

    def _save(self, net, fileDir, label):
        Ws, bs = net.get_weights(self.dtype)
        with open(fileDir, 'wb') as f:
            pickle.dump([Ws, bs], f)
        print("Save " + label + "NN parameters successfully...")

    @property
    def uv_weights(self):
        return self.uv_net.get_weights(self.dtype)[0]

    @property
    def uv_biases(self):
        return self.uv_net.get_weights(self.dtype)[1]

    # ---- loss bookkeeping
    def _total(self, terms):
        return float(sum(w * terms[i] for i, w in enumerate(self.term_weights)))

    def _in_affine(self):
        return (1, 1, 1), (0, 0, 0)

    def _evaluate_terms(self):
        self.engine.evaluate()
        return self.engine.terms_host()

    # ---- Adam loop (plate:475-506, semi:290-328, conf:373-408)
    def _adam_loop(self, iters, learning_rate, chunks, refeed=False):
        """refeed=True reproduces the reference's per-iteration host traffic (feed_dict re-upload of every point
        array on each sess.run, plate:482-497, and the loss scalars fetched back, plate:502-505): every step
        copies the point sets host->device from pinned memory and reads the loss terms back (bench `e2e`)."""
        hist_rows = []
        eng = self.engine
        self.h2d_bytes_per_step = self.d2h_bytes_per_step = 0
        if refeed:
            # double-buffered device copies of every point set + a copy stream: the upload of step i+1 overlaps the
            # kernels of step i; the loss terms of step i-1 are read back while step i runs (queue stays 1-2 steps deep).
            # The pinned host copies, the second device buffers, the stream and the events live as long as the point sets do
            # (the reference keeps its feed_dict arrays for the life of the object as well): a train() call allocates nothing.
            terms = [t for t in eng.terms if t.enabled]
            key = tuple(t.points.data_ptr() for t in terms)
            rc = getattr(self, '_refeed_cache', None)
            if rc is None or rc['key'] != key:
                rc = self._refeed_cache = dict(
                    key=key, pinned=[t.points.cpu().pin_memory() for t in terms], bufs=[(t.points, torch.empty_like(t.points)) for t in terms],
                    host_rows=torch.zeros((2, L.PE_MAX_TERMS), dtype=torch.float32).pin_memory(), copy_stream=torch.cuda.Stream(self.device),
                    up_done=[torch.cuda.Event(), torch.cuda.Event()], used_done=[torch.cuda.Event(), torch.cuda.Event()],
                    step_done=[torch.cuda.Event(), torch.cuda.Event()])
            pinned, bufs, host_rows, copy_stream = rc['pinned'], rc['bufs'], rc['host_rows'], rc['copy_stream']
            up_done, used_done, step_done = rc['up_done'], rc['used_done'], rc['step_done']   # upload into buffer b finished / kernels reading
            self.h2d_bytes_per_step = sum(h.numel() * 4 for h in pinned)                      # buffer b finished / loss row copied to the host
            self.d2h_bytes_per_step = L.PE_MAX_TERMS * 4
            main = torch.cuda.current_stream(self.device)

            def upload(b):
                copy_stream.wait_event(used_done[b])
                with torch.cuda.stream(copy_stream):
                    for t, h, bb in zip(terms, pinned, bufs):
                        bb[b].copy_(h, non_blocking=True)
                    up_done[b].record(copy_stream)
            used_done[0].record(main); used_done[1].record(main)
            upload(0)
        for (a, b) in chunks:
            if a is not None:
                eng.set_chunk(self._collo_term, a, b)
            hist = torch.zeros((iters + 1, L.PE_MAX_TERMS), dtype=torch.float32, device=self.device)
            for it in range(iters):
                if refeed:
                    cur = it & 1
                    main.wait_event(up_done[cur])
                    for t, bb in zip(terms, bufs):
                        t.points = bb[cur]
                    if it + 1 < iters:
                        upload(1 - cur)
                eng.adam_step(learning_rate, hist[it])                 # hist[it] = terms BEFORE update it
                if eng.comm is not None and (it + 1) % 1000 == 0:
                    eng.check_comm()                                   # a peer that timed out must not go unnoticed for the rest of a long run
                if refeed:
                    used_done[cur].record(main)
                    host_rows[cur].copy_(hist[it], non_blocking=True)
                    step_done[cur].record(main)
                    if it > 0:
                        step_done[1 - cur].synchronize()               # loss of the previous step is on the host
            if refeed:
                torch.cuda.current_stream(self.device).synchronize()
                for t, bb in zip(terms, bufs):
                    t.points = bb[0]
            eng.evaluate(hist[iters])                                  # terms after the last update
            eng.check_comm()
            h = hist.cpu().numpy().astype(np.float64)
            if self.verbose:                                           # same lines as plate:499-501, printed after the
                for it in range(0, iters, 10):                         # asynchronous loop has drained (no per-step host sync)
                    print(self.print_fmt % (it, self._total(h[it + 1])))
            hist_rows.append(h[1:])                                    # row it+1 = terms after update it
        H = np.concatenate(hist_rows, 0) if hist_rows else np.zeros((0, L.PE_MAX_TERMS))
        out = {name: list(H[:, i]) for i, name in enumerate(self.term_names)}
        out['loss'] = [self._total(r) for r in H]
        return out

    def _chunks(self, batch_num):
        N = self._collo_term.global_n
        return [(int(i * N / batch_num), int((i + 1) * N / batch_num)) for i in range(batch_num)]

    # ---- L-BFGS-B through SciPy, as tf.contrib.opt.ScipyOptimizerInterface does (plate:240-247,522-525)
    bfgs_driver = 'scipy'          # 'scipy': SciPy's L-BFGS-B exactly as the reference drives it; 'gpu': device-resident L-BFGS

    def _bfgs(self, options, callback, engine=None, net=None, total=None, shown=None, driver=None):
        """total(terms) is the minimised scalar; shown(terms) is what the reference passes to loss_callback
        (`fetches`, e.g. the unscaled loss_DIST while 1000*loss_DIST is minimised, plate:220,543)."""
        engine = engine or self.engine
        net = net or self.uv_net
        total = total or self._total
        shown = shown or total
        driver = driver or dict(options).pop('driver', None) or self.bfgs_driver
        if driver == 'gpu':
            return self._bfgs_gpu({k: v for k, v in dict(options).items() if k != 'driver'}, callback, engine, net, total, shown)
        if driver != 'scipy':
            raise ValueError(f"bfgs driver {driver!r}: 'scipy' or 'gpu'")
        options = {k: v for k, v in dict(options).items() if k != 'driver'}

        def fun(x):
            net.set_flat(x)
            engine.evaluate()
            terms = engine.terms_host()
            g = engine.grad_compact_host().astype(np.float64)
            f = total(terms)
            callback(shown(terms))                        # loss_callback fires on every evaluation (plate:463-465)
            return f, g

        x0 = net.get_flat().astype(np.float64)
        res = scipy.optimize.minimize(fun, x0, jac=True, method='L-BFGS-B', options=dict(options))
        net.set_flat(res.x)
        return res

    # ---- device-resident L-BFGS: same objective, same evaluation kernels, same options dictionary as
    # the SciPy path, but x, g and the (S, Y) history stay in HBM (csrc/pe_lbfgs.cu) and the host only reads the
    # scalars the line search branches on (8 loss terms, g.d, max|g|: one 40-byte D2H per evaluation).
    def _bfgs_gpu(self, options, callback, engine, net, total, shown):
        """Unconstrained L-BFGS (the reference passes no bounds, so L-BFGS-B's projection is inactive: plate:240-247) with
        a strong-Wolfe line search (sufficient decrease 1e-3, curvature 0.9: the constants of L-BFGS-B's lnsrlb).
        Honours maxiter / maxfun / maxcor / maxls / ftol / gtol with SciPy's meaning and returns an OptimizeResult.
        Iterates are not bit-identical to SciPy's (different line search, fp32 vectors); tests compare the reached loss."""
        lib, n, dev = engine.lib, net.Pp, self.device
        m = max(1, min(int(options.get('maxcor', 10)), 64))
        maxiter = int(options.get('maxiter', 15000)); maxfun = int(options.get('maxfun', 15000))
        maxls = int(options.get('maxls', 20)); ftol = float(options.get('ftol', 2.2204460492503131e-09))
        gtol = float(options.get('gtol', 1e-5))
        c1, c2 = 1e-3, 0.9
        f32 = dict(dtype=torch.float32, device=dev)
        S = torch.zeros(m, n, **f32); Y = torch.zeros(m, n, **f32); state = torch.zeros(m + 3, **f32)
        xprev = torch.empty(n, **f32); gprev = torch.empty(n, **f32); d = torch.zeros(n, **f32)
        scal = torch.zeros(10, **f32)
        if not engine._built:
            engine.build()
        x, g = net.params, engine.out[:n]
        st = engine._stream
        P = lambda t: C.c_void_p(t.data_ptr())
        nfev = 0

        def scalars(other):
            """(terms, g.other, max|g|) of the current evaluation: one small D2H."""
            scal[:8].copy_(engine.out[n:n + 8])
            L.check(lib.pe_vec_dot_max(n, P(g), P(other), P(scal[8:]), st()), 'pe_vec_dot_max')
            h = scal.cpu().numpy().astype(np.float64)
            return h[:8], float(h[8]), float(h[9])

        def feval(other):
            nonlocal nfev
            engine.evaluate()
            nfev += 1
            terms, gd, gmax = scalars(other)
            callback(shown(terms))
            return total(terms), gd, gmax

        def phi(alpha):
            L.check(lib.pe_vec_axpy(n, P(x), P(xprev), float(alpha), P(d), st()), 'pe_vec_axpy')
            return feval(d)

        scal2 = torch.zeros(12, **f32)
        import time as _time
        trace = {'enqueue': 0.0, 'sync': 0.0, 'n': 0} if os.environ.get('PE_BFGS_TRACE') else None

        def step_and_eval(alpha0):
            """direction is in d: queue  g.d -> save x, g -> x = x + alpha0 d -> evaluate -> g_new.d  and read everything back with ONE
            device->host copy: (dphi0, f(alpha0), dphi(alpha0), max|g_new|)."""
            nonlocal nfev
            L.check(lib.pe_vec_dot_max(n, P(g), P(d), P(scal2[10:]), st()), 'pe_vec_dot_max')
            xprev.copy_(x); gprev.copy_(g)
            L.check(lib.pe_vec_axpy(n, P(x), P(xprev), float(alpha0), P(d), st()), 'pe_vec_axpy')
            engine.evaluate()
            nfev += 1
            scal2[:8].copy_(engine.out[n:n + 8])
            L.check(lib.pe_vec_dot_max(n, P(g), P(d), P(scal2[8:10]), st()), 'pe_vec_dot_max')
            if trace is not None:
                t1 = _time.perf_counter()
            h = scal2.cpu().numpy().astype(np.float64)
            if trace is not None:
                t2 = _time.perf_counter()
                trace['enqueue'] += t1 - trace.get('t0', t1); trace['sync'] += t2 - t1; trace['n'] += 1
            callback(shown(h[:8]))
            return float(h[10]), (total(h[:8]), float(h[8]), float(h[9]))

        f, gg, gmax = feval(g)
        count, head, nit = 0, -1, 0
        message, success = 'STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT', False
        while nit < maxiter:
            if gmax <= gtol:
                message, success = 'CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL', True
                break
            if nfev >= maxfun:
                message = 'STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT'
                break
            if trace is not None:
                trace['t0'] = _time.perf_counter()
            L.check(lib.pe_lbfgs_direction(n, m, count, max(head, 0), P(g), P(S), P(Y), P(state), P(d), st()), 'pe_lbfgs_direction')
            alpha0 = 1.0 if count > 0 else min(1.0, 1.0 / max(np.sqrt(gg), 1e-30))   # L-BFGS-B: first step 1/||d||
            # the first trial point of the line search is evaluated right behind the direction kernel: one host round trip per iteration
            # in the common case (unit step accepted) instead of one per kernel result the host looks at
            dphi0, first = step_and_eval(alpha0)
            if not dphi0 < 0.0:
                x.copy_(xprev); g.copy_(gprev)                          # the speculative step is void
                if count == 0:
                    message = 'ABNORMAL: NOT A DESCENT DIRECTION'
                    break
                count = 0                                               # drop the history, restart from steepest descent
                _, gg, gmax = scalars(g)                                # ... whose first trial step is 1 / ||g|| of the CURRENT gradient
                continue
            ok, alpha, f_new, _, gmax_new = strong_wolfe(phi, f, dphi0, alpha0, c1, c2, maxls, budget=lambda: maxfun - nfev, first=first)
            if not ok:
                x.copy_(xprev); g.copy_(gprev)
                if nfev >= maxfun:
                    message = 'STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT'
                    break
                if count == 0:
                    message = 'ABNORMAL: LINE SEARCH FAILED'
                    break
                count = 0
                _, gg, gmax = scalars(g)
                continue
            head = (head + 1) % m
            count = min(count + 1, m)
            L.check(lib.pe_lbfgs_store_pair(n, m, head, P(x), P(xprev), P(g), P(gprev), P(S), P(Y), P(state), st()), 'pe_lbfgs_store_pair')
            nit += 1
            rel = (f - f_new) / max(abs(f), abs(f_new), 1.0)
            f, gmax = f_new, gmax_new
            if rel <= ftol:
                message, success = 'CONVERGENCE: RELATIVE REDUCTION OF F <= FTOL', True
                break
        if trace is not None and trace['n']:
            import sys as _sys
            print('[bfgs trace] iterations %d: host enqueue %.3f ms, wait for the device %.3f ms per iteration' % (trace['n'], 1e3 * trace['enqueue'] / trace['n'], 1e3 * trace['sync'] / trace['n']), file=_sys.stderr)
        return scipy.optimize.OptimizeResult(x=net.get_flat().astype(np.float64), fun=f, nit=nit, nfev=nfev, message=message,
                                             success=success, status=0 if success else 1)

    def callback(self, loss):
        self.count = self.count + 1
        if self.verbose:
            print('{} th iterations, Loss: {}'.format(self.count, loss))

    # ---- inference (plate:561-570, semi:348-370)
    def _points_tensor(self, x_star, y_star, t_star):
        X = np.concatenate([_col(x_star), _col(y_star), _col(t_star)], 1).astype(np.float32)
        return torch.from_numpy(X).to(self.device)

    def _predict_aux(self, pts):
        return None

    def predict(self, x_star, y_star, t_star):
        pts = self._points_tensor(x_star, y_star, t_star)
        sc, sh = self._in_affine()
        out = self.uv_net.forward_fields(pts, self.formulation, sc, sh, self._predict_aux(pts)).cpu().numpy().astype(self.dtype)
        return tuple(out[:, i:i + 1] for i in range(8))

    def probe(self, x_star, y_star, t_star):
        return self.predict(x_star, y_star, t_star)

    def predict_frames(self, x_star, y_star, times):
        """All output frames in ONE launch (the reference drivers call predict() once per frame, 8 sess.run each:
        plate:980-998, semi:907-922).  Returns a list with one predict()-style 8-tuple per time in `times`."""
        x = _col(x_star); y = _col(y_star)
        n, times = x.shape[0], np.asarray(times, dtype=np.float64).ravel()
        X = np.concatenate([np.tile(x, (len(times), 1)), np.tile(y, (len(times), 1)), np.repeat(times, n)[:, None]], 1)
        out = self.predict(X[:, 0:1], X[:, 1:2], X[:, 2:3])
        return [tuple(f[i * n:(i + 1) * n] for f in out) for i in range(len(times))]


# ======================================================================================= plate
class PINN(_Base):
    """Defected plate, plane stress, 5 outputs, second-order in t (PlateHoleQuarter/train/train.py:26)."""

    formulation = L.RES_F5
    term_names = ('loss_f_uv', 'loss_f_s', 'loss_HOLE')
    term_weights = (10.0, 10.0, 10.0)                       # plate:217
    print_fmt = 'It: %d, Loss: %.6e'                        # plate:501
    bfgs_options = dict(maxiter=70000, maxfun=70000, maxcor=50, maxls=50, ftol=0.00001 * np.finfo(float).eps)   # plate:243-247
    pre_options = dict(maxiter=20000, maxfun=20000, maxcor=50, maxls=50, ftol=0.00001 * np.finfo(float).eps)    # plate:223-237

    def __init__(self, Collo, HOLE, IC, LF, RT, UP, LW, DIST, uv_layers, dist_layers, part_layers, lb, ub,
                 partDir='', distDir='', uvDir='', engine=None, verbose=True, dtype=np.float64, composite=None):
        self._init_common(uv_layers, lb, ub, engine, verbose, dtype)
        self.E, self.mu, self.rho, self.hole_r = 20.0, 0.25, 1.0, 0.1           # plate:39-42
        A = lambda a: None if a is None else np.asarray(a, dtype=np.float64)
        self.Collo, self.HOLE, self.IC, self.LF, self.RT, self.UP, self.LW, self.DIST = map(A, (Collo, HOLE, IC, LF, RT, UP, LW, DIST))
        self.x_c, self.y_c, self.t_c = self.Collo[:, 0:1], self.Collo[:, 1:2], self.Collo[:, 2:3]
        self.x_HOLE, self.y_HOLE, self.t_HOLE = self.HOLE[:, 0:1], self.HOLE[:, 1:2], self.HOLE[:, 2:3]
        self.dist_layers, self.part_layers = dist_layers, part_layers
        # composite u = P + D*N is used when dist/part networks exist (the reference always builds them, plate:96-106);
        # composite=False gives the plain 5-output net (BASELINE configs 1/2/4: "5x50 mixed-variable net")
        self.composite = (dist_layers is not None and part_layers is not None) if composite is None else composite
        if self.composite:
            self.dist_net = Network(dist_layers, self.device)
            self.part_net = Network(part_layers, self.device)
            self.dist_net.set_weights(*(self.initialize_NN(dist_layers) if distDir == '' else self._loading("dist", distDir, dist_layers)))
            self.part_net.set_weights(*(self.initialize_NN(part_layers) if partDir == '' else self._loading("part", partDir, part_layers)))
        self.uv_net.set_weights(*(self.initialize_NN(self.uv_layers) if uvDir == '' else self._loading("uv", uvDir, self.uv_layers)))
        self._setup_terms()

    def _loading(self, what, d, layers):
        if self.verbose:
            print("Loading %s NN ..." % what)
        return self.load_NN(d, layers)

    def _composite_aux(self, pts, K):
        """[n][2][K][5]: jets of the frozen dist / part nets at the points (plate:361-362)."""
        D = self.dist_net.forward_jets(pts, K)
        P = self.part_net.forward_jets(pts, K)
        return torch.stack([D, P], 1).contiguous()

    def _setup_terms(self):
        eng = self.engine
        eng.terms = []
        mat = dict(E=self.E, mu=self.mu, rho=self.rho, hole_r=self.hole_r)
        aux_c = aux_h = None
        if self.composite:
            aux_c = self._composite_aux(torch.from_numpy(self.Collo[:, :3].astype(np.float32)).to(self.device), 5)
            aux_h = self._composite_aux(torch.from_numpy(self.HOLE[:, :3].astype(np.float32)).to(self.device), 1)
        self._collo_term = eng.add_term('Collo', L.RES_F5, 5, self.Collo[:, :3], terms=(0, 1), weights=(10.0, 10.0),
                                        aux=aux_c, aux_k=5 if self.composite else 0, **mat)
        self._hole_term = eng.add_term('HOLE', L.RES_TRACTION, 1, self.HOLE[:, :3], terms=(2,), weights=(10.0,),
                                       aux=aux_h, aux_k=1 if self.composite else 0, **mat)

    # ---- pre-training of the distance-function / particular-solution networks (plate:194-237, 527-559)
    def _pre_engines(self):
        if hasattr(self, 'dist_engine'):
            return
        if not self.composite or any(a is None for a in (self.IC, self.LF, self.RT, self.UP, self.LW, self.DIST)):
            raise L.PeError('train_bfgs_dist / train_bfgs_part need dist/part networks and the IC, LF, RT, UP, LW, DIST point sets')
        W = 1000.0                                                             # optimizer_dist / optimizer_part minimise 1000 * loss
        de = self.dist_engine = LossEngine(self.dist_net, 'simt')
        de.add_term('DIST', L.RES_COLS, 1, self.DIST[:, :8], cols=(0, 1, 2, 3, 4), tgts=(3, 4, 5, 6, 7), terms=(0,) * 5, weights=(W,) * 5)   # plate:194-198
        de.add_term('IC_dt', L.RES_DT, 2, self.IC[:, :3], cols=(0, 1), tgts=(-1, -1), terms=(0, 0), weights=(W, W))                          # plate:199-200
        pe_ = self.part_engine = LossEngine(self.part_net, 'simt')
        pe_.add_term('IC', L.RES_COLS, 1, self.IC[:, :3], cols=(0, 1, 2, 3, 4), tgts=(-1,) * 5, terms=(0,) * 5, weights=(W,) * 5)            # plate:201-205
        pe_.add_term('IC_dt', L.RES_DT, 2, self.IC[:, :3], cols=(0, 1), tgts=(-1, -1), terms=(0, 0), weights=(W, W))                         # plate:206-207
        pe_.add_term('LF', L.RES_COLS, 1, self.LF[:, :3], cols=(0, 4), tgts=(-1, -1), terms=(0, 0), weights=(W, W))                          # plate:208-209
        pe_.add_term('RT', L.RES_COLS, 1, self.RT[:, :4], cols=(2, 4), tgts=(3, -1), terms=(0, 0), weights=(W, W))                           # plate:210-211
        pe_.add_term('LW', L.RES_COLS, 1, self.LW[:, :3], cols=(1, 4), tgts=(-1, -1), terms=(0, 0), weights=(W, W))                          # plate:212-213
        pe_.add_term('UP', L.RES_COLS, 1, self.UP[:, :3], cols=(3, 4), tgts=(-1, -1), terms=(0, 0), weights=(W, W))                          # plate:214-215

    def callback_dist(self, loss_dist):
        self.count = self.count + 1
        if self.verbose:
            print('{} th iterations, Loss: {}'.format(self.count, loss_dist))

    callback_part = callback_dist

    def train_bfgs_dist(self, options=None):
        self._pre_engines()
        r = self._bfgs(options or self.pre_options, self.callback_dist, engine=self.dist_engine, net=self.dist_net,
                       total=lambda t: 1000.0 * t[0], shown=lambda t: t[0])
        self.refresh_composite()
        return r

    def train_bfgs_part(self, options=None):
        self._pre_engines()
        r = self._bfgs(options or self.pre_options, self.callback_part, engine=self.part_engine, net=self.part_net,
                       total=lambda t: 1000.0 * t[0], shown=lambda t: t[0])
        self.refresh_composite()
        return r

    def refresh_composite(self):
        """Re-evaluate the frozen dist/part jets (after train_bfgs_dist / train_bfgs_part changed them)."""
        if self.composite:
            self._setup_terms()

    def save_NN(self, fileDir, TYPE=''):
        net = {'UV': self.uv_net, 'DIST': getattr(self, 'dist_net', None), 'PART': getattr(self, 'part_net', None)}.get(TYPE)
        if net is None:
            raise UnboundLocalError("local variable 'uv_weights' referenced before assignment")   # plate:286-289 falls through
        self._save(net, fileDir, TYPE + ' ')

    def train(self, iter, learning_rate, refeed=False):
        r = self._adam_loop(iter, learning_rate, [(None, None)], refeed=refeed)
        return r['loss_f_uv'], r['loss_f_s'], r['loss_HOLE'], r['loss']

    def train_bfgs(self, options=None):
        return self._bfgs(options or self.bfgs_options, self.callback)

    def _predict_aux(self, pts):
        return self._composite_aux(pts, 4) if self.composite else None

    def predict_D(self, x_star, y_star, t_star):
        out = self.dist_net.forward_jets(self._points_tensor(x_star, y_star, t_star), 1).cpu().numpy().astype(self.dtype)
        return tuple(out[:, 0, i:i + 1] for i in range(5))

    def predict_P(self, x_star, y_star, t_star):
        out = self.part_net.forward_jets(self._points_tensor(x_star, y_star, t_star), 1).cpu().numpy().astype(self.dtype)
        return tuple(out[:, 0, i:i + 1] for i in range(5))

    def getloss(self):
        t = self._evaluate_terms()
        vals = {'loss_f_uv': t[0], 'loss_f_s': t[1], 'loss_HOLE': t[2], 'loss': self._total(t)}
        try:                                                                   # plate:605-606
            self._pre_engines()
            self.part_engine.evaluate(); vals['loss_PART'] = self.part_engine.terms_host()[0]
            self.dist_engine.evaluate(); vals['loss_DIST'] = self.dist_engine.terms_host()[0]
        except L.PeError:
            pass
        for k in ('loss_f_uv', 'loss_f_s', 'loss_HOLE', 'loss', 'loss_PART', 'loss_DIST'):
            if k in vals:
                print(k, vals[k])
        return vals


PhysicsInformedNN = PINN       # the name BASELINE.json's north_star uses


# ======================================================================================= waves
class DeepHPM(_Base):
    """Elastic wave, plane strain, 7 outputs (u,v,ut,vt,s11,s22,s12), first-order system.
    variant='semi': ElasticWaveSemiInfinite/ElasticWave.py:23 (loss 5,5,2,2,2; semi:127);
    variant='inf' : ElasticWaveInfinite/ElasticWave.py:21 (normalised inputs inf:191; loss 1,1,1,1, NB unused inf:119)."""

    formulation = L.RES_F7
    term_names = ('loss_f_uv', 'loss_f_s', 'loss_IC', 'loss_SRC', 'loss_NB')

    def __init__(self, Collo, SRC, IC, UP, uv_layers, lb, ub, ExistModel=0, modelDir=None, variant='semi',
                 engine=None, verbose=True, dtype=None):
        self.variant = variant
        if dtype is None:
            dtype = np.float32 if variant == 'inf' else np.float64
        self._init_common(uv_layers, lb, ub, engine, verbose, dtype)
        self.E, self.mu, self.rho = 2.5, 0.25, 1.0                         # semi:35-37
        self.loss_rec = []                                                  # semi:39
        if variant == 'semi':
            self.term_weights = (5.0, 5.0, 2.0, 2.0, 2.0)
            self.bfgs_options = dict(maxiter=1000, maxfun=1000, maxcor=50, maxls=50, ftol=0.001 * np.finfo(float).eps)      # semi:133-137
        else:
            self.term_weights = (1.0, 1.0, 1.0, 1.0, 0.0)
            self.bfgs_options = dict(maxiter=10000, maxfun=10000, maxcor=50, maxls=50, ftol=0.001 * np.finfo(float).eps)    # inf:125-129
        A = lambda a: np.asarray(a, dtype=np.float64)
        self.Collo, self.SRC, self.IC, self.UP = map(A, (Collo, SRC, IC, UP))
        self.x_c, self.y_c, self.t_c = self.Collo[:, 0:1], self.Collo[:, 1:2], self.Collo[:, 2:3]
        if ExistModel == 0:
            self.uv_net.set_weights(*self.initialize_NN(self.uv_layers))
        else:
            self.uv_net.set_weights(*self.load_NN(modelDir, self.uv_layers))
        self._setup_terms()

    def _in_affine(self):
        if self.variant != 'inf':
            return (1, 1, 1), (0, 0, 0)
        lb = np.asarray(self.lb, np.float64).ravel(); ub = np.asarray(self.ub, np.float64).ravel()
        sc = 2.0 / (ub - lb)
        return tuple(sc), tuple(-2.0 * lb / (ub - lb) - 1.0)

    def _setup_terms(self):
        eng = self.engine
        eng.terms = []
        sc, sh = self._in_affine()
        w = self.term_weights
        com = dict(E=self.E, mu=self.mu, rho=self.rho, in_scale=sc, in_shift=sh)
        self._collo_term = eng.add_term('Collo', L.RES_F7, 4, self.Collo[:, :3], terms=(0, 1), weights=(w[0], w[1]), **com)
        eng.add_term('IC', L.RES_COLS, 1, self.IC[:, :3], cols=(0, 1, 2, 3), tgts=(-1,) * 4, terms=(2,) * 4, weights=(w[2],) * 4, **com)
        eng.add_term('SRC', L.RES_COLS, 1, self.SRC[:, :5], cols=(0, 1), tgts=(3, 4), terms=(3, 3), weights=(w[3],) * 2, **com)
        eng.add_term('UP', L.RES_COLS, 1, self.UP[:, :3], cols=(5, 6), tgts=(-1, -1), terms=(4, 4), weights=(w[4],) * 2, **com)

    def save_NN(self, fileDir):
        self._save(self.uv_net, fileDir, '')

    def callback(self, loss):
        self.count = self.count + 1
        self.loss_rec.append(loss)                                          # semi:287
        if self.verbose:
            print('{} th iterations, Loss: {}'.format(self.count, loss))

    def train(self, iter, learning_rate, batch_num, refeed=False):
        r = self._adam_loop(iter, learning_rate, self._chunks(batch_num), refeed=refeed)
        return r['loss_f_uv'], r['loss_f_s'], r['loss_IC'], r['loss_SRC'], r['loss']

    def train_bfgs(self, batch_num, options=None):
        for (a, b) in self._chunks(batch_num):
            self.engine.set_chunk(self._collo_term, a, b)
            self._bfgs(options or self.bfgs_options, self.callback)

    def getloss(self):
        N = self._collo_term.global_n
        self.engine.set_chunk(self._collo_term, 0, N)
        t = self._evaluate_terms()
        loss = self._total(t)
        if self.variant == 'inf':
            return loss, t[0], t[1], t[2], t[3], t[4]                      # inf:376
        for k, v in (('loss: ', loss), ('loss_f_uv: ', t[0]), ('loss_f_s: ', t[1]), ('loss_IC: ', t[2]), ('loss_SRC: ', t[3]), ('loss_NB: ', t[4])):
            print(k, v)                                                     # semi:387-392


class DeepElasticWave(_Base):
    """Confined elastic wave, soft BCs (ElasticWaveConfined/ElasticWave.py:21).  The dist/part networks the
    reference constructs are never used by its net_uv (conf:282-294) and are not built here."""

    formulation = L.RES_F7
    term_names = ('loss_f_uv', 'loss_f_s', 'loss_SRC', 'loss_IC', 'loss_FIX')
    term_weights = (5.0, 5.0, 1.0, 1.0, 1.0)                                # conf:156
    bfgs_options = dict(maxiter=100000, maxfun=100000, maxcor=50, maxls=50, ftol=1 * np.finfo(float).eps)   # conf:162-166

    def __init__(self, Collo, SRC, IC, FIXED, DIST, uv_layers, dist_layers, part_layers, lb, ub,
                 uvDir='', partDir='', distDir='', engine=None, verbose=True, dtype=np.float64):
        self._init_common(uv_layers, lb, ub, engine, verbose, dtype)
        self.E, self.mu, self.rho = 2.5, 0.25, 1.0
        A = lambda a: None if a is None else np.asarray(a, dtype=np.float64)
        self.Collo, self.SRC, self.IC, self.FIXED, self.DIST = map(A, (Collo, SRC, IC, FIXED, DIST))
        self.x_c, self.y_c, self.t_c = self.Collo[:, 0:1], self.Collo[:, 1:2], self.Collo[:, 2:3]
        self.dist_layers, self.part_layers = dist_layers, part_layers
        if uvDir != '':
            if self.verbose:
                print("Loading uv NN ...")
            self.uv_net.set_weights(*self.load_NN(uvDir, self.uv_layers))
        else:
            self.uv_net.set_weights(*self.initialize_NN(self.uv_layers))
        eng = self.engine
        w = self.term_weights
        com = dict(E=self.E, mu=self.mu, rho=self.rho)
        self._collo_term = eng.add_term('Collo', L.RES_F7, 4, self.Collo[:, :3], terms=(0, 1), weights=(w[0], w[1]), **com)
        eng.add_term('SRC', L.RES_COLS, 1, self.SRC[:, :5], cols=(0, 1), tgts=(3, 4), terms=(2, 2), weights=(w[2],) * 2, **com)
        eng.add_term('IC', L.RES_COLS, 1, self.IC[:, :3], cols=(0, 1, 2, 3), tgts=(-1,) * 4, terms=(3,) * 4, weights=(w[3],) * 4, **com)
        eng.add_term('FIXED', L.RES_COLS, 1, self.FIXED[:, :3], cols=(0, 1), tgts=(-1, -1), terms=(4, 4), weights=(w[4],) * 2, **com)

    def save_NN(self, fileDir, TYPE=''):
        if TYPE != 'UV':
            raise UnboundLocalError("local variable 'uv_weights' referenced before assignment")
        self._save(self.uv_net, fileDir, TYPE + ' ')

    def train(self, iter, learning_rate, batch_num, refeed=False):
        r = self._adam_loop(iter, learning_rate, self._chunks(batch_num), refeed=refeed)
        return r['loss_f_uv'], r['loss_f_s'], r['loss']

    def train_bfgs(self, batch_num, options=None):
        for (a, b) in self._chunks(batch_num):
            self.engine.set_chunk(self._collo_term, a, b)
            self._bfgs(options or self.bfgs_options, self.callback)

    def getloss(self):
        N = self._collo_term.global_n
        self.engine.set_chunk(self._collo_term, 0, N)
        t = self._evaluate_terms()
        for k, v in (('loss_f_uv', t[0]), ('loss_f_s', t[1]), ('loss_SRC', t[2]), ('loss_IC', t[3]), ('loss_FIX', t[4]), ('loss', self._total(t))):
            print(k, v)
