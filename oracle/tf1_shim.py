"""Minimal TensorFlow-1 graph-mode emulator over torch (CPU)  --  TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's OWN class files (PlateHoleQuarter/train/train.py, ElasticWave*/ElasticWave.py), unmodified and
read in place from /root/reference, inside this container (which has no tensorflow), so that golden vectors can be generated
from the reference's own graph-construction code instead of from our restatement of it
(tests/golden/make_reference_golden.py -> tests/golden/reference_*.npz).  What is emulated is only the behaviour of the TF
primitives those files call; the algorithm (net_uv, net_e, net_f_sig, net_t, the loss assembly, the train loops, the
ScipyOptimizerInterface call) is the reference's code running as written.

Emulated surface (everything the four scripts touch; third-party semantics, TensorFlow 1.10 documentation):
  tf.placeholder, tf.Variable, tf.zeros, tf.truncated_normal, tf.concat, tf.matmul, tf.add, tf.multiply, tf.tanh, tf.square,
  tf.reduce_mean, tf.gradients, tensor slicing and arithmetic, tf.Session / ConfigProto / global_variables_initializer /
  set_random_seed / device, tf.train.AdamOptimizer(...).minimize, tf.contrib.opt.ScipyOptimizerInterface(...).minimize.
  * graph mode: ops build nodes; Session.run evaluates the fetched nodes with the fed placeholder values, one evaluation
    per node per run
  * tf.gradients(y, x)[0] = d sum(y) / d x, differentiable again (torch.autograd.grad(y, x, ones, create_graph=True))
  * AdamOptimizer: m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2; lr_t = lr sqrt(1-b2^t)/(1-b1^t);
    var <- var - lr_t m / (sqrt(v) + eps)          (beta1 .9, beta2 .999, epsilon 1e-8; slots persist across runs)
  * ScipyOptimizerInterface: variables packed in var_list order into one float64 vector, scipy.optimize.minimize(fun, x0,
    jac=True, method=..., options=...), loss_callback(*fetches) on every function evaluation, result assigned back
  * truncated_normal: N(0, stddev) with |x| > 2 stddev redrawn (numpy RandomState seeded by set_random_seed: the values are
    NOT TensorFlow's random stream -- golden runs load their weights from pickles through the reference's load_NN)

install() puts this module into sys.modules as `tensorflow` together with stubs for `pyDOE`, `matplotlib` and
`mpl_toolkits` (imported at the top of the reference files, used only by their plotting / __main__ code).
Nothing under pinn_elastodynamics_b200/ imports this file.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
_RNG = np.random.RandomState(0)
_ALL_VARIABLES = []


def set_random_seed(seed):
    global _RNG
    _RNG = np.random.RandomState(seed)


def reset_default_graph():
    """forget the variables created so far (a later global_variables_initializer() must not re-initialise another model's)"""
    del _ALL_VARIABLES[:]


def _np_dtype(dt):
    return np.float32 if dt == torch.float32 else np.float64


class Tensor:
    """A graph node.  `fn(*input_values)` computes its torch value from the values of `inputs`."""
    __array_ufunc__ = None          # numpy operands defer to the reflected operators below
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), dtype=None, name=None):
        self.fn, self.inputs, self.name = fn, tuple(inputs), name
        self.dtype = dtype if dtype is not None else next((i.dtype for i in self.inputs if isinstance(i, Tensor)), float64)

    # ---- arithmetic
    def _bin(self, other, f, reflected=False):
        o = other if isinstance(other, Tensor) else constant(other, self.dtype)
        a, b = (o, self) if reflected else (self, o)
        return Tensor(f, (a, b), self.dtype)

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Tensor(torch.neg, (self,), self.dtype)
    def __getitem__(self, idx): return Tensor(lambda v: v[idx], (self,), self.dtype)
    __hash__ = object.__hash__


class Operation:
    """A node run for its side effect (initialiser, Adam step); Session.run returns None for it."""

    def __init__(self, run):
        self.run = run


def constant(value, dtype=float64):
    v = torch.as_tensor(np.asarray(value, dtype=_np_dtype(dtype)))
    return Tensor(lambda: v, (), dtype)


class _Placeholder(Tensor):
    def __init__(self, dtype, shape=None, name=None):
        super().__init__(None, (), dtype, name)
        self.shape = shape


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape, name)


class Variable(Tensor):
    def __init__(self, initial_value, dtype=None, name=None, trainable=True):
        if dtype is not None:
            dt = dtype
        elif isinstance(initial_value, Tensor):
            dt = initial_value.dtype
        else:                                   # like TF: the dtype of the initial value (a float32 pickle gives float32 variables)
            dt = float32 if np.asarray(initial_value).dtype == np.float32 else float64
        super().__init__(None, (), dt, name)
        self.initial_value = initial_value
        self.value = None                       # torch leaf tensor once initialised
        _ALL_VARIABLES.append(self)

    def initialise(self, session):
        iv = self.initial_value
        arr = session._eval_plain(iv) if isinstance(iv, Tensor) else np.asarray(iv)
        self.value = torch.tensor(np.asarray(arr, dtype=_np_dtype(self.dtype))).requires_grad_(True)

    def assign_numpy(self, arr):
        with torch.no_grad():
            self.value.copy_(torch.as_tensor(np.asarray(arr, dtype=_np_dtype(self.dtype))).reshape(self.value.shape))


def zeros(shape, dtype=float32, name=None):
    return Tensor(lambda: torch.zeros(*shape, dtype=dtype), (), dtype)


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    def draw():
        x = _RNG.normal(0.0, 1.0, size=shape)
        bad = np.abs(x) > 2.0
        while bad.any():
            x[bad] = _RNG.normal(0.0, 1.0, size=int(bad.sum()))
            bad = np.abs(x) > 2.0
        return torch.as_tensor((mean + stddev * x).astype(_np_dtype(dtype)))
    return Tensor(draw, (), dtype)


def concat(values, axis, name=None):
    return Tensor(lambda *v: torch.cat(v, dim=axis), values)


def matmul(a, b, name=None): return Tensor(torch.matmul, (a, b))
def add(a, b, name=None): return a + b if isinstance(a, Tensor) else b + a
def multiply(a, b, name=None): return a * b if isinstance(a, Tensor) else b * a
def tanh(x, name=None): return Tensor(torch.tanh, (x,))
def square(x, name=None): return Tensor(torch.square, (x,))
def reduce_mean(x, axis=None, name=None): return Tensor(torch.mean if axis is None else (lambda v: torch.mean(v, dim=axis)), (x,))


class _Grad(Tensor):
    """d sum(y) / d x; evaluated by the session because it needs the run's value objects of y and x."""

    def __init__(self, y, x):
        super().__init__(None, (y, x), x.dtype)


def gradients(ys, xs, name=None):
    ys = ys if isinstance(ys, (list, tuple)) else [ys]
    xs_l = xs if isinstance(xs, (list, tuple)) else [xs]
    out = []
    for x in xs_l:
        terms = [_Grad(y, x) for y in ys]
        g = terms[0]
        for t in terms[1:]:
            g = g + t
        out.append(g)
    return out


def ConfigProto(*a, **k):
    return types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False), **k)


@contextlib.contextmanager
def device(name):
    yield


class Session:
    def __init__(self, config=None, graph=None):
        self.config = config

    # ---- evaluation of one run
    def _value(self, node, cache, feed):
        k = id(node)
        if k in cache:
            return cache[k]
        if isinstance(node, _Placeholder):
            if node not in feed:
                raise ValueError('placeholder %r was not fed' % (node.name,))
            v = torch.tensor(np.asarray(feed[node], dtype=_np_dtype(node.dtype)))
            if v.dtype.is_floating_point:
                v.requires_grad_(True)
        elif isinstance(node, Variable):
            if node.value is None:
                raise RuntimeError('uninitialised variable (run tf.global_variables_initializer() first)')
            v = node.value
        elif isinstance(node, _Grad):
            y = self._value(node.inputs[0], cache, feed)
            x = self._value(node.inputs[1], cache, feed)
            g = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True, allow_unused=True)[0]
            v = torch.zeros_like(x) if g is None else g
        else:
            v = node.fn(*[self._value(i, cache, feed) for i in node.inputs])
        cache[k] = v
        return v

    def _eval_plain(self, node):
        return self._value(node, {}, {}).detach().numpy()

    def run(self, fetches, feed_dict=None):
        feed = feed_dict or {}
        cache = {}

        def one(f):
            if isinstance(f, (list, tuple)):
                return [one(g) for g in f]
            if isinstance(f, Operation):
                f.run(self, cache, feed)
                return None
            v = self._value(f, cache, feed).detach().numpy().copy()
            return v if v.ndim else v.dtype.type(v)
        return one(fetches)

    def close(self):
        pass


def global_variables_initializer():
    todo = list(_ALL_VARIABLES)

    def run(session, cache, feed):
        for v in todo:
            v.initialise(session)
    return Operation(run)


def _grads_wrt(session, loss, var_list, cache, feed):
    lv = session._value(loss, cache, feed)
    gs = torch.autograd.grad(lv, [v.value for v in var_list], allow_unused=True)
    return lv, [torch.zeros_like(v.value) if g is None else g for g, v in zip(gs, var_list)]


class _AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, use_locking=False, name='Adam'):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss, var_list=None, global_step=None, name=None):
        var_list = list(var_list if var_list is not None else _ALL_VARIABLES)
        slots = {'t': 0, 'm': None, 'v': None}

        def run(session, cache, feed):
            if slots['m'] is None:
                slots['m'] = [torch.zeros_like(v.value) for v in var_list]
                slots['v'] = [torch.zeros_like(v.value) for v in var_list]
            _, gs = _grads_wrt(session, loss, var_list, cache, feed)
            lr = float(session._value(self.lr, cache, feed).detach()) if isinstance(self.lr, Tensor) else float(self.lr)
            slots['t'] += 1
            t = slots['t']
            lr_t = lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
            with torch.no_grad():
                for var, g, m, v in zip(var_list, gs, slots['m'], slots['v']):
                    m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                    v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                    var.value.sub_((lr_t * m / (torch.sqrt(v) + self.eps)).to(var.value.dtype))
        return Operation(run)


class _ScipyOptimizerInterface:
    def __init__(self, loss, var_list=None, equalities=None, inequalities=None, var_to_bounds=None, **optimizer_kwargs):
        self.loss = loss
        self.var_list = list(var_list if var_list is not None else _ALL_VARIABLES)
        self.optimizer_kwargs = optimizer_kwargs

    def minimize(self, session=None, feed_dict=None, fetches=None, step_callback=None, loss_callback=None, **run_kwargs):
        import scipy.optimize
        fetches = list(fetches or [])
        shapes = [tuple(v.value.shape) for v in self.var_list]
        sizes = [int(np.prod(s)) for s in shapes]

        def assign(x):
            o = 0
            for v, s, n in zip(self.var_list, shapes, sizes):
                v.assign_numpy(np.asarray(x[o:o + n]).reshape(s)); o += n

        def loss_grad(x):
            assign(x)
            cache = {}
            lv, gs = _grads_wrt(session, self.loss, self.var_list, cache, feed_dict or {})
            if loss_callback is not None:
                vals = [session._value(f, cache, feed_dict or {}).detach().numpy() for f in fetches]
                loss_callback(*[v if v.ndim else v.dtype.type(v) for v in vals])
            g = np.concatenate([gi.detach().numpy().astype(np.float64).ravel() for gi in gs])
            return float(lv), g

        x0 = np.concatenate([v.value.detach().numpy().astype(np.float64).ravel() for v in self.var_list])
        kw = dict(self.optimizer_kwargs)
        method = kw.pop('method', 'L-BFGS-B')
        options = kw.pop('options', None)
        res = scipy.optimize.minimize(loss_grad, x0, jac=True, method=method, callback=step_callback, options=options, **kw)
        assign(res.x)
        self.last_result = res
        return res


train = types.SimpleNamespace(AdamOptimizer=_AdamOptimizer)
contrib = types.SimpleNamespace(opt=types.SimpleNamespace(ScipyOptimizerInterface=_ScipyOptimizerInterface))


def _lhs(n, samples=None, criterion=None, iterations=None):
    """pyDOE.lhs stand-in (random latin hypercube); the reference calls it from __main__ only."""
    samples = samples or n
    u = np.random.rand(samples, n)
    out = np.zeros_like(u)
    for j in range(n):
        out[:, j] = (np.random.permutation(samples) + u[:, j]) / samples
    return out


class _Dummy:
    """stands in for any plotting object: every attribute and every call returns another dummy"""

    def __call__(self, *a, **k): return _Dummy()
    def __getattr__(self, name): return _Dummy()
    def __iter__(self): return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy()


def install():
    """Register this module as `tensorflow` and stub the plotting / sampling imports of the reference files."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    if 'pyDOE' not in sys.modules:
        m = types.ModuleType('pyDOE'); m.lhs = _lhs; sys.modules['pyDOE'] = m
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _StubModule(name)
    for parent, child in (('matplotlib', 'pyplot'), ('matplotlib', 'colors'), ('mpl_toolkits', 'mplot3d')):
        p, c = sys.modules.get(parent), sys.modules.get(parent + '.' + child)
        if isinstance(p, _StubModule) and c is not None:
            p.__dict__[child] = c
    return me


def load_reference_module(path, name):
    """Import one reference script (its __main__ block is guarded) under a private module name, with the shim installed."""
    import importlib.util
    install()
    del _ALL_VARIABLES[:]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
