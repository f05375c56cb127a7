"""Run on the GPU box: device time of each piece of one GPU-resident L-BFGS iteration (BASELINE config 5: plate, 200,000 collocation points).
    python tests/bench_lbfgs_phases.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import pinn_elastodynamics_b200 as pe           # noqa: E402
from pinn_elastodynamics_b200 import _lib as L  # noqa: E402
from pinn_elastodynamics_b200.models import xavier_init_lists   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
layers = [3] + 5 * [50] + [5]
Collo, HOLE = bench.make_workload(N)
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='auto')
m.uv_net.set_weights(*xavier_init_lists(layers, np.random.default_rng(1111)))
eng, net, lib = m.engine, m.uv_net, m.engine.lib
n, M = net.Pp, 50
f32 = dict(dtype=torch.float32, device='cuda')
S = torch.randn(M, n, **f32) * 1e-3; Y = torch.randn(M, n, **f32) * 1e-3; state = torch.rand(M + 3, **f32) + 0.5
d = torch.zeros(n, **f32); xp = torch.zeros(n, **f32); gp = torch.zeros(n, **f32); res = torch.zeros(2, **f32)
P = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
eng.evaluate()
g = eng.out[:n]


def timed(name, fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print('%-44s device %8.1f us   host+device %8.1f us' % (name, 1e3 * e0.elapsed_time(e1) / reps, 1e6 * (time.perf_counter() - t0) / reps), flush=True)


timed('evaluate (residual kernels + slot reduction)', lambda: eng.evaluate())
timed('evaluate + terms read-back (D2H, sync)', lambda: (eng.evaluate(), eng.out[n:n + 8].cpu()))
timed('pe_lbfgs_direction, 50 pairs', lambda: lib.pe_lbfgs_direction(n, M, M, 0, P(g), P(S), P(Y), P(state), P(d), st()))
timed('pe_lbfgs_direction, 10 pairs', lambda: lib.pe_lbfgs_direction(n, M, 10, 9, P(g), P(S), P(Y), P(state), P(d), st()))
timed('pe_lbfgs_store_pair', lambda: lib.pe_lbfgs_store_pair(n, M, 0, P(net.params), P(xp), P(g), P(gp), P(S), P(Y), P(state), st()))
timed('pe_vec_dot_max', lambda: lib.pe_vec_dot_max(n, P(g), P(d), P(res), st()))
timed('pe_vec_axpy', lambda: lib.pe_vec_axpy(n, P(net.params), P(xp), 0.5, P(d), st()))
timed('scalars(): 8 terms + dot_max + .cpu()', lambda: (lib.pe_vec_dot_max(n, P(g), P(d), P(res), st()), res.cpu()))
