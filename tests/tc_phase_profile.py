"""Run on the GPU box: per-phase cycle counters of the tcgen05 residual kernel (CTA 0, thread 0, clock64)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_torch as R
import pinn_elastodynamics_b200 as pe
from pinn_elastodynamics_b200 import _lib as L
lib = L.load()

layers = [3] + 5 * [50] + [5]
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
Collo = rng.uniform([0, 0, 0], [.5, .5, 10], (N, 3)); HOLE = rng.uniform([0, 0, 0], [.1, .1, 10], (N // 10, 3))
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tc3')
Ws, bs = R.xavier_params(layers, seed=1111); m.uv_net.set_weights(Ws, bs)
for _ in range(3): m.engine.adam_step(5e-4)
prof = torch.zeros(16, dtype=torch.int64, device='cuda')
lib.pe_debug_set_tc_profile(C.c_void_p(prof.data_ptr()))
steps = 10
for _ in range(steps): m.engine.adam_step(5e-4)
torch.cuda.synchronize()
lib.pe_debug_set_tc_profile(None)
p = prof.cpu().numpy().astype(np.float64) / steps
names = ['layer1 fwd', 'fwd img+sync', 'fwd mma issue', 'fwd mma wait', 'fwd resid stage', 'fwd epilogue (thread 0)', 'bwd img+sync', 'adj issue', 'convZ0+loadA0+adj wait',
         'dW convert+sync', 'dW issue', 'dW wait', 'dW drain', 'bwd epilogue', 'layer1 grad', 'tile start']
tiles = (N + 127) // 128 / 148
print('cycles per step (CTA 0), tiles per CTA ~%.2f' % tiles)
for n, v in zip(names, p):
    print('%-26s %10.0f  %5.1f%%' % (n, v, 100 * v / p.sum()))
print('total', p.sum(), 'cycles =', p.sum() / 1.965e3, 'us')
