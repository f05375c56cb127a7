"""Not a test: prints CUDA-vs-oracle error tables (run on the GPU box, output kept under gpurun_out/)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_torch as R, jet_numpy as J
from tests.util import *
import pinn_elastodynamics_b200 as pe
from pinn_elastodynamics_b200.engine import Network

G = lambda n: np.load(os.path.join(os.path.dirname(__file__), 'golden', n))
print(torch.cuda.get_device_name(0))
g = G('synthetic_5x50.npz')
layers = [3] + 5 * [50] + [5]
Ws, bs = R.xavier_params(layers, seed=1111)
bs = random_biases(bs, 3)
# 1. forward jets
net = Network(layers); net.set_weights(Ws, bs)
pts = torch.from_numpy(g['f5_collo'].astype(np.float32)).cuda()
for K in (1, 2, 4, 5):
    out = net.forward_jets(pts, K).cpu().numpy()
    Y, _ = J.forward_jets(g['f5_collo'], Ws, bs, K)
    print('jets K=%d' % K, [rel_err(out[:, k], Y[k]) for k in range(K)])
# 2. loss + grad
orc = R.Oracle('plate', Ws, bs)
sets = {'Collo': g['f5_collo'], 'HOLE': g['f5_hole']}
T, loss, gref = orc.loss_and_grad(sets)
m = pe.PINN(sets['Collo'], sets['HOLE'], None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=(sys.argv[1] if len(sys.argv) > 1 else 'simt'))
m.uv_net.set_weights(Ws, bs)
m.engine.evaluate(); torch.cuda.synchronize()
t = m.engine.terms_host(); gc = m.engine.grad_compact_host()
print('terms cuda', t[:3], 'ref', [T[k] for k in ('loss_f_uv', 'loss_f_s', 'loss_HOLE')])
print('grad per-layer rel err', per_layer_grad_err(gc, gref, layers))
# only collocation term / only hole term
for name in ('Collo', 'HOLE'):
    for tm in m.engine.terms: tm.enabled = (tm.name == name)
    m.engine._built = False
    m.engine.evaluate(); gc = m.engine.grad_compact_host(); tt = m.engine.terms_host()
    Tt, _ = orc.loss_terms(sets)
    l = 10 * (Tt['loss_f_uv'] + Tt['loss_f_s']) if name == 'Collo' else 10 * Tt['loss_HOLE']
    gs = torch.autograd.grad(l, orc.params())
    gr = np.concatenate([x.numpy().ravel() for x in gs])
    print(name, 'terms', tt[:3], 'grad err', per_layer_grad_err(gc, gr, layers))
for tm in m.engine.terms: tm.enabled = True
m.engine._built = False
# 3. timing at 50k points
rng = np.random.default_rng(0)
N = 50000
Collo = rng.uniform([0, 0, 0], [.5, .5, 10], (N, 3)); HOLE = rng.uniform([0, 0, 0], [.1, .1, 10], (N // 10, 3))
import sys as _s
ENG = _s.argv[1] if len(_s.argv) > 1 else 'simt'
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=ENG)
m.uv_net.set_weights(Ws, bs)
for _ in range(3): m.engine.adam_step(5e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): m.engine.adam_step(5e-4)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(ENG, '50k pts: %.3f ms/step  -> %.1f Mpts/s, %.2f TFLOP/s algorithmic' % (ms, N / ms / 1e3, N * 312000 / ms / 1e9))
