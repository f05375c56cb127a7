"""BASELINE config 3 shaped measurement (not a bench line): elastic wave in a half-space (plane strain, F7, K = 4 jet streams),
[3]+5*[50]+[7] net, 200,000 collocation points + IC/SRC/UP sets, Adam, fp32 SIMT engine (the tcgen05 engine covers F5 only)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_torch as R
import pinn_elastodynamics_b200 as pe
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(1111)
lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
P = rng.uniform(lb, ub, (int(N * 1.1), 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 2.0][:N]
IC = rng.uniform(lb, ub, (N // 12, 3)); IC[:, 2] = 0
UP = rng.uniform(lb, ub, (N // 10, 3)); UP[:, 1] = 15
th = rng.uniform(0, 2 * np.pi, N // 5); ts = rng.uniform(0, 16, N // 5)
SRC = np.stack([2 * np.cos(th), 2 * np.sin(th), ts, 0.1 * np.cos(th) * np.sin(ts), 0.1 * np.sin(th) * np.sin(ts)], 1)
layers = [3] + 5 * [50] + [7]
m = pe.DeepHPM(P, SRC, IC, UP, layers, lb, ub, verbose=False)
Ws, bs = R.xavier_params(layers, seed=1111); Ws[0] = Ws[0] * 0.1; m.uv_net.set_weights(Ws, bs)
for _ in range(3): m.engine.adam_step(5e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): m.engine.adam_step(5e-4)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print('F7 semi, %d collocation pts (+%d IC, %d SRC, %d UP): %.3f ms/step -> %.2f Mpts/s (252 kflop/pt -> %.1f TFLOP/s)' % (N, len(IC), len(SRC), len(UP), ms, N / ms / 1e3, N * 252000 / ms / 1e9))
