"""BASELINE config 5 shaped measurement (not a bench line): L-BFGS function evaluations per second -- loss + packed gradient on
200,000 collocation points, parameters host->device and [grad | terms] device->host on every evaluation, SciPy L-BFGS-B driver."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import ref_torch as R
import pinn_elastodynamics_b200 as pe
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = sys.argv[2] if len(sys.argv) > 2 else 'tc3'
Collo, HOLE = bench.make_workload(N)
layers = [3] + 5 * [50] + [5]
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=eng)
Ws, bs = R.xavier_params(layers, seed=1111); m.uv_net.set_weights(Ws, bs)
m.train_bfgs(dict(maxiter=5, maxfun=5, maxcor=50, maxls=50, ftol=1e-5 * np.finfo(float).eps))
m.count = 0
torch.cuda.synchronize(); t0 = time.perf_counter()
res = m.train_bfgs(dict(maxiter=60, maxfun=60, maxcor=50, maxls=50, ftol=1e-5 * np.finfo(float).eps))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('%s: %d L-BFGS evaluations on %d pts in %.3f s -> %.2f ms/eval, %.3e point-evals/s, loss %.4e -> %.4e' % (eng, m.count, N, dt, 1e3 * dt / m.count, N * m.count / dt, res.fun if False else 0, res.fun))
