"""BASELINE config 5 shaped measurement (not a bench line): L-BFGS function evaluations per second -- loss + gradient on 200,000
collocation points.  driver 'scipy': parameters host->device and [grad | terms] device->host on every evaluation, SciPy L-BFGS-B
on one CPU core (what ScipyOptimizerInterface does, plate:240-247); driver 'gpu': device-resident L-BFGS (csrc/pe_lbfgs.cu).
usage: python tests/bench_lbfgs.py [N] [engine] [evals]"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import ref_torch as R
import pinn_elastodynamics_b200 as pe
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = sys.argv[2] if len(sys.argv) > 2 else 'tcf'
EV = int(sys.argv[3]) if len(sys.argv) > 3 else 300
Collo, HOLE = bench.make_workload(N)
layers = [3] + 5 * [50] + [5]
for driver in ('scipy', 'gpu'):
    m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=eng)
    Ws, bs = R.xavier_params(layers, seed=1111); m.uv_net.set_weights(Ws, bs)
    m.train_bfgs(dict(maxiter=5, maxfun=5, maxcor=50, maxls=50, ftol=1e-5 * np.finfo(float).eps, driver=driver))
    m.uv_net.set_weights(Ws, bs)
    m.count = 0
    seq = []
    m.callback = lambda l: seq.append(l)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = m.train_bfgs(dict(maxiter=EV, maxfun=EV, maxcor=50, maxls=50, ftol=1e-5 * np.finfo(float).eps, driver=driver))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = len(seq)
    print('%s/%s: %d L-BFGS evaluations (%d iterations) on %d pts in %.3f s -> %.2f ms/eval, %.3e point-evals/s, loss %.4e -> %.4e [%s]'
          % (eng, driver, n, res.nit, N, dt, 1e3 * dt / n, N * n / dt, seq[0], res.fun, str(res.message)[:48]), flush=True)
