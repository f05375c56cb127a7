"""Run on the GPU box under compute-sanitizer: two Adam steps of a small plate workload (2,000 collocation + 300 hole points, 5x50 net)
on one engine, nothing else in the process.     compute-sanitizer --tool memcheck|racecheck python tests/sanitize_target.py [engine=tcf]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402  (make_workload only)
from pinn_elastodynamics_b200.models import xavier_init_lists   # noqa: E402
import pinn_elastodynamics_b200 as pe           # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else 'tcf'
layers = [3] + 5 * [50] + [5]
Collo, HOLE = bench.make_workload(2000)
m = pe.PINN(Collo, HOLE[:300], None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=engine)
Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111))
m.uv_net.set_weights(Ws, bs)
for _ in range(2):
    m.engine.adam_step(5e-4)
torch.cuda.synchronize()
print('done', engine, [float(x) for x in m.engine.terms_host()[:3]])
