"""Run on the GPU box under ncu: a few Adam steps of the bench workload (BASELINE config 2: 50,000 collocation + 5,000 hole
points, 5x50 net) on one engine, nothing else in the process.     python tests/ncu_target.py [engine=tcf] [steps=6]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402  (make_workload only)
import numpy as np                              # noqa: E402
from pinn_elastodynamics_b200.models import xavier_init_lists   # noqa: E402
import pinn_elastodynamics_b200 as pe           # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else 'tcf'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
layers = [3] + 5 * [50] + [5]
Collo, HOLE = bench.make_workload(50000)
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=engine)
Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111))
m.uv_net.set_weights(Ws, bs)
for _ in range(steps):
    m.engine.adam_step(5e-4)
torch.cuda.synchronize()
print('done', engine, steps)
