"""Run on the GPU box (not a pytest file): measured errors of the CUDA path against the reference-generated golden vectors
(tests/golden/reference_tf1shim.npz), per engine -- the numbers behind the tolerances of tests/test_gpu_reference_golden.py.
    python tests/gpu_refgold_report.py  > gpurun_out/refgold_report.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pinn_elastodynamics_b200 as pe                                          # noqa: E402
from tests.util import layers_of, per_layer_grad_err, rel_err, unpack_golden   # noqa: E402

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_tf1shim.npz'))


def uv(kind):
    Ws, bs = unpack_golden(G, f'{kind}_uv')
    return [np.asarray(w, np.float64) for w in Ws], [np.asarray(b, np.float64) for b in bs]


def relv(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))


for engine in (sys.argv[1:] or ('simt', 'tc3s')):
    for composite in (False, True):
        S = {k: G['plate_' + k] for k in ('Collo', 'HOLE', 'IC', 'LF', 'RT', 'UP', 'LW', 'DIST', 'lb', 'ub')}
        Ws, bs = uv('plate')
        layers = layers_of(Ws)
        if composite:
            di, pa = unpack_golden(G, 'plate_dist'), unpack_golden(G, 'plate_part')
            m = pe.PINN(S['Collo'], S['HOLE'], S['IC'], S['LF'], S['RT'], S['UP'], S['LW'], S['DIST'], layers, layers_of(di[0]), layers_of(pa[0]),
                        S['lb'], S['ub'], verbose=False, engine=engine)
            m.dist_net.set_weights(*di); m.part_net.set_weights(*pa); m.refresh_composite()
            tag = 'comp'
        else:
            m = pe.PINN(S['Collo'], S['HOLE'], None, None, None, None, None, None, layers, None, None, S['lb'], S['ub'], verbose=False, engine=engine)
            tag = 'plain'
        m.uv_net.set_weights(Ws, bs)
        m.engine.evaluate()
        t = m.engine.terms_host()
        ref = G[f'plate_{tag}_terms']
        row = dict(case='plate_' + tag, engine=engine, terms_rel=relv(t[:3], ref[:3]), total_rel=relv([m._total(t)], ref[3:4]),
                   grad_block_rel=max(e for _, e in per_layer_grad_err(m.engine.grad_compact_host(), G[f'plate_{tag}_grad'], layers)))
        steps = 20 if tag == 'plain' else 8
        out = m.train(steps, 5e-4)
        C = G[f'plate_{tag}_adam']
        row['adam_terms_rel'] = [relv(out[i], C[:, i]) for i in range(3)]
        row['adam_total_rel'] = relv(out[3], C[:, 3])
        row['adam_total_rel_per_step'] = [float(abs(a - b) / abs(b)) for a, b in zip(out[3], C[:, 3])]
        if tag == 'plain':
            row['params_after_rel'] = rel_err(m.uv_net.get_flat(), G['plate_plain_params_after_adam'])
        print(json.dumps(row), flush=True)
    for kind in ('semi', 'inf', 'conf'):
        S = {k: G[f'{kind}_{k}'] for k in (('Collo', 'SRC', 'IC', 'FIXED', 'lb', 'ub') if kind == 'conf' else ('Collo', 'SRC', 'IC', 'UP', 'lb', 'ub'))}
        Ws, bs = uv(kind)
        layers = layers_of(Ws)
        if kind == 'conf':
            m = pe.DeepElasticWave(S['Collo'], S['SRC'], S['IC'], S['FIXED'], None, layers, None, None, S['lb'], S['ub'], verbose=False, engine=engine)
        else:
            m = pe.DeepHPM(S['Collo'], S['SRC'], S['IC'], S['UP'], layers, S['lb'], S['ub'], variant=kind, verbose=False, engine=engine)
        m.uv_net.set_weights(Ws, bs)
        m.engine.evaluate()
        t = m.engine.terms_host()
        ref = G[f'{kind}_terms']
        n = len(ref) - 1
        row = dict(case=kind, engine=engine, terms_rel=relv(t[:n], ref[:n]), total_rel=relv([m._total(t)], ref[n:n + 1]),
                   grad_block_rel=max(e for _, e in per_layer_grad_err(m.engine.grad_compact_host(), G[f'{kind}_grad'], layers)))
        out = m.train(6, 1e-3, 2)
        C = G[f'{kind}_adam_b2']
        row['adam_total_rel'] = relv(out[-1], C[:, -1])
        row['adam_f_uv_rel'] = relv(out[0], C[:, 0])
        row['params_after_rel'] = rel_err(m.uv_net.get_flat(), G[f'{kind}_params_after_adam'])
        print(json.dumps(row), flush=True)
