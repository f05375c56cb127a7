"""Run on the GPU box (not a pytest file): A/B check of the fp16-pair tcgen05 engine ('tcf', csrc/pe_tcf.cu) on the bench workloads.
Every result is appended to gpurun_out/tcf_check.jsonl as soon as it exists (run under `timeout`).

    python tests/tcf_gpu_check.py [f5|f7|prof ...]      (default: f5 f7)      PE_CHECK_ENGINES=tc3s,tcf (default)

 f5   : BASELINE config 2 workload (50,000 collocation + 5,000 hole points, 5x50 net): terms / gradient vs the SIMT engine, ms per Adam step
 f7   : BASELINE config 3 shaped workload (half-space wave, 200,000 collocation points, [3]+5*[50]+[7])
 prof : per-phase cycle counters of the tcf kernel (CTA 0: epilogue thread 0 and the issuer)
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402  (make_workload only)
import pinn_elastodynamics_b200 as pe           # noqa: E402
from pinn_elastodynamics_b200 import _lib as L  # noqa: E402
from pinn_elastodynamics_b200.models import xavier_init_lists   # noqa: E402
from tests.util import per_layer_grad_err       # noqa: E402

lib = L.load()
ENGINES = [x for x in os.environ.get('PE_CHECK_ENGINES', 'tc3s,tcf').split(',') if x]
OUT = os.path.join(ROOT, 'gpurun_out', 'tcf_check.jsonl')
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    kw['t'] = round(time.time() - T0, 1)
    with open(OUT, 'a') as f:
        f.write(json.dumps(kw) + '\n')
    print(json.dumps(kw), flush=True)


def time_steps(eng, steps, flush):
    for _ in range(5):
        eng.adam_step(5e-4)
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.adam_step(5e-4); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = np.array([a.elapsed_time(b) for a, b in evs])
    return float(ts.mean()), float(np.median(ts)), float(ts.min())


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def compare(case, name, m, ref, layers, n_terms):
    m.engine.evaluate()
    t = m.engine.terms_host()[:n_terms]
    g = m.engine.grad_compact_host().astype(np.float64)
    row = dict(case=case, engine=name, terms=[float(x) for x in t], finite=bool(np.isfinite(g).all() and np.isfinite(t).all()))
    if ref is not None:
        row['terms_rel_vs_simt'] = [float((a - b) / b) if b else 0.0 for a, b in zip(t, ref[0])]      # signed: a bias shows as a common sign
        row['grad_rel_vs_simt'] = rel(g, ref[1])
        row['grad_block_rel_vs_simt'] = [float(e) for _, e in per_layer_grad_err(g, ref[1], layers)]
    emit(**row)
    return t, g


def f5(flush):
    layers = [3] + 5 * [50] + [5]
    Collo, HOLE = bench.make_workload(50000)
    Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111))
    ref = None
    for name in ['simt'] + ENGINES:
        m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=name)
        m.uv_net.set_weights(Ws, bs)
        r = compare('f5', name, m, ref, layers, 3)
        if name == 'simt':
            ref = r
        else:
            mean, med, mn = time_steps(m.engine, 300, flush)
            emit(case='f5', engine=name, ms_per_step_mean=mean, ms_median=med, ms_min=mn, mpts_per_s=50000 / mean / 1e3)
        del m


def f7_model(name, N=200000):
    rng = np.random.default_rng(1111)
    lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
    P = rng.uniform(lb, ub, (int(N * 1.1), 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 2.0][:N]
    IC = rng.uniform(lb, ub, (N // 12, 3)); IC[:, 2] = 0
    UP = rng.uniform(lb, ub, (N // 10, 3)); UP[:, 1] = 15
    th = rng.uniform(0, 2 * np.pi, N // 5); ts = rng.uniform(0, 16, N // 5)
    SRC = np.stack([2 * np.cos(th), 2 * np.sin(th), ts, 0.1 * np.cos(th) * np.sin(ts), 0.1 * np.sin(th) * np.sin(ts)], 1)
    layers = [3] + 5 * [50] + [7]
    Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111)); Ws[0] = Ws[0] * 0.1
    m = pe.DeepHPM(P, SRC, IC, UP, layers, lb, ub, verbose=False, engine=name)
    m.uv_net.set_weights(Ws, bs)
    return m, layers


def f7(flush):
    N = 200000
    ref = None
    for name in ['simt'] + ENGINES:
        m, layers = f7_model(name, N)
        r = compare('f7', name, m, ref, layers, 5)
        if name == 'simt':
            ref = r
        mean, med, mn = time_steps(m.engine, 60, flush)
        emit(case='f7', engine=name, collo_engine=int(m.engine.terms[0].engine), ms_per_step_mean=mean, ms_median=med, ms_min=mn, mpts_per_s=N / mean / 1e3,
             tflops_algorithmic=N * 252000 / mean / 1e9)
        del m


TCF_NAMES = ['E: tile start + layer1 fwd', 'E: fwd wait ACC[G0]', 'E: fwd epilogue G0 (tanh)', 'E: fwd wait ACC[G1]', 'E: fwd epilogue G1 (x,y)', 'E: fwd wait ACC[G2]',
             'E: fwd epilogue G2 (t,tt)', 'E: output layer wait', 'E: output/residual stage', 'E: rev wait ACC[G1] (after the slot reads of streams 0, 1)', 'E: rev stream passes (waits + work)',
             'E: rev wait dW MMAs', 'E: dW drain + publish', 'E: layer1 grad', 'E: rev wait SDONE[stream 0]', 'E: rev wait SDONE[stream 1]',
             'I: fwd wait image', 'I: fwd wait ACT[G0]', 'I: fwd issue G0', 'I: fwd wait ACT[G1]', 'I: fwd issue G1', 'I: fwd wait ACT[G2]', 'I: fwd issue G2',
             'I: fwd wait prev layer + TMA', 'I: rev wait image', 'I: rev wait ACT', 'I: adjoint issue', 'I2: wait ACT + dW loop (other waits + issue)', 'I: wait dW done + slots free', 'I: layer-1 gradient (wait + MMAs)', 'I2: wait DRAINED (after ACT)', 'I2: wait SFULL[stream 0] (bulk copy landed)']


def prof(flush):
    """per-phase cycles of the tcf kernel: epilogue thread 0 (E) and the MMA/TMA issuer (I) of CTA 0"""
    for case in ('f5', 'f7'):
        if case == 'f5':
            layers = [3] + 5 * [50] + [5]
            Collo, HOLE = bench.make_workload(50000)
            m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tcf')
            Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111))
            m.uv_net.set_weights(Ws, bs)
        else:
            m, layers = f7_model('tcf')
        for _ in range(3):
            m.engine.adam_step(5e-4)
        pr = torch.zeros(32, dtype=torch.int64, device='cuda')
        lib.pe_debug_set_tcf_profile(C.c_void_p(pr.data_ptr()))
        steps = 10
        for _ in range(steps):
            m.engine.adam_step(5e-4)
        torch.cuda.synchronize()
        lib.pe_debug_set_tcf_profile(None)
        p = pr.cpu().numpy().astype(np.float64) / steps
        emit(case='prof', workload=case, E_total=float(p[:16].sum()), I_total=float(p[16:].sum()), phases={n: float(v) for n, v in zip(TCF_NAMES, p) if n != '-'})
        del m


def fields(flush):
    """predict of a frame batch (2 M points, 5x50 plate net): SIMT fields kernel against the tensor-core forward sweep"""
    layers = [3] + 5 * [50] + [5]
    Collo, HOLE = bench.make_workload(2000)
    m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tcf')
    Ws, bs = xavier_init_lists(layers, np.random.default_rng(1111))
    m.uv_net.set_weights(Ws, bs)
    n = 2_000_000
    pts = torch.rand((n, 3), device='cuda') * torch.tensor([0.5, 0.5, 10.0], device='cuda')
    ref = None
    for name, eng in (('simt', L.ENGINE_SIMT_FP32), ('tcf', L.ENGINE_TCF)):
        lib.pe_debug_set_fields_engine(eng)
        for _ in range(3):
            out = m.uv_net.forward_fields(pts, m.formulation)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = m.uv_net.forward_fields(pts, m.formulation); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        o = out.cpu().numpy().astype(np.float64)
        row = dict(case='fields', engine=name, points=n, ms=float(np.median(ts)), mpts_per_s=n / float(np.median(ts)) / 1e3, hbm_gbs=n * 44 / float(np.median(ts)) / 1e6)
        if ref is None:
            ref = o
        else:
            row['max_rel_vs_simt'] = [float(np.abs(o[:, c] - ref[:, c]).max() / max(1.0, np.abs(ref[:, c]).max())) for c in range(8)]
        emit(**row)
    lib.pe_debug_set_fields_engine(-1)


if __name__ == '__main__':
    T0 = time.time()
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    what = sys.argv[1:] or ['f5', 'f7']
    emit(case='start', what=what, engines=ENGINES, device=torch.cuda.get_device_name(0))
    for w in what:
        {'f5': f5, 'f7': f7, 'prof': prof, 'fields': fields}[w](flush)
    emit(case='done')
