"""Run on the GPU box (not a pytest file): A/B check of the second-generation tcgen05 engine ('tc3p', csrc/pe_tcp.cu) against the
round-1 engines on the bench workloads.  Every result is appended to gpurun_out/tcp_check.jsonl as soon as it exists, so a
hang in a later variant (run this under `timeout`) does not lose the earlier ones.

    python tests/tcp_gpu_check.py [f5|f7|prof ...]      (default: all)

 f5   : BASELINE config 2 workload (50,000 collocation + 5,000 hole points, 5x50 net): terms / gradient of tc3p (serial and
        pipelined weight-gradient phase) vs tc3 and simt, and ms per Adam step of each (CUDA events, L2 flushed between steps)
 f7   : BASELINE config 3 shaped workload (half-space wave, 200,000 collocation points, [3]+5*[50]+[7]): simt vs tc3p
 prof : per-phase cycle counters of the pipelined kernel (CTA 0 / thread 0)
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
from oracle import ref_torch as R               # noqa: E402  (Xavier arrays only)
import pinn_elastodynamics_b200 as pe           # noqa: E402
from pinn_elastodynamics_b200 import _lib as L  # noqa: E402

lib = L.load()
ONLY = [x for x in os.environ.get('PE_CHECK_ONLY', '').split(',') if x]      # e.g. PE_CHECK_ONLY=tc3s
OUT = os.path.join(ROOT, 'gpurun_out', 'tcp_check.jsonl')
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
    tot = 0.0
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


def f5(flush):
    layers = [3] + 5 * [50] + [5]
    Collo, HOLE = bench.make_workload(50000)
    Ws, bs = R.xavier_params(layers, seed=1111)
    ref = {}
    for name, eng_name, pipe in (('simt', 'simt', 1), ('tc3', 'tc3', 1), ('tc3p_serial', 'tc3p', 0), ('tc3p_pipe', 'tc3p', 1), ('tc3s', 'tc3s', 1), ('tc4', 'tc4', 1)):
        if (ONLY and name not in ONLY and name not in ('simt', 'tc3')) or (name == 'tc4' and 'tc4' not in ONLY):
            continue
        lib.pe_debug_set_tcp_pipeline(pipe)
        m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=eng_name)
        m.uv_net.set_weights(Ws, bs)
        m.engine.evaluate()
        t = m.engine.terms_host()[:3]
        g = m.engine.grad_compact_host().astype(np.float64)
        ref[name] = (t, g)
        row = dict(case='f5', engine=name, terms=[float(x) for x in t])
        if name != 'simt':
            row['terms_rel_vs_simt'] = rel(t, ref['simt'][0]); row['grad_rel_vs_simt'] = rel(g, ref['simt'][1])
        if name.startswith('tc3p') or name in ('tc3s', 'tc4'):
            row['grad_rel_vs_tc3'] = rel(g, ref['tc3'][1]); row['bit_equal_tc3'] = bool(np.array_equal(g, ref['tc3'][1]))
        emit(**row)
        if name != 'simt':
            mean, med, mn = time_steps(m.engine, 300, flush)
            emit(case='f5', engine=name, ms_per_step_mean=mean, ms_median=med, ms_min=mn, mpts_per_s=50000 / mean / 1e3)
        del m
    lib.pe_debug_set_tcp_pipeline(1)


def f7(flush):
    N = 200000
    rng = np.random.default_rng(1111)
    lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
    P = rng.uniform(lb, ub, (int(N * 1.1), 3)); P = P[np.hypot(P[:, 0], P[:, 1]) > 2.0][:N]
    IC = rng.uniform(lb, ub, (N // 12, 3)); IC[:, 2] = 0
    UP = rng.uniform(lb, ub, (N // 10, 3)); UP[:, 1] = 15
    th = rng.uniform(0, 2 * np.pi, N // 5); ts = rng.uniform(0, 16, N // 5)
    SRC = np.stack([2 * np.cos(th), 2 * np.sin(th), ts, 0.1 * np.cos(th) * np.sin(ts), 0.1 * np.sin(th) * np.sin(ts)], 1)
    layers = [3] + 5 * [50] + [7]
    Ws, bs = R.xavier_params(layers, seed=1111); Ws[0] = Ws[0] * 0.1
    ref = {}
    for name, eng_name, pipe in (('simt', 'simt', 1), ('tc3p_serial', 'tc3p', 0), ('tc3p_pipe', 'tc3p', 1), ('tc3s', 'tc3s', 1), ('tc4', 'tc4', 1)):
        if (ONLY and name not in ONLY and name != 'simt') or (name == 'tc4' and 'tc4' not in ONLY):
            continue
        lib.pe_debug_set_tcp_pipeline(pipe)
        m = pe.DeepHPM(P, SRC, IC, UP, layers, lb, ub, verbose=False, engine=eng_name)
        m.uv_net.set_weights(Ws, bs)
        m.engine.evaluate()
        t = m.engine.terms_host()[:5]
        g = m.engine.grad_compact_host().astype(np.float64)
        ref[name] = (t, g)
        row = dict(case='f7', engine=name, terms=[float(x) for x in t], collo_engine=int(m.engine.terms[0].engine))
        if name != 'simt':
            row['terms_rel_vs_simt'] = rel(t, ref['simt'][0]); row['grad_rel_vs_simt'] = rel(g, ref['simt'][1])
        emit(**row)
        mean, med, mn = time_steps(m.engine, 60, flush)
        emit(case='f7', engine=name, ms_per_step_mean=mean, ms_median=med, ms_min=mn, mpts_per_s=N / mean / 1e3,
             tflops_algorithmic=N * 252000 / mean / 1e9)
        del m
    lib.pe_debug_set_tcp_pipeline(1)


def prof(flush):
    layers = [3] + 5 * [50] + [5]
    Collo, HOLE = bench.make_workload(50000)
    Ws, bs = R.xavier_params(layers, seed=1111)
    names = ['layer1 fwd', 'fwd img+sync', 'fwd mma issue', 'fwd mma wait', 'fwd resid stage', 'fwd epilogue', 'bwd epilogue(prev)+img+sync', 'adj issue',
             'convZ0+ldA0+adj wait', 'dW convert+sync (serial only)', 'dW issue (pipe: whole pipeline, issuer view)', 'dW wait', 'dW drain', 'bwd epilogue (last layer)',
             'layer1 grad (last tile)', 'layer1 grad + tile start']
    for pipe in (0, 1):
        lib.pe_debug_set_tcp_pipeline(pipe)
        m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tc3p')
        m.uv_net.set_weights(Ws, bs)
        for _ in range(3):
            m.engine.adam_step(5e-4)
        pr = torch.zeros(16, dtype=torch.int64, device='cuda')
        lib.pe_debug_set_tcp_profile(C.c_void_p(pr.data_ptr()))
        steps = 10
        for _ in range(steps):
            m.engine.adam_step(5e-4)
        torch.cuda.synchronize()
        lib.pe_debug_set_tcp_profile(None)
        p = pr.cpu().numpy().astype(np.float64) / steps
        emit(case='prof', pipe=pipe, total_cycles=float(p.sum()), phases={n: float(v) for n, v in zip(names, p)})
        del m
    lib.pe_debug_set_tcp_pipeline(1)


TCS_NAMES = ['E: tile start + layer1 fwd', 'E: fwd wait ACC[G0]', 'E: fwd epilogue G0 (tanh)', 'E: fwd wait ACC[G1]', 'E: fwd epilogue G1 (x,y)', 'E: fwd wait ACC[G2]',
             'E: fwd epilogue G2 (t,tt)', 'E: output layer wait', 'E: output/residual stage', 'E: rev sync + early conversions', 'E: rev wait adjoint MMAs',
             'E: dW converter loop', 'E: wait dW MMAs', 'E: dW drain', 'E: bwd epilogue', 'E: layer1 grad',
             'I: fwd wait image', 'I: fwd wait ACT[G0]', 'I: fwd issue G0', 'I: fwd wait ACT[G1]', 'I: fwd issue G1', 'I: fwd wait ACT[G2]', 'I: fwd issue G2',
             'I: fwd wait prev layer + TMA', 'I: rev wait image', 'I: rev wait ACT', 'I: adjoint issue', 'I: dW loop (waits + issue)', 'I: wait dW done', '-', '-', '-']


def profs(flush):
    """per-phase cycles of the stream-pipelined kernel: epilogue thread 0 (E) and the MMA/TMA issuer (I) of CTA 0"""
    for case in ('f5', 'f7'):
        if case == 'f5':
            layers = [3] + 5 * [50] + [5]
            Collo, HOLE = bench.make_workload(50000)
            m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine='tc3s')
        else:
            N = 200000
            rng = np.random.default_rng(1111)
            lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
            P = rng.uniform(lb, ub, (N, 3))
            IC = rng.uniform(lb, ub, (N // 12, 3)); UP = rng.uniform(lb, ub, (N // 10, 3))
            SRC = np.concatenate([rng.uniform(lb, ub, (N // 5, 3)), rng.standard_normal((N // 5, 2)) * 0.1], 1)
            layers = [3] + 5 * [50] + [7]
            m = pe.DeepHPM(P, SRC, IC, UP, layers, lb, ub, verbose=False, engine='tc3s')
        Ws, bs = R.xavier_params(layers, seed=1111); Ws[0] = Ws[0] * (0.1 if case == 'f7' else 1.0)
        m.uv_net.set_weights(Ws, bs)
        for _ in range(3):
            m.engine.adam_step(5e-4)
        pr = torch.zeros(32, dtype=torch.int64, device='cuda')
        lib.pe_debug_set_tcs_profile(C.c_void_p(pr.data_ptr()))
        steps = 10
        for _ in range(steps):
            m.engine.adam_step(5e-4)
        torch.cuda.synchronize()
        lib.pe_debug_set_tcs_profile(None)
        p = pr.cpu().numpy().astype(np.float64) / steps
        emit(case='profs', workload=case, E_total=float(p[:16].sum()), I_total=float(p[16:].sum()), phases={n: float(v) for n, v in zip(TCS_NAMES, p) if n != '-'})
        del m


if __name__ == '__main__':
    T0 = time.time()
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    what = sys.argv[1:] or ['f5', 'f7', 'prof']
    emit(case='start', what=what, device=torch.cuda.get_device_name(0))
    for w in what:
        {'f5': f5, 'f7': f7, 'prof': prof, 'profs': profs}[w](flush)
    emit(case='done')
