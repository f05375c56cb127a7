#!/usr/bin/env python
"""bench.py -- collocation-point residual+grad evaluations / second per Adam step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine auto|tcf|tc3s|simt|tc1s] [--points P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8d "Config 2"): defected plate, plane stress (E=20, mu=.25, rho=1),
5x50 tanh mixed-variable net [3,50,50,50,50,50,5] (plain, no dist/part composite), 50,000 collocation points PER GPU
(weak scaling; 8 GPUs x 125,000 = BASELINE config 4 is reachable with --points 125000) + N_c/10 hole-traction points,
loss 10*(f_uv + f_s + HOLE), fp32, Adam lr 5e-4 (TF1 form).  One "step" = loss terms + full d loss/d theta over all
point sets + slot reduction [+ NCCL all-reduce of [grad | terms] for N > 1] + Adam update.  value = N_c(total) * K / time.

Timing: W >= 3 warm-up steps; each timed step is bracketed by its own CUDA-event pair on the launching stream and L2 is
flushed (256 MiB memset) between steps, outside the event pairs; the K step times are summed, max over ranks; the
whole timed region is bracketed by barrier + synchronize.  `e2e` runs the same steps through PINN.train(refeed=True):
point arrays re-uploaded from pinned host memory and the loss terms read back every step, as the reference's
feed_dict / sess.run does (plate:482-505).  `--impl reference` times the float64 CPU oracle (stand-in for the TF1 CPU
path, which cannot be imported here: no tensorflow) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS = [3] + 5 * [50] + [5]
S_WEIGHTS = sum(LAYERS[i] * LAYERS[i + 1] for i in range(len(LAYERS) - 1))     # 10,400
FLOP_PER_POINT = 6 * 5 * S_WEIGHTS                                            # 6*K*S = 312,000
METRIC = 'collocation-pt residual+grad evals/sec per Adam step'


def measured_traffic(engine, points):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: read from profiles/traffic.json, which is written
    from an `ncu --set full` capture of THIS build and workload (profiles/summarize.py); None when no capture of that engine / size is
    committed (a number from another build is not a measurement of this run)."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        e = t.get(f'{engine}:{points}')
        return None if e is None else float(e['dram_bytes_read']) + float(e['dram_bytes_write'])
    except Exception:
        return None


def make_workload(n_c, seed=1111):
    """SURVEY 8d synthetic inputs: uniform in [0,.5]^2 x [0,10] minus the r<=0.1 quarter hole, resampled to exactly N."""
    rng = np.random.default_rng(seed)
    pts = np.zeros((0, 3))
    while pts.shape[0] < n_c:
        P = rng.uniform([0, 0, 0], [.5, .5, 10], (n_c, 3))
        pts = np.concatenate([pts, P[np.hypot(P[:, 0], P[:, 1]) > 0.1]])
    Collo = pts[:n_c]
    n_h = max(n_c // 10, 1)
    th = rng.uniform(0, np.pi / 2, n_h)
    HOLE = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), rng.uniform(0, 10, n_h)], 1)
    return Collo, HOLE


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    return p


class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def pick_threads(make_step):
    """The oracle is torch-CPU autograd over many small GEMMs: past a few dozen threads it slows down (oversubscribed
    intra-op pools).  Give the CPU arm its best shot: time one step at several thread counts, keep the fastest."""
    import torch
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best, best_t = cands[0], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        make_step()                      # warm
        t0 = time.perf_counter()
        make_step()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, emit):
    """--impl reference: the float64 CPU oracle (torch autograd restatement of the TF1 graph), all host threads,
    bounded sample of the workload.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import ref_torch as R
    n_s = min(args.points, args.ref_points)
    Collo, HOLE = make_workload(n_s)
    Ws, bs = R.xavier_params(LAYERS, seed=1111)
    orc = R.Oracle('plate', Ws, bs)
    sets = {'Collo': Collo, 'HOLE': HOLE}
    small = {'Collo': Collo[:2000], 'HOLE': HOLE[:200]}

    def step(s=sets):
        T, loss = orc.loss_terms(s)
        gs = torch.autograd.grad(loss, orc.params())
        orc.adam_step(gs, 5e-4)
    cores = pick_threads(lambda: step(small))
    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, args.ref_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = n_s * steps / dt
    sample = f'{steps} Adam steps (loss+grad+update) on {n_s} collocation + {HOLE.shape[0]} hole points, float64, torch {torch.__version__} CPU autograd'
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'points/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 1),
        'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'plate F5 5x50, CPU oracle sample', 'net': LAYERS, 'points_per_step': n_s},
        'cpu_baseline': {'value': v, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def cpu_baseline_leg(n_points, budget_s=12.0):
    import torch
    from oracle import ref_torch as R
    Collo, HOLE = make_workload(n_points)
    Ws, bs = R.xavier_params(LAYERS, seed=1111)
    orc = R.Oracle('plate', Ws, bs)
    sets = {'Collo': Collo, 'HOLE': HOLE}
    small = {'Collo': Collo[:2000], 'HOLE': HOLE[:200]}

    def step(s=sets):
        T, loss = orc.loss_terms(s)
        gs = torch.autograd.grad(loss, orc.params())
        orc.adam_step(gs, 5e-4)
    cores = pick_threads(lambda: step(small))
    step()
    steps, t0 = 0, time.perf_counter()
    while True:
        step()
        steps += 1
        if time.perf_counter() - t0 > budget_s or steps >= 20:
            break
    dt = time.perf_counter() - t0
    return {'value': n_points * steps / dt, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
            'host_cores': os.cpu_count(),
            'sample': f'{steps} Adam steps on {n_points} collocation + {HOLE.shape[0]} hole points, float64 torch-CPU autograd oracle, {cores} threads (fastest of 8/16/32/64/all) '
                      f'(stand-in for the TF1 CPU path; tensorflow is not importable), {dt:.1f} s'}


def main():
    # Only the final JSON line may reach stdout: libraries (NCCL prints its version banner on stdout) are diverted to stderr
    # at the file-descriptor level; the JSON is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + '\n').encode())

    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1500)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--engine', default='auto', help="auto = tcf = fp16-pair tcgen05 engine (fp32-grade, what the drop-in classes select by default), tc3s = TF32x3 tcgen05 engine (A/B partner), simt = fp32 FFMA engine, tc1s = single-pass TF32")
    ap.add_argument('--points', type=int, default=50000, help='collocation points per GPU')
    ap.add_argument('--ref-points', type=int, default=10000)
    ap.add_argument('--ref-steps', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, emit)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import pinn_elastodynamics_b200 as pe

    n_total = args.points * world
    Collo, HOLE = make_workload(n_total)
    model = pe.PINN(Collo, HOLE, None, None, None, None, None, None, LAYERS, None, None, None, None, verbose=False, engine=args.engine)
    # the package's own Xavier initialiser with the reference's seed: array-equal to the arrays the reference arm / cpu_baseline leg feed
    # to the oracle (tests/test_host.py), so nothing under oracle/ is imported on this arm
    from pinn_elastodynamics_b200.models import xavier_init_lists
    Ws, bs = xavier_init_lists(LAYERS, np.random.default_rng(1111))
    model.uv_net.set_weights(Ws, bs)
    eng = model.engine
    ename = 'tcf' if args.engine == 'auto' else args.engine
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(k, record_kernels):
        evs = []
        eng.kernel_events = {} if record_kernels else None
        for _ in range(k):
            flush.zero_()                                    # L2 flush, outside the timed pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.adam_step(5e-4)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ke = eng.kernel_events
        eng.kernel_events = None
        return sum(a.elapsed_time(b) for a, b in evs), ke

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        eng.adam_step(5e-4)
    barrier()
    l0 = eng.launches
    barrier()
    t_wall = time.perf_counter()
    total_ms, kev = timed_steps(args.steps, True)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = eng.launches - l0
    tm = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    ms_step = total_ms / args.steps
    value = n_total / (ms_step * 1e-3)
    # dominant kernel: the collocation residual kernel, CUDA events around each launch (same timed region)
    kms = float(np.mean([a.elapsed_time(b) for a, b in kev['Collo']]))
    n_local = eng.terms[0].points.shape[0]
    achieved = n_local * FLOP_PER_POINT / (kms * 1e-3) / 1e12
    pk = peaks()
    peak = pk.get('bf16_tflops_sustained', 1400.0)

    # ---- e2e: same steps through the public class API with host buffers re-fed every step
    e2e = None
    if not args.no_e2e:
        k2 = max(3, min(args.steps, 300))      # enough steps that train()'s one-time set-up (pinned buffers, final evaluation) is amortised
        model.train(3, 5e-4, refeed=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        e0.record()
        model.train(k2, 5e-4, refeed=True)
        e1.record()
        barrier()
        tw = (time.perf_counter() - tw) * 1e3
        t2 = torch.tensor([max(e0.elapsed_time(e1), tw)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        # train(k2) runs k2 updates + 1 final evaluation; count k2 steps
        e2e = {'value': n_total * k2 / (float(t2.item()) * 1e-3), 'unit': 'points/s',
               'h2d_bytes_per_step': int(model.h2d_bytes_per_step), 'd2h_bytes_per_step': int(model.d2h_bytes_per_step),
               'steps': k2, 'api': 'PINN.train(iter, lr, refeed=True)'}

    clocks = sampler.stop() if rank == 0 else None      # sampled over warm-up + timed region + e2e (all under load)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(min(args.points, args.ref_points))

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'defected plate F5 (plane stress), 5x50 tanh mixed-variable net, {args.points} collocation pts/GPU + {HOLE.shape[0] // world} hole pts/GPU, Adam lr 5e-4 (BASELINE configs[1])',
                       'net': LAYERS, 'global_collocation_points': n_total, 'engine': ename, 'parallelism': f'dp{world} (index-sharded points, 1 all-reduce of [grad|terms] per step)',
                       'allreduce': None if world == 1 else ('in-kernel over NVLink peer memory, fused with slot reduction and Adam (pe_reduce_peer)' if eng.comm is not None else 'NCCL (reduce -> all_reduce -> Adam)'),
                       'l2': 'flushed between timed steps (256 MiB memset outside the per-step event pairs)'},
            'clocks': clocks,
            'gpu_launches': launches,
            'e2e': e2e,
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': measured_traffic(ename, args.points),
                         'kernel': {'simt': 'resid_simt_kernel<5>', 'tcf': 'resid_tcf_kernel<5> (tcgen05, fp16-pair operands)',
                                    'tc3s': 'resid_tcs_kernel<5> (tcgen05, TF32x3)', 'tc1s': 'resid_tcs_kernel<5> (tcgen05, TF32)'}[ename] + ' (collocation F5)', 'kernel_ms': kms, 'kernel_share_of_step': kms / ms_step,
                         'flop_per_point': FLOP_PER_POINT,
                         'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if 'bf16_tflops_sustained' in pk else 'fallback',
                         'fp32_ffma_peak_tflops': 148 * 128 * 2 * (clocks['sm_mhz'] or 1965.0) * 1e-6},
            'cpu_baseline': cpu,
            'wall_s_timed_region': t_wall,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
