"""Multi-GPU sharding invariance (needs >= 2 GPUs; the 4- and 8-rank cases skip by device count): N NCCL ranks on the sharded point sets
reproduce the 1-rank loss terms, gradient and Adam trajectory; the in-kernel peer-memory all-reduce equals the NCCL
path; batch_num chunks are spread over the ranks (semi:299-302)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from oracle import ref_torch as R
import pinn_elastodynamics_b200 as pe
local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rng = np.random.default_rng(3)
layers = [3] + 3 * [50] + [5]
Collo = rng.uniform([0, 0, 0], [.5, .5, 10], (3001, 3)); HOLE = rng.uniform([0, 0, 0], [.1, .1, 10], (333, 3))
m = pe.PINN(Collo, HOLE, None, None, None, None, None, None, layers, None, None, None, None, verbose=False, engine=sys.argv[1])
Ws, bs = R.xavier_params(layers, seed=8); m.uv_net.set_weights(Ws, bs)
m.engine.evaluate()
t0 = m.engine.terms_host().tolist(); g0 = m.engine.grad_compact_host().astype(float)
curve = m.train(5, 5e-4)[3]
if int(os.environ.get('RANK', '0')) == 0:
    print('RESULT ' + json.dumps({'peer': m.engine.comm is not None, 'terms': t0, 'gnorm': float(np.linalg.norm(g0)), 'g': g0[::97].tolist(), 'curve': curve, 'params': m.uv_net.get_flat()[::97].astype(float).tolist()}))
if world > 1:
    dist.destroy_process_group()
''' % ROOT


def _run(world, engine, tmp_path, peer=True):
    f = tmp_path / 'w.py'
    f.write_text(WORKER)
    if world == 1:
        cmd = [sys.executable, str(f), engine]
    else:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
               '--master-port', '29611', str(f), engine]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PE_PEER_ALLREDUCE='1' if peer else '0'))
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(line[0][7:])


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('engine', ['simt', 'tcf'])
def test_n_gpus_match_one(engine, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    a, b = _run(1, engine, tmp_path), _run(world, engine, tmp_path)
    assert b['peer'], 'the one-kernel peer-memory all-reduce did not come up (CUDA IPC): the run fell back to NCCL'
    np.testing.assert_allclose(b['terms'][:3], a['terms'][:3], rtol=2e-6)
    np.testing.assert_allclose(b['g'], a['g'], rtol=1e-4, atol=1e-6 * a['gnorm'])
    np.testing.assert_allclose(b['curve'], a['curve'], rtol=1e-5)
    np.testing.assert_allclose(b['params'], a['params'], rtol=1e-5, atol=1e-7)


def test_peer_memory_allreduce_equals_nccl_path(tmp_path):
    """reduce + NVLink peer-memory exchange + Adam in one kernel (pe_reduce_peer) against reduce -> NCCL all_reduce -> Adam:
    with 2 ranks both add the two partial sums once (a + b), so the trajectories agree bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    a, b = _run(2, 'tcf', tmp_path, peer=True), _run(2, 'tcf', tmp_path, peer=False)
    assert a['peer'] and not b['peer']
    assert a['terms'] == b['terms'] and a['g'] == b['g'] and a['curve'] == b['curve'] and a['params'] == b['params']


@pytest.mark.parametrize('world', [4, 8])
def test_peer_memory_allreduce_close_to_nccl_path_at_4_and_8_ranks(world, tmp_path):
    """more than two ranks: the peer kernel adds the rows in rank order on every rank (identical bits on all ranks), NCCL's ring / tree
    order differs, so the two paths agree to fp32 summation-order error, not bit for bit"""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    a, b = _run(world, 'tcf', tmp_path, peer=True), _run(world, 'tcf', tmp_path, peer=False)
    assert a['peer'] and not b['peer']
    np.testing.assert_allclose(a['terms'][:3], b['terms'][:3], rtol=1e-6)
    np.testing.assert_allclose(a['g'], b['g'], rtol=1e-5, atol=1e-7 * a['gnorm'])
    np.testing.assert_allclose(a['curve'], b['curve'], rtol=1e-6)


CHUNK_WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import pinn_elastodynamics_b200 as pe
from pinn_elastodynamics_b200.models import xavier_init_lists
local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rng = np.random.default_rng(5)
lb, ub = np.array([-15., -15, 0]), np.array([15., 15, 16])
layers = [3] + 3 * [40] + [7]
Collo = rng.uniform(lb, ub, (2003, 3)); IC = rng.uniform(lb, ub, (211, 3)); IC[:, 2] = 0; UP = rng.uniform(lb, ub, (173, 3)); UP[:, 1] = 15
SRC = np.concatenate([rng.uniform(lb, ub, (257, 3)), rng.standard_normal((257, 2)) * 0.1], 1)
m = pe.DeepHPM(Collo, SRC, IC, UP, layers, lb, ub, verbose=False, engine=sys.argv[1])
Ws, bs = xavier_init_lists(layers, np.random.default_rng(2)); Ws[0] = Ws[0] * 0.1
m.uv_net.set_weights(Ws, bs)
out = m.train(4, 1e-3, 3)            # three chunks of the collocation set, 4 steps each (semi:299-326)
if int(os.environ.get('RANK', '0')) == 0:
    print('RESULT ' + json.dumps({'curve': out[-1], 'f_uv': out[0], 'params': m.uv_net.get_flat()[::53].astype(float).tolist()}))
if world > 1:
    dist.destroy_process_group()
''' % ROOT


@pytest.mark.parametrize('engine', ['simt', 'tcf'])
def test_batch_num_chunks_are_spread_over_the_ranks(engine, tmp_path):
    """train(iter, lr, batch_num) under 2 ranks: every chunk [int(i N / B), int((i + 1) N / B)) is sharded over both GPUs and the trajectory is
    the 1-rank trajectory (the reference is single-device; its chunk semantics are kept, semi:299-305)"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    f = tmp_path / 'c.py'
    f.write_text(CHUNK_WORKER)
    outs = []
    for world in (1, 2):
        cmd = [sys.executable, str(f), engine] if world == 1 else [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
                                                                   '--master-addr', '127.0.0.1', '--master-port', '29613', str(f), engine]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        assert line, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(json.loads(line[0][7:]))
    a, b = outs
    assert len(a['curve']) == 12
    np.testing.assert_allclose(b['curve'], a['curve'], rtol=1e-5)
    np.testing.assert_allclose(b['f_uv'], a['f_uv'], rtol=2e-5)
    np.testing.assert_allclose(b['params'], a['params'], rtol=1e-5, atol=1e-7)
