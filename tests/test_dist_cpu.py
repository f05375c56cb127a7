"""world_size-2 gloo test of the multi-GPU protocol on CPU: contiguous index sharding of every point set
(`shard_range`, the reference's chunk arithmetic), per-rank sums scaled by 1/N_global, ONE sum all-reduce of
[grad | loss terms], identical TF1-Adam update on every rank.  The per-shard compute is done by the oracle here (no GPU in
this container); the CUDA kernels implement exactly the same per-shard contract (n_local rows, n_global denominator) and
are checked against the oracle in tests/test_gpu_parity.py, and 1-vs-N GPU equality in tests/test_gpu_dist.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_torch as R


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _shard_loss_grad(orc, sets, rank, world):
    """rank-local [grad | terms]: sums over the shard rows divided by the GLOBAL row count."""
    from pinn_elastodynamics_b200.engine import shard_range
    terms, total = None, 0.0
    local = {}
    scale = {}
    for k, A in sets.items():
        a, b = shard_range(A.shape[0], rank, world)
        local[k] = A[a:b]
        scale[k] = (b - a) / A.shape[0]
    # oracle means are over local rows; rescale each term to "sum over shard / N_global"
    Tc, _ = orc.loss_terms({'Collo': local['Collo'], 'HOLE': local['HOLE']})
    t_uv, t_s, t_h = Tc['loss_f_uv'] * scale['Collo'], Tc['loss_f_s'] * scale['Collo'], Tc['loss_HOLE'] * scale['HOLE']
    loss = 10 * (t_uv + t_s + t_h)
    gs = torch.autograd.grad(loss, orc.params())
    flat = torch.cat([g.reshape(-1) for g in gs])
    return torch.cat([flat, torch.stack([t_uv, t_s, t_h]).detach()])


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    layers = [3, 12, 12, 5]
    Ws, bs = R.xavier_params(layers, seed=4)
    sets = {'Collo': rng.uniform([0, 0, 0], [.5, .5, 10], (101, 3)), 'HOLE': rng.uniform([0, 0, 0], [.1, .1, 10], (17, 3))}
    orc = R.Oracle('plate', Ws, bs)
    curve = []
    for _ in range(3):
        buf = _shard_loss_grad(orc, sets, rank, world)
        dist.all_reduce(buf)                                  # the one collective per step
        P = buf.numel() - 3
        grads, o = [], 0
        for p_ in orc.params():
            grads.append(buf[o:o + p_.numel()].reshape(p_.shape)); o += p_.numel()
        orc.adam_step(grads, 5e-4)
        curve.append(float(10 * buf[P:].sum()))
    q.put((rank, orc.flat_params(), curve))
    dist.destroy_process_group()


def test_two_rank_sharded_training_equals_single_rank():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(world):
        r, flat, curve = q.get(timeout=120)
        res[r] = (flat, curve)
    for p in procs:
        p.join(timeout=60)
    # every rank holds identical parameters (replicated update, no broadcast needed)
    np.testing.assert_array_equal(res[0][0], res[1][0])
    # and they equal the single-process run on the whole sets
    rng = np.random.default_rng(5)
    layers = [3, 12, 12, 5]
    Ws, bs = R.xavier_params(layers, seed=4)
    sets = {'Collo': rng.uniform([0, 0, 0], [.5, .5, 10], (101, 3)), 'HOLE': rng.uniform([0, 0, 0], [.1, .1, 10], (17, 3))}
    orc = R.Oracle('plate', Ws, bs)
    ref_curve = []
    for _ in range(3):
        T, loss = orc.loss_terms(sets)
        gs = torch.autograd.grad(loss, orc.params())
        ref_curve.append(float(loss))
        orc.adam_step(gs, 5e-4)
    np.testing.assert_allclose(res[0][1], ref_curve, rtol=1e-12)
    np.testing.assert_allclose(res[0][0], orc.flat_params(), rtol=1e-10, atol=1e-14)


def test_chunks_are_sharded_over_all_ranks():
    """batch_num chunking under world_size > 1 (engine.set_chunk): the ranks' parts of a chunk tile it exactly, in rank order, for ragged
    sizes too; with one rank the part is the chunk itself"""
    from pinn_elastodynamics_b200.engine import chunk_shard_range
    for N, B, world in ((2003, 3, 2), (150357, 7, 8), (10, 4, 8), (5, 1, 3)):
        for i in range(B):
            g_lo, g_hi = int(i * N / B), int((i + 1) * N / B)                 # semi:300-302
            parts = [chunk_shard_range(g_lo, g_hi, r, world) for r in range(world)]
            assert parts[0][0] == g_lo and parts[-1][1] == g_hi
            assert all(a <= b for a, b in parts) and all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
        assert chunk_shard_range(3, 11, 0, 1) == (3, 11)
