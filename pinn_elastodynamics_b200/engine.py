"""Host-side engine: owns device buffers (torch tensors) and drives the C ABI of libpinn_elasto.so.

torch is plumbing only -- device memory, streams and torch.distributed (NCCL) for the one all-reduce of
`[grad | loss terms]` per evaluation.  All arithmetic of the hot path happens in the
hand-written sm_100a kernels behind include/pinn_elasto.h.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import sys

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Network:
    """One tanh MLP resident on one GPU in the padded layout of the C ABI."""

    def __init__(self, layers, device=None):
        self.lib = L.load()
        self.layers = [int(x) for x in layers]
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise L.PeError('the PINN-elastodynamics engine runs on CUDA devices only (no CPU fallback)')
        dims = (C.c_int * len(self.layers))(*self.layers)
        self.plan = self.lib.pe_plan_create(dims, len(self.layers), self.device.index or 0)
        if not self.plan:
            raise L.PeError('pe_plan_create: ' + self.lib.pe_last_error().decode())
        self.P = self.lib.pe_plan_param_count(self.plan)
        self.Pp = self.lib.pe_plan_param_count_padded(self.plan)
        z = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device=self.device)
        self.params, self.m, self.v = z(self.Pp), z(self.Pp), z(self.Pp)
        self.step = z(2, torch.int32)            # {adam step count, ticket}

    def __del__(self):
        try:
            if getattr(self, 'plan', None):
                self.lib.pe_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- compact <-> padded (reference order: weights then biases, plate:241)
    def pack(self, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.P, (flat.size, self.P)
        out = np.zeros(self.Pp, np.float32)
        L.check(self.lib.pe_pack_params(self.plan, flat.ctypes.data, out.ctypes.data), 'pe_pack_params')
        return out

    def unpack(self, padded):
        padded = np.ascontiguousarray(padded, dtype=np.float32)
        out = np.zeros(self.P, np.float32)
        L.check(self.lib.pe_unpack_params(self.plan, padded.ctypes.data, out.ctypes.data), 'pe_unpack_params')
        return out

    def set_flat(self, flat):
        self.params.copy_(torch.from_numpy(self.pack(flat)))

    def get_flat(self):
        return self.unpack(self.params.cpu().numpy())

    def set_weights(self, Ws, bs):
        if len(Ws) != len(self.layers) - 1:
            raise AssertionError('stored model must have the same number of layers')     # plate:299
        for l, (w, b) in enumerate(zip(Ws, bs)):
            if tuple(np.shape(w)) != (self.layers[l], self.layers[l + 1]):
                raise ValueError(f'layer {l}: weight shape {np.shape(w)} != {(self.layers[l], self.layers[l + 1])}')
        self.set_flat(np.concatenate([np.asarray(w, np.float64).ravel() for w in Ws] +
                                     [np.asarray(b, np.float64).ravel() for b in bs]))

    def get_weights(self, dtype=np.float32):
        flat = self.get_flat()
        Ws, bs, o = [], [], 0
        for l in range(len(self.layers) - 1):
            n = self.layers[l] * self.layers[l + 1]
            Ws.append(flat[o:o + n].reshape(self.layers[l], self.layers[l + 1]).astype(dtype)); o += n
        for l in range(len(self.layers) - 1):
            n = self.layers[l + 1]
            bs.append(flat[o:o + n].reshape(1, n).astype(dtype)); o += n
        return Ws, bs

    def reset_optimizer(self):
        self.m.zero_(); self.v.zero_(); self.step.zero_()

    # ---- forward-only entry points
    def forward_jets(self, points, K, in_scale=None, in_shift=None):
        """points: cuda float32 [n, ld] -> [n, K, O] (pe_forward_jets)."""
        n, ld = points.shape
        O = self.layers[-1]
        out = torch.empty((n, K, O), dtype=torch.float32, device=self.device)
        sc = (C.c_float * 3)(*(in_scale if in_scale is not None else (1, 1, 1)))
        sh = (C.c_float * 3)(*(in_shift if in_shift is not None else (0, 0, 0)))
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.pe_forward_jets(self.plan, K, _ptr(points), ld, n, sc, sh, _ptr(self.params), _ptr(out), C.c_void_p(st)),
                'pe_forward_jets')
        return out

    def forward_fields(self, points, formulation, in_scale=None, in_shift=None, aux=None):
        """points [n, ld] -> [n, 8] = (u, v, s11, s22, s12, e11, e22, e12) (pe_forward_fields)."""
        n, ld = points.shape
        out = torch.empty((n, 8), dtype=torch.float32, device=self.device)
        sc = (C.c_float * 3)(*(in_scale if in_scale is not None else (1, 1, 1)))
        sh = (C.c_float * 3)(*(in_shift if in_shift is not None else (0, 0, 0)))
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.pe_forward_fields(self.plan, formulation, _ptr(points), ld, n, sc, sh, _ptr(aux), 4 if aux is not None else 0,
                                           _ptr(self.params), _ptr(out), C.c_void_p(st)), 'pe_forward_fields')
        return out


class Term:
    """One loss contribution of one point set (one pe_residual_loss_grad launch per evaluation)."""

    def __init__(self, name, kind, K, points, desc, aux=None, enabled=True):
        self.name, self.kind, self.K = name, kind, K
        self.points = points                  # local shard, cuda float32 [n_local, ld]
        self.desc = desc
        self.aux = aux
        self.enabled = enabled
        self.lo, self.hi = 0, points.shape[0]  # active local row range
        self.slots = 0
        self.slot_base = 0


def shard_range(n, rank, world):
    """Contiguous index sharding, the reference's chunk arithmetic (semi:300-302)."""
    return int(rank * n / world), int((rank + 1) * n / world)


def chunk_shard_range(g_lo, g_hi, rank, world):
    """GLOBAL rows of rank `rank` inside the batch_num chunk [g_lo, g_hi): the chunk itself is index-sharded over all ranks, so every
    GPU works on every chunk (the reference trains the chunks one after the other on one device, semi:299-305)."""
    a, b = shard_range(g_hi - g_lo, rank, world)
    return g_lo + a, g_lo + b


class LossEngine:
    """Loss terms + gradient of one trainable network on one rank."""

    def __init__(self, net: Network, engine='simt', group=None):
        self.net = net
        self.lib = net.lib
        self.device = net.device
        self.engine = L.ENGINES[engine] if isinstance(engine, str) else int(engine)
        self.group = group
        self.world = torch.distributed.get_world_size(group) if (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        self.rank = torch.distributed.get_rank(group) if self.world > 1 else 0
        self.terms = []
        self.out = torch.zeros(net.Pp + L.PE_MAX_TERMS, dtype=torch.float32, device=self.device)
        self._built = False
        self.comm = None                       # peer-memory communicator (multi-GPU, one node): see _setup_comm
        self._comm_tried = False
        self.launches = 0                      # kernels of this library launched so far (bench `gpu_launches`)
        self.kernel_events = None              # set to a dict name -> [(start, end)] to time each residual kernel with CUDA events

    # ---- term construction
    def make_desc(self, kind, n_global, ld, E=0.0, mu=0.0, rho=0.0, hole_r=0.1, in_scale=(1, 1, 1), in_shift=(0, 0, 0),
                  cols=(), tgts=(), terms=(0, 1), weights=(1.0, 1.0), aux_k=0):
        d = L.TermDesc()
        d.kind, d.n_global, d.ld = kind, int(n_global), int(ld)
        d.E, d.mu, d.rho, d.hole_r = E, mu, rho, hole_r
        for i in range(3):
            d.in_scale[i] = in_scale[i]; d.in_shift[i] = in_shift[i]
        d.ncols = len(cols)
        for i in range(L.PE_MAX_COLS):
            d.col[i] = cols[i] if i < len(cols) else 0
            d.tgt[i] = tgts[i] if i < len(tgts) else -1
            d.term[i] = terms[i] if i < len(terms) else 0
            d.w[i] = weights[i] if i < len(weights) else 0.0
        d.aux_k = aux_k
        return d

    def add_term(self, name, kind, K, points_np, **kw):
        """points_np: GLOBAL host array [N, ld]; this rank keeps rows shard_range(N, rank, world)."""
        pts = np.ascontiguousarray(points_np, dtype=np.float32)
        N, ld = pts.shape
        a, b = shard_range(N, self.rank, self.world)
        local = torch.from_numpy(pts[a:b].copy()).to(self.device)
        aux = kw.pop('aux', None)
        if aux is not None:
            aux = aux[a:b].contiguous()
        desc = self.make_desc(kind, N, ld, **kw)
        t = Term(name, kind, K, local, desc, aux)
        t.global_n = N
        t.global_lo = a
        t.host = pts                          # global rows (host, fp32): source of chunk re-sharding (set_chunk, world > 1)
        t.resident = local                    # this rank's shard of the whole set
        self.terms.append(t)
        self._built = False
        return t

    def set_chunk(self, term, g_lo, g_hi):
        """Activate GLOBAL rows [g_lo, g_hi) of a term (reference `batch_num` chunking, semi:299-305).
        The mean is over the chunk's rows; each rank processes its part of the chunk."""
        if (g_lo, g_hi) == (0, term.global_n):          # whole set: this rank's resident shard
            term.points = term.resident
            term.lo, term.hi = 0, term.points.shape[0]
            term.desc.n_global = term.global_n
        elif self.world == 1:
            term.points = term.resident
            term.lo, term.hi = g_lo, g_hi
            term.desc.n_global = g_hi - g_lo
        else:
            # every rank takes its index shard OF THE CHUNK (same arithmetic, semi:300-302), so a chunk is spread over all GPUs instead of
            # living on the ranks whose resident shard happens to contain it; one H2D copy per chunk switch (once per `iter` steps)
            if term.aux is not None:
                raise L.PeError('batch_num chunking of a composite (aux) point set with world_size > 1 is not supported')
            a, b = chunk_shard_range(g_lo, g_hi, self.rank, self.world)
            term.points = torch.from_numpy(term.host[a:b]).to(self.device)
            term.lo, term.hi = 0, b - a
            term.desc.n_global = g_hi - g_lo
        self._built = False

    def build(self):
        base, stash = 0, 1
        for t in self.terms:
            t.fused_into, t.fused = None, None
            # the requested engine where it implements the term, the fp32 SIMT engine elsewhere (data / boundary terms)
            t.engine = self.engine if self.lib.pe_engine_supported(self.net.plan, t.kind, t.K, self.engine) else L.ENGINE_SIMT_FP32
        # tensor-core engine: one primal-only set (hole traction / data term) rides the collocation launch as extra tiles
        main = next((t for t in self.terms if t.enabled and t.engine != L.ENGINE_SIMT_FP32), None)
        if main is not None:
            sec = next((t for t in self.terms if t.enabled and t is not main and t.K == 1 and t.kind in (L.RES_TRACTION, L.RES_COLS)), None)
            if sec is not None:
                main.fused, sec.fused_into = sec, main
        for t in self.terms:
            if t.fused_into is not None:
                t.slots, t.slot_base = 0, base
                continue
            # slot count = what the launch itself derives from the real row counts (csrc/pe_api.cu residual_common): an empty shard or
            # chunk still runs one CTA that zero-fills its slot
            n = t.hi - t.lo
            if t.fused is not None:
                tiles = -(-n // L.PE_TC_TILE) + -(-(t.fused.hi - t.fused.lo) // L.PE_TC_TILE)
                n = tiles * L.PE_TC_TILE
            t.slots = self.lib.pe_plan_slots(self.net.plan, n, t.K, t.engine)
            t.slot_base = base
            base += t.slots
            stash = max(stash, self.lib.pe_plan_scratch_floats(self.net.plan, n, t.K, t.engine))
        self.n_slots = base
        need = base * self.net.Pp
        if not hasattr(self, 'gpart') or self.gpart.numel() < need:
            self.gpart = torch.empty(need, dtype=torch.float32, device=self.device)
            self.tpart = torch.empty(base * L.PE_MAX_TERMS, dtype=torch.float32, device=self.device)
        if not hasattr(self, 'stash') or self.stash.numel() < stash:
            self.stash = torch.empty(stash, dtype=torch.float32, device=self.device)
        self._built = True

    # ---- evaluation
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_terms(self):
        if not self._built:
            self.build()
        st = self._stream()
        slots = 0
        for t in self.terms:
            if not t.enabled or t.fused_into is not None:
                continue
            n = t.hi - t.lo
            pts = t.points[t.lo:t.hi] if (t.lo, t.hi) != (0, t.points.shape[0]) else t.points
            aux = None if t.aux is None else (t.aux[t.lo:t.hi] if n > 0 else t.aux)      # an empty slice has no data pointer
            if self.kernel_events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if t.fused is not None:
                f = t.fused
                L.check(self.lib.pe_residual_loss_grad_fused(self.net.plan, C.byref(t.desc), t.K, t.engine, _ptr(pts), n, _ptr(aux),
                                                             C.byref(f.desc), _ptr(f.points[f.lo:f.hi]), f.hi - f.lo, _ptr(None if f.aux is None else f.aux[f.lo:f.hi]),
                                                             _ptr(self.net.params), _ptr(self.gpart), _ptr(self.tpart), _ptr(self.stash), slots, st),
                        f'pe_residual_loss_grad_fused[{t.name}+{f.name}]')
            else:
                L.check(self.lib.pe_residual_loss_grad(self.net.plan, C.byref(t.desc), t.K, t.engine,
                                                       _ptr(pts), n, _ptr(aux), _ptr(self.net.params),
                                                       _ptr(self.gpart), _ptr(self.tpart), _ptr(self.stash), slots, st),
                        f'pe_residual_loss_grad[{t.name}]')
            if self.kernel_events is not None:
                e1.record()
                self.kernel_events.setdefault(t.name, []).append((e0, e1))
            slots += t.slots
            self.launches += 1 if t.engine == L.ENGINE_SIMT_FP32 else 2      # tensor-core engine: operand-image prep + residual kernel
        return slots

    # ---- multi-GPU: one-kernel slot reduction + all-reduce over NVLink peer memory (+ Adam); NCCL when it cannot be set up
    def _setup_comm(self):
        """Collective over the group (every rank reaches it at its first multi-rank evaluation).  The IPC handles travel over
        torch.distributed; PE_PEER_ALLREDUCE=0 keeps the reduce -> NCCL all_reduce -> Adam path."""
        self._comm_tried = True
        dist = torch.distributed
        if self.world < 2 or self.world > L.PE_MAX_PEERS or os.environ.get('PE_PEER_ALLREDUCE', '1') == '0':
            return
        if dist.get_backend(self.group) != 'nccl':
            return
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=self.group)
        handle = (C.c_ubyte * L.PE_IPC_HANDLE_BYTES)()
        comm = self.lib.pe_comm_create(self.net.plan, self.rank, self.world, handle) if len(set(hosts)) == 1 else None
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)
        ok = torch.tensor([1 if comm else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            blob = torch.cat(allh).cpu().numpy().tobytes()
            rc = self.lib.pe_comm_connect(comm, blob)
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self.comm = comm
        else:                                   # some rank could not export / map the region: all ranks use NCCL
            if comm:
                self.lib.pe_comm_destroy(comm)
            if self.rank == 0:
                msg = self.lib.pe_last_error()
                print('[pinn_elasto] peer-memory all-reduce unavailable (%s): using NCCL' % (msg.decode() if msg else 'not one node'), file=sys.stderr)

    def check_comm(self):
        """Raise if a peer-memory reduction timed out (synchronises)."""
        if self.comm is not None and self.lib.pe_comm_error(self.comm) != 0:
            L.check(1, 'pe_reduce_peer')

    def close(self):
        """Collective teardown of the communicator (optional; the 2 MiB region otherwise lives until the process exits)."""
        if self.comm is not None:
            torch.cuda.synchronize(self.device)
            self.lib.pe_comm_disconnect(self.comm)
            torch.distributed.barrier(group=self.group)
            self.lib.pe_comm_destroy(self.comm)
            self.comm = None
            self._comm_tried = False

    def evaluate(self, hist_row=None):
        """loss terms + full gradient into self.out = [grad (padded) | terms(8)] (all-reduced over ranks)."""
        if self.world > 1 and not self._comm_tried:
            self._setup_comm()
        slots = self.launch_terms()
        if self.comm is not None:
            L.check(self.lib.pe_reduce_peer(self.net.plan, self.comm, _ptr(self.gpart), _ptr(self.tpart), slots, _ptr(self.out), _ptr(hist_row),
                                            None, None, None, None, 0.0, 0.0, 0.0, 0.0, self._stream()), 'pe_reduce_peer')
            self.launches += 1
            return self.out
        copy = hist_row if self.world == 1 else None
        L.check(self.lib.pe_reduce_partials(self.net.plan, _ptr(self.gpart), _ptr(self.tpart), slots, _ptr(self.out), _ptr(copy), self._stream()),
                'pe_reduce_partials')
        self.launches += 1
        if self.world > 1:
            torch.distributed.all_reduce(self.out, group=self.group)
            if hist_row is not None:
                hist_row.copy_(self.out[self.net.Pp:])
        return self.out

    def adam_step(self, lr, hist_row=None, beta1=0.9, beta2=0.999, eps=1e-8):
        """One Adam step (TF1 form).  hist_row receives the PRE-update loss terms of this step."""
        net = self.net
        if self.world > 1 and not self._comm_tried:
            self._setup_comm()
        if self.world == 1:
            slots = self.launch_terms()
            L.check(self.lib.pe_reduce_adam(net.plan, _ptr(self.gpart), _ptr(self.tpart), slots, _ptr(self.out), _ptr(hist_row),
                                            _ptr(net.params), _ptr(net.m), _ptr(net.v), _ptr(net.step),
                                            lr, beta1, beta2, eps, self._stream()), 'pe_reduce_adam')
            self.launches += 1
        elif self.comm is not None:             # slot reduction + peer-memory all-reduce + Adam: one launch
            slots = self.launch_terms()
            L.check(self.lib.pe_reduce_peer(net.plan, self.comm, _ptr(self.gpart), _ptr(self.tpart), slots, _ptr(self.out), _ptr(hist_row),
                                            _ptr(net.params), _ptr(net.m), _ptr(net.v), _ptr(net.step),
                                            lr, beta1, beta2, eps, self._stream()), 'pe_reduce_peer')
            self.launches += 1
        else:                                   # NCCL path: reduce -> all_reduce -> Adam
            self.evaluate(hist_row)
            L.check(self.lib.pe_adam_step(net.plan, _ptr(net.params), _ptr(self.out), _ptr(net.m), _ptr(net.v), _ptr(net.step),
                                          lr, beta1, beta2, eps, self._stream()), 'pe_adam_step')
            self.launches += 1

    def terms_host(self):
        self.check_comm()
        return self.out[self.net.Pp:].cpu().numpy().astype(np.float64)

    def grad_compact_host(self):
        return self.net.unpack(self.out[:self.net.Pp].cpu().numpy())
