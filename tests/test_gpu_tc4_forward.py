"""OPT-IN (PE_TEST_TC4=1): forward-only probe of the next tensor-core engine's arithmetic (csrc/pe_tc4_probe.cu) against
the SIMT fp32 forward jets.  The probe was written after round 1's GPU budget was spent and has not run on hardware yet, so it is not
part of the default `-m gpu` suite; round 2 starts with   PE_TEST_TC4=1 python -m pytest tests/test_gpu_tc4_forward.py -q   .
Expected from the CPU model of the arithmetic (tests/emulate_engine_precision.py): ~2e-6 of each stream's output scale after five layers
(the shipped TF32 split: ~5e-6); the bar below is 1e-5."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import ref_torch as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('PE_TEST_TC4') != '1', reason='experimental probe: set PE_TEST_TC4=1')]


def _run(layers, K, n, variant, lb=None, ub=None, seed=3):
    from pinn_elastodynamics_b200 import _lib as L
    from pinn_elastodynamics_b200.engine import Network
    dev = torch.device('cuda', torch.cuda.current_device())
    net = Network(layers, dev)
    Ws, bs = R.xavier_params(layers, seed=seed)
    rng = np.random.default_rng(seed)
    bs = [rng.standard_normal(b.shape) * 0.1 for b in bs]
    net.set_weights(Ws, bs)
    lo = np.array([0., 0, 0]) if lb is None else lb
    hi = np.array([.5, .5, 10.]) if ub is None else ub
    pts = torch.from_numpy(rng.uniform(lo, hi, (n, 3)).astype(np.float32)).to(dev)
    sc = None if lb is None else tuple(2.0 / (hi - lo))
    sh = None if lb is None else tuple(-2.0 * lo / (hi - lo) - 1.0)
    ref = net.forward_jets(pts, K, sc, sh).cpu().numpy().astype(np.float64)
    lib = net.lib
    scratch = torch.zeros(int(lib.pe_debug_tc4_scratch_bytes(net.plan)), dtype=torch.uint8, device=dev)
    out = torch.full((n, K, layers[-1]), float('nan'), dtype=torch.float32, device=dev)
    csc = (C.c_float * 3)(*(sc if sc is not None else (1, 1, 1)))
    csh = (C.c_float * 3)(*(sh if sh is not None else (0, 0, 0)))
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(lib.pe_debug_forward_jets_tc4(net.plan, K, C.c_void_p(pts.data_ptr()), 3, n, csc, csh, C.c_void_p(net.params.data_ptr()),
                                          C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), variant, st), 'pe_debug_forward_jets_tc4')
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64), ref


@pytest.mark.parametrize('variant', [0, 2])            # 0: 8-chunk planes with a zero pad chunk, 2: 7-chunk planes, the last K-step reads past the plane (x zero rows)
                                                       # (variant 1, LBO = 0 on the last K-step, never completes on B200: the launch hangs until the bounded wait traps)
@pytest.mark.parametrize('K,O', [(5, 5), (4, 7)])
def test_forward_jets_on_the_16_bit_split(K, O, variant):
    got, ref = _run([3] + 5 * [50] + [O], K, 1000, variant)
    assert np.isfinite(got).all()
    print('tc-forward K=%d O=%d variant %d: per-stream rel err' % (K, O, variant), ['%.2e' % (np.abs(got[:, k] - ref[:, k]).max() / max(1e-30, np.abs(ref[:, k]).max())) for k in range(K)])
    for k in range(K):      # per stream: error against that stream's output scale
        assert np.abs(got[:, k] - ref[:, k]).max() <= 1e-5 * max(1e-30, np.abs(ref[:, k]).max()), (k, np.abs(got[:, k] - ref[:, k]).max(), np.abs(ref[:, k]).max())


@pytest.mark.parametrize('n', [1, 127, 129, 128 * 149 + 3])
def test_ragged_point_counts_and_narrow_nets(n):
    got, ref = _run([3, 14, 30, 5], 5, n, 2)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def test_normalised_inputs():
    lb, ub = np.array([0., 0, 0]), np.array([30., 30, 20.])
    got, ref = _run([3] + 3 * [50] + [7], 4, 500, 2, lb, ub)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
