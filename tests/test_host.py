"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/pinn_elasto.h declares,
layout / pack / unpack logic, sharding arithmetic, and the 2-rank gloo all-reduce protocol of [grad | terms]."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from pinn_elastodynamics_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return L, L.load()


def test_library_exports_every_declared_symbol():
    L, lib = _lib()
    hdr = open(os.path.join(ROOT, 'include', 'pinn_elasto.h')).read()
    declared = set(re.findall(r'^(?:int|size_t|void|const char \*|pe_plan \*|pe_comm \*)\s*(pe_[a-z_0-9]+)\(', hdr, flags=re.M))
    assert declared, 'no declarations parsed'
    bound = {name for name, _, _ in L.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pe_version() >= 100


def test_term_desc_abi_matches_ctypes_mirror():
    L, lib = _lib()
    assert lib.pe_abi_sizeof_term_desc() == C.sizeof(L.TermDesc)


def test_reference_chunk_arithmetic_of_the_wave_classes():
    """batch_num chunk i = rows [int(i*N/B), int((i+1)*N/B)) (semi:299-302) -- the host logic of train(iter, lr, batch_num)"""
    from pinn_elastodynamics_b200.models import _Base

    class T:
        global_n = 150397
    o = _Base.__new__(_Base)
    o._collo_term = T()
    for B in (1, 3, 7):
        ch = o._chunks(B)
        assert ch[0][0] == 0 and ch[-1][1] == 150397 and all(ch[i][1] == ch[i + 1][0] for i in range(B - 1))
        assert ch == [(int(i * 150397 / B), int((i + 1) * 150397 / B)) for i in range(B)]


def test_layout_pack_unpack_roundtrip():
    L, lib = _lib()
    for layers in ([3, 50, 50, 50, 50, 50, 5], [3] + 8 * [70] + [5], [3] + 8 * [100] + [7], [3] + 6 * [140] + [7], [3, 20, 20, 20, 20, 5], [3, 7, 5]):
        dims = (C.c_int * len(layers))(*layers)
        plan = lib.pe_plan_create(dims, len(layers), -1)
        assert plan, lib.pe_last_error()
        P = lib.pe_plan_param_count(plan)
        Pp = lib.pe_plan_param_count_padded(plan)
        assert P == sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1))
        assert Pp % 4 == 0 and Pp >= P
        flat = np.arange(1, P + 1, dtype=np.float32)
        padded = np.full(Pp, np.nan, np.float32)
        assert lib.pe_pack_params(plan, flat.ctypes.data, padded.ctypes.data) == 0
        assert np.isfinite(padded).all()
        back = np.zeros(P, np.float32)
        assert lib.pe_unpack_params(plan, padded.ctypes.data, back.ctypes.data) == 0
        np.testing.assert_array_equal(back, flat)
        # W_l rows are 16-byte aligned and in reference (in, out) order
        for l in range(len(layers) - 1):
            off, ld = lib.pe_plan_weight_offset(plan, l), lib.pe_plan_weight_ld(plan, l)
            assert off % 4 == 0 and ld % 4 == 0 and ld >= layers[l + 1]
        o0 = lib.pe_plan_weight_offset(plan, 0)
        assert padded[o0 + 1] == 2.0 and padded[o0 + lib.pe_plan_weight_ld(plan, 0)] == layers[1] + 1.0
        assert padded.astype(np.float64).sum() == flat.astype(np.float64).sum()
        lib.pe_plan_destroy(plan)


def test_plan_rejects_bad_networks():
    L, lib = _lib()
    for layers in ([2, 10, 5], [3, 10], [3, 10, 11], [3] + 17 * [8] + [5]):
        dims = (C.c_int * len(layers))(*layers)
        assert not lib.pe_plan_create(dims, len(layers), -1)
        assert lib.pe_last_error()


def test_device_entry_points_fail_loudly_without_device():
    L, lib = _lib()
    dims = (C.c_int * 3)(3, 10, 5)
    plan = lib.pe_plan_create(dims, 3, -1)
    d = L.TermDesc(); d.kind = L.RES_F5; d.n_global = 1; d.ld = 3
    rc = lib.pe_residual_loss_grad(plan, C.byref(d), 5, 0, None, 0, None, None, None, None, None, 0, None)
    assert rc != 0 and b'without a device' in lib.pe_last_error()
    lib.pe_plan_destroy(plan)


def test_shard_range_is_reference_chunk_arithmetic():
    from pinn_elastodynamics_b200.engine import shard_range
    for n in (1, 5, 31, 1000, 150397):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert r == [(int(k * n / world), int((k + 1) * n / world)) for k in range(world)]      # semi:300-302


def test_models_require_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import pinn_elastodynamics_b200 as pe
    with pytest.raises(Exception):
        pe.PINN(np.zeros((4, 3)), np.zeros((4, 3)), None, None, None, None, None, None, [3, 8, 5], None, None, None, None, verbose=False)


def test_engine_support_matrix():
    """which residual kinds / networks each engine takes (pe_engine_supported): the SIMT engine everything; the tcgen05 engines the F5 (K = 5,
    5 outputs) and F7 (K = 4, 7 outputs) collocation terms of networks with hidden widths <= 56 (the fp16-pair engine: >= 2 hidden layers)."""
    L, lib = _lib()

    def sup(layers, kind, K, engine):
        dims = (C.c_int * len(layers))(*layers)
        plan = lib.pe_plan_create(dims, len(layers), -1)
        assert plan, lib.pe_last_error()
        r = lib.pe_engine_supported(plan, kind, K, L.ENGINES[engine])
        lib.pe_plan_destroy(plan)
        return r
    f5, f7, w100 = [3] + 5 * [50] + [5], [3] + 5 * [50] + [7], [3] + 8 * [100] + [7]
    assert set(L.ENGINES) == {'simt', 'tc3s', 'tc1s', 'tcf', 'tcf16', 'auto'} and L.ENGINES['auto'] == L.ENGINES['tcf']
    for eng in ('simt', 'tc3s', 'tc1s', 'tcf', 'tcf16', 'auto'):
        assert sup(f5, L.RES_F5, 5, eng) == 1 and sup(f7, L.RES_F7, 4, eng) == 1
        assert sup(f5, L.RES_TRACTION, 1, eng) == (1 if eng == 'simt' else 0)      # data terms stay on the SIMT engine
        assert sup(f7, L.RES_COLS, 1, eng) == (1 if eng == 'simt' else 0)
        assert sup(w100, L.RES_F7, 4, eng) == (1 if eng == 'simt' else 0)          # hidden width > 56: SIMT only
        assert sup(f5, L.RES_F7, 4, eng) == sup(f7, L.RES_F5, 5, eng)              # K / output count must match the formulation (0 on tensor cores)
    assert sup([3, 56, 56, 5], L.RES_F5, 5, 'tcf') == 1 and sup([3, 57, 56, 5], L.RES_F5, 5, 'tcf') == 0
    assert sup([3, 50, 5], L.RES_F5, 5, 'tc3s') == 1 and sup([3, 50, 5], L.RES_F5, 5, 'tcf') == 0      # one hidden layer
    # slots / scratch queries are engine-aware
    dims = (C.c_int * len(f5))(*f5)
    plan = lib.pe_plan_create(dims, len(f5), -1)
    assert lib.pe_plan_slots(plan, 50000, 5, L.ENGINES['tcf']) == lib.pe_plan_slots(plan, 50000, 5, L.ENGINES['tc3s']) == 148
    assert lib.pe_plan_slots(plan, 0, 5, L.ENGINES['tcf']) == 1 == lib.pe_plan_slots(plan, 0, 5, L.ENGINES['simt'])      # an empty shard still owns one (zero-filled) slot
    assert lib.pe_plan_scratch_floats(plan, 50000, 5, L.ENGINES['tcf']) == lib.pe_plan_scratch_floats(plan, 50000, 5, L.ENGINES['tc3s'])
    lib.pe_plan_destroy(plan)


def test_bench_weights_equal_the_oracle_arm_weights():
    """bench.py's GPU arm initialises with the package's own Xavier routine (nothing under oracle/ on that arm); the reference arm and
    the cpu_baseline leg feed the oracle R.xavier_params(seed=1111): both must be the same arrays."""
    from oracle import ref_torch as R
    from pinn_elastodynamics_b200.models import xavier_init_lists
    layers = [3] + 5 * [50] + [5]
    a = R.xavier_params(layers, seed=1111)
    b = xavier_init_lists(layers, np.random.default_rng(1111))
    assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0])) and all(np.array_equal(x, y) for x, y in zip(a[1], b[1]))
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py')).read()
    main_arm = src[src.index('import pinn_elastodynamics_b200 as pe'):src.index("# ---- e2e")]
    assert 'from oracle' not in main_arm and 'import oracle' not in main_arm
