"""ctypes binding of libpinn_elasto.so (the C ABI in include/pinn_elasto.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C pinn_elastodynamics_b200/csrc`.
There is NO fallback: if the shared object is missing or does not export a symbol, importing the hot
path fails loudly (a silent CPU path would void every parity claim).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# PE_LIB_PATH: A/B timing of experimental builds of the same C ABI (tests/build_variant.sh); unset in normal use
LIB_PATH = os.environ.get('PE_LIB_PATH') or os.path.join(HERE, 'libpinn_elasto.so')

PE_MAX_LAYERS = 16
PE_MAX_TERMS = 8
PE_MAX_COLS = 8
PE_MAX_PEERS = 8
PE_IPC_HANDLE_BYTES = 64
PE_TILE_POINTS = 32
PE_TC_TILE = 128

RES_F5, RES_F7, RES_COLS, RES_TRACTION, RES_DT = 0, 1, 2, 3, 4
ENGINE_SIMT_FP32, ENGINE_TCS_TF32X3, ENGINE_TCS_TF32, ENGINE_TCF, ENGINE_TCF_F16FWD = 0, 5, 6, 8, 9
# 'tcf'  : fp16-pair tcgen05 engine (csrc/pe_tcf.cu): fp32-grade (measured <= 7e-7 on loss terms, <= 5e-6 on gradient blocks, <= 8.3e-6 on the
#          reference's 20-step Adam curves: profiles/r2_refgold_report_tcf.jsonl) -- what 'auto' selects
# 'tc3s' : TF32x3 tcgen05 engine (csrc/pe_tcs.cu), measured 5e-6 .. 8e-6 / 3.3e-5: A/B partner only;  'tc1s': its single-pass TF32 mode
# 'simt' : fp32 FFMA engine, any width, every residual kind: the parity anchor
# 'tcf16': 'tcf' with single-product fp16 forward GEMMs (BASELINE config 3's 16-bit forward / fp32 gradient mode; ~1e-3 class)
ENGINES = {'simt': ENGINE_SIMT_FP32, 'tc3s': ENGINE_TCS_TF32X3, 'tc1s': ENGINE_TCS_TF32, 'tcf': ENGINE_TCF, 'tcf16': ENGINE_TCF_F16FWD,
           'auto': ENGINE_TCF}             # terms / networks the tensor-core engine does not implement run on the SIMT engine (engine.py build())


class TermDesc(C.Structure):
    """Mirror of `pe_term_desc` (include/pinn_elasto.h)."""
    _fields_ = [
        ('kind', C.c_int), ('n_global', C.c_int), ('ld', C.c_int),
        ('E', C.c_float), ('mu', C.c_float), ('rho', C.c_float), ('hole_r', C.c_float),
        ('in_scale', C.c_float * 3), ('in_shift', C.c_float * 3),
        ('ncols', C.c_int),
        ('col', C.c_int * PE_MAX_COLS), ('tgt', C.c_int * PE_MAX_COLS), ('term', C.c_int * PE_MAX_COLS),
        ('w', C.c_float * PE_MAX_COLS),
        ('aux_k', C.c_int),
    ]


# every symbol include/pinn_elasto.h declares: (name, restype, argtypes)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
SYMBOLS = [
    ('pe_version', _i, []),
    ('pe_last_error', C.c_char_p, []),
    ('pe_abi_sizeof_term_desc', _i, []),
    ('pe_plan_create', _vp, [C.POINTER(_i), _i, _i]),
    ('pe_plan_destroy', None, [_vp]),
    ('pe_plan_param_count', _i, [_vp]),
    ('pe_plan_param_count_padded', _i, [_vp]),
    ('pe_plan_weight_offset', _i, [_vp, _i]),
    ('pe_plan_bias_offset', _i, [_vp, _i]),
    ('pe_plan_weight_ld', _i, [_vp, _i]),
    ('pe_engine_supported', _i, [_vp, _i, _i, _i]),
    ('pe_plan_slots', _i, [_vp, _i, _i, _i]),
    ('pe_plan_scratch_floats', C.c_size_t, [_vp, _i, _i, _i]),
    ('pe_pack_params', _i, [_vp, _vp, _vp]),
    ('pe_unpack_params', _i, [_vp, _vp, _vp]),
    ('pe_residual_loss_grad', _i, [_vp, C.POINTER(TermDesc), _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    ('pe_residual_loss_grad_fused', _i, [_vp, C.POINTER(TermDesc), _i, _i, _vp, _i, _vp, C.POINTER(TermDesc), _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    ('pe_reduce_partials', _i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    ('pe_adam_step', _i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp]),
    ('pe_reduce_adam', _i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp]),
    ('pe_forward_fields', _i, [_vp, _i, _vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _i, _vp, _vp, _vp]),
    ('pe_forward_jets', _i, [_vp, _i, _vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp, _vp]),
    ('pe_comm_create', _vp, [_vp, _i, _i, _vp]),
    ('pe_comm_connect', _i, [_vp, _vp]),
    ('pe_comm_error', _i, [_vp]),
    ('pe_comm_disconnect', None, [_vp]),
    ('pe_comm_destroy', None, [_vp]),
    ('pe_reduce_peer', _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp]),
    ('pe_lbfgs_direction', _i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    ('pe_lbfgs_store_pair', _i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ('pe_vec_axpy', _i, [_i, _vp, _vp, _f, _vp, _vp]),
    ('pe_vec_dot_max', _i, [_i, _vp, _vp, _vp, _vp]),
    ('pe_debug_set_tcs_profile', None, [_vp]),
    ('pe_debug_set_tcf_profile', None, [_vp]),
    ('pe_debug_set_fields_engine', None, [_i]),
]

_lib = None


def build(verbose=False):
    """Compile the CUDA sources for sm_100a into libpinn_elasto.so (in-tree)."""
    cmd = ['make', '-C', os.path.join(HERE, 'csrc'), '-j4']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError('building libpinn_elasto.so failed')
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          f'(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pe_abi_sizeof_term_desc() != C.sizeof(TermDesc):
        raise ImportError(f'pe_term_desc ABI mismatch: library {lib.pe_abi_sizeof_term_desc()} bytes, ctypes mirror {C.sizeof(TermDesc)} bytes')
    _lib = lib
    return lib


class PeError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        msg = load().pe_last_error()
        raise PeError(f'{what}: {msg.decode() if msg else "error %d" % rc}')
