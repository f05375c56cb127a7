"""Hardware pins of the tcgen05 conventions pe_tc.cu relies on (exact integer GEMMs through the UMMA probe harness,
pinn_elastodynamics_b200/csrc/umma_probe.cu).  The exploratory hypotheses that FAIL by design (swapped LBO/SBO, MN-major TF32
without the 32-bit-base swizzle) live in tests/probe_umma.py; profiles/r1_umma_probe.txt records their outcome."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RELIED_ON = [
    'kmajor tf32 K=56 N=64',               # forward / adjoint A32 x Whi, Wlo (SS, K-major, no swizzle, SBO = 128, LBO = chunk stride)
    'kmajor tf32 K=56 N=16',               # output layer (N = 16)
    'bf16 TS K=64',                        # Alo (bf16 pairs in tensor memory, element 2c in the low half) x Wbf16
    'mixed tf32 SS + bf16 TS accumulate',  # both kinds into the same fp32 accumulator columns
    'mnmajor bf16 M=64 N=56',              # weight gradient: MN-major 16-bit operands [chunk of 8][point][8], M = 64 lane map
    'cp 128x256b + tf32 TS',               # tcgen05.cp operand copy + A-from-TMEM TF32 (kept for the TS experiments)
]


@pytest.mark.parametrize('name', RELIED_ON)
def test_umma_convention(name):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import probe_umma
    if not os.path.exists(probe_umma.LIB):
        pytest.skip('libumma_probe.so not built')
    idx = [n for n, _ in probe_umma.TESTS].index(name)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'probe_umma.py'), str(idx)], capture_output=True, text=True, timeout=120)
    assert r.stdout.startswith('PASS ' + name), r.stdout + r.stderr


def test_tf32_operands_are_truncated():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import probe_umma
    if not os.path.exists(probe_umma.LIB):
        pytest.skip('libumma_probe.so not built')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'probe_umma.py'), str(len(probe_umma.TESTS) - 1)], capture_output=True, text=True, timeout=120)
    assert '-> 1.0 ' in r.stdout, r.stdout + r.stderr       # the 3-term split in pe_tc.cu assumes truncation (lo = x - trunc(x) >= 0)
