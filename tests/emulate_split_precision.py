"""CPU study (not a pytest file): accuracy of split-operand tensor-core schemes for one layer GEMM Z = A W (128 points x 50 x 50),
against float64, for operand magnitudes as they occur in the forward streams (O(1)) and in the adjoint streams (1e-4 .. 1e-9).

    python tests/emulate_split_precision.py            -> table on stdout (profiles/r1_split_precision_study.txt)

Schemes (products exact, accumulation in float32 as the tensor core's accumulator; its truncating adds are not modelled):
  fp32        : plain float32 GEMM (the SIMT engine)
  tf32x3      : the current engines: trunc_tf32(A) Whi + trunc_tf32(A) Wlo + bf16(A - trunc_tf32(A)) bf16(W)
  tf32x1      : single-pass TF32 (the `tc1*` engines)
  bf16x3      : bf16 hi/mid of both operands, hh + hm + mh
  fp16x3      : fp16 hi/lo of both operands, hh + hl + lh, lo parts unscaled (fp16 subnormals below 6e-5)
  fp16x3s     : the same with the lo parts scaled by 2^11 (kept in a second accumulator, combined in the epilogue)
  f16b16x3    : fp16 hi + bf16 lo of both operands (kind::f16 takes the A and B formats independently): hh + hl + lh in ONE accumulator
Round-2 question: can one 16-bit split serve forward, adjoint and weight-gradient GEMMs without the
TF32 residue operand and the bf16 conversion pass?
"""
import numpy as np


def trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16            # round to nearest even
    return r.astype(np.uint32).view(np.float32)


def fp16(x):
    return x.astype(np.float32).astype(np.float16).astype(np.float32)


def mm(a, b):
    """exact products, float32 accumulation over k in MMA-sized steps of 8"""
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k in range(0, a.shape[1], 8):
        acc = (acc.astype(np.float64) + a[:, k:k + 8].astype(np.float64) @ b[k:k + 8].astype(np.float64)).astype(np.float32)
    return acc


def schemes(A, W):
    A = A.astype(np.float32); W = W.astype(np.float32)
    out = {'fp32': mm(A, W)}
    Ah = trunc_tf32(A); Al = bf16(A - Ah)
    Wh = trunc_tf32(W); Wl = trunc_tf32(W - Wh)
    out['tf32x3'] = (mm(Ah, Wh).astype(np.float64) + mm(Ah, Wl) + mm(Al, bf16(W))).astype(np.float32)
    out['tf32x1'] = mm(Ah, Wh)
    Ab, Wb = bf16(A), bf16(W)
    Am, Wm = bf16(A - Ab), bf16(W - Wb)
    out['bf16x3'] = (mm(Ab, Wb).astype(np.float64) + mm(Ab, Wm) + mm(Am, Wb)).astype(np.float32)
    Af, Wf = fp16(A), fp16(W)
    out['fp16x3'] = (mm(Af, Wf).astype(np.float64) + mm(Af, fp16(W - Wf)) + mm(fp16(A - Af), Wf)).astype(np.float32)
    s = np.float32(2048.0)
    Als, Wls = fp16((A - Af) * s), fp16((W - Wf) * s)
    out['fp16x3s'] = (mm(Af, Wf).astype(np.float64) + (mm(Af, Wls).astype(np.float64) + mm(Als, Wf)) / s).astype(np.float32)
    out['f16b16x3'] = (mm(Af, Wf).astype(np.float64) + mm(Af, bf16(W - Wf)) + mm(bf16(A - Af), Wf)).astype(np.float32)
    return out


def main():
    rng = np.random.default_rng(0)
    W = np.clip(rng.standard_normal((50, 50)), -2, 2) * np.sqrt(2 / 100)
    print('%-34s' % 'operand A (128 x 50)' + ''.join('%10s' % k for k in ('fp32', 'tf32x3', 'tf32x1', 'bf16x3', 'fp16x3', 'fp16x3s', 'f16b16x3')))
    cases = [('tanh activations, |a| < 1', np.tanh(rng.standard_normal((128, 50))))]
    for mag in (1e-2, 1e-4, 1e-6, 1e-8):
        cases.append(('adjoint-like, magnitude %.0e' % mag, rng.standard_normal((128, 50)) * mag))
    cases.append(('mixed magnitudes 1 .. 1e-8 per row', rng.standard_normal((128, 50)) * 10.0 ** rng.uniform(-8, 0, (128, 1))))
    for name, A in cases:
        ref = A.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64)
        res = schemes(A, W)
        # error of each output row relative to that row's max (rows = points: each point's jets live at its own scale)
        row = lambda z: float(np.max(np.abs(z - ref).max(1) / np.abs(ref).max(1)))
        print('%-34s' % name + ''.join('%10.1e' % row(res[k]) for k in ('fp32', 'tf32x3', 'tf32x1', 'bf16x3', 'fp16x3', 'fp16x3s', 'f16b16x3')))


if __name__ == '__main__':
    main()
