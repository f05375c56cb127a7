"""CPU study (not a pytest file): end-to-end loss / gradient error of the collocation residual (F5, 5x50 net) when the three GEMM
families of the kernel -- forward layer GEMMs, adjoint GEMMs, weight-gradient GEMMs -- run on split-operand tensor-core arithmetic,
everything else in float32, against the float64 jet oracle (oracle/jet_numpy.py).

    python tests/emulate_engine_precision.py [n_points] [DEPTHxWIDTH, default 5x50]      -> table on stdout (profiles/r*_engine_precision_study*.txt)

Schemes:
  fp32      : float32 GEMMs (what the SIMT engine does)
  tf32x3    : the shipped tcgen05 engines: layer / adjoint GEMMs  trunc_tf32(A) Whi + trunc_tf32(A) Wlo + bf16(A - Ahi) bf16(W),
              weight gradient on bf16 hi / mid of both operands (hh + hm + mh)
  f16b16x3  : round-2 candidate: every GEMM on fp16 hi + bf16 lo of both operands (hh + hl + lh, one
              accumulator).  `seeds x 2^k` multiplies the adjoint seeds by a power of two and divides the gradient by it at the end:
              negative k mimics larger global point counts (seeds ~ 1/N: 2^-5 ~ 50 k points, 2^-9 ~ 1 M points for this 2,048-point
              sample), positive k is the loss scaling that keeps the fp16 hi parts normal
  f16pair   : what round 2 ships (engine 'tcf', csrc/pe_tcf.cu): every operand as fp16 hi + fp16 lo x 2^11, per K-step the two cross
              products first, then the hi x hi products with the scale-input-d step (D = A B + D 2^-11), ONE accumulator; adjoint seeds of
              every 128-point tile scaled by a power of two into [1, 2), gradient tiles (hh | hl, lh) drained in two adds times 1 / sigma.
              `+ RZ`: the accumulator update D + (exact sum of the K = 16 products) is rounded toward zero, as probed on B200
              (profiles/r2_umma_probe.txt: the accumulator truncates)
The tensor core's own accumulation is modelled as float32 adds of exact K = 8 / 16 partial products (its truncating adder only in the
`+ RZ` rows).  Measured on B200: tf32x3 ~7e-6 (round 1); f16pair: loss terms <= 7e-7, gradient blocks <= 5e-6
(profiles/r2_refgold_report_tcf.jsonl) -- the `f16pair + RZ` row.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jet_numpy as J          # noqa: E402
from oracle import ref_torch as R          # noqa: E402

f32 = np.float32


def trunc_tf32(x):
    return (x.astype(f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def bf16(x):
    u = x.astype(f32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(f32)


def fp16(x):
    return x.astype(f32).astype(np.float16).astype(f32)


def mm(a, b, kstep):
    """a [..., m, k] @ b [k, n]: exact products, float32 accumulation over K steps of `kstep`"""
    acc = np.zeros(a.shape[:-1] + (b.shape[1],), f32)
    for k in range(0, a.shape[-1], kstep):
        acc = (acc.astype(np.float64) + a[..., k:k + kstep].astype(np.float64) @ b[k:k + kstep].astype(np.float64)).astype(f32)
    return acc


def to_f32_rz(x64):
    """float64 -> float32, rounded toward zero"""
    f = x64.astype(f32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(f, f32(0)), f).astype(f32)


def split16(x):
    x = x.astype(f32)
    h = fp16(x)
    return h, fp16((x - h) * f32(2048.0))


def mma_acc(acc, a, b, rz, scale_d=False):
    """acc <- a @ b + acc [* 2^-11]: exact K = 16 products, one rounding (nearest, or toward zero) per accumulator update"""
    d = acc.astype(np.float64) * (2.0 ** -11 if scale_d else 1.0) + a.astype(np.float64) @ b.astype(np.float64)
    return to_f32_rz(d) if rz else d.astype(f32)


def gemm_pair(A, W, rz):
    Ah, Al = split16(A)
    Wh, Wl = split16(W)
    acc = np.zeros(A.shape[:-1] + (W.shape[1],), f32)
    ks = range(0, A.shape[-1], 16)
    for k in ks:
        acc = mma_acc(acc, Ah[..., k:k + 16], Wl[k:k + 16], rz)
        acc = mma_acc(acc, Al[..., k:k + 16], Wh[k:k + 16], rz)
    for i, k in enumerate(ks):
        acc = mma_acc(acc, Ah[..., k:k + 16], Wh[k:k + 16], rz, scale_d=(i == 0))
    return acc


def wgrad_pair(A, Z, inv_sigma, rz, tile=128):
    """per 128-point tile: hh = Ah^T Zh, hl = Ah^T Zl, lh = Al^T Zh accumulated over the streams and the 8 K-steps of 16 points; drained
    into the float32 slot as (hh + hl / 2048) / sigma, then (lh / 2048) / sigma"""
    K, N, _ = A.shape
    out = np.zeros((A.shape[2], Z.shape[2]), f32)
    for t, p0 in enumerate(range(0, N, tile)):
        hh = np.zeros_like(out); hl = np.zeros_like(out); lh = np.zeros_like(out)
        for k in range(K):
            ah, al = split16(A[k, p0:p0 + tile])
            zh, zl = split16(Z[k, p0:p0 + tile])
            for s in range(0, ah.shape[0], 16):
                hh = mma_acc(hh, ah[s:s + 16].T, zh[s:s + 16], rz)
                hl = mma_acc(hl, ah[s:s + 16].T, zl[s:s + 16], rz)
                lh = mma_acc(lh, al[s:s + 16].T, zh[s:s + 16], rz)
        isg = f32(inv_sigma[t])
        out = (out + ((hl * f32(1 / 2048.0) + hh).astype(f32) * isg).astype(f32)).astype(f32)
        out = (out + (lh * (f32(1 / 2048.0) * isg)).astype(f32)).astype(f32)
    return out


def add32(*terms):
    out = terms[0].astype(f32)
    for t in terms[1:]:
        out = (out + t.astype(f32)).astype(f32)
    return out


def gemm(A, W, scheme):
    A = A.astype(f32); W = W.astype(f32)
    if scheme == 'fp32':
        return mm(A, W, 1)
    if scheme == 'tf32x3':
        Ah = trunc_tf32(A); Wh = trunc_tf32(W)
        return add32(mm(Ah, Wh, 8), mm(Ah, trunc_tf32(W - Wh), 8), mm(bf16(A - Ah), bf16(W), 16))
    if scheme == 'f16b16x3':
        Ah, Wh = fp16(A), fp16(W)
        return add32(mm(Ah, Wh, 16), mm(Ah, bf16(W - Wh), 16), mm(bf16(A - Ah), Wh, 16))
    raise ValueError(scheme)


def wgrad(A, Z, scheme, tile=128):
    """dW = sum_k sum_p A[k,p,i] Z[k,p,j]: per 128-point tile on the tensor core (K = 16 points per MMA, streams accumulate in the same
    tile), tiles added in float32 (the per-CTA slot); the slot reduction across CTAs is not modelled (one slot)."""
    A = A.astype(f32); Z = Z.astype(f32)
    K, N, _ = A.shape
    out = np.zeros((A.shape[2], Z.shape[2]), f32)
    for p0 in range(0, N, tile):
        a = np.concatenate([A[k, p0:p0 + tile] for k in range(K)], 0)        # [K*tile, i]
        z = np.concatenate([Z[k, p0:p0 + tile] for k in range(K)], 0)
        if scheme == 'fp32':
            t = mm(a.T, z, 1)
        elif scheme == 'tf32x3':
            ah, zh = bf16(a), bf16(z)
            t = add32(mm(ah.T, zh, 16), mm(ah.T, bf16(z - zh), 16), mm(bf16(a - ah).T, zh, 16))
        else:
            ah, zh = fp16(a), fp16(z)
            t = add32(mm(ah.T, zh, 16), mm(ah.T, bf16(z - zh), 16), mm(bf16(a - ah).T, zh, 16))
        out = (out + t).astype(f32)
    return out


def run(X, Ws, bs, scheme, seed_scale=1.0, E=20.0, mu=0.25, rho=1.0, w=10.0):
    """float32 restatement of oracle/jet_numpy.forward_jets / backward_jets with the GEMMs of layers >= 1 routed through `scheme`
    (layer 0 has three inputs and is plain FFMA in every engine)."""
    K, N = 5, X.shape[0]
    pair, pair_rz = scheme.startswith('f16pair'), scheme.endswith('RZ')
    A = np.zeros((K, N, 3), f32)
    A[0] = X.astype(f32); A[1, :, 0] = 1; A[2, :, 1] = 1; A[3, :, 2] = 1
    ins, acts = [], []
    L = len(Ws)
    for l in range(L):
        W = Ws[l].astype(f32); b = bs[l].astype(f32).reshape(1, -1)
        ins.append(A)
        Z = mm(A, W, 1) if l == 0 else (gemm_pair(A, W, pair_rz) if pair else gemm(A, W, scheme))
        Z[0] = Z[0] + b
        if l == L - 1:
            Y = Z
            break
        a = np.tanh(Z[0]).astype(f32); s = (1 - a * a).astype(f32)
        An = np.empty_like(Z)
        An[0] = a
        for k in (1, 2, 3):
            An[k] = s * Z[k]
        An[4] = s * Z[4] - 2 * a * s * Z[3] * Z[3]
        acts.append((a, s, Z))
        A = An.astype(f32)
    f = J.residual_f5(Y, f32(E), f32(mu), f32(rho)).astype(f32)
    l_uv = float((f[0].astype(np.float64) ** 2 + f[1].astype(np.float64) ** 2).sum() / N)
    l_s = float((f[2:].astype(np.float64) ** 2).sum() / N)
    fbar = (2 * f / N * w * seed_scale).astype(f32)
    Zbar = J.residual_f5_adjoint(fbar, Y, f32(E), f32(mu), f32(rho)).astype(f32)
    dWs, dbs = [None] * L, [None] * L
    inv_sigma = None
    if pair:          # per-tile power-of-two seed scale: largest |seed| of the tile in [1, 2)
        nt = (N + 127) // 128
        m = np.array([np.abs(Zbar[:, t * 128:(t + 1) * 128]).max() for t in range(nt)], f32)
        eb = np.clip((m.view(np.uint32) >> 23) & 0xFF, 2, 252).astype(np.uint32)
        eb[m == 0] = 127
        sigma = ((254 - eb) << 23).astype(np.uint32).view(f32)
        inv_sigma = (eb << 23).astype(np.uint32).view(f32)
        Zbar = (Zbar * np.repeat(sigma, 128)[:N][None, :, None]).astype(f32)
    for l in range(L - 1, -1, -1):
        Ain = ins[l]
        if pair:
            dWs[l] = wgrad_pair(Ain, Zbar, inv_sigma, pair_rz)
            ones = np.ones((1, N, 1), f32)
            dbs[l] = wgrad_pair(ones, Zbar[0:1], inv_sigma, pair_rz)[0]
        else:
            dWs[l] = wgrad(Ain, Zbar, 'fp32' if l == 0 else scheme)
            dbs[l] = Zbar[0].sum(0, dtype=f32)
        if l == 0:
            break
        Abar = gemm_pair(Zbar, Ws[l].astype(f32).T.copy(), pair_rz) if pair else gemm(Zbar, Ws[l].astype(f32).T.copy(), scheme)
        a, s, Z = acts[l - 1]
        Zb = np.empty_like(Abar)
        for k in (1, 2, 4):
            Zb[k] = s * Abar[k]
        acc = Z[1] * Abar[1] + Z[2] * Abar[2] + Z[3] * Abar[3] + Z[4] * Abar[4]
        zv = Abar[0] - 2 * a * acc
        Zb[3] = s * Abar[3] - 4 * a * s * Z[3] * Abar[4]
        zv = zv - 2 * (1 - 3 * a * a) * Z[3] * Z[3] * Abar[4]
        Zb[0] = s * zv
        Zbar = Zb.astype(f32)
    g = np.concatenate([d.astype(np.float64).ravel() for d in dWs] + [d.astype(np.float64).ravel() for d in dbs]) / seed_scale
    return l_uv, l_s, g


def block_err(g, gref, layers):
    out, o = [], 0
    for n in [layers[i] * layers[i + 1] for i in range(len(layers) - 1)] + layers[1:]:
        out.append(np.abs(g[o:o + n] - gref[o:o + n]).max() / np.abs(gref[o:o + n]).max()); o += n
    return max(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    depth, width = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '5x50').split('x'))      # e.g. 8x70: the reference's shipped plate net
    rng = np.random.default_rng(1111)
    P = rng.uniform([0, 0, 0], [.5, .5, 10], (2 * n, 3))
    X = P[np.hypot(P[:, 0], P[:, 1]) > 0.1][:n]            # the bench workload's point distribution
    layers = [3] + depth * [width] + [5]
    Ws, bs = R.xavier_params(layers, seed=1111)
    print('F5 collocation term, %dx%d net, %d points of the bench distribution; errors against the float64 jet oracle' % (depth, width, X.shape[0]))
    print('%-28s %12s %12s %16s' % ('scheme', 'loss_f_uv', 'loss_f_s', 'grad (block max)'))
    for label, Wl, bl in (('Xavier init', Ws, bs), ('Xavier x 1.5, biases 0.1', [w * 1.5 for w in Ws], [b + 0.1 for b in bs])):
        luv, ls, dW, db = J.loss_grad_residual('f5', X, Wl, bl, 10.0, 10.0, 20.0, 0.25, 1.0)
        gref = np.concatenate([d.ravel() for d in dW] + [d.ravel() for d in db])
        print('-- weights:', label, ' (loss_f_uv %.3e, loss_f_s %.3e, typical |adjoint seed| ~ %.1e)' % (luv, ls, 2 * 10.0 / n * np.sqrt(ls)))
        for scheme, sc in (('fp32', 1.0), ('tf32x3', 1.0), ('f16pair', 1.0), ('f16pair + RZ', 1.0), ('f16pair + RZ', 2.0 ** -9), ('f16b16x3', 1.0), ('f16b16x3', 2.0 ** -5), ('f16b16x3', 2.0 ** -9), ('f16b16x3', 2.0 ** -14), ('f16b16x3', 2.0 ** 6)):
            a, b, g = run(X, Wl, bl, scheme, sc)
            tag = scheme + ('' if sc == 1.0 else ', seeds x 2^%d' % int(np.log2(sc)))
            print('%-28s %12.1e %12.1e %16.1e' % (tag, abs(a - luv) / luv, abs(b - ls) / ls, block_err(g, gref, layers)))


if __name__ == '__main__':
    main()
