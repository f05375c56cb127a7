"""CPU model of the operand arithmetic of the fp16-pair tcgen05 engine (csrc/pe_tcf.cu): the split / join of an operand, the three-product
GEMM with the scale-input-d step, the per-tile power-of-two seed scale, and the tile layout + drain formula of the layer-1 gradient MMAs.
The kernel itself is checked on the GPU (tests/test_gpu_tcf.py); these tests pin the ALGEBRA the kernel's comments state, in numpy, so that
the tolerance class (fp32-grade, ~3e-7 per GEMM) and the drain formulas can be re-derived without a device."""
import numpy as np

LO = 2048.0


def f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def split(x):
    """Xh = fp16(X), Xl = fp16((X - Xh) 2^11)   (split2 in pe_tcf.cu; X - Xh is exact in fp32)"""
    x = np.asarray(x, np.float32)
    h = f16(x)
    return h, f16((x - h) * np.float32(LO))


def join(h, l):
    return (l * np.float32(1.0 / LO) + h).astype(np.float32)


def mma(acc, a, b, scale_d=False):
    """one kind::f16 MMA: exact products of fp16 values, fp32 accumulator; scale_d: D = A B + D 2^-11"""
    d = acc.astype(np.float64) * (2.0 ** -11 if scale_d else 1.0)
    return (d + a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)


def pair_gemm(A, W):
    """A W ~= Ah Wh + 2^-11 (Ah Wl + Al Wh): cross products of every K-step first, then the hi x hi products, the first of them
    in the scale-input-d form (issue_group in pe_tcf.cu)"""
    Ah, Al = split(A)
    Wh, Wl = split(W)
    acc = np.zeros((A.shape[0], W.shape[1]), np.float32)
    for k in range(0, A.shape[1], 16):
        acc = mma(acc, Ah[:, k:k + 16], Wl[k:k + 16])
        acc = mma(acc, Al[:, k:k + 16], Wh[k:k + 16])
    for i, k in enumerate(range(0, A.shape[1], 16)):
        acc = mma(acc, Ah[:, k:k + 16], Wh[k:k + 16], scale_d=(i == 0))
    return acc


def test_split_join_keeps_22_bits_inside_the_fp16_range():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(20000) * 10.0 ** rng.uniform(-3, 3, 20000)).astype(np.float32)
    x = x[(np.abs(x) > 2e-4) & (np.abs(x) < 6e4)]
    h, l = split(x)
    assert np.all(np.abs(join(h, l) - x) <= np.abs(x) * 2.0 ** -21)
    assert np.all(np.abs(l) <= 1.0 * LO * np.abs(x) * 2.0 ** -10)          # the scaled residue is itself a normal fp16 number


def test_three_product_gemm_is_fp32_grade():
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, (128, 64)).astype(np.float32)                    # activations of a tanh layer
    W = (rng.standard_normal((64, 64)) * np.sqrt(2.0 / 100)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64)
    rowmax = np.abs(ref).max(1, keepdims=True)
    e_pair = np.abs(pair_gemm(A, W) - ref) / rowmax
    e_fp32 = np.abs((A @ W).astype(np.float64) - ref) / rowmax
    e_half = np.abs(mma(np.zeros((128, 64), np.float32), f16(A), f16(W)) - ref) / rowmax
    assert e_pair.max() <= 6e-7 and e_pair.max() <= 4 * e_fp32.max() + 1e-7      # fp32 class (the CPU study measured 2.6e-7 against 1.5e-7)
    assert e_half.max() >= 1e-4                                                # single fp16 products (engine 'tcf16') are three orders worse


def test_seed_scale_is_a_power_of_two_that_brings_the_largest_seed_into_1_2():
    for m in (3e-9, 7.3e-5, 0.9, 1.0, 1.5, 6e4):
        eb = (np.float32(m).view(np.uint32) >> 23) & 0xFF
        eb = min(max(int(eb), 2), 252)
        sigma = np.uint32((254 - eb) << 23).view(np.float32)
        inv_sigma = np.uint32(eb << 23).view(np.float32)
        assert sigma * inv_sigma == 1.0
        assert 1.0 <= m * float(sigma) < 2.0


def test_layer1_gradient_tiles_and_drain_formula():
    """issuer 2: T0 = [Zh | Zl]^T [xh yh th 1 | xl yl tl 0] (value stream), T_k = [Zh_k | Zl_k]^T 1 (k = d/dx, d/dy, d/dt); drain: rows r < d1
    hold the Zh sums of unit r, rows 56 + j the Zl sums of unit j (scaled by 2^11); columns 4..6 carry the coordinate residues (2^11)."""
    rng = np.random.default_rng(2)
    n, d1 = 128, 50
    xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    Z = [(rng.standard_normal((n, d1)) * 10.0 ** rng.uniform(-2, 0)).astype(np.float32) for _ in range(4)]      # Zbar_1 of value, x, y, t streams
    sc = np.array([0.7, 1.3, 0.05], np.float32)                                                                 # input scale 2 / (ub - lb)
    ch, cl = split(xyz)
    X = np.concatenate([ch, np.ones((n, 1), np.float32), cl, np.zeros((n, 1), np.float32)], 1)                  # [n][8]
    tiles = []
    for k in range(4):
        zh, zl = split(Z[k])
        A = np.concatenate([zh, np.zeros((n, 6), np.float32), zl], 1).T                                         # rows 0..55 hi units, 56.. lo units
        B = X if k == 0 else np.ones((n, 8), np.float32)
        acc = np.zeros((A.shape[0], 8), np.float32)
        for s in range(0, n, 16):
            acc = mma(acc, A[:, s:s + 16], B[s:s + 16])
        tiles.append(acc)
    g = np.zeros((4, d1), np.float64)                                                                           # rows: dW_0[0..2][j], db_0[j]
    for r in range(d1):
        for part, row, scale in ((0, r, 1.0), (1, 56 + r, 1.0 / LO)):
            v0, v1, v2, v3 = (t[row] for t in tiles)
            gx = v0[4] / LO + v0[0] + sc[0] * v1[0]
            gy = v0[5] / LO + v0[1] + sc[1] * v2[0]
            gt = v0[6] / LO + v0[2] + sc[2] * v3[0]
            g[:, r] += scale * np.array([gx, gy, gt, v0[3]])
    Zd = [z.astype(np.float64) for z in Z]
    ref = np.stack([xyz[:, 0].astype(np.float64) @ Zd[0] + sc[0] * Zd[1].sum(0), xyz[:, 1].astype(np.float64) @ Zd[0] + sc[1] * Zd[2].sum(0),
                    xyz[:, 2].astype(np.float64) @ Zd[0] + sc[2] * Zd[3].sum(0), Zd[0].sum(0)])
    assert np.abs(g - ref).max() <= 2e-6 * np.abs(ref).max()
