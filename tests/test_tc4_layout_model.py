"""CPU model of the operand addressing of the experimental fourth-generation engine (csrc/pe_tc4.cu).

The engine has not run on hardware yet, so its index algebra is checked here against a numpy model of how tcgen05.mma reads no-swizzle
16-bit operands through shared-memory descriptors.  The descriptor semantics are the ones pinned on hardware by tests/probe_umma.py
(profiles/r1_umma_probe.txt), restated as byte addresses:
    K-major  operand, element (row r, k) of one K = 16 MMA:  start + (r // 8) * SBO + (k // 8) * LBO + (r % 8) * 16 + (k % 8) * 2
    MN-major operand, element (mn m, k)                    :  start + (m // 8) * SBO + (k // 8) * LBO + (k % 8) * 16 + (m % 8) * 2
(the probe builds its images as [chunk of 8][row][8] and passes exactly these LBO / SBO values).  Everything below mirrors, constant for
constant, the offset formulas of pe_tc4.cu: plane layout (put4), image layout (tc4_image_kernel), the K-steps of issue_streams including
the LBO = 0 step, the weight-gradient and bias-gradient MMAs.  Values are small integers (exact in fp16 and bf16), so every GEMM must come out exact.
"""
import numpy as np

F_CH, F_PLANE = 2048, 7 * 2048
F_STREAM = 2 * F_PLANE
F_IMG_HALF = 8192
F_ONES, F_ACT = 0, 2048
NS = 5
F_R = F_ACT + 5 * F_STREAM
SMEM = F_R + 16384 + 2 * F_STREAM + 4096


def put(smem, off, v):
    """store one 16-bit element (we keep its value as a float in a parallel array indexed by byte offset / 2)"""
    smem[off // 2] = v


def plane_off(k, unit, p, lo=False):
    """pe_tc4.cu put4: o = k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8 with c4 = unit // 4, element (unit % 4) of the uint2"""
    c4, e = unit // 4, unit % 4
    return F_ACT + k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8 + e * 2 + (F_PLANE if lo else 0)


def read_kmajor(smem, start, lbo, sbo, rows):
    out = np.zeros((rows, 16))
    for r in range(rows):
        for k in range(16):
            out[r, k] = smem[(start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2) // 2]
    return out


def read_mnmajor(smem, start, lbo, sbo, mn):
    out = np.zeros((mn, 16))
    for m in range(mn):
        for k in range(16):
            out[m, k] = smem[(start + (m // 8) * sbo + (k // 8) * lbo + (k % 8) * 16 + (m % 8) * 2) // 2]
    return out


def image(W, adjoint):
    """tc4_image_kernel: forward element (i, j) at (i >> 3) * 512 + j * 8 + (i & 7); adjoint at (j >> 3) * 512 + i * 8 + (j & 7) (16-bit elements)"""
    din, dout = W.shape
    img = np.zeros(64 * 64)
    for i in range(64):
        for j in range(64):
            w = W[i, j] if (i < din and j < dout) else 0.0
            img[((j >> 3) * 512 + i * 8 + (j & 7)) if adjoint else ((i >> 3) * 512 + j * 8 + (i & 7))] = w
    return img


def layer_gemm(smem, img_off, k, kdim, N=64):
    """issue_streams for one stream, hi x hi only (the hl / lh products use the same addressing with the lo plane / lo image)"""
    D = np.zeros((128, N))
    for s in range((kdim + 15) >> 4):
        lbo = F_CH if 2 * s + 1 < 7 else 0
        A = read_kmajor(smem, F_ACT + k * F_STREAM + 2 * s * F_CH, lbo, 128, 128)
        B = read_kmajor(smem, img_off + 2 * s * 1024, 1024, 128, N)
        D += A @ B.T
    return D


def test_forward_and_adjoint_gemm_addressing_including_the_lbo0_kstep():
    rng = np.random.default_rng(0)
    smem = rng.integers(-3, 4, SMEM // 2).astype(float)           # garbage everywhere: whatever is read by mistake shows up
    for din, dout in ((50, 50), (50, 5), (14, 30), (56, 56)):
        A = rng.integers(-3, 4, (NS, 128, din)).astype(float)
        W = rng.integers(-3, 4, (din, dout)).astype(float)
        for k in range(NS):
            for p in range(128):
                for u in range(56):                                # the epilogue writes all 56 units, zeros beyond the layer width
                    put(smem, plane_off(k, u, p), A[k, p, u] if u < din else 0.0)
        smem[F_R // 2:F_R // 2 + 4096] = image(W, adjoint=False)
        for k in range(NS):
            D = layer_gemm(smem, F_R, k, din)
            assert np.array_equal(D[:, :dout], A[k] @ W), (din, dout, k)
            assert not D[:, dout:].any()
        # adjoint: Abar = Zbar W^T with the adjoint image; Zbar has `dout` units (planes zero beyond, as the epilogue / output stage leave them)
        Z = rng.integers(-3, 4, (128, dout)).astype(float)
        for p in range(128):
            for u in range(56):
                put(smem, plane_off(0, u, p), Z[p, u] if u < dout else 0.0)
        smem[F_R // 2:F_R // 2 + 4096] = image(W, adjoint=True)
        D = layer_gemm(smem, F_R, 0, dout)
        assert np.array_equal(D[:, :din], Z @ W.T) and not D[:, din:].any()


def test_weight_and_bias_gradient_addressing():
    rng = np.random.default_rng(1)
    smem = rng.integers(-3, 4, SMEM // 2).astype(float)
    din, dout = 50, 50
    NZ = (dout + 7) & ~7
    A = rng.integers(-3, 4, (128, din)).astype(float)             # stashed planes of one stream, brought back into staging buffer 0
    Z = rng.integers(-3, 4, (128, dout)).astype(float)
    stg = F_R + 16384
    for p in range(128):
        for u in range(56):
            c4, e = u // 4, u % 4
            smem[(stg + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8 + e * 2) // 2] = A[p, u] if u < din else 0.0   # same plane layout, relative to the buffer
            put(smem, plane_off(0, u, p), Z[p, u] if u < dout else 0.0)
    D = np.zeros((64, NZ))
    for s in range(8):                                            # 8 K-steps of 16 points
        o = s * 256
        a = read_mnmajor(smem, stg + o, 128, F_CH, 64)            # M = 64 unit rows; rows 56..63 come from the lo plane (garbage here)
        z = read_mnmajor(smem, F_ACT + o, 128, F_CH, NZ)
        D += a @ z.T
    assert np.array_equal(D[:din, :dout], A.T @ Z)                # rows >= 56 are never drained
    assert not D[:din, dout:].any()
    # bias gradient: ones chunk (unit 0 of every point = 1, units 1..7 = 0) as chunk 0 of the A operand; chunks 1..7 alias the value planes
    for p in range(128):
        for u in range(8):
            smem[(F_ONES + p * 16 + u * 2) // 2] = 1.0 if u == 0 else 0.0
    Db = np.zeros((64, NZ))
    for s in range(8):
        o = s * 256
        a = read_mnmajor(smem, F_ONES + o, 128, F_CH, 64)
        z = read_mnmajor(smem, F_ACT + o, 128, F_CH, NZ)
        Db += a @ z.T
    assert np.array_equal(Db[0, :dout], Z.sum(0)) and not Db[1:8].any()


def test_shared_memory_map_of_the_engine_fits():
    F_STG = F_R + 16384
    F_MISC = F_R + 16384 + 2 * F_STREAM
    total = F_MISC + 256 + 128 * 16 + 4096 + 16 * 256 + 1024
    assert (F_R, F_STG, F_MISC, total) == (145408, 161792, 219136, 230656) and total <= 227 * 1024
    assert 2 * 16384 <= 16384 + 2 * F_STREAM                        # the forward image double buffer lives inside the reverse-sweep region
    assert 320 + 64 <= 384 and 384 + 56 <= 512                      # tensor-memory columns: accumulators, weight-gradient tile, bias tile
