// fp32 SIMT (FFMA) kernels: fused forward-jet MLP + residual + MSE partials + reverse sweep, and the
// forward-only field/jet kernels.  Parity anchor for every width; the tcgen05 path (pe_tc.cu) is checked
// against this one and against the oracle.
//
// Reference graph nodes replaced (paths relative to the reference root):
//   neural_net      PlateHoleQuarter/train/train.py:308-320   (inf: ElasticWaveInfinite/ElasticWave.py:188-199)
//   net_uv/net_e    train.py:358-396      net_f_sig  train.py:404-439 (F5), ElasticWaveSemiInfinite/ElasticWave.py:228-272 (F7)
//   net_t           train.py:452-461      losses     train.py:187-217, semi:112-127
//   d loss / d(uv weights, biases)        train.py:249-250 (AdamOptimizer.minimize builds the reverse graph)
//
// Mapping.  One CTA = one tile of 32 collocation points x G warps.  lane = point, warp g owns hidden units
// [10g, 10g+10) of every layer.  All K jet streams (value, d/dx, d/dy, d/dt, d2/dt2) of a point are carried
// together, so first and second derivatives fall out of one forward pass and the reverse
// sweep is the hand-derived adjoint of that pass (A.2) -- no graph replay.
//   forward  layer l : z_k[p][j] = sum_i A_k[p][i] W[i][j]      thread (p, J) keeps K*10 accumulators
//   adjoint  layer l : abar_k[p][i] = sum_j zbar_k[p][j] W[i][j]  thread (p, I) keeps K*10 accumulators
//   weight grad      : dW[i][j] = sum_{k,p} A_k[p][i] zbar_k[p][j] 4x8 register blocks, contraction over the tile
// Activations of the current layer live in shared memory as [k][p][lda] (unit index contiguous, lda/4 odd so
// that 128-bit loads of 32 lanes are bank-conflict free); all hidden activations are stashed per CTA in a
// global scratch that stays L2 resident (296 CTAs x 133 KB for the 5x50 net) and re-read by the reverse
// sweep.  Weights are read through L1 with warp-uniform 64/128-bit __ldg.  Weight-gradient blocks are
// accumulated over the tile in registers and added to the CTA's private partial buffer with red.global.add.v4.f32
// (fire and forget; exactly one owner thread per element and tile => fixed order, deterministic); pe_reduce_* sums the slots.
#include "pe_common.cuh"
#include "pe_device.cuh"

namespace {
using namespace pe_dev;


// ---- input jets of tile -> smem [k][p][lda0=4]
template <int K>
__device__ __forceinline__ void write_input_jets(float* buf, int lda0, int p, float x, float y, float t,
                                                 const float* sc, const float* sh) {
    float* r0 = buf + (0 * PE_P + p) * lda0;
    r0[0] = fmaf(x, sc[0], sh[0]); r0[1] = fmaf(y, sc[1], sh[1]); r0[2] = fmaf(t, sc[2], sh[2]); r0[3] = 0.f;
    if (K >= 4) {
        float* r1 = buf + (1 * PE_P + p) * lda0; r1[0] = sc[0]; r1[1] = 0.f; r1[2] = 0.f; r1[3] = 0.f;
        float* r2 = buf + (2 * PE_P + p) * lda0; r2[0] = 0.f; r2[1] = sc[1]; r2[2] = 0.f; r2[3] = 0.f;
        float* r3 = buf + (3 * PE_P + p) * lda0; r3[0] = 0.f; r3[1] = 0.f; r3[2] = sc[2]; r3[3] = 0.f;
    }
    if (K == 2) {
        float* r1 = buf + (1 * PE_P + p) * lda0; r1[0] = 0.f; r1[1] = 0.f; r1[2] = sc[2]; r1[3] = 0.f;
    }
    if (K == 5) {
        float* r4 = buf + (4 * PE_P + p) * lda0; r4[0] = 0.f; r4[1] = 0.f; r4[2] = 0.f; r4[3] = 0.f;
    }
}

// weights come either from the shared-memory staging buffer (WS) or straight from global through L1 (__ldg)
template <bool WS> __device__ __forceinline__ float2 wld2(const float* p) { return WS ? *reinterpret_cast<const float2*>(p) : __ldg(reinterpret_cast<const float2*>(p)); }
template <bool WS> __device__ __forceinline__ float4 wld4(const float* p) { return WS ? *reinterpret_cast<const float4*>(p) : __ldg(reinterpret_cast<const float4*>(p)); }
template <bool WS> __device__ __forceinline__ float wld1(const float* p) { return WS ? *p : __ldg(p); }

// ---- forward GEMM of one layer for thread (p, units j0..j0+9): acc[k][u] = sum_i in[k][p][i] * W[i][j0+u]
// packed fp32 FMAs (fma.rn.f32x2 / FFMA2, sm_100): the activation is the broadcast operand, unit pairs are packed
template <int K, bool WS>
__device__ __forceinline__ void gemm_fwd(float (&acc)[K][PE_UJ], const float* __restrict__ in, int lda_in, int din,
                                         const float* __restrict__ W, int ldw, int j0, int p) {
    float2 acc2[K][PE_UJ / 2];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int u = 0; u < PE_UJ / 2; ++u) acc2[k][u] = make_float2(0.f, 0.f);
    const float* wcol = W + j0;
    int i = 0;
    for (; i + 4 <= din; i += 4) {
        float a[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float4 v = ld4(in + (k * PE_P + p) * lda_in + i);
            a[k][0] = v.x; a[k][1] = v.y; a[k][2] = v.z; a[k][3] = v.w;
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float2 w[PE_UJ / 2];
            const float* wr = wcol + (size_t)(i + ii) * ldw;
#pragma unroll
            for (int u2 = 0; u2 < PE_UJ / 2; ++u2) w[u2] = wld2<WS>(wr + 2 * u2);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float2 ak = make_float2(a[k][ii], a[k][ii]);
#pragma unroll
                for (int u2 = 0; u2 < PE_UJ / 2; ++u2) acc2[k][u2] = __ffma2_rn(ak, w[u2], acc2[k][u2]);
            }
        }
    }
    for (; i < din; ++i) {
        float2 w[PE_UJ / 2];
        const float* wr = wcol + (size_t)i * ldw;
#pragma unroll
        for (int u2 = 0; u2 < PE_UJ / 2; ++u2) w[u2] = wld2<WS>(wr + 2 * u2);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float av = in[(k * PE_P + p) * lda_in + i];
            const float2 ak = make_float2(av, av);
#pragma unroll
            for (int u2 = 0; u2 < PE_UJ / 2; ++u2) acc2[k][u2] = __ffma2_rn(ak, w[u2], acc2[k][u2]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int u2 = 0; u2 < PE_UJ / 2; ++u2) { acc[k][2 * u2] = acc2[k][u2].x; acc[k][2 * u2 + 1] = acc2[k][u2].y; }
}

// ---- adjoint GEMM: ab[k][u] = sum_j zb[k][p][j] * W[i0+u][j]
template <int K, bool WS>
__device__ __forceinline__ void gemm_adj(float (&ab)[K][PE_UJ], const float* __restrict__ zb, int lda_out, int dout,
                                         const float* __restrict__ W, int ldw, int i0, int p) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int u = 0; u < PE_UJ; ++u) ab[k][u] = 0.f;
    const float* wrow = W + (size_t)i0 * ldw;
    int j = 0;
    for (; j + 4 <= dout; j += 4) {
        float4 z[K];
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = ld4(zb + (k * PE_P + p) * lda_out + j);
#pragma unroll
        for (int u = 0; u < PE_UJ; ++u) {
            float4 w = wld4<WS>(wrow + (size_t)u * ldw + j);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float s = ab[k][u];
                s = fmaf(z[k].x, w.x, s); s = fmaf(z[k].y, w.y, s); s = fmaf(z[k].z, w.z, s); s = fmaf(z[k].w, w.w, s);
                ab[k][u] = s;
            }
        }
    }
    for (; j < dout; ++j) {
        float z[K];
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = zb[(k * PE_P + p) * lda_out + j];
#pragma unroll
        for (int u = 0; u < PE_UJ; ++u) {
            float w = wld1<WS>(wrow + (size_t)u * ldw + j);
#pragma unroll
            for (int k = 0; k < K; ++k) ab[k][u] = fmaf(z[k], w, ab[k][u]);
        }
    }
}

// MAXG = max warps per CTA this instantiation is launched with; MINB = CTAs/SM the register budget allows
// cooperative copy of weight matrix m (+ its bias) into a staging buffer: one batch of coalesced 128-bit loads per
// thread instead of din dependent L2 round trips inside the GEMM loop
__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ params, const PeLayout& lay, int m, int tid, int nthr) {
    const float4* src = reinterpret_cast<const float4*>(params + lay.woff[m]);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int n4 = (lay.d[m] * lay.ldw[m]) >> 2;
    for (int i = tid; i < n4; i += nthr) d4[i] = __ldg(src + i);
    const float4* sb = reinterpret_cast<const float4*>(params + lay.boff[m]);
    float4* db = reinterpret_cast<float4*>(dst + lay.wmat_floats);
    for (int i = tid; i < (lay.ldw[m] >> 2); i += nthr) db[i] = __ldg(sb + i);
}

template <int K, int MAXG, int MINB, bool WS>
__global__ void __launch_bounds__(PE_P * MAXG, MINB)
resid_simt_kernel(const PeResidArgs args) {
    extern __shared__ __align__(16) float smem[];
    const PeLayout& lay = args.lay;
    const pe_term_desc& T = args.term;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int p = tid & 31, g = tid >> 5;
    const int L = lay.L;
    const int bufsz = K * PE_P * lay.max_lda + 16;
    float* buf0 = smem;
    float* buf1 = smem + bufsz;
    float* wst = smem + 2 * bufsz;                 // WS: two weight staging buffers of lay.wstage_floats
    for (int i = tid; i < 2 * bufsz; i += nthr) smem[i] = 0.f;
    // weight matrix m of the current layer: staged copy (buffer m & 1) or global
    auto Wm = [&](int m) -> const float* { return WS ? wst + (m & 1) * lay.wstage_floats : args.params + lay.woff[m]; };
    auto Bm = [&](int m) -> const float* { return WS ? wst + (m & 1) * lay.wstage_floats + lay.wmat_floats : args.params + lay.boff[m]; };

    const int slot = args.slot_base + blockIdx.x;
    float* gpart = args.grad_partials + (size_t)slot * lay.total;
    float* stash = args.stash + (size_t)blockIdx.x * args.stash_floats;   // scratch is per launch, not per slot
    const float* __restrict__ params = args.params;
    for (int i = tid; i < lay.total; i += nthr) __stcg(gpart + i, 0.f);    // slot fully overwritten (also when this CTA gets no tile)
    float tsum[PE_MAX_TERMS];
#pragma unroll
    for (int i = 0; i < PE_MAX_TERMS; ++i) tsum[i] = 0.f;
    __syncthreads();

    const int ntiles = (args.n + PE_P - 1) / PE_P;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pt = tile * PE_P + p;
        const bool valid = pt < args.n;
        const float* row = args.points + (size_t)(valid ? pt : 0) * T.ld;
        float x = 0.f, y = 0.f, t = 0.f;
        if (valid) { x = row[0]; y = row[1]; t = row[2]; }
        float* cur = buf0;
        float* oth = buf1;
        if (g == 0) write_input_jets<K>(cur, lay.lda[0], p, x, y, t, T.in_scale, T.in_shift);
        if (WS) stage_weights(wst, params, lay, 0, tid, nthr);
        __syncthreads();

        // ------------------------------------------------ forward, hidden layers
        for (int l = 1; l < L; ++l) {
            const int din = lay.d[l - 1], dout = lay.d[l];
            const int j0 = g * PE_UJ;
            if (WS) stage_weights(wst + (l & 1) * lay.wstage_floats, params, lay, l, tid, nthr);   // next layer's matrix, overlaps this GEMM
            if (j0 < dout) {
                float acc[K][PE_UJ];
                gemm_fwd<K, WS>(acc, cur, lay.lda[l - 1], din, Wm(l - 1), lay.ldw[l - 1], j0, p);
                const float* bias = Bm(l - 1) + j0;
                float* st = stash + (size_t)K * PE_P * lay.soff[l];
                const int ldo = lay.lda[l];
#pragma unroll
                for (int u = 0; u < PE_UJ; u += 2) {
                    if (j0 + u < dout) {
                        float z0[K], z1[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) { z0[k] = acc[k][u]; z1[k] = acc[k][u + 1]; }
                        act_fwd<K>(z0, wld1<WS>(bias + u));
                        const bool two = (j0 + u + 1 < dout);
                        if (two) act_fwd<K>(z1, wld1<WS>(bias + u + 1));
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            float2 v = make_float2(z0[k], two ? z1[k] : 0.f);
                            *reinterpret_cast<float2*>(oth + (k * PE_P + p) * ldo + j0 + u) = v;
                            __stcg(reinterpret_cast<float2*>(st + (k * PE_P + p) * ldo + j0 + u), v);
                        }
                    }
                }
            }
            __syncthreads();
            float* tmp = cur; cur = oth; oth = tmp;
        }
        // ------------------------------------------------ output layer + residual + seeds (warp 0)
        {
            const int din = lay.d[L - 1], dout = lay.d[L];
            const int ldo = lay.lda[L];
            if (g == 0) {
                float Y[K][PE_UJ];
                gemm_fwd<K, WS>(Y, cur, lay.lda[L - 1], din, Wm(L - 1), lay.ldw[L - 1], 0, p);
                const float* bias = Bm(L - 1);
#pragma unroll
                for (int u = 0; u < PE_UJ; ++u) {
                    if (u < dout) Y[0][u] += wld1<WS>(bias + u);
                    else {
#pragma unroll
                        for (int k = 0; k < K; ++k) Y[k][u] = 0.f;
                    }
                }
                const float* aux_row = args.aux ? args.aux + (size_t)(valid ? pt : 0) * (2 * T.aux_k * 5) : nullptr;
                residual_stage<K>(Y, T, aux_row, row, valid, args.inv_n, tsum);
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u)
                        if (u < dout) oth[(k * PE_P + p) * ldo + u] = Y[k][u];
            }
            __syncthreads();
        }
        float* bufZ = oth;     // zbar of layer l, [k][p][lda[l]]
        float* bufA = cur;     // inputs of layer l, [k][p][lda[l-1]]  (already holds A^{L-1})

        // ------------------------------------------------ reverse sweep
        for (int l = L; l >= 1; --l) {
            const int din = lay.d[l - 1], dout = lay.d[l];
            const int ldi = lay.lda[l - 1], ldo = lay.lda[l];
            // matrices L-1 and L-2 are still staged from the forward pass; deeper ones are prefetched one layer ahead
            if (WS && l <= L - 1 && l >= 3) stage_weights(wst + ((l - 2) & 1) * lay.wstage_floats, params, lay, l - 2, tid, nthr);
            if (l < L) {
                if (l == 1) {
                    if (g == 0) write_input_jets<K>(bufA, ldi, p, x, y, t, T.in_scale, T.in_shift);
                } else {
                    const float4* src = reinterpret_cast<const float4*>(stash + (size_t)K * PE_P * lay.soff[l - 1]);
                    float4* dst = reinterpret_cast<float4*>(bufA);
                    const int n4 = K * PE_P * ldi / 4;
                    for (int i = tid; i < n4; i += nthr) dst[i] = __ldcg(src + i);
                }
                __syncthreads();
            }
            // ---- weight gradient blocks 4(i) x 8(j), contraction over (k, p)
            {
                const int nI = (din + 3) >> 2, nJ = (dout + 7) >> 3;
                const int ldw = lay.ldw[l - 1];
                float* gW = gpart + lay.woff[l - 1];
                for (int task = tid; task < nI * nJ; task += nthr) {
                    const int ib = task / nJ, jb = task - ib * nJ;
                    const int i0 = ib * 4, j0 = jb * 8;
                    float2 acc2[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc2[r][c] = make_float2(0.f, 0.f);
                    const float* pa = bufA + i0;
                    const float* pz = bufZ + j0;
#pragma unroll 4
                    for (int kp = 0; kp < K * PE_P; ++kp) {
                        float4 a = ld4(pa + kp * ldi);
                        float4 z0 = ld4(pz + kp * ldo);
                        float4 z1 = ld4(pz + kp * ldo + 4);
                        const float av[4] = {a.x, a.y, a.z, a.w};
                        const float2 zp[4] = {make_float2(z0.x, z0.y), make_float2(z0.z, z0.w), make_float2(z1.x, z1.y), make_float2(z1.z, z1.w)};
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float2 ar = make_float2(av[r], av[r]);
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc2[r][c] = __ffma2_rn(ar, zp[c], acc2[r][c]);
                        }
                    }
                    float acc[4][8];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {     // pad columns (j >= dout) of the padded layout keep a zero gradient
                            acc[r][2 * c] = (j0 + 2 * c < dout) ? acc2[r][c].x : 0.f;
                            acc[r][2 * c + 1] = (j0 + 2 * c + 1 < dout) ? acc2[r][c].y : 0.f;
                        }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = i0 + r;
                        if (i < din) {
                            float* dstp = gW + (size_t)i * ldw + j0;
                            // red.global.add.v4.f32 (no return value, nothing to wait for); one owner thread per element and tile => fixed order
                            if (j0 < ldw) atomicAdd(reinterpret_cast<float4*>(dstp), make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
                            if (j0 + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(dstp + 4), make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
                        }
                    }
                }
                // bias gradient: value stream only
                float* gB = gpart + lay.boff[l - 1];
                for (int j = tid; j < dout; j += nthr) {
                    float s = 0.f;
#pragma unroll 8
                    for (int pp = 0; pp < PE_P; ++pp) s += bufZ[pp * ldo + j];
                    atomicAdd(gB + j, s);
                }
            }
            if (l == 1) { __syncthreads(); break; }
            // ---- adjoint of the layer inputs, then through tanh
            const int i0 = g * PE_UJ;
            float ab[K][PE_UJ];
            const bool active = i0 < din;
            if (active) gemm_adj<K, WS>(ab, bufZ, ldo, dout, Wm(l - 1), lay.ldw[l - 1], i0, p);
            __syncthreads();
            if (active) {
#pragma unroll
                for (int u = 0; u < PE_UJ; u += 2) {
                    if (i0 + u < din) {
                        float A0[K], A1[K], b0[K], b1[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            float2 v = ld2(bufA + (k * PE_P + p) * ldi + i0 + u);
                            A0[k] = v.x; A1[k] = v.y; b0[k] = ab[k][u]; b1[k] = ab[k][u + 1];
                        }
                        act_bwd<K>(b0, A0);
                        const bool two = (i0 + u + 1 < din);
                        if (two) act_bwd<K>(b1, A1);
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            *reinterpret_cast<float2*>(bufZ + (k * PE_P + p) * ldi + i0 + u) = make_float2(b0[k], two ? b1[k] : 0.f);
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- loss-term partial sums of this CTA (warp 0 holds them, indexed by residual position c -> term[c])
    if (g == 0) {
        float tot[PE_MAX_TERMS];
#pragma unroll
        for (int i = 0; i < PE_MAX_TERMS; ++i) tot[i] = warp_sum(tsum[i]);
        if (p == 0) {
            float* tp = args.term_partials + (size_t)slot * PE_MAX_TERMS;
#pragma unroll
            for (int i = 0; i < PE_MAX_TERMS; ++i) tp[i] = 0.f;
            const int nres = (T.kind == PE_RES_F5 || T.kind == PE_RES_F7) ? 2 : (T.kind == PE_RES_TRACTION ? 1 : T.ncols);
#pragma unroll
            for (int c = 0; c < PE_MAX_TERMS; ++c)
                if (c < nres) tp[T.term[c]] += tot[c] * args.inv_n;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward-only: fields (predict) and raw jets.  Same tile mapping, no stash, no reverse sweep.
template <int K>
__global__ void __launch_bounds__(PE_P * 16, 1)
fields_simt_kernel(const PeFieldsArgs args) {
    extern __shared__ __align__(16) float smem[];
    const PeLayout& lay = args.lay;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int p = tid & 31, g = tid >> 5;
    const int L = lay.L;
    const int bufsz = K * PE_P * lay.max_lda + 16;
    float* buf0 = smem;
    float* buf1 = smem + bufsz;
    for (int i = tid; i < 2 * bufsz; i += nthr) smem[i] = 0.f;
    __syncthreads();
    const float* __restrict__ params = args.params;
    const int ntiles = (args.n + PE_P - 1) / PE_P;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pt = tile * PE_P + p;
        const bool valid = pt < args.n;
        const float* row = args.points + (size_t)(valid ? pt : 0) * args.ld;
        float x = 0.f, y = 0.f, t = 0.f;
        if (valid) { x = row[0]; y = row[1]; t = row[2]; }
        float* cur = buf0;
        float* oth = buf1;
        if (g == 0) write_input_jets<K>(cur, lay.lda[0], p, x, y, t, args.in_scale, args.in_shift);
        __syncthreads();
        for (int l = 1; l < L; ++l) {
            const int din = lay.d[l - 1], dout = lay.d[l];
            const int j0 = g * PE_UJ;
            if (j0 < dout) {
                float acc[K][PE_UJ];
                gemm_fwd<K, false>(acc, cur, lay.lda[l - 1], din, params + lay.woff[l - 1], lay.ldw[l - 1], j0, p);
                const float* bias = params + lay.boff[l - 1] + j0;
                const int ldo = lay.lda[l];
#pragma unroll
                for (int u = 0; u < PE_UJ; u += 2) {
                    if (j0 + u < dout) {
                        float z0[K], z1[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) { z0[k] = acc[k][u]; z1[k] = acc[k][u + 1]; }
                        act_fwd<K>(z0, __ldg(bias + u));
                        const bool two = (j0 + u + 1 < dout);
                        if (two) act_fwd<K>(z1, __ldg(bias + u + 1));
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            *reinterpret_cast<float2*>(oth + (k * PE_P + p) * ldo + j0 + u) = make_float2(z0[k], two ? z1[k] : 0.f);
                    }
                }
            }
            __syncthreads();
            float* tmp = cur; cur = oth; oth = tmp;
        }
        if (g == 0) {
            const int din = lay.d[L - 1], dout = lay.d[L];
            float Y[K][PE_UJ];
            gemm_fwd<K, false>(Y, cur, lay.lda[L - 1], din, params + lay.woff[L - 1], lay.ldw[L - 1], 0, p);
            const float* bias = params + lay.boff[L - 1];
#pragma unroll
            for (int u = 0; u < PE_UJ; ++u) if (u < dout) Y[0][u] += __ldg(bias + u);
            if (valid) {
                if (args.mode == 1) {
                    float* o = args.out + (size_t)pt * K * dout;
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int u = 0; u < PE_UJ; ++u) if (u < dout) o[k * dout + u] = Y[k][u];
                } else if (K == 4) {
                    if (args.aux_k) {       // composite on (value, x, y, t) streams, 5 outputs
                        const float* ar = args.aux + (size_t)pt * (2 * args.aux_k * 5);
#pragma unroll
                        for (int o = 0; o < 5; ++o) {
                            float D0 = ar[o], N0 = Y[0][o];
#pragma unroll
                            for (int k = 3; k >= 1; --k) Y[k][o] = ar[args.aux_k * 5 + k * 5 + o] + ar[k * 5 + o] * N0 + D0 * Y[k][o];
                            Y[0][o] = fmaf(D0, N0, ar[args.aux_k * 5 + o]);
                        }
                    }
                    float* o = args.out + (size_t)pt * 8;
                    const bool f7 = args.formulation == PE_RES_F7;
                    o[0] = Y[0][0]; o[1] = Y[0][1];
                    o[2] = f7 ? Y[0][4] : Y[0][2];
                    o[3] = f7 ? Y[0][5] : Y[0][3];
                    o[4] = f7 ? Y[0][6] : Y[0][4];
                    o[5] = Y[1][0];                 // e11 = u_x     (plate:393)
                    o[6] = Y[2][1];                 // e22 = v_y     (plate:394)
                    o[7] = Y[2][0] + Y[1][1];       // e12 = u_y+v_x (plate:395)
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int pe_simt_smem_bytes(const PeLayout& lay, int K) {
    return 2 * (K * PE_P * lay.max_lda + 16) * (int)sizeof(float);
}

// weights are staged in shared memory (double buffered) whenever that still leaves the same number of CTAs per SM
bool pe_simt_stage_weights(const pe_plan* plan, int K) {
    const PeLayout& lay = plan->lay;
    int base = pe_simt_smem_bytes(lay, K), extra = 2 * lay.wstage_floats * (int)sizeof(float);
    int limit = plan->smem_optin > 0 ? plan->smem_optin : 227 * 1024;
    int by_reg = lay.groups <= 5 ? 2 : 1;
    return (base + extra + 1024) * by_reg <= limit + 1024 * by_reg && base + extra <= limit;
}

template <int K, int MAXG, int MINB, bool WS>
static int launch_resid_g(const PeResidArgs& a, int slots, cudaStream_t st) {
    int smem = pe_simt_smem_bytes(a.lay, K) + (WS ? 2 * a.lay.wstage_floats * (int)sizeof(float) : 0);
    cudaError_t e = cudaFuncSetAttribute(resid_simt_kernel<K, MAXG, MINB, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_simt<%d>, smem=%d): %s", K, smem, cudaGetErrorString(e)); return 2; }
    resid_simt_kernel<K, MAXG, MINB, WS><<<slots, PE_P * a.lay.groups, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_simt<%d>: %s", K, cudaGetErrorString(e)); return 3; }
    return 0;
}

template <int K>
static int launch_resid(const pe_plan* plan, const PeResidArgs& a, int slots, cudaStream_t st) {
    const bool ws = pe_simt_stage_weights(plan, K);
    if (a.lay.groups <= 5) return ws ? launch_resid_g<K, 5, 2, true>(a, slots, st) : launch_resid_g<K, 5, 2, false>(a, slots, st);
    if (a.lay.groups <= 8) return ws ? launch_resid_g<K, 8, 1, true>(a, slots, st) : launch_resid_g<K, 8, 1, false>(a, slots, st);
    return ws ? launch_resid_g<K, 16, 1, true>(a, slots, st) : launch_resid_g<K, 16, 1, false>(a, slots, st);
}

int pe_launch_resid_simt(const pe_plan* plan, const PeResidArgs& a, int K, int slots, cudaStream_t st) {
    switch (K) {
        case 1: return launch_resid<1>(plan, a, slots, st);
        case 2: return launch_resid<2>(plan, a, slots, st);
        case 4: return launch_resid<4>(plan, a, slots, st);
        case 5: return launch_resid<5>(plan, a, slots, st);
    }
    pe_set_error("unsupported stream count K=%d", K);
    return 1;
}

template <int K>
static int launch_fields(const pe_plan* plan, const PeFieldsArgs& a, cudaStream_t st) {
    int smem = pe_simt_smem_bytes(a.lay, K);
    cudaError_t e = cudaFuncSetAttribute(fields_simt_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(fields_simt<%d>): %s", K, cudaGetErrorString(e)); return 2; }
    int ntiles = (a.n + PE_P - 1) / PE_P;
    int grid = ntiles < plan->sms * 4 ? ntiles : plan->sms * 4;
    if (grid < 1) grid = 1;
    fields_simt_kernel<K><<<grid, PE_P * a.lay.groups, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch fields_simt<%d>: %s", K, cudaGetErrorString(e)); return 3; }
    return 0;
}

int pe_launch_fields(const pe_plan* plan, const PeFieldsArgs& a, int K, cudaStream_t st) {
    switch (K) {
        case 1: return launch_fields<1>(plan, a, st);
        case 2: return launch_fields<2>(plan, a, st);
        case 4: return launch_fields<4>(plan, a, st);
        case 5: return launch_fields<5>(plan, a, st);
    }
    pe_set_error("unsupported stream count K=%d", K);
    return 1;
}

int pe_simt_ctas_per_sm(const pe_plan* plan, int K) {
    int smem = pe_simt_smem_bytes(plan->lay, K) + (pe_simt_stage_weights(plan, K) ? 2 * plan->lay.wstage_floats * (int)sizeof(float) : 0);
    int by_smem = (plan->smem_optin > 0 ? plan->smem_optin + 1024 : 228 * 1024) / (smem + 1024);
    int by_reg = plan->lay.groups <= 5 ? 2 : 1;     // register budgets of the <K, MAXG, MINB> instantiations
    int n = by_smem < by_reg ? by_smem : by_reg;
    return n < 1 ? 1 : n;
}
