// Pipelined tcgen05 / TMEM engine for the collocation residual: second generation of pe_tc.cu.
//
//   * NS jet streams as a template parameter: NS = 5 (plate formulation F5: value, d/dx, d/dy, d/dt, d2/dt2; plate:404-439)
//     and NS = 4 (wave formulation F7: value, d/dx, d/dy, d/dt, 7 outputs; semi:228-272, inf:221-265, conf:304-348).
//   * the weight-gradient phase runs as a producer / consumer pipeline: warps 1..7 convert the two MN-major bf16 (hi, mid)
//     operands (stashed activations A_k, adjoints Zbar_k) half a tile (64 points) at a time and hand each half to warp 0
//     through an mbarrier; warp 0 issues the tcgen05.mma of that half and hands the buffer half back with tcgen05.commit.
//     In pe_tc.cu these two activities alternate (convert | sync | MMA | wait) and cost 12 % + 12 % of the kernel
//     (profiles/r1_tc3_phase_cycles.txt); here the MMAs of half X run under the conversion of half 1-X.
//   * the next tile's coordinates are pulled into L2 while the current tile computes.
//
// Math, operand layouts, TMEM map and the 3-term fp32 split are those of pe_tc.cu (see its header); reference lines per
// stage are cited in pe_simt.cu / pe_device.cuh.  PIPE = false reproduces the pe_tc.cu phase order (A/B measurements).
#include <cstring>
#include "pe_device.cuh"
#include "pe_tc_common.cuh"

namespace {
using namespace pe_dev;
using namespace pe_tcc;

constexpr int CONV_THREADS = TC_THREADS - 32;    // warps 1..7 convert, warp 0 issues

#define TCP_PROF(slot) do { if (PROF) { if (args.prof && blockIdx.x == 0 && tid == 0) { long long now_ = clock64(); prof_acc[slot] += (unsigned long long)(now_ - prof_t); prof_t = now_; } } } while (0)

// MMAs of one layer GEMM for all NS streams (one thread).  K-steps are issued stream-interleaved: consecutive MMAs never touch
// the same accumulator (a dependent chain runs at pipeline latency, pe_tc.cu).
template <int NS>
__device__ __forceinline__ void issue_layer(uint32_t tbase, uint32_t act_s, uint32_t wimg_s, int N, int ksteps, int kb, int fast) {
    const uint32_t id32 = idesc_tf32(N), id16 = idesc_bf16(N);
    const uint32_t nrow = (uint32_t)N * 16u;
    const uint64_t a_step = (uint64_t)((2u * TC_CH) >> 4), b_step = (uint64_t)((2u * nrow) >> 4);
    const uint64_t b_hi = sdesc(wimg_s + TC_IMG_HI, nrow, 128), b_lo = sdesc(wimg_s + TC_IMG_LO, nrow, 128), b_bf = sdesc(wimg_s + TC_IMG_BF, nrow, 128);
    const uint64_t a0 = sdesc(act_s, TC_CH, 128);
    const uint64_t a_stream = (uint64_t)(TC_ACT_STREAM >> 4);
    const uint32_t d0 = tbase + TM_ACC;
#pragma unroll
    for (int s = 0; s < 7; ++s)
        if (s < ksteps) {
#pragma unroll
            for (int k = 0; k < NS; ++k) mma_tf32_ss(d0 + 64u * k, a0 + k * a_stream + s * a_step, b_hi + s * b_step, id32, s > 0);
        }
    if (!fast) {
#pragma unroll
        for (int s = 0; s < 7; ++s)
            if (s < ksteps) {
#pragma unroll
                for (int k = 0; k < NS; ++k) mma_tf32_ss(d0 + 64u * k, a0 + k * a_stream + s * a_step, b_lo + s * b_step, id32, 1);
            }
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < kb) {
#pragma unroll
                for (int k = 0; k < NS; ++k) mma_bf16_ts(d0 + 64u * k, tbase + TM_LO + 32u * k + 8u * s, b_bf + s * b_step, id16, 1);
            }
    }
}

template <int NS, bool PIPE, bool PROF>
__global__ void __launch_bounds__(TC_THREADS, 1) resid_tcp_kernel(const TcpArgs args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STASH_LAYER = NS * TC_STASH_STREAM;          // bytes per stashed layer
    const PeResidArgs& A = args.r;
    const PeLayout& lay = A.lay;
    const pe_term_desc& T = A.term;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = 32 * (warp & 3) + lane;            // TMEM lane = point of the tile
    const int h = warp >> 2;                         // unit half: chunks [7h, 7h+7)
    const int L = lay.L;
    const int fast = args.fast;
    uint8_t* act = smem + SM_ACT;
    uint8_t* wimg = smem + SM_WIMG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_MISC + MB_TMEM);
    float* coord = reinterpret_cast<float*>(smem + SM_COORD);       // [128][4]: a0x, a0y, a0t, valid
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);
    float* sw0 = reinterpret_cast<float*>(smem + SM_W0);
    const uint32_t act_s = smem_u32(act), wimg_s = smem_u32(wimg), bar_s = smem_u32(smem + SM_MISC + MB_MAIN);
    const uint32_t bar_full = bar_s + MB_FULL, bar_empty = bar_s + MB_EMPTY;

    for (int i = tid; i < SM_TOTAL / 16; i += TC_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int slot = A.slot_base + blockIdx.x;
    float* gpart = A.grad_partials + (size_t)slot * lay.total;
    float* stash = A.stash + (size_t)blockIdx.x * A.stash_floats;
    const float* __restrict__ params = A.params;
    for (int i = tid; i < lay.total; i += TC_THREADS) __stcg(gpart + i, 0.f);
    __syncthreads();
    if (tid < 64) {                                  // first-layer weights (3 x d1) and bias -> smem, once per launch
        const bool in = tid < lay.d[1];
        sw0[tid] = in ? __ldg(params + lay.woff[0] + tid) : 0.f;
        sw0[64 + tid] = in ? __ldg(params + lay.woff[0] + lay.ldw[0] + tid) : 0.f;
        sw0[128 + tid] = in ? __ldg(params + lay.woff[0] + 2 * lay.ldw[0] + tid) : 0.f;
        sw0[192 + tid] = in ? __ldg(params + lay.boff[0] + tid) : 0.f;
    }
    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_full, CONV_THREADS);
        mbar_init(bar_full + 8, CONV_THREADS);
        mbar_init(bar_empty, 1);
        mbar_init(bar_empty + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    // zero the bf16 "lo" operand columns once (units 56..63 are never written afterwards and must stay zero)
    for (int c = h * 80; c < h * 80 + 80; c += 2) tm_st2(tlane + TM_LO + c, 0u, 0u);
    tm_wait_st();
    uint32_t parity = 0;
    uint32_t pfull = 0, pempty = 0;                  // phase parities of the operand-half barriers (bit X); per-thread bookkeeping
    unsigned long long prof_acc[PROF ? 16 : 1];
#pragma unroll
    for (int i = 0; i < (PROF ? 16 : 1); ++i) prof_acc[i] = 0ull;
    long long prof_t = PROF ? clock64() : 0;
    (void)prof_t;
    float tsum[PE_MAX_TERMS];
#pragma unroll
    for (int i = 0; i < PE_MAX_TERMS; ++i) tsum[i] = 0.f;
    fence_before();
    __syncthreads();
    fence_after();

    float bias_pre = (tid < 64 && tid < lay.d[2]) ? __ldg(params + lay.boff[1] + tid) : 0.f;   // bias of the next TC layer, prefetched like the images
    float4 img[9];                                   // one operand-image set (36,864 B / 256 threads), prefetched a phase ahead
#pragma unroll
    for (int i = 0; i < 9; ++i) img[i] = __ldg(reinterpret_cast<const float4*>(args.images + (size_t)1 * TC_IMG_LAYER) + tid + i * TC_THREADS);
    const pe_term_desc& T2 = args.term2;
    float tsum2[PE_MAX_TERMS];
#pragma unroll
    for (int i = 0; i < PE_MAX_TERMS; ++i) tsum2[i] = 0.f;
    const int ntiles_main = (A.n + TC_P - 1) / TC_P;
    const int ntiles = ntiles_main + (args.n2 + TC_P - 1) / TC_P;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const bool sec = tile >= ntiles_main;                          // tile of the fused primal-only set (CTA-uniform)
        const pe_term_desc& Tc = sec ? T2 : T;
        const int pt = (sec ? tile - ntiles_main : tile) * TC_P + p;
        const bool valid = pt < (sec ? args.n2 : A.n);
        const float* row = (sec ? args.points2 : A.points) + (size_t)(valid ? pt : 0) * Tc.ld;
        if (h == 0) {
            float x = 0.f, y = 0.f, t = 0.f;
            if (valid) { x = row[0]; y = row[1]; t = row[2]; }
            *reinterpret_cast<float4*>(coord + 4 * p) = make_float4(fmaf(x, Tc.in_scale[0], Tc.in_shift[0]), fmaf(y, Tc.in_scale[1], Tc.in_shift[1]),
                                                                    fmaf(t, Tc.in_scale[2], Tc.in_shift[2]), valid ? 1.f : 0.f);
        } else {
            // pull this CTA's next tile (coordinates, and the composite jets if any) towards L2: the loads above are the
            // only DRAM-latency loads on the tile's critical path
            const int nt = tile + (int)gridDim.x;
            if (nt < ntiles) {
                const bool nsec = nt >= ntiles_main;
                const int npt = (nsec ? nt - ntiles_main : nt) * TC_P + p;
                if (npt < (nsec ? args.n2 : A.n)) {
                    const float* nrow = (nsec ? args.points2 : A.points) + (size_t)npt * (nsec ? T2.ld : T.ld);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow));
                    if (!nsec && A.aux) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.aux + (size_t)npt * 50));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.aux + (size_t)npt * 50 + 32));
                    }
                }
            }
        }
        __syncthreads();
        TCP_PROF(15);
        // ================================================================ layer 1 (3 -> d1): per-thread FFMA
        {
            const float4 c4 = *reinterpret_cast<const float4*>(coord + 4 * p);
            const int dout = lay.d[1];
            float* st = stash;                                         // stash layer index 0 = outputs of layer 1
#pragma unroll 1
            for (int c = 7 * h; c < 7 * h + 7; ++c) {
                float o[NS][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    float z[NS];
#pragma unroll
                    for (int k = 0; k < NS; ++k) z[k] = 0.f;
                    if (j < dout) {
                        const float w0 = sw0[j], w1 = sw0[64 + j], w2 = sw0[128 + j];
                        z[0] = fmaf(c4.x, w0, fmaf(c4.y, w1, c4.z * w2));
                        z[1] = Tc.in_scale[0] * w0; z[2] = Tc.in_scale[1] * w1; z[3] = Tc.in_scale[2] * w2;
                        act_fwd<NS, true>(z, sw0[192 + j]);
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) o[k][u] = z[k];
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const float4 v = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                    __stcg(reinterpret_cast<float4*>(st + (size_t)k * (TC_STASH_STREAM / 4) + c * 512 + p * 4), v);
                    tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(o[k][0], o[k][1]), lo_pair(o[k][2], o[k][3]));
                }
            }
        }
        TCP_PROF(0);
        // ================================================================ forward: hidden layers 2..L-1 and the output layer L
        for (int l = 2; l <= L; ++l) {
            const int m = l - 1;                                       // weight matrix index
            const int dout = lay.d[l];
            const int NF = (dout <= 16) ? 16 : 64;
#pragma unroll
            for (int i = 0; i < 9; ++i) reinterpret_cast<float4*>(wimg)[tid + i * TC_THREADS] = img[i];
            if (tid < 64) sbias[tid] = bias_pre;
            tm_wait_st();
            fence_async_smem();
            fence_before();
            __syncthreads();
            TCP_PROF(1);
            if (tid == 0) {
                fence_after();
                issue_layer<NS>(tbase, act_s, wimg_s, NF, (lay.d[l - 1] + 7) >> 3, (lay.d[l - 1] + 15) >> 4, fast);
                mma_commit(bar_s);
            }
            TCP_PROF(2);
            {   // prefetch the next operand image while the MMAs run: forward image of the next matrix, or the adjoint image of the last one
                const uint8_t* nsrc = (l < L) ? args.images + (size_t)(m + 1) * TC_IMG_LAYER : args.images + (size_t)m * TC_IMG_LAYER + TC_IMG_SET;
#pragma unroll
                for (int i = 0; i < 9; ++i) img[i] = __ldg(reinterpret_cast<const float4*>(nsrc) + tid + i * TC_THREADS);
                if (l < L) bias_pre = (tid < 64 && tid < lay.d[l + 1]) ? __ldg(params + lay.boff[l] + tid) : 0.f;
            }
            mbar_wait(bar_s, parity);
            parity ^= 1;
            fence_after();
            TCP_PROF(3);
            if (l < L) {
                float* st = stash + (size_t)(l - 1) * (STASH_LAYER / 4);
#pragma unroll 1
                for (int c = 7 * h; c < 7 * h + 7; ++c) {
                    float z[NS][4];
#pragma unroll
                    for (int k = 0; k < NS; ++k) tm_ld4(tlane + TM_ACC + 64 * k + 4 * c, z[k]);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c + u;
                        float zz[NS];
#pragma unroll
                        for (int k = 0; k < NS; ++k) zz[k] = z[k][u];
                        if (j < dout) act_fwd<NS, true>(zz, sbias[j]);
                        else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) zz[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < NS; ++k) z[k][u] = zz[k];
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        const float4 v = make_float4(z[k][0], z[k][1], z[k][2], z[k][3]);
                        *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                        __stcg(reinterpret_cast<float4*>(st + (size_t)k * (TC_STASH_STREAM / 4) + c * 512 + p * 4), v);
                        tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(z[k][0], z[k][1]), lo_pair(z[k][2], z[k][3]));
                    }
                }
                if (h == 1) {      // units 56..63 of the lo operands stay zero
#pragma unroll
                    for (int k = 0; k < NS; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
                TCP_PROF(5);
            } else if (h == 0) {
                // ---------------- outputs -> residuals -> loss partials -> seeds (adjoint of the outputs)
                float Y[NS][PE_UJ];
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    float v[8];
                    tm_ld8(tlane + TM_ACC + 64 * k, v);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) Y[k][u] = (u < 8 && u < dout) ? v[u < 8 ? u : 0] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < PE_UJ; ++u) if (u < dout) Y[0][u] += sbias[u];
                if (!sec) {
                    const float* aux_row = (NS == 5 && A.aux) ? A.aux + (size_t)(valid ? pt : 0) * 50 : nullptr;
                    residual_stage<NS>(Y, T, aux_row, row, valid, A.inv_n, tsum);
                } else {    // primal-only set: residual on the value stream, zero seeds for the derivative streams
                    float Y1[1][PE_UJ];
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) Y1[0][u] = Y[0][u];
                    const float* aux_row = args.aux2 ? args.aux2 + (size_t)(valid ? pt : 0) * 10 : nullptr;
                    residual_stage<1>(Y1, T2, aux_row, row, valid, args.inv_n2, tsum2);
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) {
                        Y[0][u] = Y1[0][u];
#pragma unroll
                        for (int k = 1; k < NS; ++k) Y[k][u] = 0.f;
                    }
                }
                // seeds of up to 8 outputs (F5: 5, F7: 7) -> adjoint operand, units 8..15 zero (read by the K = 16 bf16 step)
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const float4 v0 = make_float4(Y[k][0], Y[k][1], Y[k][2], Y[k][3]);
                    const float4 v1 = make_float4(Y[k][4], Y[k][5], Y[k][6], Y[k][7]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + 0 * TC_CH + p * 16) = v0;
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + 1 * TC_CH + p * 16) = v1;
                    tm_st2(tlane + TM_LO + 32 * k + 0, lo_pair(v0.x, v0.y), lo_pair(v0.z, v0.w));
                    tm_st2(tlane + TM_LO + 32 * k + 2, lo_pair(v1.x, v1.y), lo_pair(v1.z, v1.w));
                    tm_st2(tlane + TM_LO + 32 * k + 4, 0u, 0u);
                    tm_st2(tlane + TM_LO + 32 * k + 6, 0u, 0u);
                }
            }
        }
        TCP_PROF(4);
        // ================================================================ reverse sweep, layers L .. 2 on tensor cores
        for (int l = L; l >= 2; --l) {
            const int m = l - 1;
            const int din = lay.d[l - 1], dout = lay.d[l];
#pragma unroll
            for (int i = 0; i < 9; ++i) reinterpret_cast<float4*>(wimg)[tid + i * TC_THREADS] = img[i];     // adjoint image of matrix m
            tm_wait_st();
            fence_async_smem();
            fence_before();
            __syncthreads();
            TCP_PROF(6);
            if (tid == 0) {
                fence_after();
                issue_layer<NS>(tbase, act_s, wimg_s, 64, (dout + 7) >> 3, (dout + 15) >> 4, fast);
                mma_commit(bar_s);
            }
            TCP_PROF(7);
            {   // prefetch: adjoint image of the next (shallower) matrix, or the first forward image of the next tile
                const uint8_t* nsrc = (l > 2) ? args.images + (size_t)(m - 1) * TC_IMG_LAYER + TC_IMG_SET : args.images + (size_t)1 * TC_IMG_LAYER;
#pragma unroll
                for (int i = 0; i < 9; ++i) img[i] = __ldg(reinterpret_cast<const float4*>(nsrc) + tid + i * TC_THREADS);
                if (l == 2) bias_pre = (tid < 64 && tid < lay.d[2]) ? __ldg(params + lay.boff[1] + tid) : 0.f;   // next tile's first TC layer
            }
            // ---- weight / bias gradient of layer l on the tensor cores (bf16 hi/mid operands)
            const float* stash_in = stash + (size_t)(l - 2) * (STASH_LAYER / 4);         // outputs of layer l-1 = inputs A of layer l
            const int NZ = (dout + 7) & ~7;                                              // N of the dW tile
            const int zc8 = NZ >> 3;                                                     // 8-unit chunks of Zbar
            // hh and hm in ONE MMA: the Zhi and Zmid images are contiguous (7 + 7 chunks of 8 units), so a B operand with
            // N = 56 + NZ starting at Zhi yields D[:, 0:56] = Ahi^T Zhi and D[:, 56:56+NZ] = Ahi^T Zmid; mh goes into D[:, 0:NZ].
            if (l >= 3) {   // pull the stash layer of the next (shallower) iteration towards L2 while this layer's MMAs run
                const char* nxt = reinterpret_cast<const char*>(stash + (size_t)(l - 3) * (STASH_LAYER / 4));
                for (int i = tid; i < STASH_LAYER / 128; i += TC_THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)i * 128));
            }
            if constexpr (PIPE) {
                // ------------------------------------------------------------ producer / consumer pipeline over (stream k, point half X)
                const int ct = tid - 32;                                                 // converter index 0..223 (warps 1..7)
                float4 pre[2][2];
                auto ldA = [&](int k, int X) {                                           // stash (L2) -> registers: tasks (c8 < 7, 64 points)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int t = ct + i * CONV_THREADS;
                        const int c8 = t >> 6, pp = 64 * X + (t & 63);
                        const float* src = stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (2 * c8) * 512 + pp * 4;
                        pre[i][0] = __ldcg(reinterpret_cast<const float4*>(src));
                        pre[i][1] = __ldcg(reinterpret_cast<const float4*>(src + 512));
                    }
                };
                auto stA = [&](int k, int X) {                                           // registers -> bf16 hi/mid images of A_k, half X
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int t = ct + i * CONV_THREADS;
                        const int c8 = t >> 6, pp = 64 * X + (t & 63);
                        uint4 hi, mid;
                        split8(pre[i][0], pre[i][1], hi, mid);
                        *reinterpret_cast<uint4*>(smem + DW_AHI + c8 * 2048 + pp * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + DW_AMID + c8 * 2048 + pp * 16) = mid;
                    }
                    if (ct < 64) {                                                       // chunk 7: units 56..63, unit 63 = ones row of the value stream
                        const int pp = 64 * X + ct;
                        const uint32_t one_hi = (k == 0) ? 0x3F800000u : 0u;             // bf16(1.0) in the high half = element 7
                        *reinterpret_cast<uint4*>(smem + DW_AHI + 7 * 2048 + pp * 16) = make_uint4(0u, 0u, 0u, one_hi);
                        *reinterpret_cast<uint4*>(smem + DW_AMID + 7 * 2048 + pp * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                };
                auto cvZ = [&](int k, int X) {                                           // ACT[k] (fp32 smem) -> bf16 hi/mid images of Zbar_k, half X
                    for (int t = ct; t < zc8 * 64; t += CONV_THREADS) {
                        const int c8 = t >> 6, pp = 64 * X + (t & 63);
                        const uint8_t* src = act + k * TC_ACT_STREAM + (2 * c8) * TC_CH + pp * 16;
                        uint4 hi, mid;
                        split8(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + TC_CH), hi, mid);
                        *reinterpret_cast<uint4*>(smem + DW_ZHI + c8 * 2048 + pp * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + DW_ZMID + c8 * 2048 + pp * 16) = mid;
                    }
                };
                if (warp > 0) {
                    ldA(0, 0);
                    cvZ(0, 0);                               // STAGE region: free while the adjoint MMAs read WIMG / ACT / LO
                    cvZ(0, 1);
                }
                mbar_wait(bar_s, parity);                    // adjoint MMAs done: WIMG region and the LO columns are free now
                parity ^= 1;
                fence_after();
                TCP_PROF(8);
                if (warp > 0) {
#pragma unroll 1
                    for (int k = 0; k < NS; ++k) {
#pragma unroll
                        for (int X = 0; X < 2; ++X) {
                            if (k > 0) {                     // the MMAs of (k-1, X) have read this half of the four images
                                mbar_wait(bar_empty + 8 * X, (pempty >> X) & 1u);
                                pempty ^= 1u << X;
                            }
                            stA(k, X);
                            if (X == 0) ldA(k, 1);           // next half's stash loads: issued as early as the registers are free
                            else if (k + 1 < NS) ldA(k + 1, 0);
                            if (k > 0) cvZ(k, X);
                            fence_async_smem();              // generic-proxy writes -> visible to the tensor core's async-proxy reads
                            mbar_arrive(bar_full + 8 * X);
                        }
                    }
                } else {
                    if (lane == 0) {
                        const uint32_t id2 = idesc_bf16_mn(64, 56 + NZ), id1 = idesc_bf16_mn(64, NZ);
                        const uint32_t d = tbase + TM_LO;
                        const uint64_t ahi = sdesc(smem_u32(smem + DW_AHI), 128, 2048), amid = sdesc(smem_u32(smem + DW_AMID), 128, 2048);
                        const uint64_t zhi = sdesc(smem_u32(smem + DW_ZHI), 128, 2048);
#pragma unroll 1
                        for (int k = 0; k < NS; ++k) {
#pragma unroll
                            for (int X = 0; X < 2; ++X) {
                                mbar_wait(bar_full + 8 * X, (pfull >> X) & 1u);
                                pfull ^= 1u << X;
                                fence_after();
#pragma unroll
                                for (int s = 0; s < 4; ++s) {        // 16 points per MMA: start address += 256 B
                                    const int s8 = 4 * X + s;
                                    const uint64_t o = (uint64_t)(s8 * 16);
                                    mma_bf16_ss(d, ahi + o, zhi + o, id2, (k > 0 || s8 > 0) ? 1u : 0u);
                                    mma_bf16_ss(d, amid + o, zhi + o, id1, 1u);
                                }
                                if (k < NS - 1) mma_commit(bar_empty + 8 * X);
                            }
                        }
                        mma_commit(bar_s);                   // everything issued so far, i.e. the whole weight-gradient tile
                    }
                    __syncwarp();
                }
                TCP_PROF(10);
                mbar_wait(bar_s, parity);
                parity ^= 1;
                TCP_PROF(11);
            } else {
                // ------------------------------------------------------------ pe_tc.cu order: convert | sync | MMA | wait, per stream
                float4 pre[4][2];
                auto load_A = [&](int k) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int t = tid + i * TC_THREADS;
                        if (t < 7 * TC_P) {
                            const int c8 = t >> 7, pp = t & 127;
                            const float* src = stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (2 * c8) * 512 + pp * 4;
                            pre[i][0] = __ldcg(reinterpret_cast<const float4*>(src));
                            pre[i][1] = __ldcg(reinterpret_cast<const float4*>(src + 512));
                        }
                    }
                };
                auto store_A = [&](int k) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int t = tid + i * TC_THREADS;
                        if (t < 7 * TC_P) {
                            const int c8 = t >> 7, pp = t & 127;
                            uint4 hi, mid;
                            split8(pre[i][0], pre[i][1], hi, mid);
                            *reinterpret_cast<uint4*>(smem + DW_AHI + c8 * 2048 + pp * 16) = hi;
                            *reinterpret_cast<uint4*>(smem + DW_AMID + c8 * 2048 + pp * 16) = mid;
                        }
                    }
                    if (tid < TC_P) {
                        const uint32_t one_hi = (k == 0) ? 0x3F800000u : 0u;
                        *reinterpret_cast<uint4*>(smem + DW_AHI + 7 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, one_hi);
                        *reinterpret_cast<uint4*>(smem + DW_AMID + 7 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                };
                auto conv_Z = [&](int k) {
                    for (int t = tid; t < zc8 * TC_P; t += TC_THREADS) {
                        const int c8 = t >> 7, pp = t & 127;
                        const uint8_t* src = act + k * TC_ACT_STREAM + (2 * c8) * TC_CH + pp * 16;
                        uint4 hi, mid;
                        split8(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + TC_CH), hi, mid);
                        *reinterpret_cast<uint4*>(smem + DW_ZHI + c8 * 2048 + pp * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + DW_ZMID + c8 * 2048 + pp * 16) = mid;
                    }
                };
                load_A(0);
                conv_Z(0);
                mbar_wait(bar_s, parity);
                parity ^= 1;
                fence_after();
                TCP_PROF(8);
#pragma unroll 1
                for (int k = 0; k < NS; ++k) {
                    store_A(k);
                    if (k < NS - 1) load_A(k + 1);
                    if (k > 0) conv_Z(k);
                    fence_async_smem();
                    fence_before();
                    __syncthreads();
                    TCP_PROF(9);
                    if (tid == 0) {
                        fence_after();
                        const uint32_t id2 = idesc_bf16_mn(64, 56 + NZ), id1 = idesc_bf16_mn(64, NZ);
                        const uint32_t d = tbase + TM_LO;
                        const uint64_t ahi = sdesc(smem_u32(smem + DW_AHI), 128, 2048), amid = sdesc(smem_u32(smem + DW_AMID), 128, 2048);
                        const uint64_t zhi = sdesc(smem_u32(smem + DW_ZHI), 128, 2048);
#pragma unroll
                        for (int s8 = 0; s8 < 8; ++s8) {
                            const uint64_t o = (uint64_t)(s8 * 16);
                            mma_bf16_ss(d, ahi + o, zhi + o, id2, (k > 0 || s8 > 0) ? 1u : 0u);
                            mma_bf16_ss(d, amid + o, zhi + o, id1, 1u);
                        }
                        mma_commit(bar_s);
                    }
                    TCP_PROF(10);
                    mbar_wait(bar_s, parity);
                    parity ^= 1;
                    TCP_PROF(11);
                }
            }
            fence_after();
            {   // drain the dW tile: rows i = 16*quadrant + lane (lane < 16), row 63 = bias gradient; h selects the column half
                const int quad = warp & 3;
                const int i = 16 * quad + lane;
                const int ldw = lay.ldw[m];
                float* gW = gpart + lay.woff[m];
                float* gB = gpart + lay.boff[m];
                const int c_lo = h ? 32 : 0, c_hi = h ? 56 : 32;
                for (int c = c_lo; c < c_hi; c += 8) {
                    float v[8], v2[8];
                    tm_ld8(tlane + TM_LO + c, v);
                    tm_ld8(tlane + TM_LO + 56 + c, v2);          // the hm block
                    tm_wait_ld();
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += v2[q];
                    if (lane < 16 && c < NZ) {
                        float* dst = (i < din) ? gW + (size_t)i * ldw + c : ((i == 63) ? gB + c : nullptr);
                        if (dst) {
                            if (c < ldw) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                            if (c + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4], v[5], v[6], v[7]));
                        }
                    }
                }
                // the tile aliased the lo-operand columns of streams 0..3 (320..431) including zero pads (units 56..63): restore them
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
            }
            TCP_PROF(12);
            // ---- through tanh of layer l-1: zbar^{l-1} from abar^{l-1} (TMEM) and the stashed outputs
            float4 Anext[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) Anext[k] = __ldcg(reinterpret_cast<const float4*>(stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (7 * h) * 512 + p * 4));
#pragma unroll 1
            for (int c = 7 * h; c < 7 * h + 7; ++c) {
                float ab[NS][4];
                float4 Av[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) Av[k] = Anext[k];
                if (c + 1 < 7 * h + 7) {       // prefetch the next chunk's stashed activations (L2 latency hidden behind this chunk's math)
#pragma unroll
                    for (int k = 0; k < NS; ++k) Anext[k] = __ldcg(reinterpret_cast<const float4*>(stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (c + 1) * 512 + p * 4));
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) tm_ld4(tlane + TM_ACC + 64 * k + 4 * c, ab[k]);
                tm_wait_ld();
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    float b[NS], Aa[NS];
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        b[k] = ab[k][u];
                        Aa[k] = (u == 0) ? Av[k].x : (u == 1) ? Av[k].y : (u == 2) ? Av[k].z : Av[k].w;
                    }
                    if (j < din) act_bwd<NS>(b, Aa);
                    else {
#pragma unroll
                        for (int k = 0; k < NS; ++k) b[k] = 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) ab[k][u] = b[k];
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const float4 v = make_float4(ab[k][0], ab[k][1], ab[k][2], ab[k][3]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                    tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(v.x, v.y), lo_pair(v.z, v.w));
                }
            }
            if (h == 1) {
#pragma unroll
                for (int k = 0; k < NS; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
            }
        }
        TCP_PROF(13);
        // ================================================================ layer 1 gradient (3 x d1 + bias): FFMA, fixed-order reduce
        __syncthreads();
        {
            const int d1 = lay.d[1];
            const int j = tid & 63, qq = tid >> 6;                    // 4 point quarters x 64 units
            float g0 = 0.f, g1 = 0.f, g2 = 0.f, gb = 0.f;
            if (j < d1) {
                const uint8_t* base = act + (j >> 2) * TC_CH + (j & 3) * 4;
#pragma unroll 4
                for (int s = 0; s < 32; ++s) {
                    const int pp = 32 * qq + s;
                    const float4 c4 = *reinterpret_cast<const float4*>(coord + 4 * pp);
                    const float zv = *reinterpret_cast<const float*>(base + pp * 16);
                    const float zx = *reinterpret_cast<const float*>(base + 1 * TC_ACT_STREAM + pp * 16);
                    const float zy = *reinterpret_cast<const float*>(base + 2 * TC_ACT_STREAM + pp * 16);
                    const float zt = *reinterpret_cast<const float*>(base + 3 * TC_ACT_STREAM + pp * 16);
                    g0 = fmaf(c4.x, zv, fmaf(Tc.in_scale[0], zx, g0));
                    g1 = fmaf(c4.y, zv, fmaf(Tc.in_scale[1], zy, g1));
                    g2 = fmaf(c4.z, zv, fmaf(Tc.in_scale[2], zt, g2));
                    gb += zv;
                }
            }
            *reinterpret_cast<float4*>(red + (qq * 64 + j) * 4) = make_float4(g0, g1, g2, gb);
            __syncthreads();
            if (tid < 64 && tid < d1) {
                float4 s = *reinterpret_cast<float4*>(red + tid * 4);
#pragma unroll
                for (int r = 1; r < 4; ++r) {
                    const float4 v = *reinterpret_cast<float4*>(red + (r * 64 + tid) * 4);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                float* gW = gpart + lay.woff[0];
                const int ldw = lay.ldw[0];
                __stcg(gW + tid, __ldcg(gW + tid) + s.x);
                __stcg(gW + ldw + tid, __ldcg(gW + ldw + tid) + s.y);
                __stcg(gW + 2 * ldw + tid, __ldcg(gW + 2 * ldw + tid) + s.z);
                float* gB = gpart + lay.boff[0];
                __stcg(gB + tid, __ldcg(gB + tid) + s.w);
            }
            __syncthreads();
        }
    }
    TCP_PROF(14);
    if (PROF) {
        if (args.prof && blockIdx.x == 0 && tid == 0)
            for (int i = 0; i < 16; ++i) atomicAdd(args.prof + i, prof_acc[PROF ? i : 0]);
    }
    // ---- loss-term partial sums (threads with h == 0 hold them): warp reduce, then 4 warps through smem (fixed order)
    {
        float tot[2 + PE_MAX_TERMS];
        tot[0] = warp_sum(tsum[0]);
        tot[1] = warp_sum(tsum[1]);
#pragma unroll
        for (int c = 0; c < PE_MAX_TERMS; ++c) tot[2 + c] = warp_sum(tsum2[c]);
        __syncthreads();
        if (h == 0 && lane == 0) {
#pragma unroll
            for (int c = 0; c < 2 + PE_MAX_TERMS; ++c) red[(2 + PE_MAX_TERMS) * warp + c] = tot[c];
        }
        __syncthreads();
        if (tid == 0) {
            float* tp = A.term_partials + (size_t)slot * PE_MAX_TERMS;
#pragma unroll
            for (int i = 0; i < PE_MAX_TERMS; ++i) tp[i] = 0.f;
            auto S = [&](int c) { const int st = 2 + PE_MAX_TERMS; return red[c] + red[st + c] + red[2 * st + c] + red[3 * st + c]; };
            tp[T.term[0]] += S(0) * A.inv_n;
            tp[T.term[1]] += S(1) * A.inv_n;
            if (args.n2 > 0) {
                const int nres2 = (T2.kind == PE_RES_TRACTION) ? 1 : T2.ncols;
                for (int c = 0; c < nres2; ++c) tp[T2.term[c]] += S(2 + c) * args.inv_n2;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int NS, bool PIPE, bool PROF>
int launch_variant(const TcpArgs& t, int slots, cudaStream_t st) {
    auto kern = resid_tcp_kernel<NS, PIPE, PROF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_tcp, %d): %s", SM_TOTAL + 1024, cudaGetErrorString(e)); return 2; }
    kern<<<slots, TC_THREADS, SM_TOTAL + 1024, st>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_tcp<%d,%d>: %s", NS, (int)PIPE, cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // namespace

// K = 5 with 5 outputs (F5) or K = 4 with 7 outputs (F7); at least one hidden->hidden matrix; hidden widths <= 56 (K = 56 operands)
int pe_tcp_supported(const pe_plan* plan, int K) {
    const PeLayout& lay = plan->lay;
    const int O = lay.d[lay.L];
    if (!((K == 5 && O == 5) || (K == 4 && O == 7)) || lay.L < 2) return 0;
    for (int l = 1; l < lay.L; ++l)
        if (lay.d[l] > 56) return 0;
    return 1;
}

static unsigned long long* g_tcp_prof = nullptr;
static int g_tcp_pipe = 1;
extern "C" void pe_debug_set_tcp_profile(unsigned long long* d_counters16) { g_tcp_prof = d_counters16; }
extern "C" void pe_debug_set_tcp_pipeline(int on) { g_tcp_pipe = on ? 1 : 0; }

size_t pe_tc_stash_floats_per_slot(const pe_plan* plan);

int pe_launch_resid_tcp(const pe_plan* plan, const PeResidArgs& a, int K, int fast, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2) {
    TcpArgs t;
    t.r = a;
    t.prof = g_tcp_prof;
    t.n2 = 0; t.points2 = nullptr; t.aux2 = nullptr; t.inv_n2 = 0.f;
    memset(&t.term2, 0, sizeof(t.term2));
    if (term2 && n2 > 0) {
        t.term2 = *term2; t.points2 = points2; t.n2 = n2; t.aux2 = term2->aux_k ? aux2 : nullptr;
        t.inv_n2 = 1.0f / (float)term2->n_global;
    }
    t.fast = fast;
    // scratch layout: [slots x stash floats][weight images]  (stash sized for 5 streams; K = 4 uses 4/5 of it)
    t.r.stash_floats = (int)pe_tc_stash_floats_per_slot(plan);
    uint8_t* images = reinterpret_cast<uint8_t*>(a.stash + (size_t)slots * t.r.stash_floats);
    t.images = images;
    tcp_prep_kernel<<<plan->lay.L * 16, 256, 0, st>>>(a.params, a.lay, images);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tcp_prep_kernel: %s", cudaGetErrorString(e)); return 3; }
    const bool prof = g_tcp_prof != nullptr;
    if (K == 5) {
        if (prof) return g_tcp_pipe ? launch_variant<5, true, true>(t, slots, st) : launch_variant<5, false, true>(t, slots, st);
        return g_tcp_pipe ? launch_variant<5, true, false>(t, slots, st) : launch_variant<5, false, false>(t, slots, st);
    }
    if (K == 4) {
        if (prof) return g_tcp_pipe ? launch_variant<4, true, true>(t, slots, st) : launch_variant<4, false, true>(t, slots, st);
        return g_tcp_pipe ? launch_variant<4, true, false>(t, slots, st) : launch_variant<4, false, false>(t, slots, st);
    }
    pe_set_error("pipelined tensor-core engine: K = %d not instantiated (4 or 5)", K);
    return 1;
}
