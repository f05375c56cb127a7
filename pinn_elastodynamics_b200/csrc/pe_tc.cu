// tcgen05 / TMEM tensor-core engine for the collocation residual (F5, K = 5 jet streams, hidden width <= 56).
//
// Same math and same C-ABI contract as the SIMT engine (pe_simt.cu); the layer GEMMs of the forward jet pass and of
// the adjoint pass run on the 5th-generation tensor cores:
//
//   forward  layer l : Z_k[128 pts x 64] = A_k[128 x 56] * W_l            (k = value, d/dx, d/dy, d/dt, d2/dt2)
//   adjoint  layer l : Abar_k[128 x 64] = Zbar_k[128 x 56] * W_l^T
//
// fp32 parity on TF32 tensor cores.  kind::tf32 TRUNCATES fp32 operands to 19 bits (measured, profiles/r1_umma_probe.txt),
// so each product is split A*W ~= A32*Whi + A32*Wlo + Alo*Wbf16 with
//     A32  = the fp32 activation in shared memory (hardware reads its top 19 bits = Ahi)
//     Alo  = A - Ahi rounded to bf16, kept in TENSOR MEMORY (A-from-TMEM operand, kind::f16, 2 per column)
//     Whi/Wlo = tf32 split of the weights, Wbf16 = bf16(W)   (operand images built once per step by tc_prep_kernel)
// all three accumulate into the same fp32 TMEM accumulator (mixed-kind accumulation, measured exact).  Error per
// product ~2^-20, i.e. fp32-class (tests hold the engine to the same tolerances as the SIMT engine).
// PE_ENGINE_TC_TF32 drops the two correction terms (fast mode, ~5e-4).
//
// Operand layout (all K-major, no swizzle -- the conventions pinned by tests/probe_umma.py):
//   activations  ACT[k][chunk c = unit/4][point p][4 fp32]   chunk stride 2064 B (2048 + 16 so that the FFMA
//                weight-gradient loops below are bank-conflict free), descriptor LBO = 2064, SBO = 128
//   weights      [chunk = k/4][n][4]  (n = output unit, forward) or (n = input unit, adjoint), LBO = N*16, SBO = 128
// TMEM (512 columns): accumulators of the 5 streams at columns 64k (k < 5); bf16 Alo operands at 320 + 32k.
//
// One CTA (256 threads) per SM, persistent over tiles of 128 points.  Thread (p, h): p = TMEM lane = point,
// h = warp/4 selects the unit half (chunks 7h..7h+6): all 8 warps run the epilogues (tanh + jet chain rule / its
// adjoint) straight out of tensor memory with tcgen05.ld and write the next operand back
// (st.shared 16 B + tcgen05.st), plus the stash the reverse sweep needs (global, L2-resident).
// The weight-gradient contraction dW_l = sum_{k,p} A_k[p][i] Zbar_k[p][j] (contraction over POINTS) would need
// MN-major TF32 operands, which tcgen05 only accepts in the SWIZZLE_128B_BASE32B layout (probe); in this round it runs
// on the FFMA pipe (8x8 register blocks, contraction split over 4 point-quarters inside a warp, shuffle-reduced)
// concurrently with the asynchronous adjoint MMAs of the same layer.
#include <cuda_bf16.h>
#include <cstring>
#include "pe_common.cuh"
#include "pe_device.cuh"

namespace {
using namespace pe_dev;

constexpr int TC_P = 128;
constexpr int TC_NCH = 14;                       // 4-float chunks per activation row (K = 56)
constexpr int TC_CH = 2064;                      // chunk stride in ACT / STAGE (bytes)
constexpr int TC_ACT_STREAM = TC_NCH * TC_CH;    // 28,896
constexpr int TC_THREADS = 256;
constexpr int TC_IMG_HI = 0, TC_IMG_LO = 14336, TC_IMG_BF = 28672, TC_IMG_SET = 36864, TC_IMG_LAYER = 2 * TC_IMG_SET;
constexpr int TC_STASH_STREAM = TC_NCH * 2048;   // stash is dense: [k][c][p][4]
constexpr int TC_STASH_LAYER = 5 * TC_STASH_STREAM;

constexpr int SM_ACT = 0;
constexpr int SM_WIMG = SM_ACT + 5 * TC_ACT_STREAM;          // 144,480
constexpr int SM_STAGE = SM_WIMG + TC_IMG_SET;               // 181,344: bf16 hi/mid images of Zbar for the weight-gradient MMAs
constexpr int SM_MISC = SM_STAGE + TC_ACT_STREAM;            // 210,240
constexpr int SM_COORD = SM_MISC + 64;                       // 128 x 4 floats
constexpr int SM_RED = SM_COORD + 128 * 16;                  // 4 x 64 x 4 floats scratch (layer-1 gradient) / term sums
constexpr int SM_BIAS = SM_RED + 4096;                        // 64 floats: bias of the current layer
constexpr int SM_W0 = SM_BIAS + 256;                          // [4][64] floats: first-layer weight rows 0..2 and bias
constexpr int SM_TOTAL = SM_W0 + 1024;

constexpr uint32_t TM_ACC = 0, TM_LO = 320;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ uint32_t idesc_bf16(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }

__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
// smem -> TMEM copy of one K-step of an activation operand (128 rows x 32 B -> 8 columns); ordered with the MMAs of the issuing thread
__device__ __forceinline__ void tm_cp_128x256b(uint32_t taddr, uint64_t sd) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sd) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t addr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bf16 pair of the truncation residues (x - tf32_trunc(x)) of two values, element 0 in the low half
__device__ __forceinline__ uint32_t lo_pair(float x0, float x1) {
    float l0 = x0 - __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
    float l1 = x1 - __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
    __nv_bfloat162 b = __floats2bfloat162_rn(l0, l1);
    return *reinterpret_cast<uint32_t*>(&b);
}

// ---------------------------------------------------------------------------------------------- weight images
// One block per weight matrix.  Builds, from the padded fp32 parameters, the six tensor-core operand images of the
// matrix: forward B operand [n = out unit j][k = in unit i] and adjoint B operand [n = i][k = j], each as tf32-hi,
// tf32-lo (4-float chunks) and bf16 (8-element chunks).  Zero padded to K = 56 (64 for bf16), N = 64.
__global__ void tc_prep_kernel(const float* __restrict__ params, PeLayout lay, uint8_t* __restrict__ images) {
    const int m = blockIdx.x >> 4;                  // matrix index 0..L-1; 16 blocks of 256 elements per matrix (latency-bound otherwise)
    const int din = lay.d[m], dout = lay.d[m + 1], ldw = lay.ldw[m];
    const float* W = params + lay.woff[m];
    uint8_t* img = images + (size_t)m * TC_IMG_LAYER;
    float* fhi = reinterpret_cast<float*>(img + TC_IMG_HI);
    float* flo = reinterpret_cast<float*>(img + TC_IMG_LO);
    __nv_bfloat16* fbf = reinterpret_cast<__nv_bfloat16*>(img + TC_IMG_BF);
    float* ahi = reinterpret_cast<float*>(img + TC_IMG_SET + TC_IMG_HI);
    float* alo = reinterpret_cast<float*>(img + TC_IMG_SET + TC_IMG_LO);
    __nv_bfloat16* abf = reinterpret_cast<__nv_bfloat16*>(img + TC_IMG_SET + TC_IMG_BF);
    const int NF = (dout <= 16) ? 16 : 64;          // forward N
    {
        const int e = (blockIdx.x & 15) * 256 + threadIdx.x;
        const int i = e >> 6, j = e & 63;
        const float w = (i < din && j < dout) ? W[(size_t)i * ldw + j] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
        const float lo = w - hi;
        const __nv_bfloat16 bf = __float2bfloat16_rn(w);
        if (j < NF) {                               // forward image: rows n = j, K index i
            if (i < 56) { const int o = (i >> 2) * (NF * 4) + j * 4 + (i & 3); fhi[o] = hi; flo[o] = lo; }
            fbf[(i >> 3) * (NF * 8) + j * 8 + (i & 7)] = bf;
        }
        // adjoint image: rows n = i, K index j
        if (j < 56) { const int o = (j >> 2) * 256 + i * 4 + (j & 3); ahi[o] = hi; alo[o] = lo; }
        abf[(j >> 3) * 512 + i * 8 + (j & 7)] = bf;
    }
}

struct TcArgs {
    PeResidArgs r;
    const uint8_t* images;
    int fast;            // 1 = single-pass TF32
    // optional second, primal-only point set (traction / data term, K = 1) fused into the same launch as extra tiles
    pe_term_desc term2;
    const float* points2;
    const float* aux2;
    int n2;
    float inv_n2;
    unsigned long long* prof;   // optional: 16 phase-cycle counters written by CTA 0 / thread 0 (pe_debug_set_tc_profile)
};

// phase timing (debug): CTA 0, thread 0 accumulates clock64 deltas per phase
#define TC_PROF(slot) do { if (args.prof && blockIdx.x == 0 && tid == 0) { long long now_ = clock64(); prof_acc[slot] += (unsigned long long)(now_ - prof_t); prof_t = now_; } } while (0)

// ---- issue the MMAs of one layer GEMM for all 5 streams (one thread).  ksteps = K/8 (tf32), kb = K/16 (bf16).
// Back-to-back MMAs that accumulate into the SAME TMEM tile form a dependent chain and run at pipeline LATENCY (~100 cycles
// per 128x64x8 MMA measured, against a 32-cycle throughput floor; moving the A operand to tensor memory did not change that).
// The 5 jet streams have 5 independent accumulators, so the K-steps are issued stream-interleaved: consecutive MMAs never
// touch the same accumulator.
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
            for (int k = 0; k < 5; ++k) mma_tf32_ss(d0 + 64u * k, a0 + k * a_stream + s * a_step, b_hi + s * b_step, id32, s > 0);
        }
    if (!fast) {
#pragma unroll
        for (int s = 0; s < 7; ++s)
            if (s < ksteps) {
#pragma unroll
                for (int k = 0; k < 5; ++k) mma_tf32_ss(d0 + 64u * k, a0 + k * a_stream + s * a_step, b_lo + s * b_step, id32, 1);
            }
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < kb) {
#pragma unroll
                for (int k = 0; k < 5; ++k) mma_bf16_ts(d0 + 64u * k, tbase + TM_LO + 32u * k + 8u * s, b_bf + s * b_step, id16, 1);
            }
    }
}

// ---------------------------------------------------------------------------------------------- weight gradient on tensor cores
// dW[i][j] = sum_{k,p} A_k[p][i] * Zbar_k[p][j] contracts over POINTS: both operands are "MN-major" (unit-contiguous).
// tcgen05 accepts MN-major 16-bit operands in the no-swizzle layout [chunk of 8 units][point][8] (probe), but MN-major
// TF32 only in SWIZZLE_128B_BASE32B.  So the two operands are split into bf16 (hi, mid) pairs (16 mantissa bits) and the
// product is hh + hm + mh on kind::f16 (error ~4e-7 of the block maximum, tests/ + /oracle emulation), accumulated over
// the 5 streams in one M=64 x N<=56 fp32 TMEM tile that aliases the (dead at that point) bf16 "lo" operand columns.
// Row 63 of the A operand of the value stream is a row of ones: D[63][j] = sum_p Zbar_0[p][j] = bias gradient, for free.
constexpr int DW_AHI = SM_WIMG, DW_AMID = SM_WIMG + 16384;        // [8 chunks][128 points][8 bf16]
constexpr int DW_ZHI = SM_STAGE, DW_ZMID = SM_STAGE + 14336;      // [7 chunks][128 points][8 bf16]

__device__ __forceinline__ uint32_t idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
// 8 fp32 -> 8 bf16 (hi) and 8 bf16 (mid = bf16(x - hi)), packed as two uint4
__device__ __forceinline__ void split8(const float4& v0, const float4& v1, uint4& hi, uint4& mid) {
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t h[4], m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hb);
        __nv_bfloat162 mb = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<uint32_t*>(&hb);
        m[i] = *reinterpret_cast<uint32_t*>(&mb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    mid = make_uint4(m[0], m[1], m[2], m[3]);
}

__global__ void __launch_bounds__(TC_THREADS, 1) resid_tc_kernel(const TcArgs args) {
    extern __shared__ __align__(1024) uint8_t smem[];
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
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + SM_MISC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_MISC + 16);
    float* coord = reinterpret_cast<float*>(smem + SM_COORD);       // [128][4]: a0x, a0y, a0t, valid
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);   // biases come from smem: with 217 KB of smem there is no L1 left and a
    float* sw0 = reinterpret_cast<float*>(smem + SM_W0);       // global bias load in the tanh dependency chain costs an L2 round trip
    const uint32_t act_s = smem_u32(act), wimg_s = smem_u32(wimg), bar_s = smem_u32(mbar);

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
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(1));
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
    unsigned long long prof_acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) prof_acc[i] = 0ull;
    long long prof_t = clock64();
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
        }
        __syncthreads();
        TC_PROF(15);
        // ================================================================ layer 1 (3 -> d1): per-thread FFMA
        {
            const float4 c4 = *reinterpret_cast<const float4*>(coord + 4 * p);
            const int dout = lay.d[1];
            float* st = stash;                                         // stash layer index 0 = outputs of layer 1
#pragma unroll 1
            for (int c = 7 * h; c < 7 * h + 7; ++c) {
                float o[5][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    float z[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    if (j < dout) {
                        const float w0 = sw0[j], w1 = sw0[64 + j], w2 = sw0[128 + j];
                        z[0] = fmaf(c4.x, w0, fmaf(c4.y, w1, c4.z * w2));
                        z[1] = Tc.in_scale[0] * w0; z[2] = Tc.in_scale[1] * w1; z[3] = Tc.in_scale[2] * w2; z[4] = 0.f;
                        act_fwd<5, true>(z, sw0[192 + j]);
                    }
#pragma unroll
                    for (int k = 0; k < 5; ++k) o[k][u] = z[k];
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 v = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                    __stcg(reinterpret_cast<float4*>(st + (size_t)k * (TC_STASH_STREAM / 4) + c * 512 + p * 4), v);
                    tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(o[k][0], o[k][1]), lo_pair(o[k][2], o[k][3]));
                }
            }
        }
        TC_PROF(0);
        // ================================================================ forward: hidden layers 2..L-1 and the output layer L
        for (int l = 2; l <= L; ++l) {
            const int m = l - 1;                                       // weight matrix index
            const int dout = lay.d[l];
            const int NF = (dout <= 16) ? 16 : 64;
            // operand images of matrix m (prefetched into registers one phase ahead) -> smem, and the layer's bias
#pragma unroll
            for (int i = 0; i < 9; ++i) reinterpret_cast<float4*>(wimg)[tid + i * TC_THREADS] = img[i];
            if (tid < 64) sbias[tid] = bias_pre;
            tm_wait_st();
            fence_async_smem();
            fence_before();
            __syncthreads();
            TC_PROF(1);
            if (tid == 0) {
                fence_after();
                issue_layer(tbase, act_s, wimg_s, NF, (lay.d[l - 1] + 7) >> 3, (lay.d[l - 1] + 15) >> 4, fast);
                mma_commit(bar_s);
            }
            TC_PROF(2);
            {   // prefetch the next operand image while the MMAs run: forward image of the next matrix, or the adjoint image of the last one
                const uint8_t* nsrc = (l < L) ? args.images + (size_t)(m + 1) * TC_IMG_LAYER : args.images + (size_t)m * TC_IMG_LAYER + TC_IMG_SET;
#pragma unroll
                for (int i = 0; i < 9; ++i) img[i] = __ldg(reinterpret_cast<const float4*>(nsrc) + tid + i * TC_THREADS);
                if (l < L) bias_pre = (tid < 64 && tid < lay.d[l + 1]) ? __ldg(params + lay.boff[l] + tid) : 0.f;
            }
            mbar_wait(bar_s, parity);
            parity ^= 1;
            fence_after();
            TC_PROF(3);
            if (l < L) {
                float* st = stash + (size_t)(l - 1) * (TC_STASH_LAYER / 4);
#pragma unroll 1
                for (int c = 7 * h; c < 7 * h + 7; ++c) {
                    float z[5][4];
#pragma unroll
                    for (int k = 0; k < 5; ++k) tm_ld4(tlane + TM_ACC + 64 * k + 4 * c, z[k]);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c + u;
                        float zz[5];
#pragma unroll
                        for (int k = 0; k < 5; ++k) zz[k] = z[k][u];
                        if (j < dout) act_fwd<5, true>(zz, sbias[j]);
                        else {
#pragma unroll
                            for (int k = 0; k < 5; ++k) zz[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < 5; ++k) z[k][u] = zz[k];
                    }
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const float4 v = make_float4(z[k][0], z[k][1], z[k][2], z[k][3]);
                        *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                        __stcg(reinterpret_cast<float4*>(st + (size_t)k * (TC_STASH_STREAM / 4) + c * 512 + p * 4), v);
                        tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(z[k][0], z[k][1]), lo_pair(z[k][2], z[k][3]));
                    }
                }
                if (h == 1) {      // units 56..63 of the lo operands: stream 0's were used as the TS operand window, keep them finite/zero
#pragma unroll
                    for (int k = 0; k < 5; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
                TC_PROF(5);
            } else if (h == 0) {
                // ---------------- outputs -> residuals -> loss partials -> seeds (adjoint of the outputs)
                float Y[5][PE_UJ];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    float v[8];
                    tm_ld8(tlane + TM_ACC + 64 * k, v);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) Y[k][u] = (u < 8 && u < dout) ? v[u < 8 ? u : 0] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < PE_UJ; ++u) if (u < dout) Y[0][u] += sbias[u];
                if (!sec) {
                    const float* aux_row = A.aux ? A.aux + (size_t)(valid ? pt : 0) * 50 : nullptr;
                    residual_stage<5>(Y, T, aux_row, row, valid, A.inv_n, tsum);
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
                        for (int k = 1; k < 5; ++k) Y[k][u] = 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 v0 = make_float4(Y[k][0], Y[k][1], Y[k][2], Y[k][3]);
                    const float4 v1 = make_float4(Y[k][4], 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + 0 * TC_CH + p * 16) = v0;
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + 1 * TC_CH + p * 16) = v1;
                    tm_st2(tlane + TM_LO + 32 * k + 0, lo_pair(v0.x, v0.y), lo_pair(v0.z, v0.w));
                    tm_st2(tlane + TM_LO + 32 * k + 2, lo_pair(v1.x, 0.f), 0u);
                    tm_st2(tlane + TM_LO + 32 * k + 4, 0u, 0u);        // units 8..15: read by the K = 16 bf16 step of the adjoint MMA,
                    tm_st2(tlane + TM_LO + 32 * k + 6, 0u, 0u);        // may hold the TS operand window of the forward pass
                }
            }
        }
        TC_PROF(4);
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
            TC_PROF(6);
            if (tid == 0) {
                fence_after();
                issue_layer(tbase, act_s, wimg_s, 64, (dout + 7) >> 3, (dout + 15) >> 4, fast);
                mma_commit(bar_s);
            }
            TC_PROF(7);
            {   // prefetch: adjoint image of the next (shallower) matrix, or the first forward image of the next tile
                const uint8_t* nsrc = (l > 2) ? args.images + (size_t)(m - 1) * TC_IMG_LAYER + TC_IMG_SET : args.images + (size_t)1 * TC_IMG_LAYER;
#pragma unroll
                for (int i = 0; i < 9; ++i) img[i] = __ldg(reinterpret_cast<const float4*>(nsrc) + tid + i * TC_THREADS);
                if (l == 2) bias_pre = (tid < 64 && tid < lay.d[2]) ? __ldg(params + lay.boff[1] + tid) : 0.f;   // next tile's first TC layer
            }
            // ---- weight / bias gradient of layer l on the tensor cores (bf16 hi/mid operands, see above)
            const float* stash_in = stash + (size_t)(l - 2) * (TC_STASH_LAYER / 4);      // outputs of layer l-1 = inputs A of layer l
            const int NZ = (dout + 7) & ~7;                                              // N of the dW tile
            const int zc8 = NZ >> 3;                                                     // 8-unit chunks of Zbar
            float4 pre[4][2];
            auto load_A = [&](int k) {                                                   // stash (L2) -> registers, tasks (c8 < 7, p)
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
            auto store_A = [&](int k) {                                                  // registers -> bf16 hi/mid operand images
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
                if (tid < TC_P) {                                                        // chunk 7: units 56..63, unit 63 = ones row of the value stream
                    const uint32_t one_hi = (k == 0) ? 0x3F800000u : 0u;                 // bf16(1.0) in the high half = element 7
                    *reinterpret_cast<uint4*>(smem + DW_AHI + 7 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, one_hi);
                    *reinterpret_cast<uint4*>(smem + DW_AMID + 7 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
            };
            auto conv_Z = [&](int k) {                                                   // ACT[k] (fp32, smem) -> bf16 hi/mid images
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
            if (l >= 3) {   // pull the stash layer of the next (shallower) iteration towards L2 while this layer's MMAs run
                const char* nxt = reinterpret_cast<const char*>(stash + (size_t)(l - 3) * (TC_STASH_LAYER / 4));
                for (int i = tid; i < TC_STASH_LAYER / 128; i += TC_THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)i * 128));
            }
            conv_Z(0);                                   // STAGE region: free while the adjoint MMAs read WIMG / ACT / LO
            mbar_wait(bar_s, parity);                    // adjoint MMAs done: WIMG region and the LO columns are free now
            parity ^= 1;
            fence_after();
            TC_PROF(8);
#pragma unroll 1
            for (int k = 0; k < 5; ++k) {
                store_A(k);
                if (k < 4) load_A(k + 1);                // next stream's stash loads: issued as early as the registers are free
                if (k > 0) conv_Z(k);
                fence_async_smem();
                fence_before();
                __syncthreads();
                TC_PROF(9);
                if (tid == 0) {
                    fence_after();
                    // hh and hm in ONE MMA: the Zhi and Zmid images are contiguous (7 + 7 chunks of 8 units), so a B operand with
                    // N = 112 starting at Zhi yields D[:, 0:56] = Ahi^T Zhi and D[:, 56:112] = Ahi^T Zmid; mh goes into D[:, 0:56].
                    const uint32_t id2 = idesc_bf16_mn(64, 56 + NZ), id1 = idesc_bf16_mn(64, NZ);
                    const uint32_t d = tbase + TM_LO;
                    const uint64_t ahi = sdesc(smem_u32(smem + DW_AHI), 128, 2048), amid = sdesc(smem_u32(smem + DW_AMID), 128, 2048);
                    const uint64_t zhi = sdesc(smem_u32(smem + DW_ZHI), 128, 2048);
#pragma unroll
                    for (int s8 = 0; s8 < 8; ++s8) {     // 16 points per MMA: start address += 256 B
                        const uint64_t o = (uint64_t)(s8 * 16);
                        mma_bf16_ss(d, ahi + o, zhi + o, id2, (k > 0 || s8 > 0) ? 1u : 0u);
                        mma_bf16_ss(d, amid + o, zhi + o, id1, 1u);
                    }
                    mma_commit(bar_s);
                }
                TC_PROF(10);
                mbar_wait(bar_s, parity);
                parity ^= 1;
                TC_PROF(11);
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
                // the two tiles aliased the lo-operand columns of streams 0..3 (320..439) including zero pads (units 56..63): restore them
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
            }
            TC_PROF(12);
            // ---- through tanh of layer l-1: zbar^{l-1} from abar^{l-1} (TMEM) and the stashed outputs
            float4 Anext[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) Anext[k] = __ldcg(reinterpret_cast<const float4*>(stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (7 * h) * 512 + p * 4));
#pragma unroll 1
            for (int c = 7 * h; c < 7 * h + 7; ++c) {
                float ab[5][4];
                float4 Av[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) Av[k] = Anext[k];
                if (c + 1 < 7 * h + 7) {       // prefetch the next chunk's stashed activations (L2 latency hidden behind this chunk's math)
#pragma unroll
                    for (int k = 0; k < 5; ++k) Anext[k] = __ldcg(reinterpret_cast<const float4*>(stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (c + 1) * 512 + p * 4));
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) tm_ld4(tlane + TM_ACC + 64 * k + 4 * c, ab[k]);
                tm_wait_ld();
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    float b[5], Aa[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        b[k] = ab[k][u];
                        Aa[k] = (u == 0) ? Av[k].x : (u == 1) ? Av[k].y : (u == 2) ? Av[k].z : Av[k].w;
                    }
                    if (j < din) act_bwd<5>(b, Aa);
                    else {
#pragma unroll
                        for (int k = 0; k < 5; ++k) b[k] = 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 5; ++k) ab[k][u] = b[k];
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 v = make_float4(ab[k][0], ab[k][1], ab[k][2], ab[k][3]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                    tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(v.x, v.y), lo_pair(v.z, v.w));
                }
            }
            if (h == 1) {
#pragma unroll
                for (int k = 0; k < 5; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
            }
        }
        TC_PROF(13);
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
    TC_PROF(14);
    if (args.prof && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(args.prof + i, prof_acc[i]);
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

}  // namespace

int pe_tc_supported(const pe_plan* plan, int K, int engine) {
    (void)engine;
    const PeLayout& lay = plan->lay;
    if (K != 5 || lay.d[lay.L] != 5 || lay.L < 2) return 0;
    for (int l = 1; l < lay.L; ++l)
        if (lay.d[l] > 56) return 0;     // K = 56 operands, 7 x 8-unit row blocks in the FFMA weight gradient
    return 1;
}

int pe_tc_slots(const pe_plan* plan, int n_points) {      // for a fused launch pass 128 * (tiles of set 1 + tiles of set 2)
    int ntiles = (n_points + TC_P - 1) / TC_P;
    int s = ntiles < plan->sms ? ntiles : plan->sms;
    return s < 1 ? 1 : s;
}

size_t pe_tc_stash_floats_per_slot(const pe_plan* plan) { return (size_t)(plan->lay.L - 1) * (TC_STASH_LAYER / 4) + 64; }
size_t pe_tc_image_floats(const pe_plan* plan) { return (size_t)plan->lay.L * (TC_IMG_LAYER / 4); }

static unsigned long long* g_tc_prof = nullptr;
extern "C" void pe_debug_set_tc_profile(unsigned long long* d_counters16) { g_tc_prof = d_counters16; }

int pe_launch_resid_tc(const pe_plan* plan, const PeResidArgs& a, int K, int engine, int slots, cudaStream_t st,
                       const pe_term_desc* term2, const float* points2, int n2, const float* aux2) {
    (void)K;
    TcArgs t;
    t.r = a;
    t.prof = g_tc_prof;
    t.n2 = 0; t.points2 = nullptr; t.aux2 = nullptr; t.inv_n2 = 0.f;
    memset(&t.term2, 0, sizeof(t.term2));
    if (term2 && n2 > 0) {
        t.term2 = *term2; t.points2 = points2; t.n2 = n2; t.aux2 = term2->aux_k ? aux2 : nullptr;
        t.inv_n2 = 1.0f / (float)term2->n_global;
    }
    t.fast = (engine == PE_ENGINE_TC_TF32) ? 1 : 0;
    // scratch layout: [slots x stash floats][weight images]
    t.r.stash_floats = (int)pe_tc_stash_floats_per_slot(plan);
    uint8_t* images = reinterpret_cast<uint8_t*>(a.stash + (size_t)slots * t.r.stash_floats);
    t.images = images;
    tc_prep_kernel<<<plan->lay.L * 16, 256, 0, st>>>(a.params, a.lay, images);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tc_prep_kernel: %s", cudaGetErrorString(e)); return 3; }
    e = cudaFuncSetAttribute(resid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL + 1024);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_tc, %d): %s", SM_TOTAL + 1024, cudaGetErrorString(e)); return 2; }
    resid_tc_kernel<<<slots, TC_THREADS, SM_TOTAL + 1024, st>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_tc: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}
