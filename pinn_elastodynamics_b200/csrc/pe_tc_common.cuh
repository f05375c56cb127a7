// tcgen05 / TMEM building blocks of the pipelined tensor-core engine (pe_tcp.cu): shared-memory map, UMMA descriptors,
// PTX wrappers and the operand-image kernel.  Conventions (descriptor fields, no-swizzle K-major / MN-major layouts,
// mixed-kind accumulation) are the ones pinned on hardware by tests/probe_umma.py (profiles/r1_umma_probe.txt) and
// documented at the top of pe_tc.cu, which keeps its own private copies (the round-1 verified engine is left untouched).
#pragma once
#include <cuda_bf16.h>
#include "pe_common.cuh"

namespace pe_tcc {

constexpr int TC_P = 128;
constexpr int TC_NCH = 14;                       // 4-float chunks per activation row (K = 56)
constexpr int TC_CH = 2064;                      // chunk stride in ACT / STAGE (bytes)
constexpr int TC_ACT_STREAM = TC_NCH * TC_CH;    // 28,896
constexpr int TC_THREADS = 256;
constexpr int TC_IMG_HI = 0, TC_IMG_LO = 14336, TC_IMG_BF = 28672, TC_IMG_SET = 36864, TC_IMG_LAYER = 2 * TC_IMG_SET;
constexpr int TC_STASH_STREAM = TC_NCH * 2048;   // stash is dense: [k][c][p][4]
constexpr int TC_MAX_STREAMS = 5;                // smem / stash are laid out for 5 jet streams; K = 4 leaves the last one idle

constexpr int SM_ACT = 0;
constexpr int SM_WIMG = SM_ACT + TC_MAX_STREAMS * TC_ACT_STREAM;   // 144,480
constexpr int SM_STAGE = SM_WIMG + TC_IMG_SET;               // 181,344: bf16 hi/mid images of Zbar for the weight-gradient MMAs
constexpr int SM_MISC = SM_STAGE + TC_ACT_STREAM;            // 210,240: mbarriers + TMEM base slot
constexpr int SM_COORD = SM_MISC + 64;                       // 128 x 4 floats
constexpr int SM_RED = SM_COORD + 128 * 16;                  // 4 x 64 x 4 floats scratch (layer-1 gradient) / term sums
constexpr int SM_BIAS = SM_RED + 4096;                       // 64 floats: bias of the current layer
constexpr int SM_W0 = SM_BIAS + 256;                         // [4][64] floats: first-layer weight rows 0..2 and bias
constexpr int SM_TOTAL = SM_W0 + 1024;

// byte offsets inside SM_MISC
constexpr int MB_MAIN = 0;        // layer-GEMM / "all weight-gradient MMAs done" barrier (count 1, tcgen05.commit)
constexpr int MB_TMEM = 16;       // TMEM base address written by tcgen05.alloc
constexpr int MB_FULL = 24;       // [2] operand half X converted (count = converter threads)
constexpr int MB_EMPTY = 40;      // [2] MMAs of operand half X complete (count 1, tcgen05.commit)

constexpr uint32_t TM_ACC = 0, TM_LO = 320;

// weight-gradient operand images (bf16 hi / mid, MN-major, [chunk of 8 units][128 points][8])
constexpr int DW_AHI = SM_WIMG, DW_AMID = SM_WIMG + 16384;
constexpr int DW_ZHI = SM_STAGE, DW_ZMID = SM_STAGE + 14336;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ uint32_t idesc_bf16(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ uint32_t idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
// kind::f16 with scale-input-d: D = A B + D * 2^-11 (always accumulates)
__device__ __forceinline__ void mma_f16_ss_scaled11(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded: a phase that does not complete within ~2 s of SM clock (a tile takes ~100 us) is a protocol error -- trap (the launch
// fails with an error the host reports) instead of spinning on the SM forever.  (A poll count is not a time bound: one try_wait may
// suspend the thread for a hardware-defined interval; round 2 saw a 2^26-poll bound outlive a 300 s test timeout.)
// The retries carry a suspend-time hint: without one a failed try_wait returns after a few cycles and the waiting warps poll -- ncu
// counted half of the kernel's executed instructions in this loop, on the schedulers the epilogue warps need; with the hint the warp
// sleeps in hardware until the phase completes (wake-up ~60 cycles) or the hint expires.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (!ok && clock64() - t0 > (1ll << 32)) __trap();
    }
}
// release-arrive of one thread (the generic-proxy writes before it were made visible to the async proxy by fence_async_smem)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
// ---- warp-specialised kernels (pe_tcs.cu): named barrier over a subset of warps, 1-D bulk TMA global -> shared
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bytes: multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src_global, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_global), "r"(bytes), "r"(bar) : "memory");
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

// Operand images: one set per weight matrix, forward B operand [n = out unit j][k = in unit i] and adjoint B operand
// [n = i][k = j], each as tf32-hi, tf32-lo (4-float chunks) and bf16 (8-element chunks); zero padded to K = 56 (64 for
// bf16), N = 64 (16 for an output layer of <= 16 units).  16 blocks of 256 elements per matrix.
static __global__ void tcp_prep_kernel(const float* __restrict__ params, PeLayout lay, uint8_t* __restrict__ images) {
    const int m = blockIdx.x >> 4;
    const int din = lay.d[m], dout = lay.d[m + 1], ldw = lay.ldw[m];
    const float* W = params + lay.woff[m];
    uint8_t* img = images + (size_t)m * TC_IMG_LAYER;
    float* fhi = reinterpret_cast<float*>(img + TC_IMG_HI);
    float* flo = reinterpret_cast<float*>(img + TC_IMG_LO);
    __nv_bfloat16* fbf = reinterpret_cast<__nv_bfloat16*>(img + TC_IMG_BF);
    float* ahi = reinterpret_cast<float*>(img + TC_IMG_SET + TC_IMG_HI);
    float* alo = reinterpret_cast<float*>(img + TC_IMG_SET + TC_IMG_LO);
    __nv_bfloat16* abf = reinterpret_cast<__nv_bfloat16*>(img + TC_IMG_SET + TC_IMG_BF);
    const int NF = (dout <= 16) ? 16 : 64;
    const int e = (blockIdx.x & 15) * 256 + threadIdx.x;
    const int i = e >> 6, j = e & 63;
    const float w = (i < din && j < dout) ? W[(size_t)i * ldw + j] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const float lo = w - hi;
    const __nv_bfloat16 bf = __float2bfloat16_rn(w);
    if (j < NF) {
        if (i < 56) { const int o = (i >> 2) * (NF * 4) + j * 4 + (i & 3); fhi[o] = hi; flo[o] = lo; }
        fbf[(i >> 3) * (NF * 8) + j * 8 + (i & 7)] = bf;
    }
    if (j < 56) { const int o = (j >> 2) * 256 + i * 4 + (j & 3); ahi[o] = hi; alo[o] = lo; }
    abf[(j >> 3) * 512 + i * 8 + (j & 7)] = bf;
}

struct TcpArgs {
    PeResidArgs r;
    const uint8_t* images;
    int fast;            // 1 = single-pass TF32
    // optional second, primal-only point set (traction / data term, K = 1) fused into the same launch as extra tiles
    pe_term_desc term2;
    const float* points2;
    const float* aux2;
    int n2;
    float inv_n2;
    unsigned long long* prof;   // PROF instantiations only: 16 phase-cycle counters (CTA 0 / thread 0)
};

}  // namespace pe_tcc
