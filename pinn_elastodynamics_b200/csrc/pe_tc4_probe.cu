// Forward-only PROBE of the next tensor-core engine's arithmetic -- experimental, opt-in, NOT on any product path and
// NOT yet run on hardware (written at the end of round 1 after the GPU budget was spent; tests/test_gpu_tc4_forward.py runs it only when
// PE_TEST_TC4=1).  It answers, in one launch, the questions the full engine depends on:
//   * one 16-bit split for every operand:  X = Xhi + Xlo,  Xhi = fp16(X),  Xlo = bf16(X - Xhi)
//   * layer GEMM  Z = A W  ~=  Ahi Whi + Ahi Wlo + Alo Whi  as three kind::f16 MMAs per K-step of 16 into ONE fp32 accumulator, with
//     MIXED operand formats (fp16 x bf16) selected per MMA through the A / B format fields of the instruction descriptor
//   * activation planes  [chunk of 8 units][128 points][8 x 16 bit]  (K-major, no swizzle, LBO = 2,048 B, SBO = 128 B), 7 chunks = 56 units;
//     the fourth K-step reads chunk 6 twice (LBO = 0) against weight rows 56..63 that are zero, instead of a chunk 7 that does not exist
// CPU model of this arithmetic: tests/emulate_engine_precision.py (loss 1.2e-6, gradient blocks 2-3e-6 end to end).
// Structure is deliberately the simplest that can work (no warp specialisation, no TMA, images copied by all threads, one MMA issue
// thread that waits for each layer): it is a correctness probe of formats and descriptors, not a performance kernel.
// Output: d_out[n][K][O] like pe_forward_jets (reference lines: neural_net plate:308-320 / inf:188-199; jets SURVEY A.1).
#include <cuda_fp16.h>
#include <cstring>
#include "pe_device.cuh"
#include "pe_tc_common.cuh"

namespace {
using namespace pe_dev;
using namespace pe_tcc;

constexpr int Q_THREADS = 256;
constexpr int Q_CH = 2048;                      // one chunk of 8 units: 128 points x 16 B
constexpr int Q_PLANE7 = 7 * Q_CH;              // 14,336: 56 units (the full engine's plane: no room for more)
constexpr int Q_PLANE8 = 8 * Q_CH;              // 16,384: 56 units + one chunk that stays zero (control variant of this probe)
constexpr int Q_IMG_HALF = 8 * 64 * 16;         // [8 K-chunks][64 rows n][8 x 16 bit] = 8,192
constexpr int Q_IMG = 2 * Q_IMG_HALF;           // Whi (fp16) then Wlo (bf16)
constexpr int QS_ACT = 0;
constexpr int QS_IMG = TC_MAX_STREAMS * 2 * Q_PLANE8;      // 163,840 (sized for the 8-chunk control variant)
constexpr int QS_MISC = QS_IMG + Q_IMG;                    // mbarrier + TMEM slot
constexpr int QS_TOTAL = QS_MISC + 64;

// instruction descriptor of kind::f16 with separate A / B formats (0 = F16, 1 = BF16), fp32 accumulate, K-major operands, M = 128
__device__ __forceinline__ uint32_t idesc_16(int afmt, int bfmt, int N) {
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}

// 4 floats -> 4 fp16 (hi) and 4 bf16 (lo = bf16(x - hi)), each packed into a uint2
__device__ __forceinline__ void split4(const float (&x)[4], uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(x[0], x[1]), h23 = __floats2half2_rn(x[2], x[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((x[0] - f01.x) * 2048.f, (x[1] - f01.y) * 2048.f), l23 = __floats2half2_rn((x[2] - f23.x) * 2048.f, (x[3] - f23.y) * 2048.f);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// forward operand images of every matrix: element (n = out unit j, k = in unit i) at [(k >> 3)][n][k & 7]; zero padded to K = 64, N = 64
__global__ void tc4_prep_kernel(const float* __restrict__ params, PeLayout lay, uint8_t* __restrict__ images) {
    const int m = blockIdx.x >> 4;
    const int din = lay.d[m], dout = lay.d[m + 1], ldw = lay.ldw[m];
    const float* W = params + lay.woff[m];
    __half* whi = reinterpret_cast<__half*>(images + (size_t)m * Q_IMG);
    __half* wlo = reinterpret_cast<__half*>(images + (size_t)m * Q_IMG + Q_IMG_HALF);
    const int e = (blockIdx.x & 15) * 256 + threadIdx.x;     // 4,096 elements: i = e >> 6 (0..63), j = e & 63
    const int i = e >> 6, j = e & 63;
    const float w = (i < din && j < dout) ? W[(size_t)i * ldw + j] : 0.f;
    const __half h = __float2half_rn(w);
    const int o = (i >> 3) * 512 + j * 8 + (i & 7);
    whi[o] = h;
    wlo[o] = __float2half_rn((w - __half2float(h)) * 2048.f);
}

struct Tc4Args {
    PeLayout lay;
    const float* points;
    const float* params;
    const uint8_t* images;
    float* out;                 // [n][NS][O]
    int n, ld;
    int lbo_trick;              // 1: 7-chunk planes, fourth K-step with LBO = 0 (what the full engine needs); 0: 8-chunk planes with a zero chunk
    float in_scale[3], in_shift[3];
};

template <int NS>
__global__ void __launch_bounds__(Q_THREADS, 1) fwd_tc4_kernel(const Tc4Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PeLayout& lay = a.lay;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = lay.L;
    const int Q_PLANE = a.lbo_trick ? Q_PLANE7 : Q_PLANE8;      // hi plane (fp16), then the lo plane (fp16, scaled by 2^11) of the same stream
    const int Q_STREAM = 2 * Q_PLANE;
    uint8_t* act = smem + QS_ACT;
    uint8_t* img = smem + QS_IMG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + QS_MISC + 16);
    const uint32_t act_s = smem_u32(act), img_s = smem_u32(img), bar = smem_u32(smem + QS_MISC);
    for (int i = tid; i < QS_TOTAL / 16; i += Q_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *tmem_slot;
    const int p = 32 * (warp & 3) + lane;            // TMEM lane = point of the tile
    const int h = warp >> 2;                         // unit half: 4-unit groups [7h, 7h+7)
    const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phase = 0;
    const int ntiles = (a.n + TC_P - 1) / TC_P;
    // store 4 units (group c4 = units 4 c4 .. 4 c4 + 3) of stream k for this thread's point: 8 B into the hi plane, 8 B into the lo plane
    auto store4 = [&](int k, int c4, const float (&v)[4]) {
        uint2 hi, lo;
        split4(v, hi, lo);
        uint8_t* dst = act + k * Q_STREAM + (c4 >> 1) * Q_CH + p * 16 + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(dst) = hi;
        *reinterpret_cast<uint2*>(dst + Q_PLANE) = lo;
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pt = tile * TC_P + p;
        const bool valid = pt < a.n;
        const float* row = a.points + (size_t)(valid ? pt : 0) * a.ld;
        float x = 0.f, y = 0.f, t = 0.f;
        if (valid) { x = row[0]; y = row[1]; t = row[2]; }
        const float c0 = fmaf(x, a.in_scale[0], a.in_shift[0]), c1 = fmaf(y, a.in_scale[1], a.in_shift[1]), c2 = fmaf(t, a.in_scale[2], a.in_shift[2]);
        // ---- layer 1 (3 -> d1): per-thread FFMA, all streams (plate:316 with X = concat(x, y, t))
        {
            const int dout = lay.d[1];
            const float* W0 = a.params + lay.woff[0];
            const float* b0 = a.params + lay.boff[0];
            const int ldw = lay.ldw[0];
            for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                float o[NS][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c4 + u;
                    float z[NS];
#pragma unroll
                    for (int k = 0; k < NS; ++k) z[k] = 0.f;
                    if (j < dout) {
                        const float w0 = __ldg(W0 + j), w1 = __ldg(W0 + ldw + j), w2 = __ldg(W0 + 2 * ldw + j);
                        z[0] = fmaf(c0, w0, fmaf(c1, w1, c2 * w2));
                        if (NS >= 4) { z[1] = a.in_scale[0] * w0; z[2] = a.in_scale[1] * w1; z[3] = a.in_scale[2] * w2; }
                        if (L > 1) act_fwd<NS, true>(z, __ldg(b0 + j));
                        else z[0] += __ldg(b0 + j);
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) o[k][u] = z[k];
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) store4(k, c4, o[k]);
            }
        }
        // ---- layers 2..L on the tensor core
        for (int l = 2; l <= L; ++l) {
            const int m = l - 1;
            const int din = lay.d[m], dout = lay.d[l];
            // operand image of matrix m -> smem (all threads; the previous layer's MMAs are complete: everybody waited for them)
            {
                const float4* src = reinterpret_cast<const float4*>(a.images + (size_t)m * Q_IMG);
                for (int i = tid; i < Q_IMG / 16; i += Q_THREADS) reinterpret_cast<float4*>(img)[i] = __ldg(src + i);
            }
            fence_async_smem();          // planes and image written through the generic proxy -> visible to the MMAs
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                const int ksteps = (din + 15) >> 4;
                const uint32_t ihh = idesc_16(0, 0, 64);
                for (int k = 0; k < NS; ++k) {
                    const uint32_t d = tbase + 64u * k;
                    const uint32_t hi_s = act_s + (uint32_t)(k * Q_STREAM), lo_s = hi_s + (uint32_t)Q_PLANE;
                    // K-step s covers unit chunks 2s and 2s + 1.  Chunk 7 does not exist in the 7-chunk planes: variant 1 re-reads chunk 2s
                    // (LBO = 0), variant 2 reads whatever follows the plane (finite 16-bit patterns); either way its weights (rows 56..63 of the
                    // image) are zero.  Order: all lo x hi / hi x lo products first (they carry a factor 2^11), then the first hi x hi K-step
                    // scales the accumulator by 2^-11 (scale-input-d), the other hi x hi K-steps accumulate plainly.
                    for (int s = 0; s < ksteps; ++s) {
                        const uint32_t lbo = (2 * s + 1 < 7 || a.lbo_trick != 1) ? Q_CH : 0u;
                        const uint64_t ahi = sdesc(hi_s + 2 * s * Q_CH, lbo, 128), alo = sdesc(lo_s + 2 * s * Q_CH, lbo, 128);
                        const uint64_t bhi = sdesc(img_s + 2 * s * 1024, 1024, 128), blo = sdesc(img_s + Q_IMG_HALF + 2 * s * 1024, 1024, 128);
                        mma_bf16_ss(d, ahi, blo, ihh, s > 0);
                        mma_bf16_ss(d, alo, bhi, ihh, 1u);
                    }
                    for (int s = 0; s < ksteps; ++s) {
                        const uint32_t lbo = (2 * s + 1 < 7 || a.lbo_trick != 1) ? Q_CH : 0u;
                        const uint64_t ahi = sdesc(hi_s + 2 * s * Q_CH, lbo, 128);
                        const uint64_t bhi = sdesc(img_s + 2 * s * 1024, 1024, 128);
                        if (s == 0) mma_f16_ss_scaled11(d, ahi, bhi, ihh);
                        else mma_bf16_ss(d, ahi, bhi, ihh, 1u);
                    }
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            fence_after();
            const float* bl = a.params + lay.boff[m];
            if (l < L) {
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float zk[NS][4];
#pragma unroll
                    for (int k = 0; k < NS; ++k) tm_ld4(tlane + 64 * k + 4 * c4, zk[k]);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        float z[NS];
#pragma unroll
                        for (int k = 0; k < NS; ++k) z[k] = zk[k][u];
                        if (j < dout) act_fwd<NS, true>(z, __ldg(bl + j));
                        else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) z[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < NS; ++k) zk[k][u] = z[k];
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) store4(k, c4, zk[k]);
                }
                fence_before();
            } else if (h == 0 && valid) {                 // output layer: Y[k][o] (+ bias on the value stream) -> d_out[pt][k][o]
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    float v[8];
                    tm_ld8(tlane + 64 * k, v);
                    tm_wait_ld();
                    for (int o = 0; o < dout && o < 8; ++o)
                        a.out[((size_t)pt * NS + k) * dout + o] = v[o] + (k == 0 ? __ldg(bl + o) : 0.f);
                }
            }
            fence_before();
            __syncthreads();             // all TMEM reads of this layer are done before the next layer's MMAs overwrite the accumulators
            fence_after();
        }
        if (L == 1 && h == 0 && valid) {                   // degenerate single-matrix net: the FFMA layer is the output
            for (int k = 0; k < NS; ++k)
                for (int o = 0; o < lay.d[1] && o < 8; ++o) {
                    const __half hv = reinterpret_cast<const __half*>(act + k * Q_STREAM + (o >> 3) * Q_CH + p * 16)[o & 7];
                    const __half lv = reinterpret_cast<const __half*>(act + k * Q_STREAM + Q_PLANE + (o >> 3) * Q_CH + p * 16)[o & 7];
                    a.out[((size_t)pt * NS + k) * lay.d[1] + o] = __half2float(hv) + __half2float(lv) * (1.f / 2048.f);
                }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int NS>
int launch_tc4(const Tc4Args& t, int ctas, cudaStream_t st) {
    auto kern = fwd_tc4_kernel<NS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QS_TOTAL);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(fwd_tc4, %d): %s", QS_TOTAL, cudaGetErrorString(e)); return 2; }
    kern<<<ctas, Q_THREADS, QS_TOTAL, st>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch fwd_tc4<%d>: %s", NS, cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // namespace

extern "C" size_t pe_debug_tc4_scratch_bytes(const pe_plan* plan) { return plan ? (size_t)plan->lay.L * Q_IMG : 0; }

// Forward jets d_out[n][K][O] (K = 4 or 5, hidden widths <= 56, O <= 8) on the fp16-hi + bf16-lo split; d_scratch: pe_debug_tc4_scratch_bytes.
// variant bit 0: 1 = 7-chunk planes + LBO = 0 on the last K-step, 0 = 8-chunk planes with a zero pad chunk.
extern "C" int pe_debug_forward_jets_tc4(const pe_plan* plan, int K, const float* d_points, int ld, int n, const float* in_scale, const float* in_shift,
                                         const float* d_params, void* d_scratch, float* d_out, int variant, void* stream) {
    if (!plan || plan->device < 0) { pe_set_error("tc4 probe: plan without a device"); return 1; }
    const PeLayout& lay = plan->lay;
    if (K != 4 && K != 5) { pe_set_error("tc4 probe: K = %d (4 or 5)", K); return 1; }
    if (lay.L < 2 || lay.maxw > 56 || lay.d[lay.L] > 8) { pe_set_error("tc4 probe: needs >= 2 matrices, hidden widths <= 56, <= 8 outputs"); return 1; }
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    Tc4Args t;
    t.lbo_trick = variant & 3;
    t.lay = lay; t.points = d_points; t.params = d_params; t.images = (const uint8_t*)d_scratch; t.out = d_out; t.n = n; t.ld = ld;
    for (int i = 0; i < 3; ++i) { t.in_scale[i] = in_scale ? in_scale[i] : 1.f; t.in_shift[i] = in_shift ? in_shift[i] : 0.f; }
    tc4_prep_kernel<<<lay.L * 16, 256, 0, st>>>(d_params, lay, (uint8_t*)d_scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tc4_prep_kernel: %s", cudaGetErrorString(e)); return 3; }
    const int tiles = (n + TC_P - 1) / TC_P;
    const int ctas = tiles < plan->sms ? tiles : plan->sms;
    return K == 5 ? launch_tc4<5>(t, ctas, st) : launch_tc4<4>(t, ctas, st);
}
