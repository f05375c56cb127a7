// Fourth-generation tcgen05 / TMEM engine for the collocation residual (engine id PE_ENGINE_TC4, name 'tc4'): ONE 16-bit operand split for
// all three GEMM families.  EXPERIMENTAL: written at the end of round 1 after the GPU budget was spent; it compiles for
// sm_100a but has NOT run on hardware yet.  It is opt-in only (never selected by 'auto'), and its GPU tests (tests/test_gpu_tc4.py) are
// skipped unless PE_TEST_TC4=1.  The forward arithmetic alone can be checked first with csrc/pe_tc4_probe.cu.
//
// Why (measured on the third generation, pe_tcs.cu: per 128-point tile the tensor pipe is busy 40 k of 215 k cycles while the epilogue warps
// carry 157 k cycles of SIMT work, 23 % of it the bf16 hi/mid conversion of the weight-gradient operands, ~10 % the TF32 residue operand):
//   * every operand X (activation, adjoint, weight) is kept as  Xhi = fp16(X),  Xlo = bf16(X - Xhi);  every GEMM is
//         A B ~= Ahi Bhi + Ahi Blo + Alo Bhi      three kind::f16 MMAs (K = 16) per K-step into one fp32 accumulator,
//     the A / B formats chosen per MMA in the instruction descriptor.  CPU model (tests/emulate_engine_precision.py): loss 1.2e-6,
//     gradient blocks 2-3e-6 end to end -- better than the TF32x3 + bf16x3 scheme of the earlier generations (3e-6 / 4e-6 emulated).
//   * 12 MMAs per stream-layer instead of 18; no TF32 residue operand in tensor memory;
//   * the activation planes written by the epilogue,  [hi: 7 chunks of 8 units][128 points][8 x fp16] | [lo: same, bf16],  are K-major
//     operands of the layer GEMMs AND, read MN-major, operands of the weight-gradient GEMM: they are stashed as they are (4 B per element,
//     as many bytes as the fp32 stash before) and come back by one bulk-TMA copy per stream.  No conversion pass, no converter hand-shake:
//     the issuer thread drives the whole weight-gradient phase.
//   * the weight-gradient tile (56 columns) and a bias-gradient tile (56 columns, fed by a "ones" chunk as A operand) have their own
//     tensor-memory columns (320 accumulator + 56 + 56 of 512): nothing is aliased.
// Roles as in pe_tcs.cu: 8 epilogue warps (thread (p, h): TMEM lane p = point, unit half h = 4-unit groups [7h, 7h + 7)), control warp 8
// (lane 0 issues every MMA / commit / TMA), warps 9..11 idle.  Forward sweep: the jet streams travel as three groups G0 = {value},
// G1 = {d/dx, d/dy}, G2 = {d/dt[, d2/dt2]} with ACT[g] / ACC[g] mbarrier pairs (see pe_tcs.cu).  Reverse sweep per layer l:
//   issuer : adjoint image landed, Zbar_l published -> adjoint MMAs (all streams) -> commit ACC -> per stream k: stashed planes of A_{l-1,k}
//            landed in staging buffer k & 1 -> 8 K-steps x (hh, hl, lh) into the weight-gradient tile -> commit EMPTY[k & 1] -> refill;
//            ones x (Zhi, Zlo) of the value stream into the bias tile -> commit DW
//   epilogue warps: wait ACC, wait DW -> drain both tiles into this CTA's gradient slot -> adjoint of tanh (accumulators + stashed planes)
//            -> Zbar_{l-1} planes -> publish.
// Reference lines: see pe_simt.cu / pe_device.cuh (the epilogue algebra is shared with every other engine).
#include <cuda_fp16.h>
#include <cstring>
#include "pe_device.cuh"
#include "pe_tc_common.cuh"

namespace {
using namespace pe_dev;
using namespace pe_tcc;

constexpr int F_THREADS = 384, F_EPI = 256;
constexpr int F_CH = 2048;                        // one chunk of 8 units: 128 points x 16 B
constexpr int F_PLANE = 7 * F_CH;                 // 14,336: 56 units
constexpr int F_STREAM = 2 * F_PLANE;             // 28,672: hi plane (fp16) then lo plane (bf16)
constexpr int F_IMG_HALF = 8 * 64 * 16;           // [8 K-chunks][64 rows][8 x 16 bit] = 8,192
constexpr int F_IMG = 2 * F_IMG_HALF;             // Whi (fp16) then Wlo (bf16): 16,384
constexpr int F_IMG_LAYER = 2 * F_IMG;            // forward image, adjoint image
// shared-memory map
constexpr int F_ONES = 0;                                     // [128 points][8 fp16], unit 0 = 1: A operand of the bias-gradient MMAs
constexpr int F_ACT = F_ONES + F_CH;                          // 2,048: NS x (hi plane | lo plane)
constexpr int F_R = F_ACT + TC_MAX_STREAMS * F_STREAM;        // 145,408: forward: two weight images; reverse: adjoint image + two staging buffers
constexpr int F_STG = F_R + F_IMG;                            // staging buffer s at F_STG + s * F_STREAM
constexpr int F_MISC = F_R + F_IMG + 2 * F_STREAM;            // 219,136: mbarriers + TMEM base slot
constexpr int F_COORD = F_MISC + 256;
constexpr int F_RED = F_COORD + 128 * 16;
constexpr int F_BIAS = F_RED + 4096;                          // [PE_MAX_LAYERS][64] floats
constexpr int F_W0 = F_BIAS + PE_MAX_LAYERS * 256;            // [4][64] floats
constexpr int F_TOTAL = F_W0 + 1024;                          // 230,656
static_assert(2 * F_IMG <= F_IMG + 2 * F_STREAM, "the forward image double buffer lives inside the reverse-sweep region");
static_assert(F_TOTAL <= 227 * 1024, "shared memory map exceeds the 227 KB opt-in limit");
constexpr int B_ACC = 0, B_ACT = 24, B_IMG = 48, B_SFULL = 64, B_SEMPTY = 80, B_DW = 96, B_TMEM = 112;   // byte offsets in F_MISC
constexpr uint32_t T_ACC = 0, T_DW = 320, T_BIAS = 384;       // tensor-memory columns

template <int NS> __device__ __forceinline__ int grp_first(int g) { return g == 0 ? 0 : (g == 1 ? 1 : 3); }
template <int NS> __device__ __forceinline__ int grp_count(int g) { return g == 0 ? 1 : (g == 1 ? 2 : NS - 3); }

// kind::f16 instruction descriptors with separate A / B formats (0 = F16, 1 = BF16), fp32 accumulate
__device__ __forceinline__ uint32_t idesc_km(int afmt, int bfmt, int N) {           // K-major operands, M = 128
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}
__device__ __forceinline__ uint32_t idesc_mn(int afmt, int bfmt, int N) {           // MN-major operands, M = 64 (weight / bias gradient)
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (4u << 24);
}

// 4 floats -> 4 fp16 (hi) and 4 bf16 (lo = bf16(x - hi)), each packed into a uint2
__device__ __forceinline__ void split4(const float (&x)[4], uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(x[0], x[1]), h23 = __floats2half2_rn(x[2], x[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(x[0] - f01.x, x[1] - f01.y), l23 = __floats2bfloat162_rn(x[2] - f23.x, x[3] - f23.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}
__device__ __forceinline__ void join4(const uint2& hi, const uint2& lo, float (&x)[4]) {
    const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
    const float2 l01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.x)), l23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.y));
    x[0] = h01.x + l01.x; x[1] = h01.y + l01.y; x[2] = h23.x + l23.x; x[3] = h23.y + l23.y;
}

// Operand images per matrix m: forward B operand [n = out unit j][k = in unit i] at [(i >> 3)][j][i & 7] and adjoint B operand
// [n = i][k = j] at [(j >> 3)][i][j & 7], each as fp16 hi and bf16 lo, zero padded to 64 x 64.  16 blocks of 256 elements per matrix.
__global__ void tc4_image_kernel(const float* __restrict__ params, PeLayout lay, uint8_t* __restrict__ images) {
    const int m = blockIdx.x >> 4;
    const int din = lay.d[m], dout = lay.d[m + 1], ldw = lay.ldw[m];
    const float* W = params + lay.woff[m];
    uint8_t* img = images + (size_t)m * F_IMG_LAYER;
    __half* fhi = reinterpret_cast<__half*>(img);
    __nv_bfloat16* flo = reinterpret_cast<__nv_bfloat16*>(img + F_IMG_HALF);
    __half* ahi = reinterpret_cast<__half*>(img + F_IMG);
    __nv_bfloat16* alo = reinterpret_cast<__nv_bfloat16*>(img + F_IMG + F_IMG_HALF);
    const int e = (blockIdx.x & 15) * 256 + threadIdx.x;
    const int i = e >> 6, j = e & 63;
    const float w = (i < din && j < dout) ? W[(size_t)i * ldw + j] : 0.f;
    const __half h = __float2half_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __half2float(h));
    const int of = (i >> 3) * 512 + j * 8 + (i & 7), oa = (j >> 3) * 512 + i * 8 + (j & 7);
    fhi[of] = h; flo[of] = l;
    ahi[oa] = h; alo[oa] = l;
}

struct Tc4Args {
    PeResidArgs r;
    const uint8_t* images;
    pe_term_desc term2;
    const float* points2;
    const float* aux2;
    int n2;
    float inv_n2;
};

// MMAs of one layer GEMM for streams [k0, k0 + nk): three products per K-step of 16, K-steps interleaved across the streams of the call
// (consecutive MMAs never accumulate into the same tile).  When the second chunk of a K-step would be chunk 7 (which does not exist) the
// K-step re-reads its first chunk (LBO = 0): the matching weight rows 56..63 of the image are zero.
__device__ __forceinline__ void issue_streams(int k0, int nk, uint32_t tbase, uint32_t act_s, uint32_t img_s, int N, int kdim) {
    const int ksteps = (kdim + 15) >> 4;
    const uint32_t ihh = idesc_km(0, 0, N), ihl = idesc_km(0, 1, N), ilh = idesc_km(1, 0, N);
#pragma unroll 1
    for (int s = 0; s < ksteps; ++s) {
        const uint32_t lbo = (2 * s + 1 < 7) ? (uint32_t)F_CH : 0u;
        const uint64_t bhi = sdesc(img_s + 2 * s * 1024, 1024, 128), blo = sdesc(img_s + F_IMG_HALF + 2 * s * 1024, 1024, 128);
#pragma unroll
        for (int k = 0; k < TC_MAX_STREAMS; ++k)
            if (k < nk) {
                const uint32_t hi_s = act_s + (uint32_t)((k0 + k) * F_STREAM + 2 * s * F_CH);
                mma_bf16_ss(tbase + T_ACC + 64u * (k0 + k), sdesc(hi_s, lbo, 128), bhi, ihh, s > 0);
            }
#pragma unroll
        for (int k = 0; k < TC_MAX_STREAMS; ++k)
            if (k < nk) {
                const uint32_t hi_s = act_s + (uint32_t)((k0 + k) * F_STREAM + 2 * s * F_CH);
                mma_bf16_ss(tbase + T_ACC + 64u * (k0 + k), sdesc(hi_s, lbo, 128), blo, ihl, 1u);
            }
#pragma unroll
        for (int k = 0; k < TC_MAX_STREAMS; ++k)
            if (k < nk) {
                const uint32_t lo_s = act_s + (uint32_t)((k0 + k) * F_STREAM + F_PLANE + 2 * s * F_CH);
                mma_bf16_ss(tbase + T_ACC + 64u * (k0 + k), sdesc(lo_s, lbo, 128), bhi, ilh, 1u);
            }
    }
}

template <int NS>
__global__ void __launch_bounds__(F_THREADS, 1) resid_tc4_kernel(const Tc4Args args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STASH_LAYER = NS * F_STREAM;                 // bytes per stashed layer: NS x (hi plane | lo plane)
    const PeResidArgs& A = args.r;
    const PeLayout& lay = A.lay;
    const pe_term_desc& T = A.term;
    const pe_term_desc& T2 = args.term2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = lay.L;
    uint8_t* act = smem + F_ACT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + F_MISC + B_TMEM);
    float* coord = reinterpret_cast<float*>(smem + F_COORD);
    float* red = reinterpret_cast<float*>(smem + F_RED);
    float* sbias = reinterpret_cast<float*>(smem + F_BIAS);
    float* sw0 = reinterpret_cast<float*>(smem + F_W0);
    const uint32_t act_s = smem_u32(act), r_s = smem_u32(smem + F_R), stg_s = smem_u32(smem + F_STG), ones_s = smem_u32(smem + F_ONES);
    const uint32_t bar0 = smem_u32(smem + F_MISC);
    const uint32_t bar_acc = bar0 + B_ACC, bar_act = bar0 + B_ACT, bar_img = bar0 + B_IMG;
    const uint32_t bar_sfull = bar0 + B_SFULL, bar_sempty = bar0 + B_SEMPTY, bar_dw = bar0 + B_DW;

    for (int i = tid; i < F_TOTAL / 16; i += F_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int slot = A.slot_base + blockIdx.x;
    float* gpart = A.grad_partials + (size_t)slot * lay.total;
    uint8_t* stash = reinterpret_cast<uint8_t*>(A.stash + (size_t)blockIdx.x * A.stash_floats);
    const float* __restrict__ params = A.params;
    for (int i = tid; i < lay.total; i += F_THREADS) __stcg(gpart + i, 0.f);
    __syncthreads();
    if (tid < 128) reinterpret_cast<__half*>(smem + F_ONES)[tid * 8] = __float2half_rn(1.f);      // ones chunk: unit 0 of every point
    if (tid < 64) {                                  // first-layer weights (3 x d1) and bias -> smem, once per launch
        const bool in = tid < lay.d[1];
        sw0[tid] = in ? __ldg(params + lay.woff[0] + tid) : 0.f;
        sw0[64 + tid] = in ? __ldg(params + lay.woff[0] + lay.ldw[0] + tid) : 0.f;
        sw0[128 + tid] = in ? __ldg(params + lay.woff[0] + 2 * lay.ldw[0] + tid) : 0.f;
        sw0[192 + tid] = in ? __ldg(params + lay.boff[0] + tid) : 0.f;
    }
    for (int i = tid; i < L * 64; i += F_THREADS) {  // all biases
        const int m = i >> 6, j = i & 63;
        sbias[i] = (j < lay.d[m + 1]) ? __ldg(params + lay.boff[m] + j) : 0.f;
    }
    if (tid == 0) {
        for (int g = 0; g < 3; ++g) { mbar_init(bar_acc + 8 * g, 1); mbar_init(bar_act + 8 * g, F_EPI); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_img + 8 * b, 1); mbar_init(bar_sfull + 8 * b, 1); mbar_init(bar_sempty + 8 * b, 1); }
        mbar_init(bar_dw, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();                              // zero-filled regions and the ones chunk before any async-proxy access
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *tmem_slot;
    const int ntiles_main = (A.n + TC_P - 1) / TC_P;
    const int ntiles = ntiles_main + (args.n2 + TC_P - 1) / TC_P;
    // forward image of matrix i-1 (i = 2..L) lives in buffer i & 1 of the region F_R
    auto fwd_src = [&](int i) { return args.images + (size_t)(i - 1) * F_IMG_LAYER; };
    auto adj_src = [&](int l) { return args.images + (size_t)(l - 1) * F_IMG_LAYER + F_IMG; };      // adjoint image of matrix l-1 (layer l)

    if (warp >= 8) {
        // ============================================================================================ control warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        if (warp == 8 && lane == 0) {
            uint32_t pact = 0, pimg = 0, psfull = 0, psempty = 0;     // parity bits of the phases this thread waits for next
            uint32_t n_acc2 = 0, n_dw = 0;
            auto load_fwd = [&](int i) {
                const uint32_t b = (uint32_t)(i & 1);
                mbar_expect_tx(bar_img + 8 * b, F_IMG);
                tma_load_1d(r_s + b * F_IMG, fwd_src(i), F_IMG, bar_img + 8 * b);
            };
            auto wait_img = [&](uint32_t b) { mbar_wait(bar_img + 8 * b, (pimg >> b) & 1u); pimg ^= 1u << b; };
            auto wait_act = [&](int g) { mbar_wait(bar_act + 8 * g, (pact >> g) & 1u); pact ^= 1u << g; };
            load_fwd(2);
            if (L >= 3) load_fwd(3);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // ---------------------------------------------------------------- forward: layers 2..L, group by group
                for (int l = 2; l <= L; ++l) {
                    const uint32_t b = (uint32_t)(l & 1);
                    const int NF = (lay.d[l] <= 16) ? 16 : 64;
                    wait_img(b);
#pragma unroll 1
                    for (int g = 0; g < 3; ++g) {
                        wait_act(g);
                        fence_after();
                        issue_streams(grp_first<NS>(g), grp_count<NS>(g), tbase, act_s, r_s + b * F_IMG, NF, lay.d[l - 1]);
                        mma_commit(bar_acc + 8 * g);
                        if (g == 2) ++n_acc2;
                        if (g == 0 && l >= 3 && l + 1 <= L) {
                            // every MMA of layer l-1 precedes G0 of layer l in the pipe: once its last group is complete its image buffer
                            // (= the buffer of image l + 1) is free
                            mbar_wait(bar_acc + 16, (n_acc2 - 1) & 1u);
                            load_fwd(l + 1);
                        }
                    }
                }
                // ---------------------------------------------------------------- reverse sweep: layers L..2
                // the region F_R changes hands: wait until the output-layer MMAs (the last readers of a forward image) are complete
                mbar_wait(bar_acc + 16, (n_acc2 - 1) & 1u);
                mbar_expect_tx(bar_img, F_IMG);
                tma_load_1d(r_s, adj_src(L), F_IMG, bar_img);
                for (int l = L; l >= 2; --l) {
                    const int dout = lay.d[l];
                    const uint8_t* stash_in = stash + (size_t)(l - 2) * STASH_LAYER;     // outputs of layer l-1 = inputs A of layer l
                    auto load_stage = [&](int k) {
                        const uint32_t sb = (uint32_t)(k & 1);
                        mbar_expect_tx(bar_sfull + 8 * sb, F_STREAM);
                        tma_load_1d(stg_s + sb * F_STREAM, stash_in + (size_t)k * F_STREAM, F_STREAM, bar_sfull + 8 * sb);
                    };
                    load_stage(0);                          // both staging buffers are free: the previous layer's DW phase is complete
                    load_stage(1);
                    wait_img(0);
                    wait_act(0); wait_act(1); wait_act(2);
                    fence_after();
                    issue_streams(0, NS, tbase, act_s, r_s, 64, dout);
                    mma_commit(bar_acc);
                    mma_commit(bar_acc + 8);
                    mma_commit(bar_acc + 16);
                    ++n_acc2;
                    // ---- weight gradient  dW = sum_k A_k^T Zbar_k  (K = 128 points, 8 K-steps of 16) and bias gradient  ones^T Zbar_0
                    const int NZ = (dout + 7) & ~7;
                    const uint32_t ghh = idesc_mn(0, 0, NZ), ghl = idesc_mn(0, 1, NZ), glh = idesc_mn(1, 0, NZ);
#pragma unroll 1
                    for (int k = 0; k < NS; ++k) {
                        const uint32_t sb = (uint32_t)(k & 1);
                        mbar_wait(bar_sfull + 8 * sb, (psfull >> sb) & 1u);
                        psfull ^= 1u << sb;
                        const uint32_t a_hi = stg_s + sb * F_STREAM, a_lo = a_hi + F_PLANE;
                        const uint32_t z_hi = act_s + (uint32_t)(k * F_STREAM), z_lo = z_hi + F_PLANE;
#pragma unroll 1
                        for (int s = 0; s < 8; ++s) {
                            const uint32_t o = (uint32_t)(s * 256);          // 16 points x 16 B
                            const uint64_t dah = sdesc(a_hi + o, 128, F_CH), dal = sdesc(a_lo + o, 128, F_CH);
                            const uint64_t dzh = sdesc(z_hi + o, 128, F_CH), dzl = sdesc(z_lo + o, 128, F_CH);
                            mma_bf16_ss(tbase + T_DW, dah, dzh, ghh, (k > 0 || s > 0) ? 1u : 0u);
                            mma_bf16_ss(tbase + T_DW, dah, dzl, ghl, 1u);
                            mma_bf16_ss(tbase + T_DW, dal, dzh, glh, 1u);
                        }
                        if (k + 2 < NS) mma_commit(bar_sempty + 8 * sb);       // this buffer is refilled (stream k + 2) once its MMAs are complete
                        if (k >= 1 && k + 1 < NS) {                          // ... which is waited for one stream later, behind the next stream's MMAs
                            const uint32_t ob = sb ^ 1u;
                            mbar_wait(bar_sempty + 8 * ob, (psempty >> ob) & 1u);
                            psempty ^= 1u << ob;
                            load_stage(k + 1);
                        }
                    }
                    {   // bias gradient: unit chunk 0 of the A operand is the ones chunk (row 0 of the tile = column sums of Zbar_0); its chunks
                        // 1..7 fall on the first value-stream chunks that follow it in shared memory (rows 8..63 of the tile: ignored)
                        const uint32_t z_hi = act_s, z_lo = act_s + F_PLANE;
#pragma unroll 1
                        for (int s = 0; s < 8; ++s) {
                            const uint32_t o = (uint32_t)(s * 256);
                            const uint64_t d1 = sdesc(ones_s + o, 128, F_CH);
                            mma_bf16_ss(tbase + T_BIAS, d1, sdesc(z_hi + o, 128, F_CH), ghh, s > 0 ? 1u : 0u);
                            mma_bf16_ss(tbase + T_BIAS, d1, sdesc(z_lo + o, 128, F_CH), ghl, 1u);
                        }
                    }
                    mma_commit(bar_dw);
                    ++n_dw;
                    mbar_wait(bar_dw, (n_dw - 1) & 1u);      // adjoint image, staging buffers and the Zbar planes are free again
                    if (l > 2) {
                        mbar_expect_tx(bar_img, F_IMG);
                        tma_load_1d(r_s, adj_src(l - 1), F_IMG, bar_img);
                    } else if (tile + (int)gridDim.x < ntiles) {
                        load_fwd(2);
                        if (L >= 3) load_fwd(3);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ============================================================================================ epilogue warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int p = 32 * (warp & 3) + lane;            // TMEM lane = point of the tile
        const int h = warp >> 2;                         // unit half: 4-unit groups [7h, 7h+7)
        const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t pacc = 0, pdw = 0;
        auto wait_acc = [&](int g) { mbar_wait(bar_acc + 8 * g, (pacc >> g) & 1u); pacc ^= 1u << g; };
        auto publish = [&](int g) { mbar_arrive(bar_act + 8 * g); };
        // planes in shared memory and in the stash (global) were written through the generic proxy; their next readers are MMAs and bulk
        // copies (async proxy)
        auto publish_fences = [&]() { asm volatile("fence.proxy.async;" ::: "memory"); fence_before(); };
        float tsum[PE_MAX_TERMS], tsum2[PE_MAX_TERMS];
#pragma unroll
        for (int i = 0; i < PE_MAX_TERMS; ++i) { tsum[i] = 0.f; tsum2[i] = 0.f; }
        // 4 units (group c4) of stream k of this thread's point -> hi / lo planes in shared memory [and in the stash layer `st`]
        auto put4 = [&](uint8_t* st, int k, int c4, const float (&v)[4]) {
            uint2 hi, lo;
            split4(v, hi, lo);
            const int o = k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
            *reinterpret_cast<uint2*>(act + o) = hi;
            *reinterpret_cast<uint2*>(act + o + F_PLANE) = lo;
            if (st) {
                __stcg(reinterpret_cast<uint2*>(st + o), hi);
                __stcg(reinterpret_cast<uint2*>(st + o + F_PLANE), lo);
            }
        };
        auto get4 = [&](int k, int c4, float (&v)[4]) {          // this thread's own entries of the planes in shared memory
            const int o = k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
            join4(*reinterpret_cast<const uint2*>(act + o), *reinterpret_cast<const uint2*>(act + o + F_PLANE), v);
        };

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
            named_bar_sync(1, F_EPI);
            // ================================================================ layer 1 (3 -> d1): per-thread FFMA, all streams
            {
                const float4 c4v = *reinterpret_cast<const float4*>(coord + 4 * p);
                const int dout = lay.d[1];
#pragma unroll 1
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float o[NS][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        float z[NS];
#pragma unroll
                        for (int k = 0; k < NS; ++k) z[k] = 0.f;
                        if (j < dout) {
                            const float w0 = sw0[j], w1 = sw0[64 + j], w2 = sw0[128 + j];
                            z[0] = fmaf(c4v.x, w0, fmaf(c4v.y, w1, c4v.z * w2));
                            z[1] = Tc.in_scale[0] * w0; z[2] = Tc.in_scale[1] * w1; z[3] = Tc.in_scale[2] * w2;
                            act_fwd<NS, true>(z, sw0[192 + j]);
                        }
#pragma unroll
                        for (int k = 0; k < NS; ++k) o[k][u] = z[k];
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) put4(stash, k, c4, o[k]);                 // stash layer 0 = outputs of layer 1
                }
                publish_fences();
                publish(0); publish(1); publish(2);
            }
            // ================================================================ forward: hidden layers 2..L-1, one stream group at a time
            for (int l = 2; l < L; ++l) {
                const int dout = lay.d[l];
                const float* bl = sbias + (l - 1) * 64;
                uint8_t* st = stash + (size_t)(l - 1) * STASH_LAYER;
                // ---- G0: a = tanh(z_0 + b)
                wait_acc(0);
                fence_after();
#pragma unroll 1
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float z[4];
                    tm_ld4(tlane + T_ACC + 4 * c4, z);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        z[u] = (j < dout) ? tanh_branchfree(z[u] + bl[j]) : 0.f;
                    }
                    put4(st, 0, c4, z);
                }
                publish_fences();
                publish(0);
                // ---- G1: a_x = s z_x, a_y = s z_y   (a re-read from this thread's own entries of the value planes: hi + lo, 19 bits --
                //      the CPU model shows no change of the end-to-end error against an exact fp32 a)
                wait_acc(1);
                fence_after();
#pragma unroll 1
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float av[4], z1[4], z2[4];
                    get4(0, c4, av);
                    tm_ld4(tlane + T_ACC + 64 + 4 * c4, z1);
                    tm_ld4(tlane + T_ACC + 128 + 4 * c4, z2);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        const float s = fmaf(-av[u], av[u], 1.f);
                        z1[u] = (j < dout) ? s * z1[u] : 0.f;
                        z2[u] = (j < dout) ? s * z2[u] : 0.f;
                    }
                    put4(st, 1, c4, z1);
                    put4(st, 2, c4, z2);
                }
                publish_fences();
                publish(1);
                // ---- G2: a_t = s z_t [, a_tt = s z_tt - 2 a a_t z_t]
                wait_acc(2);
                fence_after();
#pragma unroll 1
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float av[4], z3[4], z4[4];
                    get4(0, c4, av);
                    tm_ld4(tlane + T_ACC + 192 + 4 * c4, z3);
                    if (NS == 5) tm_ld4(tlane + T_ACC + 256 + 4 * c4, z4);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        const float a = av[u];
                        const float s = fmaf(-a, a, 1.f);
                        const float zt = z3[u];
                        const float at = s * zt;
                        z3[u] = (j < dout) ? at : 0.f;
                        if (NS == 5) z4[u] = (j < dout) ? fmaf(s, z4[u], -2.f * a * at * zt) : 0.f;
                    }
                    put4(st, 3, c4, z3);
                    if (NS == 5) put4(st, 4, c4, z4);
                }
                publish_fences();
                publish(2);
            }
            // ================================================================ output layer L: residuals, loss partials, seeds
            {
                const int dout = lay.d[L];
                const float* bl = sbias + (L - 1) * 64;
                wait_acc(0); wait_acc(1); wait_acc(2);
                fence_after();
                if (h == 0) {
                    float Y[NS][PE_UJ];
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        float v[8];
                        tm_ld8(tlane + T_ACC + 64 * k, v);
                        tm_wait_ld();
#pragma unroll
                        for (int u = 0; u < PE_UJ; ++u) Y[k][u] = (u < 8 && u < dout) ? v[u < 8 ? u : 0] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) if (u < dout) Y[0][u] += bl[u];
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
                    // seeds Zbar_L: units 0..7 in groups 0 and 1; groups 2 and 3 (the second chunk of the only adjoint K-step) are zeroed
                    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        const float v0[4] = {Y[k][0], Y[k][1], Y[k][2], Y[k][3]}, v1[4] = {Y[k][4], Y[k][5], Y[k][6], Y[k][7]};
                        put4(nullptr, k, 0, v0);
                        put4(nullptr, k, 1, v1);
                        put4(nullptr, k, 2, zero4);
                        put4(nullptr, k, 3, zero4);
                    }
                }
                publish_fences();
                publish(0); publish(1); publish(2);
            }
            // ================================================================ reverse sweep, layers L .. 2
            for (int l = L; l >= 2; --l) {
                const int m = l - 1;
                const int din = lay.d[l - 1], dout = lay.d[l];
                const uint8_t* stash_in = stash + (size_t)(l - 2) * STASH_LAYER;          // outputs of layer l-1 = inputs A of layer l
                const int NZ = (dout + 7) & ~7;
                wait_acc(0); wait_acc(1); wait_acc(2);       // adjoint MMAs done: abar^{l-1} in the accumulators
                mbar_wait(bar_dw, pdw);                      // weight / bias gradient MMAs done: the Zbar_l planes may be overwritten
                pdw ^= 1u;
                fence_after();
                {   // drain: weight-gradient rows i = 16*quadrant + lane (lane < 16); bias gradient = row 0 of its tile; h selects the column half
                    const int quad = warp & 3;
                    const int i = 16 * quad + lane;
                    const int ldw = lay.ldw[m];
                    float* gW = gpart + lay.woff[m];
                    float* gB = gpart + lay.boff[m];
                    const int c_lo = h ? 32 : 0, c_hi = h ? 56 : 32;
                    for (int c = c_lo; c < c_hi; c += 8) {
                        float v[8], vb[8];
                        tm_ld8(tlane + T_DW + c, v);
                        tm_ld8(tlane + T_BIAS + c, vb);
                        tm_wait_ld();
                        if (c < NZ) {
                            if (lane < 16 && i < din) {
                                float* dst = gW + (size_t)i * ldw + c;
                                if (c < ldw) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                                if (c + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4], v[5], v[6], v[7]));
                            }
                            if (quad == 0 && lane == 0) {
                                if (c < ldw) atomicAdd(reinterpret_cast<float4*>(gB + c), make_float4(vb[0], vb[1], vb[2], vb[3]));
                                if (c + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(gB + c + 4), make_float4(vb[4], vb[5], vb[6], vb[7]));
                            }
                        }
                    }
                }
                // ---- through tanh of layer l-1: zbar^{l-1} from abar^{l-1} (TMEM) and the stashed outputs (hi + lo)
                auto ldst = [&](int k, int c4, uint2& hi, uint2& lo) {
                    const uint8_t* src = stash_in + (size_t)k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
                    hi = __ldcg(reinterpret_cast<const uint2*>(src));
                    lo = __ldcg(reinterpret_cast<const uint2*>(src + F_PLANE));
                };
                uint2 nh[NS], nl[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) ldst(k, 7 * h, nh[k], nl[k]);
#pragma unroll 1
                for (int c4 = 7 * h; c4 < 7 * h + 7; ++c4) {
                    float ab[NS][4], Av[NS][4];
#pragma unroll
                    for (int k = 0; k < NS; ++k) join4(nh[k], nl[k], Av[k]);
                    if (c4 + 1 < 7 * h + 7) {
#pragma unroll
                        for (int k = 0; k < NS; ++k) ldst(k, c4 + 1, nh[k], nl[k]);
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) tm_ld4(tlane + T_ACC + 64 * k + 4 * c4, ab[k]);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c4 + u;
                        float b[NS], Aa[NS];
#pragma unroll
                        for (int k = 0; k < NS; ++k) { b[k] = ab[k][u]; Aa[k] = Av[k][u]; }
                        if (j < din) act_bwd<NS>(b, Aa);
                        else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) b[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < NS; ++k) ab[k][u] = b[k];
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) put4(nullptr, k, c4, ab[k]);
                }
                if (l > 2) {
                    publish_fences();
                    publish(0); publish(1); publish(2);
                }
            }
            // ================================================================ layer 1 gradient (3 x d1 + bias): FFMA, fixed-order reduce
            named_bar_sync(1, F_EPI);
            {
                const int d1 = lay.d[1];
                const int j = tid & 63, qq = tid >> 6;                    // 4 point quarters x 64 units
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, gb = 0.f;
                if (j < d1) {
                    const uint8_t* base = act + (j >> 3) * F_CH + (j & 7) * 2;
#pragma unroll 4
                    for (int s = 0; s < 32; ++s) {
                        const int pp = 32 * qq + s;
                        const float4 c4v = *reinterpret_cast<const float4*>(coord + 4 * pp);
                        auto val = [&](int k) {
                            const uint8_t* q = base + k * F_STREAM + pp * 16;
                            return __half2float(*reinterpret_cast<const __half*>(q)) + __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(q + F_PLANE));
                        };
                        const float zv = val(0), zx = val(1), zy = val(2), zt = val(3);
                        g0 = fmaf(c4v.x, zv, fmaf(Tc.in_scale[0], zx, g0));
                        g1 = fmaf(c4v.y, zv, fmaf(Tc.in_scale[1], zy, g1));
                        g2 = fmaf(c4v.z, zv, fmaf(Tc.in_scale[2], zt, g2));
                        gb += zv;
                    }
                }
                *reinterpret_cast<float4*>(red + (qq * 64 + j) * 4) = make_float4(g0, g1, g2, gb);
                named_bar_sync(1, F_EPI);
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
                named_bar_sync(1, F_EPI);
            }
        }
        // ---- loss-term partial sums (threads with h == 0 hold them): warp reduce, then 4 warps through smem (fixed order)
        {
            float tot[2 + PE_MAX_TERMS];
            tot[0] = warp_sum(tsum[0]);
            tot[1] = warp_sum(tsum[1]);
#pragma unroll
            for (int c = 0; c < PE_MAX_TERMS; ++c) tot[2 + c] = warp_sum(tsum2[c]);
            named_bar_sync(1, F_EPI);
            if (h == 0 && lane == 0) {
#pragma unroll
                for (int c = 0; c < 2 + PE_MAX_TERMS; ++c) red[(2 + PE_MAX_TERMS) * warp + c] = tot[c];
            }
            named_bar_sync(1, F_EPI);
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
    }
    fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int NS>
int launch_tc4(const Tc4Args& t, int slots, cudaStream_t st) {
    auto kern = resid_tc4_kernel<NS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F_TOTAL);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_tc4, %d): %s", F_TOTAL, cudaGetErrorString(e)); return 2; }
    kern<<<slots, F_THREADS, F_TOTAL, st>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_tc4<%d>: %s", NS, cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // namespace

size_t pe_tc_stash_floats_per_slot(const pe_plan* plan);

// Scratch layout (d_stash of the C ABI): [slots][stash floats per slot] then the operand images.  The per-slot stash of the earlier
// tensor-core generations ((L-1) layers x 5 streams x 14 chunks x 2,048 B) is exactly (L-1) x 5 x F_STREAM bytes: reused as it is.
int pe_launch_resid_tc4(const pe_plan* plan, const PeResidArgs& a, int K, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2) {
    Tc4Args t;
    t.r = a;
    t.n2 = 0; t.points2 = nullptr; t.aux2 = nullptr; t.inv_n2 = 0.f;
    memset(&t.term2, 0, sizeof(t.term2));
    if (term2 && n2 > 0) {
        t.term2 = *term2; t.points2 = points2; t.n2 = n2; t.aux2 = term2->aux_k ? aux2 : nullptr;
        t.inv_n2 = 1.0f / (float)term2->n_global;
    }
    if (plan->lay.L < 3) { pe_set_error("tc4 engine: needs at least two hidden layers"); return 1; }
    t.r.stash_floats = (int)pe_tc_stash_floats_per_slot(plan);
    uint8_t* images = reinterpret_cast<uint8_t*>(a.stash + (size_t)slots * t.r.stash_floats);
    t.images = images;
    tc4_image_kernel<<<plan->lay.L * 16, 256, 0, st>>>(a.params, a.lay, images);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tc4_image_kernel: %s", cudaGetErrorString(e)); return 3; }
    if (K == 5) return launch_tc4<5>(t, slots, st);
    if (K == 4) return launch_tc4<4>(t, slots, st);
    pe_set_error("tc4 engine: K = %d not instantiated (4 or 5)", K);
    return 1;
}
