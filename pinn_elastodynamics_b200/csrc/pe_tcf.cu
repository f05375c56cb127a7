// fp16-pair tcgen05 / TMEM engine for the collocation residual (engine id PE_ENGINE_TCF, name 'tcf'; what 'auto' selects).
//
// Arithmetic.  Every GEMM operand X (activation jets, adjoints, weights) is kept as an fp16 PAIR
//         Xh = fp16(X),   Xl = fp16((X - Xh) * 2^11)            X = Xh + 2^-11 Xl  to ~22 bits
// and every product is   A B ~= Ah Bh + 2^-11 (Ah Bl + Al Bh)   on kind::f16 MMAs (K = 16) into ONE fp32 accumulator:
// the two cross products of every K-step are accumulated first, then the first Ah Bh MMA is issued in the scale-input-d form
// (D = A B + D * 2^-11, probed: tests/probe_umma.py 'f16 scale-input-d'), the other Ah Bh K-steps accumulate plainly.  12 MMAs per
// stream and layer (the TF32x3 scheme of the earlier generations: 18) at fp32-grade accuracy (CPU study
// profiles/r1_split_precision_study.txt column fp16x3s: 2.6e-7 per GEMM for operand magnitudes in [1e-4, 6e4]; fp32: 1.5e-7).
// Mixed fp16 x bf16 operands in one kind::f16 MMA -- the round-1 plan -- are an illegal instruction on B200 (profiles/r2_umma_probe.txt).
// Range.  fp16 covers 6e-5 .. 65504 at full precision.  Forward jets of sane networks live there.  The adjoint seeds carry 2 w / N
// and do not: each 128-point tile scales its seeds by a power of two sigma so that their largest magnitude is in [1, 2); the weight /
// bias gradient tiles are multiplied by 1 / sigma when they are drained.  Nothing else depends on sigma (the reverse sweep is linear).
// An overflow (|x| > 65504) turns into inf / NaN in the loss or the gradient, never into a silently wrong finite number.
//
// Layout.  Activation planes, per stream  [hi: 7 chunks of 8 units][128 points][8 x fp16] | [lo: same]  = 28,672 B: K-major operand of
// the layer GEMMs (LBO 2,048, SBO 128; the fourth K-step's second chunk is whatever follows the plane -- finite 16-bit patterns --
// against weight rows 56..63 that are zero) and, read MN-major, the [Zh | Zl] operand (N = 112) of the weight-gradient GEMM.  The forward
// epilogue stashes the planes as they are (4 B per element) and the second issuer brings them back with one bulk-TMA copy per stream: no
// conversion pass.  Tensor memory (columns): 5 x 64 accumulators | weight-gradient tile 112 (rows 0..55: Ah^T [Zh | Zl], rows 56..111:
// Al^T Zh) | bias tile 8 | layer-1 gradient tiles 4 x 8 (value stream x coordinate block, column sums of d/dx, d/dy, d/dt).
//
// Roles (warp-specialised): F_EW epilogue warps (thread (p, h): TMEM lane p = point, unit group h) and one control warpgroup with two issuing
// threads -- issuer 1: every forward / adjoint tcgen05.mma, their commits, the operand-image bulk copies; issuer 2: the stash bulk copies and
// the weight / bias / layer-1 gradient MMAs.  Forward: the jet streams travel as three groups G0 = {value}, G1 = {d/dx, d/dy},
// G2 = {d/dt[, d2/dt2]} with ACT[g] / ACC[g] mbarrier pairs, so the tensor pipe runs G1 / G2 of a layer while the epilogue warps apply tanh
// to G0.  Reverse, per layer l:
//   issuer 1: adjoint image landed, Zbar_l group published -> adjoint MMAs of that group (G1, G2, G0: the order the passes consume them) -> commit
//   issuer 2: per stream k: stashed planes of A_{l-1,k} landed in staging slot k & 1 -> 8 K-steps of 16 points x {Ah x [Zh|Zl], Al x Zh} into the
//             weight-gradient tile -> commit SDONE[slot] -> refill the slot once the epilogue warps have read it; Zh^T 1, Zl^T 1 into the bias
//             tile with the last stream -> commit DW; after layer 2: the layer-1 gradient MMAs -> commit L1
//   epilogue warps: per stream k: wait SDONE (and the stream's adjoint group) -> A_{l-1,k} from the slot, abar_k from tensor memory -> Zbar_{l-1,k}
//             over the Zbar_l,k plane -> publish (G1 early in the five-stream kernel); then wait DW -> drain both tiles into this CTA's gradient slot.
// FWD instantiation: forward sweep only (predict): no stash, the output stage writes the eight fields of every point.
// The launch is a programmatic dependent launch (pe_common.cuh): shared-memory set-up, TMEM allocation and barrier initialisation run while
// the operand-image kernel is still working; everything that reads its output sits behind pe_grid_dep_wait().
// Reference lines: see pe_simt.cu / pe_device.cuh (the epilogue algebra is shared with the SIMT engine).
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include "pe_device.cuh"
#include "pe_tc_common.cuh"

namespace {
using namespace pe_dev;
using namespace pe_tcc;

#ifndef TCF_EW
#define TCF_EW 12
#endif
// unroll factor of the forward epilogue loops over the 4-unit groups of a thread (A/B knob)
#ifndef TCF_FWD_UNROLL
#define TCF_FWD_UNROLL 1
#endif
#define TCF_PRAGMA_(x) _Pragma(#x)
#define TCF_PRAGMA(x) TCF_PRAGMA_(x)
// Two issuing threads: lane 0 of the second control warp drives the weight / bias gradient phases of the reverse sweep -- bulk copies of the
// stash, its MMAs, the slot hand-shakes -- while lane 0 of the first one issues the adjoint MMAs: one thread's tcgen05.mma issue costs ~64
// cycles per instruction, two threads together reach ~40 (tests/probe_umma_timing.py; same-box A/B against one issuer: -4.8 % per step).
static_assert(TCF_EW == 8 || TCF_EW == 12 || TCF_EW == 16, "TCF_EW: 8, 12 or 16 epilogue warps");
// register split of the 64 K registers (setmaxnreg, per warpgroup): control warpgroup / epilogue warps
#if TCF_EW == 8
#define TCF_REG_CTRL "104"
#define TCF_REG_EPI "200"
#elif TCF_EW == 12
#define TCF_REG_CTRL "104"
#define TCF_REG_EPI "136"
#else      // 16 epilogue warps: 640 threads are launched with 96 registers; 64 / 112 needs the whole file and did not start on hardware
#define TCF_REG_CTRL "64"
#define TCF_REG_EPI "104"
#endif
constexpr int F_EPI = 32 * TCF_EW, F_THREADS = F_EPI + 128;
constexpr int F_NH = TCF_EW / 4;                  // unit groups per TMEM lane quadrant
constexpr int F_MAXG = (14 + F_NH - 1) / F_NH;    // 4-unit groups one epilogue thread owns at most (14 per 56-unit plane)
constexpr int F_CTRL = TCF_EW;                    // index of the control warp
constexpr int F_CH = 2048;                        // one chunk of 8 units: 128 points x 16 B
constexpr int F_PLANE = 7 * F_CH;                 // 14,336: 56 units
constexpr int F_STREAM = 2 * F_PLANE;             // 28,672: hi plane then lo plane
constexpr int F_IMG_HALF = 8 * 64 * 16;           // [8 K-chunks][64 rows][8 x fp16] = 8,192
constexpr int F_IMG = 2 * F_IMG_HALF;             // Wh then Wl: 16,384
constexpr int F_IMG_LAYER = 2 * F_IMG;            // forward image, adjoint image
// shared-memory map
constexpr int F_ONES = 0;                                     // 1,024 B of fp16 ones: B operand of the bias-gradient MMAs
constexpr int F_ACT = 1024;                                   // NS x (hi plane | lo plane)
constexpr int F_R = F_ACT + TC_MAX_STREAMS * F_STREAM;        // 144,384: forward: two weight images; reverse: adjoint image + two staging buffers
constexpr int F_STG = F_R + F_IMG;                            // two staging slots of one stream (hi plane | lo plane) each at F_STG + s * F_STREAM
constexpr int F_MISC = F_R + F_IMG + 2 * F_STREAM;            // 218,112: mbarriers + TMEM base slot + tile scale
constexpr int F_COORD = F_MISC + 256;
constexpr int F_RED = F_COORD + 128 * 16;
constexpr int F_BIAS = F_RED + 4096;                          // [PE_MAX_LAYERS][64] floats
constexpr int F_W0 = F_BIAS + PE_MAX_LAYERS * 256;            // [4][64] floats
constexpr int F_XBLK = F_W0 + 1024;                           // [128 points][8 x fp16]: (x, y, t, 1) hi | (x, y, t, 0) lo of the tile: B operand of the layer-1 gradient MMAs
constexpr int F_TOTAL = F_XBLK + 2048;                        // 231,680
static_assert(2 * F_IMG <= F_IMG + 2 * F_STREAM, "the forward image double buffer lives inside the reverse-sweep region");
static_assert(F_TOTAL <= 227 * 1024, "shared memory map exceeds the 227 KB opt-in limit");
constexpr int B_ACC = 0, B_ACT = 24, B_IMG = 48, B_SFULL = 64, B_SDONE = 80, B_SFREE = 96, B_DW = 128, B_TMEM = 136, B_SCALE = 144, B_DRAINED = 160, B_REVGO = 168, B_REVDONE = 176, B_L1 = 184;   // byte offsets in F_MISC
constexpr uint32_t T_ACC = 0, T_DW = 320, T_BIAS = 432;       // tensor-memory columns (dW tile: 128 lanes x 112; bias tile: 128 lanes x 8)
constexpr uint32_t T_L1 = 448;                                // layer-1 gradient: 4 tiles of 8 columns (value stream x [x y t 1 | lo parts], column sums of d/dx, d/dy, d/dt)
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

template <int NS> __device__ __forceinline__ int grp_first(int g) { return g == 0 ? 0 : (g == 1 ? 1 : 3); }
template <int NS> __device__ __forceinline__ int grp_count(int g) { return g == 0 ? 1 : (g == 1 ? 2 : NS - 3); }

// kind::f16 instruction descriptors, fp16 x fp16, fp32 accumulate
__device__ __forceinline__ uint32_t idesc_km(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }                        // K-major, M = 128
__device__ __forceinline__ uint32_t idesc_mn(int N) { return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }  // MN-major, M = 128

// ---- epilogue arithmetic on PAIRS of units: sm_100 executes add / mul / fma on two packed fp32 values per instruction (FADD2 / FMUL2 /
// FFMA2), which halves the instruction count of everything that is element-wise in the unit index: the operand split and join, the
// polynomial branch of tanh, the chain rule and its adjoint.
typedef float2 f2;
__device__ __forceinline__ f2 F2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 F2(float a) { return make_float2(a, a); }
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// two values -> fp16 pair (hi = fp16(x), lo = fp16((x - hi) 2^11); x - hi is exact in fp32)
__device__ __forceinline__ void split2(f2 x, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x.x, x.y);
    const f2 d = __fmul2_rn(__ffma2_rn(__half22float2(h), F2(-1.f), x), F2(LO_SCALE));
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(d.x, d.y));
}
__device__ __forceinline__ f2 join2(uint32_t hi, uint32_t lo) {
    return __ffma2_rn(__half22float2(bits_h2(lo)), F2(LO_INV), __half22float2(bits_h2(hi)));
}
__device__ __forceinline__ void split4(f2 x01, f2 x23, uint2& hi, uint2& lo) { split2(x01, hi.x, lo.x); split2(x23, hi.y, lo.y); }

// tanh of two values: pe_dev::tanh_branchfree with the odd polynomial evaluated on the packed pair (same coefficients, same 0.3 switch)
__device__ __forceinline__ f2 tanh2(f2 x) {
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fabsf(x.x) * 2.885390081777927f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fabsf(x.y) * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(e0 + 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(e1 + 1.0f));
    const float b0 = copysignf(fmaf(-2.0f, r0, 1.0f), x.x), b1 = copysignf(fmaf(-2.0f, r1, 1.0f), x.y);
    const f2 x2 = __fmul2_rn(x, x);
    f2 p = F2(-1382.0f / 155925.0f);
    p = __ffma2_rn(p, x2, F2(62.0f / 2835.0f));
    p = __ffma2_rn(p, x2, F2(-17.0f / 315.0f));
    p = __ffma2_rn(p, x2, F2(2.0f / 15.0f));
    p = __ffma2_rn(p, x2, F2(-1.0f / 3.0f));
    const f2 sm = __ffma2_rn(__fmul2_rn(x, x2), p, x);
    return F2(fabsf(x.x) < 0.3f ? sm.x : b0, fabsf(x.y) < 0.3f ? sm.y : b1);
}

// second-time-derivative terms of the value stream's adjoint (pe_dev::act_bwd, K = 5) for two units:
//   zv -= 2 a (s z_tt) abar_tt + 2 (1 - 3 a^2) A_t z_t abar_tt,   s z_tt = a_tt + 2 a A_t z_t,   z_t = A_t / s  (approximate reciprocal, guarded)
__device__ __forceinline__ f2 tt_terms(f2 a, f2 s, f2 At, f2 Att, f2 btt, f2 zv) {
    float i0, i1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(s.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(s.y));
    const f2 zt = __fmul2_rn(At, F2(s.x > 0.f ? i0 : 0.f, s.y > 0.f ? i1 : 0.f));
    const f2 aAt = __fmul2_rn(a, At);
    const f2 sztt = __ffma2_rn(__fmul2_rn(aAt, F2(2.f)), zt, Att);
    zv = __ffma2_rn(__fmul2_rn(__fmul2_rn(a, sztt), F2(-2.f)), btt, zv);
    const f2 c = F2(fmaf(-3.f * a.x, a.x, 1.f), fmaf(-3.f * a.y, a.y, 1.f));
    return __ffma2_rn(__fmul2_rn(__fmul2_rn(__fmul2_rn(c, At), zt), F2(-2.f)), btt, zv);
}

// adjoint of the tanh layer (pe_dev::act_bwd) for two units at once.  A = stashed outputs (a, a_x, ..), ab = adjoints of the
// outputs; returns the adjoints of the pre-activations in ab.  1 / s through the approximate reciprocal (1 ulp), guarded for s = 0.
template <int K>
__device__ __forceinline__ void act_bwd2(f2 (&ab)[K], const f2 (&A)[K]) {
    const f2 a = A[0];
    const f2 s = F2(fmaf(-a.x, a.x, 1.f), fmaf(-a.y, a.y, 1.f));
    f2 acc = __fmul2_rn(A[1], ab[1]);
    acc = __ffma2_rn(A[2], ab[2], acc);
    acc = __ffma2_rn(A[3], ab[3], acc);                                             // s z_k = A_k
    f2 zv = __ffma2_rn(__fmul2_rn(a, acc), F2(-2.f), __fmul2_rn(s, ab[0]));
    if (K == 5) {
        float i0, i1;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(s.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(s.y));
        const f2 inv_s = F2(s.x > 0.f ? i0 : 0.f, s.y > 0.f ? i1 : 0.f);
        const f2 zt = __fmul2_rn(A[3], inv_s);                                      // z_t
        const f2 aA3 = __fmul2_rn(a, A[3]);
        const f2 sztt = __ffma2_rn(__fmul2_rn(aA3, F2(2.f)), zt, A[K - 1]);         // s z_tt = a_tt + 2 a s z_t^2
        zv = __ffma2_rn(__fmul2_rn(__fmul2_rn(a, sztt), F2(-2.f)), ab[K - 1], zv);
        const f2 c = F2(fmaf(-3.f * a.x, a.x, 1.f), fmaf(-3.f * a.y, a.y, 1.f));
        zv = __ffma2_rn(__fmul2_rn(__fmul2_rn(__fmul2_rn(c, A[3]), zt), F2(-2.f)), ab[K - 1], zv);
        const f2 zb3 = __ffma2_rn(s, ab[3], __fmul2_rn(__fmul2_rn(aA3, F2(-4.f)), ab[K - 1]));
        ab[K - 1] = __fmul2_rn(s, ab[K - 1]);
        ab[3] = zb3;
        ab[1] = __fmul2_rn(s, ab[1]); ab[2] = __fmul2_rn(s, ab[2]);
    } else {
#pragma unroll
        for (int k = 1; k < K; ++k) ab[k] = __fmul2_rn(s, ab[k]);
    }
    ab[0] = zv;
}

// Operand images per matrix m: forward B operand [n = out unit j][k = in unit i] at [(i >> 3)][j][i & 7] and adjoint B operand
// [n = i][k = j] at [(j >> 3)][i][j & 7], each as an fp16 pair (hi image, then lo image), zero padded to 64 x 64.  16 blocks of 256 per matrix.
__global__ void tcf_image_kernel(const float* __restrict__ params, PeLayout lay, uint8_t* __restrict__ images) {
    pe_grid_dep_wait();          // the parameters come from the previous step's Adam kernel
    pe_grid_dep_trigger();       // the residual kernel may start its prologue (it waits for this grid before it touches the images)
    const int m = blockIdx.x >> 4;
    const int din = lay.d[m], dout = lay.d[m + 1], ldw = lay.ldw[m];
    const float* W = params + lay.woff[m];
    uint8_t* img = images + (size_t)m * F_IMG_LAYER;
    __half* fhi = reinterpret_cast<__half*>(img);
    __half* flo = reinterpret_cast<__half*>(img + F_IMG_HALF);
    __half* ahi = reinterpret_cast<__half*>(img + F_IMG);
    __half* alo = reinterpret_cast<__half*>(img + F_IMG + F_IMG_HALF);
    const int e = (blockIdx.x & 15) * 256 + threadIdx.x;
    const int i = e >> 6, j = e & 63;
    const float w = (i < din && j < dout) ? W[(size_t)i * ldw + j] : 0.f;
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn((w - __half2float(h)) * LO_SCALE);
    const int of = (i >> 3) * 512 + j * 8 + (i & 7), oa = (j >> 3) * 512 + i * 8 + (j & 7);
    fhi[of] = h; flo[of] = l;
    ahi[oa] = h; alo[oa] = l;
}

struct TcfArgs {
    PeResidArgs r;
    const uint8_t* images;
    pe_term_desc term2;
    const float* points2;
    const float* aux2;
    int n2;
    float inv_n2;
    int fast;                   // 1 = 16-bit forward mode: forward layer GEMMs as single fp16 products (adjoint / weight-gradient GEMMs stay fp16 pairs)
    unsigned long long* prof;   // PROF instantiation: 32 cycle counters (0..15 epilogue thread 0, 16..31 issuer) of CTA prof_cta
    int prof_cta;               // $PE_PROF_CTA (default 0)
    float* fields_out;          // FWD instantiation (forward sweep only, `predict`): [n][8] = u, v, s11, s22, s12, e11, e22, e12
    int fields_aux_k;           // ... composite u = P + D N on the (value, x, y, t) streams: r.aux = [n][2][4][5] (0: plain outputs)
};

#define TCF_PROF(slot) do { if (PROF) { if (prof_on) { const long long now_ = clock64(); atomicAdd(args.prof + (slot), (unsigned long long)(now_ - prof_t)); prof_t = now_; } } } while (0)

// ---- lean MMA issue.  One thread issues every tcgen05.mma of the kernel, and that thread executes ordinary (non-uniform) code: each MMA costs
// it the moves of its operands into uniform registers plus whatever integer work builds the descriptors.  Measured (tests/probe_umma_timing.py):
// with descriptors rebuilt by shifts / masks per MMA the ISSUE takes ~77 cycles per instruction for every shape up to N = 128 -- more than the
// tensor pipe needs to execute them -- so all loops over streams and K-steps are fully unrolled here and a descriptor is one 32-bit add of a
// compile-time constant to a base word: the matrix start address (bits 0..13, 16-byte units; all offsets stay below 2^14) with the leading-
// dimension offset in bits 16..29; the upper word (stride offset, version bit) is shared by all descriptors of a family.
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }

// MMAs of one layer GEMM for streams [K0, K0 + NK): per K-step of 16 the two cross products, then the hi x hi products (the first of them
// scales the accumulated cross products by 2^-11); K-steps interleaved across the streams of the call, so that consecutive MMAs do not
// accumulate into the same tile when the group has more than one stream.  a_lo / b_lo: base words of the activation planes (stream 0, hi
// plane, K-step 0) and of the weight image (hi half); km_hi: upper word of K-major descriptors (SBO = 128).
// fast: 16-bit forward mode (BASELINE config 3, "16-bit forward / fp32 gradient"): only the hi x hi products, 4 MMAs per stream and layer.
template <int K0, int NK>
__device__ __forceinline__ void issue_group(uint32_t tbase, uint32_t a_lo, uint32_t b_lo, uint32_t km_hi, uint32_t id, int ksteps, bool fast = false) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (s < ksteps && !fast) {
            const uint64_t bhi = mk_desc(b_lo + (uint32_t)((2 * s * 1024) >> 4), km_hi), blo = mk_desc(b_lo + (uint32_t)((F_IMG_HALF + 2 * s * 1024) >> 4), km_hi);
#pragma unroll
            for (int k = K0; k < K0 + NK; ++k)
                mma_bf16_ss(tbase + T_ACC + 64u * k, mk_desc(a_lo + (uint32_t)((k * F_STREAM + 2 * s * F_CH) >> 4), km_hi), blo, id, s > 0 ? 1u : 0u);
#pragma unroll
            for (int k = K0; k < K0 + NK; ++k)
                mma_bf16_ss(tbase + T_ACC + 64u * k, mk_desc(a_lo + (uint32_t)((k * F_STREAM + F_PLANE + 2 * s * F_CH) >> 4), km_hi), bhi, id, 1u);
        }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (s < ksteps) {
            const uint64_t bhi = mk_desc(b_lo + (uint32_t)((2 * s * 1024) >> 4), km_hi);
#pragma unroll
            for (int k = K0; k < K0 + NK; ++k) {
                const uint64_t a = mk_desc(a_lo + (uint32_t)((k * F_STREAM + 2 * s * F_CH) >> 4), km_hi);
                if (s == 0 && !fast) mma_f16_ss_scaled11(tbase + T_ACC + 64u * k, a, bhi, id);
                else mma_bf16_ss(tbase + T_ACC + 64u * k, a, bhi, id, s > 0 ? 1u : 0u);
            }
        }
}

// FWD: forward sweep only (the reference's predict, plate:449-461: u, v, stresses and strains of a point batch) -- no stash, no seeds, no
// reverse sweep; the output stage writes the eight fields of every point.
template <int NS, bool PROF, bool FWD>
__global__ void __launch_bounds__(F_THREADS, 1) resid_tcf_kernel(const TcfArgs args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STASH_LAYER = NS * F_STREAM;                 // bytes per stashed layer: NS x (hi plane | lo plane)
    const PeResidArgs& A = args.r;
    const PeLayout& lay = A.lay;
    const pe_term_desc& T = A.term;
    const pe_term_desc& T2 = args.term2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = lay.L;
    const bool prof_on = PROF && args.prof != nullptr && (int)blockIdx.x == args.prof_cta && (tid == 0 || tid == F_EPI || tid == F_EPI + 32);
    long long prof_t = 0;
    (void)prof_on; (void)prof_t;
    uint8_t* act = smem + F_ACT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + F_MISC + B_TMEM);
    float* tile_scale = reinterpret_cast<float*>(smem + F_MISC + B_SCALE);      // [0..3]: per-warp seed maxima of the tile
    float* coord = reinterpret_cast<float*>(smem + F_COORD);
    float* red = reinterpret_cast<float*>(smem + F_RED);
    float* sbias = reinterpret_cast<float*>(smem + F_BIAS);
    float* sw0 = reinterpret_cast<float*>(smem + F_W0);
    const uint32_t act_s = smem_u32(act), r_s = smem_u32(smem + F_R), stg_s = smem_u32(smem + F_STG), ones_s = smem_u32(smem + F_ONES);
    const uint32_t bar0 = smem_u32(smem + F_MISC);
    const uint32_t bar_acc = bar0 + B_ACC, bar_act = bar0 + B_ACT, bar_img = bar0 + B_IMG;
    const uint32_t bar_sfull = bar0 + B_SFULL, bar_sdone = bar0 + B_SDONE, bar_sfree = bar0 + B_SFREE, bar_dw = bar0 + B_DW, bar_drained = bar0 + B_DRAINED;
    const uint32_t bar_revgo = bar0 + B_REVGO, bar_revdone = bar0 + B_REVDONE, bar_l1 = bar0 + B_L1;

    pe_grid_dep_trigger();                           // the reduction kernel's CTAs may be scheduled as SMs drain (they wait for this grid)
    for (int i = tid; i < F_TOTAL / 16; i += F_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pe_grid_dep_wait();                              // operand images (previous kernel) and, through it, the parameters of the last Adam step
    const int slot = A.slot_base + blockIdx.x;
    float* gpart = A.grad_partials + (size_t)slot * lay.total;
    uint8_t* stash = reinterpret_cast<uint8_t*>(A.stash + (size_t)blockIdx.x * A.stash_floats);
    const float* __restrict__ params = A.params;
    if (!FWD) for (int i = tid; i < lay.total; i += F_THREADS) __stcg(gpart + i, 0.f);
    __syncthreads();
    for (int i = tid; i < 512; i += F_THREADS) reinterpret_cast<__half*>(smem + F_ONES)[i] = __float2half_rn(1.f);
    if (tid < 64) {                                  // first-layer weights (3 x d1) and bias -> smem, once per launch
        const bool in = tid < lay.d[1];
        sw0[tid] = in ? __ldg(params + lay.woff[0] + tid) : 0.f;
        sw0[64 + tid] = in ? __ldg(params + lay.woff[0] + lay.ldw[0] + tid) : 0.f;
        sw0[128 + tid] = in ? __ldg(params + lay.woff[0] + 2 * lay.ldw[0] + tid) : 0.f;
        sw0[192 + tid] = in ? __ldg(params + lay.boff[0] + tid) : 0.f;
    }
    for (int i = tid; i < L * 64; i += F_THREADS) {  // all biases (pads zero)
        const int m = i >> 6, j = i & 63;
        sbias[i] = (j < lay.d[m + 1]) ? __ldg(params + lay.boff[m] + j) : 0.f;
    }
    if (tid == 0) {
        for (int g = 0; g < 3; ++g) { mbar_init(bar_acc + 8 * g, 1); mbar_init(bar_act + 8 * g, F_EPI); }
        for (int b = 0; b < 2; ++b) mbar_init(bar_img + 8 * b, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar_sfull + 8 * b, 1); mbar_init(bar_sdone + 8 * b, 1); mbar_init(bar_sfree + 8 * b, F_EPI); }
        mbar_init(bar_dw, 1);
        mbar_init(bar_drained, F_EPI);
        mbar_init(bar_revgo, 1); mbar_init(bar_revdone, 1); mbar_init(bar_l1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == F_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();                              // zero-filled regions and the ones block before any async-proxy access
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *tmem_slot;
    const int ntiles_main = (A.n + TC_P - 1) / TC_P;
    const int ntiles = ntiles_main + (args.n2 + TC_P - 1) / TC_P;
    // forward image of matrix i-1 (i = 2..L) lives in buffer i & 1 of the region F_R
    auto fwd_src = [&](int i) { return args.images + (size_t)(i - 1) * F_IMG_LAYER; };
    auto adj_src = [&](int l) { return args.images + (size_t)(l - 1) * F_IMG_LAYER + F_IMG; };      // adjoint image of matrix l-1 (layer l)

    if (warp >= F_CTRL) {
        // ============================================================================================ control warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " TCF_REG_CTRL ";");
        if (warp == F_CTRL && lane == 0) {
            uint32_t pact = 0, pimg = 0, psfull = 0, psfree = 0;      // parity bits of the phases this thread waits for next
            uint32_t prevdone = 0;
            (void)psfull; (void)psfree;
            uint32_t n_acc2 = 0;
            const bool fast = args.fast != 0;
            auto load_fwd = [&](int i) {
                const uint32_t b = (uint32_t)(i & 1);
                mbar_expect_tx(bar_img + 8 * b, F_IMG);
                tma_load_1d(r_s + b * F_IMG, fwd_src(i), F_IMG, bar_img + 8 * b);
            };
            auto wait_img = [&](uint32_t b) { mbar_wait(bar_img + 8 * b, (pimg >> b) & 1u); pimg ^= 1u << b; };
            auto wait_act = [&](int g) { mbar_wait(bar_act + 8 * g, (pact >> g) & 1u); pact ^= 1u << g; };
            // descriptor base words (see issue_group): K-major activation planes / weight images, MN-major planes / staging slots / ones block
            const uint32_t km_hi = desc_hi(128), mn_hi = desc_hi(F_CH);
            const uint32_t a_lo = desc_lo(act_s, F_CH), b_lo0 = desc_lo(r_s, 1024), b_lo1 = desc_lo(r_s + F_IMG, 1024);
            const uint32_t z_lo = desc_lo(act_s, 128), g_lo = desc_lo(stg_s, 128);
            const uint64_t d_ones = mk_desc(desc_lo(ones_s, 128), desc_hi(256));
            load_fwd(2);
            if (L >= 3) load_fwd(3);
            if (PROF) prof_t = clock64();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // tiles of the fused primal-only set carry the value stream alone: no MMAs for the derivative streams (their seeds are zero);
                // the hand-shakes of the other stream groups still happen, with nothing in between
                const bool sec = tile >= ntiles_main;
                // ---------------------------------------------------------------- forward: layers 2..L, group by group
                for (int l = 2; l <= L; ++l) {
                    const uint32_t b = (uint32_t)(l & 1);
                    const int NF = (lay.d[l] <= 16) ? 16 : 64;
                    wait_img(b);
                    TCF_PROF(16);
#pragma unroll 1
                    for (int g = 0; g < 3; ++g) {
                        wait_act(g);
                        TCF_PROF(17 + 2 * g);
                        fence_after();
                        {
                            const uint32_t b_lo = b ? b_lo1 : b_lo0, id = idesc_km(NF);
                            const int ksteps = (lay.d[l - 1] + 15) >> 4;
                            if (g == 0) issue_group<0, 1>(tbase, a_lo, b_lo, km_hi, id, ksteps, fast);
                            else if (sec) { }
                            else if (g == 1) issue_group<1, 2>(tbase, a_lo, b_lo, km_hi, id, ksteps, fast);
                            else issue_group<3, NS - 3>(tbase, a_lo, b_lo, km_hi, id, ksteps, fast);
                        }
                        mma_commit(bar_acc + 8 * g);
                        TCF_PROF(18 + 2 * g);
                        if (g == 2) ++n_acc2;
                        if (g == 0 && l >= 3 && l + 1 <= L) {
                            // every MMA of layer l-1 precedes G0 of layer l in the pipe: once its last group is complete its image buffer
                            // (= the buffer of image l + 1) is free
                            mbar_wait(bar_acc + 16, (n_acc2 - 1) & 1u);
                            load_fwd(l + 1);
                            TCF_PROF(23);
                        }
                    }
                }
                // ---------------------------------------------------------------- reverse sweep: layers L..2
                // the region F_R changes hands: wait until the output-layer MMAs (the last readers of a forward image) are complete
                mbar_wait(bar_acc + 16, (n_acc2 - 1) & 1u);
                if (FWD) {                                   // forward only: the image buffers go back to layers 2 and 3 of the next tile
                    if (tile + (int)gridDim.x < ntiles) {
                        load_fwd(2);
                        if (L >= 3) load_fwd(3);
                    }
                    continue;
                }
                mbar_expect_tx(bar_img, F_IMG);
                tma_load_1d(r_s, adj_src(L), F_IMG, bar_img);
                mbar_arrive(bar_revgo);                      // the second issuer may start this tile's gradient phases (the staging slots are free)
                for (int l = L; l >= 2; --l) {
                    const int dout = lay.d[l];
                    wait_img(0);
                    TCF_PROF(24);
                    // every jet stream is its own product abar_k = Zbar_k W^T and the epilogue warps take the streams one at a time (d/dx first, the
                    // value stream last): group by group in that order, one commit each.  They publish Zbar_{l,1..2} (G1) as soon as those two
                    // planes are written, two passes before the rest: the 24 MMAs of G1 run under those passes (same-box A/B of the group order
                    // alone: -0.45 % per step; profiles/r2_ab_shot21.txt)
                    {
                        const int ks = (dout + 15) >> 4;
                        const uint32_t id = idesc_km(64);
                        wait_act(1);
                        TCF_PROF(25);
                        fence_after();
                        if (!sec) issue_group<1, 2>(tbase, a_lo, b_lo0, km_hi, id, ks);
                        mma_commit(bar_acc + 8);
                        wait_act(2);
                        fence_after();
                        if (!sec) issue_group<3, NS - 3>(tbase, a_lo, b_lo0, km_hi, id, ks);
                        mma_commit(bar_acc + 16);
                        wait_act(0);
                        fence_after();
                        issue_group<0, 1>(tbase, a_lo, b_lo0, km_hi, id, ks);
                        mma_commit(bar_acc);
                    }
                    ++n_acc2;
                    TCF_PROF(26);
                    if (l > 2) {                             // the adjoint image is free once this layer's adjoint MMAs are complete (G0 is the last group)
                        mbar_wait(bar_acc, (n_acc2 - 1) & 1u);
                        mbar_expect_tx(bar_img, F_IMG);
                        tma_load_1d(r_s, adj_src(l - 1), F_IMG, bar_img);
                    }
                }
                mbar_wait(bar_revdone, prevdone);            // every gradient MMA of the tile is complete and both staging slots are free
                prevdone ^= 1u;
                pact ^= 7u;                                  // the Zbar_1 publish (layer-1 gradient MMAs) is the second issuer's phase; it is complete by now
                if (tile + (int)gridDim.x < ntiles) {
                    load_fwd(2);
                    if (L >= 3) load_fwd(3);
                }
            }
        }
        else if (!FWD && warp == F_CTRL + 1 && lane == 0) {
            // ---------------------------------------------------------------- second issuer: weight / bias gradient phases of the reverse sweep
            uint32_t pact = 0, psfull = 0, psfree = 0, pdrained = 1, prevgo = 0, n_dw = 0;
            auto wait_act = [&](int g) { mbar_wait(bar_act + 8 * g, (pact >> g) & 1u); pact ^= 1u << g; };
            const uint32_t mn_hi = desc_hi(F_CH);
            const uint32_t z_lo = desc_lo(act_s, 128), g_lo = desc_lo(stg_s, 128);
            const uint64_t d_ones = mk_desc(desc_lo(ones_s, 128), desc_hi(256));
            if (PROF) prof_t = clock64();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const bool sec = tile >= ntiles_main;
                // the L - 1 phases of every ACT barrier that belong to the forward sweep (layers 2..L) are the first issuer's
                if ((L - 1) & 1) pact ^= 7u;
                mbar_wait(bar_revgo, prevgo);                // forward MMAs complete: the region F_R changes hands
                prevgo ^= 1u;
                for (int l = L; l >= 2; --l) {
                    const int dout = lay.d[l];
                    const uint8_t* stash_in = stash + (size_t)(l - 2) * STASH_LAYER;     // outputs of layer l-1 = inputs A of layer l
                    // the stashed planes of A_{l-1} come back one stream (hi plane | lo plane, 28,672 B) per bulk copy into two staging slots
                    auto load_stage = [&](int k) {
                        const uint32_t sb = (uint32_t)(k & 1);
                        mbar_expect_tx(bar_sfull + 8 * sb, F_STREAM);
                        tma_load_1d(stg_s + sb * F_STREAM, stash_in + (size_t)k * F_STREAM, F_STREAM, bar_sfull + 8 * sb);
                    };
                    const int nst = sec ? 1 : NS;               // streams of this tile
                    load_stage(0);                              // both slots are free: the previous layer's DW phase is complete
                    if (!sec) load_stage(1);
                    if (l > 2) {                                // pull the planes of the next (shallower) layer towards L2 while this layer runs
                        const uint8_t* nxt = stash + (size_t)(l - 3) * STASH_LAYER;
#pragma unroll
                        for (int k = 0; k < NS; ++k)
                            if (k < nst) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + (size_t)k * F_STREAM), "r"(F_STREAM) : "memory");
                    }
                    wait_act(0); wait_act(1); wait_act(2);
                    fence_after();
                    // ---- weight gradient  dW = sum_k A_k^T Zbar_k  (K = 128 points, 8 K-steps of 16) as ONE M = 128 MMA per K-step: the A operand is
                    //      the staged stream read MN-major, rows 0..55 = Ah units, rows 56..111 = Al units (rows 112..127: whatever follows the slot,
                    //      ignored); the B operand [Zh | Zl] (N = 112).  Tile: rows i, columns j: Ah^T Zh | rows i, columns 56 + j: Ah^T Zl | rows 56 + i,
                    //      columns j: Al^T Zh (the cross products carry 2^11 and are resolved when the tile is drained).  The MMA count, not the MMA
                    //      size, is what the tensor pipe charges for at these shapes (tests/probe_umma_timing.py: ~64 cycles per instruction up to N = 128).
                    // the gradient tiles of the previous layer (or tile) must have been drained before they are overwritten; the epilogue warps
                    // publish Zbar first and drain afterwards, behind the adjoint MMAs issued above
                    TCF_PROF(27);
                    mbar_wait(bar_drained, pdrained);
                    pdrained ^= 1u;
                    TCF_PROF(30);
                    const int nzc = (dout + 7) >> 3;        // unit chunks of Zbar_l that hold data
                    const uint32_t id112 = idesc_mn(112), idz = idesc_mn(8 * nzc), id8 = idesc_mn(8);
#pragma unroll 1
                    for (int k = 0; k < nst; ++k) {
                        const uint32_t sb = (uint32_t)(k & 1);
                        if (k == 0) TCF_PROF(27);
                        mbar_wait(bar_sfull + 8 * sb, (psfull >> sb) & 1u);
                        psfull ^= 1u << sb;
                        if (k == 0) TCF_PROF(31);
                        const uint32_t ga = g_lo + sb * (uint32_t)(F_STREAM >> 4);
                        const uint32_t zh = z_lo + (uint32_t)k * (uint32_t)(F_STREAM >> 4), zl = zh + (uint32_t)(F_PLANE >> 4);
#pragma unroll
                        for (int s = 0; s < 8; ++s) {                        // K-steps of 16 points = 256 B
                            const uint64_t da = mk_desc(ga + 16u * s, mn_hi);
                            const uint32_t first = (k > 0 || s > 0) ? 1u : 0u;
                            if (nzc == 7) {                                  // the two Z planes are contiguous: one N = 112 MMA
                                mma_bf16_ss(tbase + T_DW, da, mk_desc(zh + 16u * s, mn_hi), id112, first);
                            } else {
                                mma_bf16_ss(tbase + T_DW, da, mk_desc(zh + 16u * s, mn_hi), idz, first);
                                mma_bf16_ss(tbase + T_DW + 56, da, mk_desc(zl + 16u * s, mn_hi), idz, first);
                            }
                        }
                        if (k == nst - 1) {
                            // bias gradient: column 0 of  [Zh | Zl]^T 1: the value-stream planes as MN-major A operand (M = 128: rows 0..55 sums
                            // of Zh, rows 56..111 sums of Zl), a block of ones as B operand (N = 8).  With the LAST stream (whose completion the pass
                            // that overwrites the value-stream plane waits for): the first passes of the epilogue warps start 8 MMAs earlier
#pragma unroll
                            for (int s = 0; s < 8; ++s) mma_bf16_ss(tbase + T_BIAS, mk_desc(z_lo + 16u * s, mn_hi), d_ones, id8, s > 0 ? 1u : 0u);
                        }
                        mma_commit(bar_sdone + 8 * sb);        // the epilogue warps' pass over stream k starts (staged planes + accumulators -> Zbar_{l-1,k})
                        if (k >= 1 && k + 1 < nst) {           // slot of stream k - 1: refilled (stream k + 1) once the epilogue warps have read it
                            const uint32_t ob = sb ^ 1u;
                            mbar_wait(bar_sfree + 8 * ob, (psfree >> ob) & 1u);
                            psfree ^= 1u << ob;
                            load_stage(k + 1);
                        }
                    }
                    mma_commit(bar_dw);
                    ++n_dw;
                    TCF_PROF(27);
                    mbar_wait(bar_dw, (n_dw - 1) & 1u);      // every MMA of the layer is complete: the adjoint image is free
                    mbar_wait(bar_sfree, psfree & 1u);       // ... and so are both staging slots once the last passes of the epilogue warps are through
                    psfree ^= 1u;
                    if (!sec) { mbar_wait(bar_sfree + 8, (psfree >> 1) & 1u); psfree ^= 2u; }
                    TCF_PROF(28);
                }
                // ---- layer-1 gradient  dW_0 = sum_k X_k^T Zbar_1,k, db_0 = Zbar_1,0^T 1  on the tensor pipe: the input jets are X_0 = (x, y, t) and constant
                // unit vectors times the input scale for d/dx, d/dy, d/dt (zero for d2/dt2), so the products are  [Zh | Zl]^T [xh yh th 1 xl yl tl 0]
                // (value stream; the B operand is the tile's coordinate block, one 16-byte row per point) and the column sums of the three first-
                // derivative streams ([Zh | Zl]^T 1, the ones block).  The epilogue warps spent 7 % of a tile on these sums in scalar code.
                wait_act(0); wait_act(1); wait_act(2);
                fence_after();
                {
                    const uint32_t id8 = idesc_mn(8);
                    const uint32_t x_lo = desc_lo(smem_u32(smem + F_XBLK), 128);
#pragma unroll
                    for (int s = 0; s < 8; ++s) mma_bf16_ss(tbase + T_L1, mk_desc(z_lo + 16u * s, mn_hi), mk_desc(x_lo + 16u * s, desc_hi(256)), id8, s > 0 ? 1u : 0u);
                    if (!sec) {
#pragma unroll
                        for (int k = 1; k < 4; ++k)
#pragma unroll
                            for (int s = 0; s < 8; ++s)
                                mma_bf16_ss(tbase + T_L1 + 8u * k, mk_desc(z_lo + (uint32_t)k * (uint32_t)(F_STREAM >> 4) + 16u * s, mn_hi), d_ones, id8, s > 0 ? 1u : 0u);
                    }
                    mma_commit(bar_l1);
                }
                TCF_PROF(29);
                mbar_arrive(bar_revdone);                    // (after the wait above: the first issuer skips that ACT phase and must find it complete)
            }
        }
        __syncwarp();
    } else {
        // ============================================================================================ epilogue warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " TCF_REG_EPI ";");
        const int p = 32 * (warp & 3) + lane;            // TMEM lane = point of the tile
        const int h = warp >> 2;                         // unit group: 4-unit groups [n4 h / F_NH, n4 (h + 1) / F_NH) of a layer with n4 groups
        const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t pacc = 0, pdw = 0, psdone = 0, pl1 = 0;
        auto wait_acc = [&](int g) { mbar_wait(bar_acc + 8 * g, (pacc >> g) & 1u); pacc ^= 1u << g; };
        auto publish = [&](int g) { mbar_arrive(bar_act + 8 * g); };
        // planes in shared memory were written through the generic proxy and are read next by MMAs (async proxy): a shared-memory proxy fence
        // per publish.  The stashed copies in global memory are read by bulk copies only in the reverse sweep: ONE all-state-space proxy fence
        // at the end of the forward sweep (publish_fences_global) covers them -- a global fence per publish makes every epilogue phase wait for
        // its stash stores to be acknowledged by L2 (measured: the forward epilogue phases were bound by it, not by their instructions).
        auto publish_fences = [&]() { fence_async_smem(); fence_before(); };
        auto publish_fences_global = [&]() { asm volatile("fence.proxy.async;" ::: "memory"); fence_before(); };
        auto c4_lo = [&](int d) { return ((d + 3) >> 2) * h / F_NH; };
        auto c4_hi = [&](int d) { return ((d + 3) >> 2) * (h + 1) / F_NH; };
        float tsum[PE_MAX_TERMS], tsum2[PE_MAX_TERMS];
#pragma unroll
        for (int i = 0; i < PE_MAX_TERMS; ++i) { tsum[i] = 0.f; tsum2[i] = 0.f; }
        // 4 units (group c4) of stream k of this thread's point -> hi / lo planes in shared memory [and in the stash layer `st`]
        auto put4 = [&](uint8_t* st, int k, int c4, f2 v01, f2 v23) {
            uint2 hi, lo;
            split4(v01, v23, hi, lo);
            const int o = k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
            *reinterpret_cast<uint2*>(act + o) = hi;
            *reinterpret_cast<uint2*>(act + o + F_PLANE) = lo;
            if (st) {
                __stcg(reinterpret_cast<uint2*>(st + o), hi);
                __stcg(reinterpret_cast<uint2*>(st + o + F_PLANE), lo);
            }
        };
        auto get4 = [&](int k, int c4, f2& v01, f2& v23) {       // this thread's own entries of the planes in shared memory
            const int o = k * F_STREAM + (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
            const uint2 hi = *reinterpret_cast<const uint2*>(act + o), lo = *reinterpret_cast<const uint2*>(act + o + F_PLANE);
            v01 = join2(hi.x, lo.x); v23 = join2(hi.y, lo.y);
        };

        if (PROF) prof_t = clock64();
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const bool sec = tile >= ntiles_main;                          // tile of the fused primal-only set (CTA-uniform)
            const pe_term_desc& Tc = sec ? T2 : T;
            const int pt = (sec ? tile - ntiles_main : tile) * TC_P + p;
            const bool valid = pt < (sec ? args.n2 : A.n);
            const float* row = (sec ? args.points2 : A.points) + (size_t)(valid ? pt : 0) * Tc.ld;
            if (h == 0) {
                float x = 0.f, y = 0.f, t = 0.f;
                if (valid) { x = row[0]; y = row[1]; t = row[2]; }
                const float xn = fmaf(x, Tc.in_scale[0], Tc.in_shift[0]), yn = fmaf(y, Tc.in_scale[1], Tc.in_shift[1]), tn = fmaf(t, Tc.in_scale[2], Tc.in_shift[2]);
                *reinterpret_cast<float4*>(coord + 4 * p) = make_float4(xn, yn, tn, valid ? 1.f : 0.f);
                uint32_t h01, l01, h23, l23;                 // coordinate block row of this point: xh yh th 1 | xl yl tl 0  (fp16 pairs)
                split2(F2(xn, yn), h01, l01);
                split2(F2(tn, 1.f), h23, l23);
                *reinterpret_cast<uint4*>(smem + F_XBLK + 16 * p) = make_uint4(h01, h23, l01, l23);
            } else if (h == 1) {
                const int nt = tile + (int)gridDim.x;                      // pull this CTA's next tile towards L2
                if (nt < ntiles) {
                    const bool nsec = nt >= ntiles_main;
                    const int npt = (nsec ? nt - ntiles_main : nt) * TC_P + p;
                    if (npt < (nsec ? args.n2 : A.n)) {
                        const float* nrow = (nsec ? args.points2 : A.points) + (size_t)npt * (nsec ? T2.ld : T.ld);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow));
                        if (!FWD && !nsec && A.aux) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.aux + (size_t)npt * 50));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.aux + (size_t)npt * 50 + 32));
                        }
                    }
                }
            }
            named_bar_sync(1, F_EPI);
            // ================================================================ layer 1 (3 -> d1): per-thread FFMA, all streams
            {
                const float4 c4v = *reinterpret_cast<const float4*>(coord + 4 * p);
                const int d1 = lay.d[1];
                // two units per instruction (packed fp32 pairs), the algebra of the hidden-layer epilogues: a = tanh(z + b), s = 1 - a^2,
                // a_k = s z_k, a_tt = s z_tt - 2 a a_t z_t with z_x = sx w0, z_y = sy w1, z_t = st w2, z_tt = 0 (the input is linear in x, y, t)
                const f2 cx = F2(c4v.x), cy = F2(c4v.y), ct = F2(c4v.z);
                const f2 sx = F2(Tc.in_scale[0]), sy = F2(Tc.in_scale[1]), stt = F2(Tc.in_scale[2]);
#pragma unroll 1
                for (int c4 = c4_lo(d1); c4 < c4_hi(d1); ++c4) {
                    f2 o[NS][2];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int j = 4 * c4 + 2 * hh;                      // pads: zero weights and bias -> tanh(0) = 0, zero derivatives
                        const f2 w0 = *reinterpret_cast<const f2*>(sw0 + j), w1 = *reinterpret_cast<const f2*>(sw0 + 64 + j);
                        const f2 w2 = *reinterpret_cast<const f2*>(sw0 + 128 + j), bb = *reinterpret_cast<const f2*>(sw0 + 192 + j);
                        const f2 z = __ffma2_rn(cx, w0, __ffma2_rn(cy, w1, __ffma2_rn(ct, w2, bb)));
                        const f2 a = tanh2(z);
                        const f2 sd = F2(fmaf(-a.x, a.x, 1.f), fmaf(-a.y, a.y, 1.f));
                        const f2 zt = __fmul2_rn(stt, w2);
                        const f2 at = __fmul2_rn(sd, zt);
                        o[0][hh] = a;
                        o[1][hh] = __fmul2_rn(sd, __fmul2_rn(sx, w0));
                        o[2][hh] = __fmul2_rn(sd, __fmul2_rn(sy, w1));
                        o[3][hh] = at;
                        if (NS == 5) o[NS - 1][hh] = __fmul2_rn(__fmul2_rn(__fmul2_rn(a, at), zt), F2(-2.f));
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k)
                        if (k == 0 || !sec) put4(FWD ? nullptr : stash, k, c4, o[k][0], o[k][1]);      // stash layer 0 = outputs of layer 1
                }
                publish_fences();
                publish(0); publish(1); publish(2);
                TCF_PROF(0);
            }
            // ================================================================ forward: hidden layers 2..L-1, one stream group at a time
            for (int l = 2; l < L; ++l) {
                const int lo4 = c4_lo(lay.d[l]), hi4 = c4_hi(lay.d[l]);
                const float* bl = sbias + (l - 1) * 64;
                uint8_t* st = FWD ? nullptr : stash + (size_t)(l - 1) * STASH_LAYER;
                // ---- G0: a = tanh(z_0 + b)      (pad units: zero weight columns and bias -> exactly 0)
                wait_acc(0);
                TCF_PROF(1);
                fence_after();
TCF_PRAGMA(unroll TCF_FWD_UNROLL)
                for (int c4 = lo4; c4 < hi4; ++c4) {
                    float z[4];
                    tm_ld4(tlane + T_ACC + 4 * c4, z);
                    const float4 b4 = *reinterpret_cast<const float4*>(bl + 4 * c4);
                    tm_wait_ld();
                    put4(st, 0, c4, tanh2(__fadd2_rn(F2(z[0], z[1]), F2(b4.x, b4.y))), tanh2(__fadd2_rn(F2(z[2], z[3]), F2(b4.z, b4.w))));
                }
                publish_fences();
                publish(0);
                TCF_PROF(2);
                // ---- G1: a_x = s z_x, a_y = s z_y   (a re-read from this thread's own entries of the value planes)
                wait_acc(1);
                TCF_PROF(3);
                fence_after();
TCF_PRAGMA(unroll TCF_FWD_UNROLL)
                for (int c4 = lo4; c4 < (sec ? lo4 : hi4); ++c4) {
                    float z1[4], z2[4];
                    f2 a01, a23;
                    tm_ld4(tlane + T_ACC + 64 + 4 * c4, z1);
                    tm_ld4(tlane + T_ACC + 128 + 4 * c4, z2);
                    get4(0, c4, a01, a23);
                    const f2 s01 = F2(fmaf(-a01.x, a01.x, 1.f), fmaf(-a01.y, a01.y, 1.f)), s23 = F2(fmaf(-a23.x, a23.x, 1.f), fmaf(-a23.y, a23.y, 1.f));
                    tm_wait_ld();
                    put4(st, 1, c4, __fmul2_rn(s01, F2(z1[0], z1[1])), __fmul2_rn(s23, F2(z1[2], z1[3])));
                    put4(st, 2, c4, __fmul2_rn(s01, F2(z2[0], z2[1])), __fmul2_rn(s23, F2(z2[2], z2[3])));
                }
                publish_fences();
                publish(1);
                TCF_PROF(4);
                // ---- G2: a_t = s z_t [, a_tt = s z_tt - 2 a a_t z_t]
                wait_acc(2);
                TCF_PROF(5);
                fence_after();
TCF_PRAGMA(unroll TCF_FWD_UNROLL)
                for (int c4 = lo4; c4 < (sec ? lo4 : hi4); ++c4) {
                    float z3[4], z4[4];
                    f2 a01, a23;
                    tm_ld4(tlane + T_ACC + 192 + 4 * c4, z3);
                    if (NS == 5) tm_ld4(tlane + T_ACC + 256 + 4 * c4, z4);
                    get4(0, c4, a01, a23);
                    const f2 s01 = F2(fmaf(-a01.x, a01.x, 1.f), fmaf(-a01.y, a01.y, 1.f)), s23 = F2(fmaf(-a23.x, a23.x, 1.f), fmaf(-a23.y, a23.y, 1.f));
                    tm_wait_ld();
                    const f2 zt01 = F2(z3[0], z3[1]), zt23 = F2(z3[2], z3[3]);
                    const f2 at01 = __fmul2_rn(s01, zt01), at23 = __fmul2_rn(s23, zt23);
                    put4(st, 3, c4, at01, at23);
                    if (NS == 5) {       // a_tt = s z_tt - 2 a a_t z_t
                        const f2 q01 = __fmul2_rn(__fmul2_rn(a01, at01), zt01), q23 = __fmul2_rn(__fmul2_rn(a23, at23), zt23);
                        put4(st, 4, c4, __ffma2_rn(q01, F2(-2.f), __fmul2_rn(s01, F2(z4[0], z4[1]))), __ffma2_rn(q23, F2(-2.f), __fmul2_rn(s23, F2(z4[2], z4[3]))));
                    }
                }
                publish_fences();
                publish(2);
                TCF_PROF(6);
            }
            // ================================================================ output layer L: residuals, loss partials, seeds
            float inv_sigma = 1.f;
            {
                const int dout = lay.d[L];
                const float* bl = sbias + (L - 1) * 64;
                wait_acc(0); wait_acc(1); wait_acc(2);
                TCF_PROF(7);
                fence_after();
                float Y[NS][PE_UJ];
                float amax = 0.f;
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        float v[8];
                        tm_ld8(tlane + T_ACC + 64 * k, v);
                        tm_wait_ld();
#pragma unroll
                        for (int u = 0; u < PE_UJ; ++u) Y[k][u] = (u < 8 && u < dout && (k == 0 || !sec)) ? v[u < 8 ? u : 0] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < PE_UJ; ++u) if (u < dout) Y[0][u] += bl[u];
                    if (FWD) {                       // predict: fields of this point (same composition and column order as fields_simt_kernel)
                        if (valid) {
                            if (args.fields_aux_k) {       // composite on (value, x, y, t) streams, 5 outputs (plate:382-387)
                                const float* ar = A.aux + (size_t)pt * (2 * args.fields_aux_k * 5);
#pragma unroll
                                for (int o = 0; o < 5; ++o) {
                                    const float D0 = ar[o], N0 = Y[0][o];
#pragma unroll
                                    for (int k = 2; k >= 1; --k) Y[k][o] = ar[args.fields_aux_k * 5 + k * 5 + o] + ar[k * 5 + o] * N0 + D0 * Y[k][o];
                                    Y[0][o] = fmaf(D0, N0, ar[args.fields_aux_k * 5 + o]);
                                }
                            }
                            const bool f7 = T.kind == PE_RES_F7;
                            float4* o4 = reinterpret_cast<float4*>(args.fields_out + (size_t)pt * 8);
                            o4[0] = make_float4(Y[0][0], Y[0][1], f7 ? Y[0][4] : Y[0][2], f7 ? Y[0][5] : Y[0][3]);
                            o4[1] = make_float4(f7 ? Y[0][6] : Y[0][4], Y[1][0], Y[2][1], Y[2][0] + Y[1][1]);      // e11 = u_x, e22 = v_y, e12 = u_y + v_x (plate:393-395)
                        }
                    } else
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
#pragma unroll
                    for (int k = 0; k < NS; ++k)
#pragma unroll
                        for (int u = 0; u < 8; ++u) amax = fmaxf(amax, fabsf(Y[k][u]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                    if (lane == 0) tile_scale[warp] = amax;
                }
                if (FWD) { TCF_PROF(8); continue; }          // next tile (its layer-1 planes are written behind the barrier at the tile start)
                // the tile's seed scale: sigma = 2^-e with max |seed| * sigma in [1, 2)  (1 when all seeds vanish); identical in every thread
                named_bar_sync(1, F_EPI);
                const float m4 = fmaxf(fmaxf(tile_scale[0], tile_scale[1]), fmaxf(tile_scale[2], tile_scale[3]));
                uint32_t eb = (__float_as_uint(m4) >> 23) & 0xFFu;
                eb = (m4 > 0.f) ? min(max(eb, 2u), 252u) : 127u;
                const float sigma = __uint_as_float((254u - eb) << 23);
                inv_sigma = __uint_as_float(eb << 23);
                if (h == 0) {
                    // seeds Zbar_L: units 0..7 in groups 0 and 1 (the only chunk the adjoint / weight-gradient MMAs of the output layer use)
#pragma unroll
                    for (int k = 0; k < NS; ++k)
                        if (k == 0 || !sec) {
                            put4(nullptr, k, 0, F2(Y[k][0] * sigma, Y[k][1] * sigma), F2(Y[k][2] * sigma, Y[k][3] * sigma));
                            put4(nullptr, k, 1, F2(Y[k][4] * sigma, Y[k][5] * sigma), F2(Y[k][6] * sigma, Y[k][7] * sigma));
                        }
                }
                publish_fences_global();                     // the whole stash of this tile before the reverse sweep's bulk copies
                publish(0); publish(1); publish(2);
                TCF_PROF(8);
            }
            // ================================================================ reverse sweep, layers L .. 2
            for (int l = L; l >= 2; --l) {
                const int m = l - 1;
                const int din = lay.d[l - 1], dout = lay.d[l];
                const int lo4 = c4_lo(din), hi4 = c4_hi(din);
                // (the wait for the adjoint MMAs -- abar^{l-1} in the accumulators -- sits inside the passes, behind the slot reads of streams 0 and 1
                // that do not need them: the staging slots are handed back, and the bulk copies of streams 2 and 3 start, while those MMAs still run)
                // ---- through tanh of layer l-1, one jet stream at a time, while the weight-gradient phase runs: as soon as the MMAs of stream k are
                // complete (SDONE) this thread reads its entries of A_{l-1,k} from the staging slot the bulk copy filled -- the stash is read from
                // L2 once, by the copy -- takes abar_k from tensor memory and overwrites the Zbar_l,k plane (no MMA reads it any more) with
                // Zbar_{l-1,k}.  The value stream's adjoint needs all streams: a (pass 0) and the running sum  sum_k A_k abar_k  stay in registers,
                // its plane is written by the last pass (which re-reads A_3 from its slot for the second-time-derivative terms).
                {
                    f2 av[F_MAXG][2], acc[F_MAXG][2];
                    const uint8_t* slot3 = smem + F_STG + F_STREAM;                        // stream 3 lives in slot 3 & 1
                    if (sec) {                       // primal-only tile: one stream, zbar_0 = s abar_0
                        mbar_wait(bar_sdone, psdone & 1u);
                        psdone ^= 1u;
                        wait_acc(0); wait_acc(1); wait_acc(2);
                        TCF_PROF(9);
                        fence_after();
                        const uint8_t* slot = smem + F_STG;
#pragma unroll 1
                        for (int c4 = lo4; c4 < hi4; ++c4) {
                            const int o = (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
                            const uint2 hA = *reinterpret_cast<const uint2*>(slot + o), lA = *reinterpret_cast<const uint2*>(slot + F_PLANE + o);
                            float b0[4];
                            tm_ld4(tlane + T_ACC + 4 * c4, b0);
                            const f2 a0 = join2(hA.x, lA.x), a1 = join2(hA.y, lA.y);
                            const f2 s0 = F2(fmaf(-a0.x, a0.x, 1.f), fmaf(-a0.y, a0.y, 1.f)), s1 = F2(fmaf(-a1.x, a1.x, 1.f), fmaf(-a1.y, a1.y, 1.f));
                            tm_wait_ld();
                            put4(nullptr, 0, c4, __fmul2_rn(s0, F2(b0[0], b0[1])), __fmul2_rn(s1, F2(b0[2], b0[3])));
                        }
                        mbar_arrive(bar_sfree);
                        const uint8_t* dead = stash + (size_t)(l - 2) * STASH_LAYER;
                        for (int i = tid; i < F_STREAM / 128; i += F_EPI)
                            asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead + (size_t)i * 128) : "memory");
                    } else
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        const uint32_t sb = (uint32_t)(k & 1);
                        if (k <= 1) TCF_PROF(10);
                        mbar_wait(bar_sdone + 8 * sb, (psdone >> sb) & 1u);
                        psdone ^= 1u << sb;
                        if (k == 0) TCF_PROF(14);    // from the end of the previous drain to the first stream's weight-gradient MMAs complete
                        if (k == 1) TCF_PROF(15);    // ... the second stream's
                        const uint8_t* slot = smem + F_STG + sb * F_STREAM;
                        // this thread's entries of the slot -> registers, then the slot is released at once (the bulk copy of stream k + 2 is the
                        // longest link of the per-slot chain copy -> MMAs -> pass; the arithmetic below runs while it is in flight)
                        // (not in the last pass: nothing waits for its slots within the layer, and it holds the most registers)
                        const bool EARLY = (k < NS - 1);
                        uint2 hA[F_MAXG], lA[F_MAXG];
                        if (EARLY) {
#pragma unroll
                            for (int g = 0; g < F_MAXG; ++g) {
                                const int c4 = lo4 + g;
                                if (c4 < hi4) {
                                    const int o = (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
                                    hA[g] = *reinterpret_cast<const uint2*>(slot + o); lA[g] = *reinterpret_cast<const uint2*>(slot + F_PLANE + o);
                                }
                            }
                            if (!(NS == 5 && k == 3)) mbar_arrive(bar_sfree + 8 * sb);     // stream 3's slot is read again by the last pass
                        }
                        // adjoint MMAs of this stream's group done: abar_k^{l-1} in the accumulators (the issuer commits G1, G2, G0 in this order)
                        if (k == 1) { wait_acc(1); TCF_PROF(9); fence_after(); }
                        if (k == 3) { wait_acc(2); fence_after(); }
                        if (k == NS - 1) { wait_acc(0); fence_after(); }
#pragma unroll
                        for (int g = 0; g < F_MAXG; ++g) {
                            const int c4 = lo4 + g;
                            if (c4 < hi4) {
                                const int o = (c4 >> 1) * F_CH + p * 16 + (c4 & 1) * 8;
                                if (!EARLY) { hA[g] = *reinterpret_cast<const uint2*>(slot + o); lA[g] = *reinterpret_cast<const uint2*>(slot + F_PLANE + o); }
                                const f2 A0 = join2(hA[g].x, lA[g].x), A1 = join2(hA[g].y, lA[g].y);
                                if (k == 0) {
                                    av[g][0] = A0; av[g][1] = A1;
                                } else {
                                    float bk[4], b4[4], b0[4];
                                    tm_ld4(tlane + T_ACC + 64 * k + 4 * c4, bk);
                                    if (NS == 5 && k == 3) tm_ld4(tlane + T_ACC + 256 + 4 * c4, b4);
                                    if (k == NS - 1) tm_ld4(tlane + T_ACC + 4 * c4, b0);
                                    const f2 a0 = av[g][0], a1 = av[g][1];
                                    const f2 s0 = F2(fmaf(-a0.x, a0.x, 1.f), fmaf(-a0.y, a0.y, 1.f)), s1 = F2(fmaf(-a1.x, a1.x, 1.f), fmaf(-a1.y, a1.y, 1.f));
                                    tm_wait_ld();
                                    const f2 q0 = F2(bk[0], bk[1]), q1 = F2(bk[2], bk[3]);
                                    if (k == 1) { acc[g][0] = __fmul2_rn(A0, q0); acc[g][1] = __fmul2_rn(A1, q1); }
                                    else if (k <= 3) { acc[g][0] = __ffma2_rn(A0, q0, acc[g][0]); acc[g][1] = __ffma2_rn(A1, q1, acc[g][1]); }      // s z_k = A_k
                                    f2 z0 = __fmul2_rn(s0, q0), z1 = __fmul2_rn(s1, q1);
                                    if (NS == 5 && k == 3) {                              // zbar_t = s abar_t - 4 a A_t abar_tt
                                        z0 = __ffma2_rn(__fmul2_rn(__fmul2_rn(a0, A0), F2(-4.f)), F2(b4[0], b4[1]), z0);
                                        z1 = __ffma2_rn(__fmul2_rn(__fmul2_rn(a1, A1), F2(-4.f)), F2(b4[2], b4[3]), z1);
                                    }
                                    put4(nullptr, k, c4, z0, z1);
                                    if (k == NS - 1) {                                    // value stream:  s abar_0 - 2 a sum_k A_k abar_k  [ - tt terms ]
                                        f2 v0 = __ffma2_rn(__fmul2_rn(a0, acc[g][0]), F2(-2.f), __fmul2_rn(s0, F2(b0[0], b0[1])));
                                        f2 v1 = __ffma2_rn(__fmul2_rn(a1, acc[g][1]), F2(-2.f), __fmul2_rn(s1, F2(b0[2], b0[3])));
                                        if (NS == 5) {                                    // here A0 / A1 = a_tt, q0 / q1 = abar_tt
                                            const uint2 h3 = *reinterpret_cast<const uint2*>(slot3 + o), l3 = *reinterpret_cast<const uint2*>(slot3 + F_PLANE + o);
                                            v0 = tt_terms(a0, s0, join2(h3.x, l3.x), A0, q0, v0);
                                            v1 = tt_terms(a1, s1, join2(h3.y, l3.y), A1, q1, v1);
                                        }
                                        put4(nullptr, 0, c4, v0, v1);
                                    }
                                }
                            }
                        }
                        if (!EARLY) {
                            if (NS == 5) { mbar_arrive(bar_sfree); mbar_arrive(bar_sfree + 8); }
                            else mbar_arrive(bar_sfree + 8 * sb);
                        }
                        // Zbar_{l-1,1..2} written: the next adjoint GEMM's first group may start under the last two passes (five streams only: with four
                        // there is one pass left and the early MMAs just compete with the weight-gradient MMAs -- same-box A/B: K = 5 -0.75 %, K = 4 +1 %)
                        if (NS == 5 && k == 2) { publish_fences(); publish(1); }
                        // the stashed planes of stream k have been copied to shared memory and nobody reads them again: drop their (dirty) L2 lines
                        // instead of letting them be written back to HBM -- the stash is scratch that the next tile overwrites.  Without this the
                        // whole stash (5.6 KB per point) goes to DRAM and the live part no longer fits L2 (ncu: 286 MB written per 55 k points).
                        {
                            const uint8_t* dead = stash + (size_t)(l - 2) * STASH_LAYER + (size_t)k * F_STREAM;
                            for (int i = tid; i < F_STREAM / 128; i += F_EPI)
                                asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead + (size_t)i * 128) : "memory");
                        }
                    }
                }
                TCF_PROF(10);
                // Zbar_{l-1} is complete: the next layer's adjoint MMAs (l > 2) / the layer-1 gradient MMAs (l = 2) may start while the tiles are drained
                // (five streams: G1 was published behind pass 2 unless this is a primal-only tile)
                publish_fences();
                publish(0); if (sec || NS != 5) publish(1); publish(2);
                mbar_wait(bar_dw, pdw);                      // every weight / bias gradient MMA of the layer is complete
                pdw ^= 1u;
                TCF_PROF(11);
                fence_after();
                {   // drain.  Tile rows r = 32 * quadrant + lane.  Rows r < din carry Ah^T Zh (columns j) and Ah^T Zl (columns 56 + j) of input unit
                    // i = r; rows 56 <= r < 56 + din carry Al^T Zh of unit i = r - 56.  Two passes with a barrier in between, so that the two adds
                    // every gradient element receives per tile always arrive in the same order (bitwise reproducible sums).  The unit groups of a
                    // quadrant share the columns.  Bias gradient: column 0 of its tile, rows j (hi sums) and 56 + j (lo sums).
                    // (tcgen05.ld is warp-collective: the loads sit behind warp-uniform conditions only, the lane's row decides about the adds)
                    const int r0 = 32 * (warp & 3), r = r0 + lane;
                    const int ldw = lay.ldw[m];
                    float* gW = gpart + lay.woff[m];
                    float* gB = gpart + lay.boff[m];
                    const int nc8 = (dout + 7) >> 3;
                    const float sc_lo = LO_INV * inv_sigma;
                    if (r0 < din) {
                        for (int c8 = h; c8 < nc8; c8 += F_NH) {
                            const int c = 8 * c8;
                            float v[8], vx[8];
                            tm_ld8(tlane + T_DW + c, v);
                            tm_ld8(tlane + T_DW + 56 + c, vx);
                            tm_wait_ld();
                            if (r < din) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] = fmaf(vx[q], LO_INV, v[q]) * inv_sigma;
                                float* dst = gW + (size_t)r * ldw + c;
                                if (c < ldw) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                                if (c + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4], v[5], v[6], v[7]));
                            }
                        }
                    }
                    if (h == F_NH - 1 && r0 < dout) {
                        float vb[8];
                        tm_ld8(tlane + T_BIAS, vb);
                        tm_wait_ld();
                        if (r < dout) atomicAdd(gB + r, vb[0] * inv_sigma);
                    }
                    named_bar_sync(1, F_EPI);
                    if (r0 + 32 > 56 && r0 < 56 + din) {
                        for (int c8 = h; c8 < nc8; c8 += F_NH) {
                            const int c = 8 * c8;
                            float v[8];
                            tm_ld8(tlane + T_DW + c, v);
                            tm_wait_ld();
                            if (r >= 56 && r < 56 + din) {
                                float* dst = gW + (size_t)(r - 56) * ldw + c;
                                if (c < ldw) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0] * sc_lo, v[1] * sc_lo, v[2] * sc_lo, v[3] * sc_lo));
                                if (c + 4 < ldw) atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4] * sc_lo, v[5] * sc_lo, v[6] * sc_lo, v[7] * sc_lo));
                            }
                        }
                    }
                    if (h == F_NH - 1 && r0 + 32 > 56 && r0 < 56 + dout) {
                        float vb[8];
                        tm_ld8(tlane + T_BIAS, vb);
                        tm_wait_ld();
                        if (r >= 56 && r < 56 + dout) atomicAdd(gB + (r - 56), vb[0] * sc_lo);
                    }
                }
                fence_before();
                mbar_arrive(bar_drained);                    // the issuer may overwrite the gradient tiles
                TCF_PROF(12);
            }
            // ================================================================ layer 1 gradient (3 x d1 + bias): drain of the four 8-column tiles
            // Rows r < d1: sums with Zh of unit r; rows 56 <= r < 56 + d1: sums with Zl of unit r - 56 (carry 2^11).  Tile 0 columns: Z^T xh, Z^T yh,
            // Z^T th, Z^T 1, Z^T xl, Z^T yl, Z^T tl (carry 2^11), -; tiles 1..3: column 0 = column sums of the d/dx, d/dy, d/dt streams.  Two passes
            // (hi rows, then lo rows) with a barrier in between: every element is updated in a fixed order.
            mbar_wait(bar_l1, pl1);
            pl1 ^= 1u;
            fence_after();
            {
                const int d1 = lay.d[1];
                const int r0 = 32 * (warp & 3), r = r0 + lane;
                float* gW = gpart + lay.woff[0];
                float* gB = gpart + lay.boff[0];
                const int ldw = lay.ldw[0];
                float gx = 0.f, gy = 0.f, gt = 0.f, gb = 0.f;
                if (h == 0) {
                    float v0[8], v1[8], v2[8], v3[8];
                    tm_ld8(tlane + T_L1, v0);
                    if (!sec) { tm_ld8(tlane + T_L1 + 8, v1); tm_ld8(tlane + T_L1 + 16, v2); tm_ld8(tlane + T_L1 + 24, v3); }
                    tm_wait_ld();
                    gx = fmaf(v0[4], LO_INV, v0[0]); gy = fmaf(v0[5], LO_INV, v0[1]); gt = fmaf(v0[6], LO_INV, v0[2]); gb = v0[3];
                    if (!sec) { gx = fmaf(Tc.in_scale[0], v1[0], gx); gy = fmaf(Tc.in_scale[1], v2[0], gy); gt = fmaf(Tc.in_scale[2], v3[0], gt); }
                    if (r < d1) {
                        __stcg(gW + r, __ldcg(gW + r) + gx * inv_sigma);
                        __stcg(gW + ldw + r, __ldcg(gW + ldw + r) + gy * inv_sigma);
                        __stcg(gW + 2 * ldw + r, __ldcg(gW + 2 * ldw + r) + gt * inv_sigma);
                        __stcg(gB + r, __ldcg(gB + r) + gb * inv_sigma);
                    }
                }
                fence_before();
                named_bar_sync(1, F_EPI);
                if (h == 0 && r >= 56 && r < 56 + d1) {
                    const int j = r - 56;
                    const float sc = LO_INV * inv_sigma;
                    __stcg(gW + j, __ldcg(gW + j) + gx * sc);
                    __stcg(gW + ldw + j, __ldcg(gW + ldw + j) + gy * sc);
                    __stcg(gW + 2 * ldw + j, __ldcg(gW + 2 * ldw + j) + gt * sc);
                    __stcg(gB + j, __ldcg(gB + j) + gb * sc);
                }
                // (the next update of these elements is a tile away, behind the barrier at the start of the next tile)
            }
            TCF_PROF(13);
        }
        // ---- loss-term partial sums (threads with h == 0 hold them): warp reduce, then 4 warps through smem (fixed order)
        if (!FWD) {
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
    if (warp == F_CTRL) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int NS, bool PROF, bool FWD = false>
int launch_tcf(const TcfArgs& t, int slots, cudaStream_t st) {
    auto kern = resid_tcf_kernel<NS, PROF, FWD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F_TOTAL);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_tcf, %d): %s", F_TOTAL, cudaGetErrorString(e)); return 2; }
    e = pe_launch_pdl(kern, dim3(slots), dim3(F_THREADS), F_TOTAL, st, t);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_tcf<%d>: %s", NS, cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // namespace

size_t pe_tc_stash_floats_per_slot(const pe_plan* plan);

static unsigned long long* g_tcf_prof = nullptr;
extern "C" void pe_debug_set_tcf_profile(unsigned long long* d_counters32) { g_tcf_prof = d_counters32; }

// Scratch layout (d_stash of the C ABI): [slots][stash floats per slot] then the operand images.  Per slot: (L-1) layers x 5 streams x
// F_STREAM bytes (the planes are stashed as they are).
int pe_launch_resid_tcf(const pe_plan* plan, const PeResidArgs& a, int K, int fast, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2) {
    TcfArgs t;
    t.r = a;
    t.n2 = 0; t.points2 = nullptr; t.aux2 = nullptr; t.inv_n2 = 0.f;
    memset(&t.term2, 0, sizeof(t.term2));
    if (term2 && n2 > 0) {
        t.term2 = *term2; t.points2 = points2; t.n2 = n2; t.aux2 = term2->aux_k ? aux2 : nullptr;
        t.inv_n2 = 1.0f / (float)term2->n_global;
    }
    if (plan->lay.L < 3) { pe_set_error("tcf engine: needs at least two hidden layers"); return 1; }
    t.prof = g_tcf_prof;
    { const char* e = getenv("PE_PROF_CTA"); t.prof_cta = e ? atoi(e) : 0; }
    t.fields_out = nullptr; t.fields_aux_k = 0;
    t.fast = fast;
    t.r.stash_floats = (int)pe_tc_stash_floats_per_slot(plan);
    uint8_t* images = reinterpret_cast<uint8_t*>(a.stash + (size_t)slots * t.r.stash_floats);
    t.images = images;
    cudaError_t e = pe_launch_pdl(tcf_image_kernel, dim3(plan->lay.L * 16), dim3(256), 0, st, a.params, a.lay, images);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tcf_image_kernel: %s", cudaGetErrorString(e)); return 3; }
    if (K == 5) return g_tcf_prof ? launch_tcf<5, true>(t, slots, st) : launch_tcf<5, false>(t, slots, st);
    if (K == 4) return g_tcf_prof ? launch_tcf<4, true>(t, slots, st) : launch_tcf<4, false>(t, slots, st);
    pe_set_error("tcf engine: K = %d not instantiated (4 or 5)", K);
    return 1;
}

// ---- predict on the tensor-core engine: forward sweep of the four streams (value, d/dx, d/dy, d/dt) in 128-point tiles, fields written by
// the output stage.  The operand images live in a buffer owned by the plan (allocated at the first call).  Same contract as the SIMT fields
// kernel (pe_launch_fields, mode 0); hidden widths <= 56 and at least two hidden layers.
int pe_launch_fields_tcf(const pe_plan* plan, const PeFieldsArgs& a, cudaStream_t st) {
    const PeLayout& lay = plan->lay;
    pe_plan* mp = const_cast<pe_plan*>(plan);
    if (!mp->d_tc_images) {
        cudaError_t e = cudaMalloc(&mp->d_tc_images, (size_t)lay.L * F_IMG_LAYER);
        if (e != cudaSuccess) { pe_set_error("tcf fields: cudaMalloc(operand images): %s", cudaGetErrorString(e)); mp->d_tc_images = nullptr; return 2; }
    }
    TcfArgs t;
    memset(&t, 0, sizeof(t));
    t.r.lay = lay;
    t.r.term.kind = a.formulation;
    t.r.term.ld = a.ld;
    for (int i = 0; i < 3; ++i) { t.r.term.in_scale[i] = a.in_scale[i]; t.r.term.in_shift[i] = a.in_shift[i]; }
    t.r.points = a.points; t.r.aux = a.aux; t.r.params = a.params;
    t.r.n = a.n;
    t.images = static_cast<const uint8_t*>(mp->d_tc_images);
    t.fields_out = a.out; t.fields_aux_k = a.aux_k;
    cudaError_t e = pe_launch_pdl(tcf_image_kernel, dim3(lay.L * 16), dim3(256), 0, st, a.params, lay, static_cast<uint8_t*>(mp->d_tc_images));
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tcf_image_kernel: %s", cudaGetErrorString(e)); return 3; }
    const int ntiles = (a.n + TC_P - 1) / TC_P;
    const int ctas = ntiles < plan->sms ? ntiles : plan->sms;
    return launch_tcf<4, false, true>(t, ctas, st);
}
