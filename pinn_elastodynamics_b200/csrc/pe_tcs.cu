// Stream-pipelined, warp-specialised tcgen05 / TMEM engine for the collocation residual (third generation).
//
// Why: in pe_tc.cu / pe_tcp.cu the tensor pipe and the SIMT pipes strictly alternate -- thread 0 issues the ~90 MMAs of a
// layer (and blocks on the MMA queue) while 255 threads wait, then all threads run the tanh / chain-rule epilogue while the
// tensor pipe idles (profiles/r1_tc3p_phase_cycles.txt: forward MMA issue 16 %, forward epilogues 12 %).  A second tile in
// flight would hide that, but one tile already fills TMEM (320 accumulator + 160 operand columns) and shared memory.
//
// The jet streams themselves are the independent work: through a layer GEMM every stream k (value, d/dx, d/dy, d/dt, d2/dt2)
// is its own matrix product Z_k = A_k W, and the tanh epilogue couples them only through a = tanh(z_0):
//     a_x = s z_x,  a_y = s z_y,  a_t = s z_t,  a_tt = s z_tt - 2 a a_t z_t,   s = 1 - a^2        
// So the streams are split into three groups  G0 = {value},  G1 = {d/dx, d/dy},  G2 = {d/dt[, d2/dt2]}  that travel through the
// forward pass as three GEMM pipelines: while the epilogue warps apply tanh to G0 of layer l, the tensor pipe runs G1 and G2 of
// layer l; when G0's activations are written, the issuer may already start G0 of layer l+1, and so on.  No extra TMEM or
// shared memory: each group owns its accumulator columns, its activation operand planes and its bf16 "lo" operand columns.
//
// Roles (512 threads = 4 warpgroups; registers re-balanced with setmaxnreg: 136 per epilogue thread, 104 per control thread):
//   warps 0..11  epilogue / converter warps: thread (p, h) owns TMEM lane p (a point) and the unit group h = warp / 4 (chunks 0-4, 5-9, 10-13)
//   warp 12      control warp: TMEM alloc/dealloc; lane 0 issues every tcgen05.mma / tcgen05.commit and the bulk-TMA copies
//   warps 13..15 idle (they only complete the fourth warpgroup that setmaxnreg needs)
//   (TCS_EW = 8 builds the original 384-thread split: 8 epilogue warps with two unit halves, control warp 8)
// Hand-shakes (mbarriers in shared memory, phase parities tracked per role):
//   ACT[g]   (one arrival per epilogue thread)  epilogue warps wrote group g's next operand (smem fp32 plane + TMEM lo columns)  -> issuer
//   ACC[g]   (tcgen05.commit) the MMAs of group g of the current layer are complete                          -> epilogue warps
//   IMG[b]   (TMA complete_tx) weight operand image in buffer b has landed                                   -> issuer
//   FULL[X] / EMPTY[X] / DW   the weight-gradient operand pipeline of pe_tcp.cu (now fed by all converter warps)
// Weight operand images (36,864 B per matrix and direction, built per step by tcp_prep_kernel) are double buffered and
// fetched with cp.async.bulk (1-D TMA) two layers ahead, instead of being staged through 36 registers per thread.
// In the reverse sweep the weight-gradient phase needs both image buffers as scratch (bf16 hi/mid operand images), so the
// sweep keeps the phase order of pe_tcp.cu per layer (adjoint GEMM -> weight gradient -> adjoint of tanh); its adjoint
// images arrive by TMA while the previous layer's accumulators are drained.
//
// Arithmetic is operation-for-operation that of pe_tc.cu (same MMA sequence per accumulator, same epilogue expressions), so
// results are bit-identical to engines tc3 / tc3p (tests/test_gpu_tcs.py).  Reference lines: see pe_simt.cu / pe_device.cuh.
#include <cstring>
#include "pe_device.cuh"
#include "pe_tc_common.cuh"

namespace {
using namespace pe_dev;
using namespace pe_tcc;

// TCS_EW = number of epilogue / converter warps: 12 (default: three unit groups of 5 / 5 / 4 chunks per TMEM lane quadrant, i.e. three
// warps per scheduler for the latency-bound epilogues; 512-thread CTA, setmaxnreg 136 per epilogue thread / 104 per control thread --
// the only split of the 64 K registers that leaves the issuer thread unspilled) or 8 (two unit halves of 7 chunks, 384 threads, 200 / 104).
// Same arithmetic per element either way: bit-identical results.  Measured on B200 (profiles/r1_tc3s_ew12_ab.log): F5 0.3743 -> 0.3651
// ms/step, F7 1.540 -> 1.510 ms/step with 12 warps.
#ifndef TCS_EW
#define TCS_EW 12
#endif
static_assert(TCS_EW == 8 || TCS_EW == 12, "TCS_EW: 8 or 12 epilogue warps");
constexpr int S_EPI = 32 * TCS_EW, S_THREADS = S_EPI + 128;  // + the control warpgroup (issuer warp + three idle warps)
constexpr int S_NH = TCS_EW / 4;                             // unit groups per lane quadrant
constexpr int S_CTRL = TCS_EW;                               // index of the control warp
constexpr int S_ACT = 0;
constexpr int S_IMG0 = TC_MAX_STREAMS * TC_ACT_STREAM;       // 144,480: image buffer 0; weight-gradient phase: bf16 hi/mid images of A
constexpr int S_IMG1 = S_IMG0 + TC_IMG_SET;                  // 181,344: image buffer 1; weight-gradient phase: bf16 hi/mid images of Zbar
constexpr int S_MISC = S_IMG1 + TC_IMG_SET;                  // 218,208: mbarriers + TMEM base slot
constexpr int S_COORD = S_MISC + 256;                        // 128 x 4 floats
constexpr int S_RED = S_COORD + 128 * 16;                    // scratch (layer-1 gradient) / term sums
constexpr int S_BIAS = S_RED + 4096;                         // [PE_MAX_LAYERS][64] floats: all biases, loaded once per launch
constexpr int S_W0 = S_BIAS + PE_MAX_LAYERS * 256;           // [4][64] floats: first-layer weight rows 0..2 and bias
constexpr int S_TOTAL = S_W0 + 1024;                         // 229,728
static_assert(S_TOTAL <= 227 * 1024, "shared memory map exceeds the 227 KB opt-in limit");

constexpr int B_ACC = 0, B_ACT = 24, B_IMG = 48, B_FULL = 64, B_EMPTY = 80, B_DW = 96, B_TMEM = 112;   // byte offsets in S_MISC
constexpr int SW_AHI = S_IMG0, SW_AMID = S_IMG0 + 16384;     // [8 chunks][128 points][8 bf16]
constexpr int SW_ZHI = S_IMG1, SW_ZMID = S_IMG1 + 14336;     // [7 chunks][128 points][8 bf16], contiguous (N = 56 + NZ trick)

template <int NS> __device__ __forceinline__ int grp_first(int g) { return g == 0 ? 0 : (g == 1 ? 1 : 3); }
template <int NS> __device__ __forceinline__ int grp_count(int g) { return g == 0 ? 1 : (g == 1 ? 2 : NS - 3); }

struct TcsArgs {
    PeResidArgs r;
    const uint8_t* images;
    int fast;
    pe_term_desc term2;
    const float* points2;
    const float* aux2;
    int n2;
    float inv_n2;
    unsigned long long* prof;   // PROF instantiation: 32 cycle counters (0..15 epilogue thread 0, 16..31 issuer) of CTA 0
};

// phase timing (PROF instantiations only): the calling thread adds the cycles since its previous mark to counter `slot`
#define TCS_PROF(slot) do { if (PROF) { if (prof_on) { const long long now_ = clock64(); atomicAdd(args.prof + (slot), (unsigned long long)(now_ - prof_t)); prof_t = now_; } } } while (0)

// MMAs of one layer GEMM for streams [k0, k0 + nk): K-steps interleaved across the streams of the call.  The K-step loops are
// deliberately not unrolled: the issuing thread lives in the register-starved control warpgroup, and the few integer adds per
// descriptor disappear behind the ~80 cycles each MMA occupies the tensor pipe.
__device__ __forceinline__ void issue_streams(int k0, int nk, uint32_t tbase, uint32_t act_s, uint32_t img_s, int N, int ksteps, int kb, int fast) {
    const uint32_t id32 = idesc_tf32(N), id16 = idesc_bf16(N);
    const uint32_t nrow = (uint32_t)N * 16u;
    const uint64_t a_step = (uint64_t)((2u * TC_CH) >> 4), b_step = (uint64_t)((2u * nrow) >> 4);
    const uint64_t a_stream = (uint64_t)(TC_ACT_STREAM >> 4);
    const uint64_t a0 = sdesc(act_s, TC_CH, 128) + (uint64_t)k0 * a_stream;
    const uint32_t d0 = tbase + TM_ACC + 64u * k0;
    const uint32_t l0 = tbase + TM_LO + 32u * k0;
    {
        uint64_t a = a0, b = sdesc(img_s + TC_IMG_HI, nrow, 128);
#pragma unroll 1
        for (int s = 0; s < ksteps; ++s, a += a_step, b += b_step) {
#pragma unroll
            for (int k = 0; k < TC_MAX_STREAMS; ++k)
                if (k < nk) mma_tf32_ss(d0 + 64u * k, a + k * a_stream, b, id32, s > 0);
        }
    }
    if (!fast) {
        uint64_t a = a0, b = sdesc(img_s + TC_IMG_LO, nrow, 128);
#pragma unroll 1
        for (int s = 0; s < ksteps; ++s, a += a_step, b += b_step) {
#pragma unroll
            for (int k = 0; k < TC_MAX_STREAMS; ++k)
                if (k < nk) mma_tf32_ss(d0 + 64u * k, a + k * a_stream, b, id32, 1);
        }
        b = sdesc(img_s + TC_IMG_BF, nrow, 128);
#pragma unroll 1
        for (int s = 0; s < kb; ++s, b += b_step) {
#pragma unroll
            for (int k = 0; k < TC_MAX_STREAMS; ++k)
                if (k < nk) mma_bf16_ts(d0 + 64u * k, l0 + 32u * k + 8u * s, b, id16, 1);
        }
    }
}

template <int NS, bool PROF>
__global__ void __launch_bounds__(S_THREADS, 1) resid_tcs_kernel(const TcsArgs args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STASH_LAYER = NS * TC_STASH_STREAM;          // bytes per stashed layer
    const PeResidArgs& A = args.r;
    const PeLayout& lay = A.lay;
    const pe_term_desc& T = A.term;
    const pe_term_desc& T2 = args.term2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = lay.L;
    const int fast = args.fast;
    const bool prof_on = PROF && args.prof != nullptr && blockIdx.x == 0 && (tid == 0 || tid == S_EPI);
    long long prof_t = 0;
    (void)prof_on; (void)prof_t;
    uint8_t* act = smem + S_ACT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S_MISC + B_TMEM);
    float* coord = reinterpret_cast<float*>(smem + S_COORD);
    float* red = reinterpret_cast<float*>(smem + S_RED);
    float* sbias = reinterpret_cast<float*>(smem + S_BIAS);     // sbias[m * 64 + j] = bias j of matrix m
    float* sw0 = reinterpret_cast<float*>(smem + S_W0);
    const uint32_t act_s = smem_u32(act), img_s0 = smem_u32(smem + S_IMG0), bar0 = smem_u32(smem + S_MISC);
    const uint32_t bar_acc = bar0 + B_ACC, bar_act = bar0 + B_ACT, bar_img = bar0 + B_IMG;
    const uint32_t bar_full = bar0 + B_FULL, bar_empty = bar0 + B_EMPTY, bar_dw = bar0 + B_DW;

    for (int i = tid; i < S_TOTAL / 16; i += S_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int slot = A.slot_base + blockIdx.x;
    float* gpart = A.grad_partials + (size_t)slot * lay.total;
    float* stash = A.stash + (size_t)blockIdx.x * A.stash_floats;
    const float* __restrict__ params = A.params;
    for (int i = tid; i < lay.total; i += S_THREADS) __stcg(gpart + i, 0.f);
    __syncthreads();
    if (tid < 64) {                                  // first-layer weights (3 x d1) and bias -> smem, once per launch
        const bool in = tid < lay.d[1];
        sw0[tid] = in ? __ldg(params + lay.woff[0] + tid) : 0.f;
        sw0[64 + tid] = in ? __ldg(params + lay.woff[0] + lay.ldw[0] + tid) : 0.f;
        sw0[128 + tid] = in ? __ldg(params + lay.woff[0] + 2 * lay.ldw[0] + tid) : 0.f;
        sw0[192 + tid] = in ? __ldg(params + lay.boff[0] + tid) : 0.f;
    }
    for (int i = tid; i < L * 64; i += S_THREADS) {  // all biases
        const int m = i >> 6, j = i & 63;
        sbias[i] = (j < lay.d[m + 1]) ? __ldg(params + lay.boff[m] + j) : 0.f;
    }
    if (tid == 0) {
        for (int g = 0; g < 3; ++g) { mbar_init(bar_acc + 8 * g, 1); mbar_init(bar_act + 8 * g, S_EPI); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_img + 8 * b, 1); mbar_init(bar_full + 8 * b, S_EPI); mbar_init(bar_empty + 8 * b, 1); }
        mbar_init(bar_dw, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == S_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();                              // zero-filled image / operand regions before any async-proxy access
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *tmem_slot;
    const int ntiles_main = (A.n + TC_P - 1) / TC_P;
    const int ntiles = ntiles_main + (args.n2 + TC_P - 1) / TC_P;
    // image index i = 2..L is the forward image of matrix i-1, i = L+1 the adjoint image of matrix L-1; buffer (L+1-i)&1, so
    // that the first adjoint image always lands in buffer 0 (buffer 1 is converter scratch from the start of the reverse sweep)
    auto img_buf = [&](int i) { return (L + 1 - i) & 1; };
    auto img_src = [&](int i) { return (i <= L) ? args.images + (size_t)(i - 1) * TC_IMG_LAYER : args.images + (size_t)(L - 1) * TC_IMG_LAYER + TC_IMG_SET; };

    if (warp >= S_CTRL) {
        // ============================================================================================ control warpgroup
#if TCS_EW == 8
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
#else
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
#endif
        if (warp == S_CTRL && lane == 0) {
            uint32_t pact = 0, pimg = 0, pfull = 0;           // parity bits of the phases this thread waits for next
            uint32_t n_acc2 = 0, n_dw = 0;                   // commits issued so far on ACC[2] / DW (their completion parity = count & 1)
            auto load_img = [&](int i) {
                const uint32_t b = (uint32_t)img_buf(i);
                mbar_expect_tx(bar_img + 8 * b, TC_IMG_SET);
                tma_load_1d(img_s0 + b * TC_IMG_SET, img_src(i), TC_IMG_SET, bar_img + 8 * b);
            };
            auto wait_img = [&](uint32_t b) { mbar_wait(bar_img + 8 * b, (pimg >> b) & 1u); pimg ^= 1u << b; };
            auto wait_act = [&](int g) { mbar_wait(bar_act + 8 * g, (pact >> g) & 1u); pact ^= 1u << g; };
            load_img(2);
            load_img(3);
            if (PROF) prof_t = clock64();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // ---------------------------------------------------------------- forward: layers 2..L, group by group
                for (int l = 2; l <= L; ++l) {
                    const uint32_t b = (uint32_t)img_buf(l);
                    const int NF = (lay.d[l] <= 16) ? 16 : 64;
                    const int ksteps = (lay.d[l - 1] + 7) >> 3, kb = (lay.d[l - 1] + 15) >> 4;
                    wait_img(b);
                    TCS_PROF(16);
#pragma unroll 1
                    for (int g = 0; g < 3; ++g) {
                        wait_act(g);
                        TCS_PROF(17 + 2 * g);
                        fence_after();
                        issue_streams(grp_first<NS>(g), grp_count<NS>(g), tbase, act_s, img_s0 + b * TC_IMG_SET, NF, ksteps, kb, fast);
                        mma_commit(bar_acc + 8 * g);
                        TCS_PROF(18 + 2 * g);
                        if (g == 2) ++n_acc2;
                        if (g == 0 && l >= 3) {
                            // every MMA of layer l-1 precedes G0 of layer l in the pipe: once its last group is complete, its image
                            // buffer (= the buffer of image l+1) is free
                            mbar_wait(bar_acc + 16, (n_acc2 - 1) & 1u);
                            load_img(l + 1);
                            TCS_PROF(23);
                        }
                    }
                }
                // ---------------------------------------------------------------- reverse sweep: layers L..2
                for (int l = L; l >= 2; --l) {
                    const int dout = lay.d[l];
                    wait_img(0);
                    TCS_PROF(24);
                    wait_act(0); wait_act(1); wait_act(2);
                    TCS_PROF(25);
                    fence_after();
                    issue_streams(0, NS, tbase, act_s, img_s0, 64, (dout + 7) >> 3, (dout + 15) >> 4, fast);
                    mma_commit(bar_acc);
                    mma_commit(bar_acc + 8);
                    mma_commit(bar_acc + 16);
                    ++n_acc2;
                    TCS_PROF(26);
                    // weight gradient: consumer side of the operand pipeline (see pe_tcp.cu)
                    const int NZ = (dout + 7) & ~7;
                    const uint32_t id2 = idesc_bf16_mn(64, 56 + NZ), id1 = idesc_bf16_mn(64, NZ);
                    const uint32_t d = tbase + TM_LO;
                    const uint64_t ahi = sdesc(smem_u32(smem + SW_AHI), 128, 2048), amid = sdesc(smem_u32(smem + SW_AMID), 128, 2048);
                    const uint64_t zhi = sdesc(smem_u32(smem + SW_ZHI), 128, 2048);
#pragma unroll 1
                    for (int k = 0; k < NS; ++k) {
#pragma unroll
                        for (int X = 0; X < 2; ++X) {
                            mbar_wait(bar_full + 8 * X, (pfull >> X) & 1u);
                            pfull ^= 1u << X;
                            fence_after();
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const int s8 = 4 * X + s;
                                const uint64_t o = (uint64_t)(s8 * 16);
                                mma_bf16_ss(d, ahi + o, zhi + o, id2, (k > 0 || s8 > 0) ? 1u : 0u);
                                mma_bf16_ss(d, amid + o, zhi + o, id1, 1u);
                            }
                            if (k < NS - 1) mma_commit(bar_empty + 8 * X);
                        }
                    }
                    mma_commit(bar_dw);
                    ++n_dw;
                    TCS_PROF(27);
                    mbar_wait(bar_dw, (n_dw - 1) & 1u);      // both buffers are free again
                    TCS_PROF(28);
                    if (l > 2) {
                        mbar_expect_tx(bar_img, TC_IMG_SET);
                        tma_load_1d(img_s0, args.images + (size_t)(l - 2) * TC_IMG_LAYER + TC_IMG_SET, TC_IMG_SET, bar_img);
                    } else if (tile + (int)gridDim.x < ntiles) {
                        load_img(2);
                        load_img(3);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ============================================================================================ epilogue / converter warps
#if TCS_EW == 8
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
#else
        asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
#endif
        const int p = 32 * (warp & 3) + lane;            // TMEM lane = point of the tile
        const int h = warp >> 2;                         // unit group: chunks [c_lo, c_hi) = [7h, 7h+7) (two groups) or [5h, min(5h+5, 14)) (three)
        const int c_lo = (S_NH == 2) ? 7 * h : 5 * h;
        const int c_hi = (S_NH == 2) ? 7 * h + 7 : (h == 2 ? TC_NCH : 5 * h + 5);
        const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
        // zero the bf16 "lo" operand columns once (units 56..63 are never written afterwards and must stay zero)
        for (int c = 2 * h; c < 160; c += 2 * S_NH) tm_st2(tlane + TM_LO + c, 0u, 0u);
        tm_wait_st();
        uint32_t pacc = 0, pempty = 0, pdw = 0;
        auto wait_acc = [&](int g) { mbar_wait(bar_acc + 8 * g, (pacc >> g) & 1u); pacc ^= 1u << g; };
        // operands of group g written: TMEM stores retired, smem stores visible to the async proxy, then release-arrive
        auto publish = [&](int g) { mbar_arrive(bar_act + 8 * g); };
        auto publish_fences = [&]() { tm_wait_st(); fence_async_smem(); fence_before(); };
        float tsum[PE_MAX_TERMS], tsum2[PE_MAX_TERMS];
#pragma unroll
        for (int i = 0; i < PE_MAX_TERMS; ++i) { tsum[i] = 0.f; tsum2[i] = 0.f; }

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
                *reinterpret_cast<float4*>(coord + 4 * p) = make_float4(fmaf(x, Tc.in_scale[0], Tc.in_shift[0]), fmaf(y, Tc.in_scale[1], Tc.in_shift[1]),
                                                                        fmaf(t, Tc.in_scale[2], Tc.in_shift[2]), valid ? 1.f : 0.f);
            } else if (h == 1) {
                const int nt = tile + (int)gridDim.x;                      // pull this CTA's next tile towards L2
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
            named_bar_sync(1, S_EPI);
            // ================================================================ layer 1 (3 -> d1): per-thread FFMA, all streams
            {
                const float4 c4 = *reinterpret_cast<const float4*>(coord + 4 * p);
                const int dout = lay.d[1];
                float* st = stash;                                         // stash layer index 0 = outputs of layer 1
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
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
                publish_fences();
                publish(0); publish(1); publish(2);
                TCS_PROF(0);
            }
            // ================================================================ forward: hidden layers 2..L-1, one stream group at a time
            for (int l = 2; l < L; ++l) {
                const int dout = lay.d[l];
                const float* bl = sbias + (l - 1) * 64;
                float* st = stash + (size_t)(l - 1) * (STASH_LAYER / 4);
                auto store_stream = [&](int k, int c, const float (&v4)[4]) {
                    const float4 v = make_float4(v4[0], v4[1], v4[2], v4[3]);
                    *reinterpret_cast<float4*>(act + k * TC_ACT_STREAM + c * TC_CH + p * 16) = v;
                    __stcg(reinterpret_cast<float4*>(st + (size_t)k * (TC_STASH_STREAM / 4) + c * 512 + p * 4), v);
                    tm_st2(tlane + TM_LO + 32 * k + 2 * c, lo_pair(v4[0], v4[1]), lo_pair(v4[2], v4[3]));
                };
                auto zero_pads = [&](int k) { if (h == S_NH - 1) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); } };
                // ---- G0: a = tanh(z_0 + b)
                wait_acc(0);
                TCS_PROF(1);
                fence_after();
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float z[4];
                    tm_ld4(tlane + TM_ACC + 4 * c, z);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c + u;
                        z[u] = (j < dout) ? tanh_branchfree(z[u] + bl[j]) : 0.f;
                    }
                    store_stream(0, c, z);
                }
                zero_pads(0);
                publish_fences();
                publish(0);
                TCS_PROF(2);
                // ---- G1: a_x = s z_x, a_y = s z_y   (a re-read from this thread's own entries of the value plane)
                wait_acc(1);
                TCS_PROF(3);
                fence_after();
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    const float4 a4 = *reinterpret_cast<const float4*>(act + c * TC_CH + p * 16);
                    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                    float z1[4], z2[4];
                    tm_ld4(tlane + TM_ACC + 64 + 4 * c, z1);
                    tm_ld4(tlane + TM_ACC + 128 + 4 * c, z2);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c + u;
                        const float s = fmaf(-av[u], av[u], 1.f);
                        z1[u] = (j < dout) ? s * z1[u] : 0.f;
                        z2[u] = (j < dout) ? s * z2[u] : 0.f;
                    }
                    store_stream(1, c, z1);
                    store_stream(2, c, z2);
                }
                zero_pads(1); zero_pads(2);
                publish_fences();
                publish(1);
                TCS_PROF(4);
                // ---- G2: a_t = s z_t [, a_tt = s z_tt - 2 a a_t z_t]
                wait_acc(2);
                TCS_PROF(5);
                fence_after();
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    const float4 a4 = *reinterpret_cast<const float4*>(act + c * TC_CH + p * 16);
                    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                    float z3[4], z4[4];
                    tm_ld4(tlane + TM_ACC + 192 + 4 * c, z3);
                    if (NS == 5) tm_ld4(tlane + TM_ACC + 256 + 4 * c, z4);
                    tm_wait_ld();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * c + u;
                        const float a = av[u];
                        const float s = fmaf(-a, a, 1.f);
                        const float zt = z3[u];
                        const float at = s * zt;
                        z3[u] = (j < dout) ? at : 0.f;
                        if (NS == 5) z4[u] = (j < dout) ? fmaf(s, z4[u], -2.f * a * at * zt) : 0.f;
                    }
                    store_stream(3, c, z3);
                    if (NS == 5) store_stream(4, c, z4);
                }
                zero_pads(3);
                if (NS == 5) zero_pads(4);
                publish_fences();
                publish(2);
                TCS_PROF(6);
            }
            // ================================================================ output layer L: residuals, loss partials, seeds
            {
                const int dout = lay.d[L];
                const float* bl = sbias + (L - 1) * 64;
                wait_acc(0); wait_acc(1); wait_acc(2);
                TCS_PROF(7);
                fence_after();
                if (h == 0) {
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
                publish_fences();
                publish(0); publish(1); publish(2);
                TCS_PROF(8);
            }
            // ================================================================ reverse sweep, layers L .. 2
            for (int l = L; l >= 2; --l) {
                const int m = l - 1;
                const int din = lay.d[l - 1], dout = lay.d[l];
                const float* stash_in = stash + (size_t)(l - 2) * (STASH_LAYER / 4);     // outputs of layer l-1 = inputs A of layer l
                const int NZ = (dout + 7) & ~7;
                const int zc8 = NZ >> 3;
                named_bar_sync(1, S_EPI);                    // Zbar_l (written by other threads of these warps) is complete
                float4 pre[2][2];
                auto ldA = [&](int k, int X) {               // stash (L2) -> registers: tasks (c8 < 7, 64 points)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int t = tid + i * S_EPI;
                        if (t < 7 * 64) {
                            const int c8 = t >> 6, pp = 64 * X + (t & 63);
                            const float* src = stash_in + (size_t)k * (TC_STASH_STREAM / 4) + (2 * c8) * 512 + pp * 4;
                            pre[i][0] = __ldcg(reinterpret_cast<const float4*>(src));
                            pre[i][1] = __ldcg(reinterpret_cast<const float4*>(src + 512));
                        }
                    }
                };
                auto stA = [&](int k, int X) {               // registers -> bf16 hi/mid images of A_k, half X
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int t = tid + i * S_EPI;
                        if (t < 7 * 64) {
                            const int c8 = t >> 6, pp = 64 * X + (t & 63);
                            uint4 hi, mid;
                            split8(pre[i][0], pre[i][1], hi, mid);
                            *reinterpret_cast<uint4*>(smem + SW_AHI + c8 * 2048 + pp * 16) = hi;
                            *reinterpret_cast<uint4*>(smem + SW_AMID + c8 * 2048 + pp * 16) = mid;
                        }
                    }
                    if (tid >= S_EPI - 64) {                 // chunk 7: units 56..63, unit 63 = ones row of the value stream (threads with one task)
                        const int pp = 64 * X + (tid - (S_EPI - 64));
                        const uint32_t one_hi = (k == 0) ? 0x3F800000u : 0u;
                        *reinterpret_cast<uint4*>(smem + SW_AHI + 7 * 2048 + pp * 16) = make_uint4(0u, 0u, 0u, one_hi);
                        *reinterpret_cast<uint4*>(smem + SW_AMID + 7 * 2048 + pp * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                };
                auto cvZ = [&](int k, int X) {               // ACT[k] (fp32 smem) -> bf16 hi/mid images of Zbar_k, half X
                    for (int t = tid; t < zc8 * 64; t += S_EPI) {
                        const int c8 = t >> 6, pp = 64 * X + (t & 63);
                        const uint8_t* src = act + k * TC_ACT_STREAM + (2 * c8) * TC_CH + pp * 16;
                        uint4 hi, mid;
                        split8(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + TC_CH), hi, mid);
                        *reinterpret_cast<uint4*>(smem + SW_ZHI + c8 * 2048 + pp * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + SW_ZMID + c8 * 2048 + pp * 16) = mid;
                    }
                };
                ldA(0, 0);
                if (l >= 3) {   // pull the stash layer of the next (shallower) iteration towards L2
                    const char* nxt = reinterpret_cast<const char*>(stash + (size_t)(l - 3) * (STASH_LAYER / 4));
                    for (int i = tid; i < STASH_LAYER / 128; i += S_EPI)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)i * 128));
                }
                cvZ(0, 0);                                   // buffer 1 is free while the adjoint MMAs read buffer 0 / ACT / LO
                cvZ(0, 1);
                TCS_PROF(9);
                wait_acc(0); wait_acc(1); wait_acc(2);       // adjoint MMAs done: buffer 0 and the LO columns are free now
                TCS_PROF(10);
                fence_after();
#pragma unroll 1
                for (int k = 0; k < NS; ++k) {
#pragma unroll
                    for (int X = 0; X < 2; ++X) {
                        if (k > 0) {
                            mbar_wait(bar_empty + 8 * X, (pempty >> X) & 1u);
                            pempty ^= 1u << X;
                        }
                        stA(k, X);
                        if (X == 0) ldA(k, 1);
                        else if (k + 1 < NS) ldA(k + 1, 0);
                        if (k > 0) cvZ(k, X);
                        fence_async_smem();
                        mbar_arrive(bar_full + 8 * X);
                    }
                }
                TCS_PROF(11);
                mbar_wait(bar_dw, pdw);
                pdw ^= 1u;
                TCS_PROF(12);
                fence_after();
                {   // drain the dW tile: rows i = 16*quadrant + lane (lane < 16), row 63 = bias gradient; h selects the column half
                    const int quad = warp & 3;
                    const int i = 16 * quad + lane;
                    const int ldw = lay.ldw[m];
                    float* gW = gpart + lay.woff[m];
                    float* gB = gpart + lay.boff[m];
                    const int dr_lo = (S_NH == 2) ? (h ? 32 : 0) : (h == 0 ? 0 : (h == 1 ? 24 : 40));
                    const int dr_hi = (S_NH == 2) ? (h ? 56 : 32) : (h == 0 ? 24 : (h == 1 ? 40 : 56));
                    for (int c = dr_lo; c < dr_hi; c += 8) {
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
                }
                // the unit groups (warps w, w + 4[, w + 8]) share the TMEM lanes: the tile must be drained by all of them before any of them
                // overwrites the aliased lo-operand columns below
                named_bar_sync(1, S_EPI);
                if (h == 0) {   // the tile aliased lo-operand columns including zero pads (units 56..63): restore them
#pragma unroll
                    for (int k = 0; k < 4; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
                TCS_PROF(13);
                // ---- through tanh of layer l-1: zbar^{l-1} from abar^{l-1} (TMEM) and the stashed outputs
                float4 Anext[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) Anext[k] = __ldcg(reinterpret_cast<const float4*>(stash_in + (size_t)k * (TC_STASH_STREAM / 4) + c_lo * 512 + p * 4));
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float ab[NS][4];
                    float4 Av[NS];
#pragma unroll
                    for (int k = 0; k < NS; ++k) Av[k] = Anext[k];
                    if (c + 1 < c_hi) {
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
                if (h == S_NH - 1) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) { tm_st2(tlane + TM_LO + 32 * k + 28, 0u, 0u); tm_st2(tlane + TM_LO + 32 * k + 30, 0u, 0u); }
                }
                if (l > 2) {
                    publish_fences();
                    publish(0); publish(1); publish(2);
                }
                TCS_PROF(14);
            }
            // ================================================================ layer 1 gradient (3 x d1 + bias): FFMA, fixed-order reduce
            tm_wait_st();
            named_bar_sync(1, S_EPI);
            {
                const int d1 = lay.d[1];
                const int j = tid & 63, qq = tid >> 6;                    // 4 point quarters x 64 units
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, gb = 0.f;
                if (j < d1 && tid < 256) {
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
                if (tid < 256) *reinterpret_cast<float4*>(red + (qq * 64 + j) * 4) = make_float4(g0, g1, g2, gb);
                named_bar_sync(1, S_EPI);
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
                named_bar_sync(1, S_EPI);
            }
            TCS_PROF(15);
        }
        // ---- loss-term partial sums (threads with h == 0 hold them): warp reduce, then 4 warps through smem (fixed order)
        {
            float tot[2 + PE_MAX_TERMS];
            tot[0] = warp_sum(tsum[0]);
            tot[1] = warp_sum(tsum[1]);
#pragma unroll
            for (int c = 0; c < PE_MAX_TERMS; ++c) tot[2 + c] = warp_sum(tsum2[c]);
            named_bar_sync(1, S_EPI);
            if (h == 0 && lane == 0) {
#pragma unroll
                for (int c = 0; c < 2 + PE_MAX_TERMS; ++c) red[(2 + PE_MAX_TERMS) * warp + c] = tot[c];
            }
            named_bar_sync(1, S_EPI);
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
    if (warp == S_CTRL) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int NS, bool PROF>
int launch_tcs(const TcsArgs& t, int slots, cudaStream_t st) {
    auto kern = resid_tcs_kernel<NS, PROF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL);
    if (e != cudaSuccess) { pe_set_error("cudaFuncSetAttribute(resid_tcs, %d): %s", S_TOTAL, cudaGetErrorString(e)); return 2; }
    kern<<<slots, S_THREADS, S_TOTAL, st>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("launch resid_tcs<%d>: %s", NS, cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // namespace

size_t pe_tc_stash_floats_per_slot(const pe_plan* plan);

static unsigned long long* g_tcs_prof = nullptr;
extern "C" void pe_debug_set_tcs_profile(unsigned long long* d_counters32) { g_tcs_prof = d_counters32; }

int pe_launch_resid_tcs(const pe_plan* plan, const PeResidArgs& a, int K, int fast, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2) {
    TcsArgs t;
    t.r = a;
    t.n2 = 0; t.points2 = nullptr; t.aux2 = nullptr; t.inv_n2 = 0.f;
    memset(&t.term2, 0, sizeof(t.term2));
    if (term2 && n2 > 0) {
        t.term2 = *term2; t.points2 = points2; t.n2 = n2; t.aux2 = term2->aux_k ? aux2 : nullptr;
        t.inv_n2 = 1.0f / (float)term2->n_global;
    }
    t.fast = fast;
    t.prof = g_tcs_prof;
    t.r.stash_floats = (int)pe_tc_stash_floats_per_slot(plan);
    uint8_t* images = reinterpret_cast<uint8_t*>(a.stash + (size_t)slots * t.r.stash_floats);
    t.images = images;
    tcp_prep_kernel<<<plan->lay.L * 16, 256, 0, st>>>(a.params, a.lay, images);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("tcp_prep_kernel: %s", cudaGetErrorString(e)); return 3; }
    if (K == 5) return g_tcs_prof ? launch_tcs<5, true>(t, slots, st) : launch_tcs<5, false>(t, slots, st);
    if (K == 4) return g_tcs_prof ? launch_tcs<4, true>(t, slots, st) : launch_tcs<4, false>(t, slots, st);
    pe_set_error("stream-pipelined tensor-core engine: K = %d not instantiated (4 or 5)", K);
    return 1;
}
