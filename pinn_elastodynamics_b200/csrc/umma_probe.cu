// UMMA probe: a tiny harness that runs host-described tcgen05.mma sequences on host-built shared-memory /
// tensor-memory images and dumps the accumulator columns.  Used by tests/test_gpu_umma_probe.py to pin the
// descriptor / layout conventions the tensor-core engine (pe_tc.cu) relies on (K-major and MN-major no-swizzle
// operands, M=64 vs M=128 accumulator lane maps, A-from-TMEM bf16 packing, mixed-kind accumulation) against
// exact integer GEMMs.  Test infrastructure for the kernels, not part of the hot path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

struct ProbeArgs {
    const uint8_t* smem_image; int smem_bytes;
    const uint32_t* tmem_image; int tmem_img_cols; int tmem_img_col0;
    int n_mma;
    const uint64_t* a_desc; const uint64_t* b_desc;   // kind >= 2: a_desc low 32 bits = TMEM column of A
    const uint32_t* idesc; const uint32_t* d_col; const uint32_t* accum; const uint32_t* kind;
    int out_cols; uint32_t* out; uint32_t sentinel;
    long long* cycles;       // optional: SM cycles from the first MMA issue to the completion of the commit (thread 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    // the MMA list is staged in shared memory first, so that the issue loop is not paced by global loads (the cycle count is then the
    // tensor pipe's time for the sequence, as long as issuing one MMA takes less time than executing it)
    __shared__ uint64_t s_a[256], s_b[256];
    __shared__ uint32_t s_i[256], s_d[256], s_acc[256], s_k[256];
    for (int i = threadIdx.x; i < a.n_mma && i < 256; i += 128) {
        s_a[i] = a.a_desc[i]; s_b[i] = a.b_desc[i]; s_i[i] = a.idesc[i]; s_d[i] = a.d_col[i]; s_acc[i] = a.accum[i]; s_k[i] = a.kind[i];
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a.smem_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(a.smem_image)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    // sentinel fill of the dumped columns
    for (int c = 0; c < a.out_cols; c += 8) {
        uint32_t s = a.sentinel;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_addr + c), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
    }
    // optional TMEM image (A operand from tensor memory): [128 lanes][tmem_img_cols]
    for (int c = 0; c < a.tmem_img_cols; c += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = a.tmem_image[(size_t)tid * a.tmem_img_cols + c + j];
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_addr + a.tmem_img_col0 + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t_start = 0;
    if (tid == 0) {
        const uint64_t base16 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
        t_start = clock64();
        for (int i = 0; i < a.n_mma; ++i) {
            uint64_t da = s_a[i], db = s_b[i];
            db = (db & ~0x3FFFull) | (((db & 0x3FFF) + base16) & 0x3FFF);
            uint32_t d = tbase + s_d[i], id = s_i[i], acc = s_acc[i], kind = s_k[i];
            if (kind < 2 || kind == 6) da = (da & ~0x3FFFull) | (((da & 0x3FFF) + base16) & 0x3FFF);
            if (kind == 4 || kind == 5) db = 0;
            uint32_t ta = tbase + (uint32_t)(da & 0xFFFFFFFFu);
            if (kind == 0)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
            else if (kind == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
            else if (kind == 2)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(ta), "l"(db), "r"(id), "r"(acc) : "memory");
            else if (kind == 3)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(ta), "l"(db), "r"(id), "r"(acc) : "memory");
            else if (kind == 6)        // kind::f16 SS with scale-input-d: D = A B + D * 2^-11
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n}" ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
            else if (kind == 4) {      // smem -> TMEM copy, 128 lanes x 256 bit; a_desc = smem matrix descriptor, d_col = target column
                uint64_t dc = (s_a[i] & ~0x3FFFull) | (((s_a[i] & 0x3FFF) + base16) & 0x3FFF);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(d), "l"(dc) : "memory");
            } else {                   // 128 lanes x 128 bit
                uint64_t dc = (s_a[i] & ~0x3FFFull) | (((s_a[i] & 0x3FFF) + base16) & 0x3FFF);
                asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(d), "l"(dc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // everyone waits for the MMAs (phase 0)
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    if (tid == 0 && a.cycles) *a.cycles = clock64() - t_start;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < a.out_cols; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(lane_addr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) a.out[(size_t)tid * a.out_cols + c + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, 256, "%s: %s", #x, cudaGetErrorString(e_)); return 1; } } while (0)
static char err[256];

static long long g_last_cycles = 0;
extern "C" const char* umma_probe_error(void) { return err; }
extern "C" long long umma_probe_last_cycles(void) { return g_last_cycles; }

// All pointers are HOST pointers; returns 0 on success.  out: [128][out_cols] uint32.
extern "C" int umma_probe(const void* smem_image, int smem_bytes, const uint32_t* tmem_image, int tmem_img_cols, int tmem_img_col0,
                          int n_mma, const uint64_t* a_desc, const uint64_t* b_desc, const uint32_t* idesc, const uint32_t* d_col,
                          const uint32_t* accum, const uint32_t* kind, int out_cols, uint32_t sentinel, uint32_t* out) {
    ProbeArgs a{};
    void *d_s = nullptr, *d_t = nullptr, *d_a, *d_b, *d_i, *d_c, *d_acc, *d_k, *d_o;
    CK(cudaMalloc(&d_s, smem_bytes)); CK(cudaMemcpy(d_s, smem_image, smem_bytes, cudaMemcpyHostToDevice));
    if (tmem_img_cols > 0) { CK(cudaMalloc(&d_t, 128 * tmem_img_cols * 4)); CK(cudaMemcpy(d_t, tmem_image, 128 * tmem_img_cols * 4, cudaMemcpyHostToDevice)); }
    CK(cudaMalloc(&d_a, n_mma * 8)); CK(cudaMemcpy(d_a, a_desc, n_mma * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_b, n_mma * 8)); CK(cudaMemcpy(d_b, b_desc, n_mma * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_i, n_mma * 4)); CK(cudaMemcpy(d_i, idesc, n_mma * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_c, n_mma * 4)); CK(cudaMemcpy(d_c, d_col, n_mma * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_acc, n_mma * 4)); CK(cudaMemcpy(d_acc, accum, n_mma * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_k, n_mma * 4)); CK(cudaMemcpy(d_k, kind, n_mma * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_o, 128 * out_cols * 4));
    void* d_cyc = nullptr;
    CK(cudaMalloc(&d_cyc, 8));
    a.cycles = (long long*)d_cyc;
    a.smem_image = (const uint8_t*)d_s; a.smem_bytes = smem_bytes;
    a.tmem_image = (const uint32_t*)d_t; a.tmem_img_cols = tmem_img_cols; a.tmem_img_col0 = tmem_img_col0;
    a.n_mma = n_mma; a.a_desc = (const uint64_t*)d_a; a.b_desc = (const uint64_t*)d_b; a.idesc = (const uint32_t*)d_i;
    a.d_col = (const uint32_t*)d_c; a.accum = (const uint32_t*)d_acc; a.kind = (const uint32_t*)d_k;
    a.out_cols = out_cols; a.out = (uint32_t*)d_o; a.sentinel = sentinel;
    CK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    umma_probe_kernel<<<1, 128, smem_bytes + 1024, 0>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d_o, 128 * out_cols * 4, cudaMemcpyDeviceToHost));
    { long long cyc = 0; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); g_last_cycles = cyc; cudaFree(d_cyc); }
    cudaFree(d_s); cudaFree(d_t); cudaFree(d_a); cudaFree(d_b); cudaFree(d_i); cudaFree(d_c); cudaFree(d_acc); cudaFree(d_k); cudaFree(d_o);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------- timing
// n_warps issuing warps (1, 2 or 4) each issue n_mma MMAs of ONE shape from registers (descriptors advance by constant steps with a
// period, accumulators rotate; warp w uses accumulator columns from w * 128), so the loop is a handful of integer adds per MMA: the cycle
// count is the tensor pipe's time unless the shape is faster than the issue.  uniform = 1: every lane of the issuing warp runs the loop
// (warp-uniform values, the MMA itself behind elect.sync) instead of lane 0 alone.
struct TimingArgs {
    int smem_bytes, n_mma, ts, n_warps, uniform;
    uint32_t idesc;
    uint64_t a_desc0; uint32_t a_step16; int a_period;     // SS: descriptor start address advances by a_step16 (16-byte units); TS: TMEM column step
    uint64_t b_desc0; uint32_t b_step16; int b_period;
    uint32_t d_stride; int n_acc;
    long long* cycles;
};

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
    return pred;
}

template <int TS, int P, int NACC>
__global__ void __launch_bounds__(128, 1) umma_timing_kernel(TimingArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_end[4], t_issue[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a.smem_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;      // fp16 ones
    if (tid == 0) {
        for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[w])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 8) {
        uint32_t s = 0x3C003C00u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_addr + c), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t0 = clock64();
    if (warp < a.n_warps && (a.uniform || lane == 0)) {
        const uint64_t base16 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
        const uint64_t a0 = a.a_desc0 + base16, b0 = a.b_desc0 + base16;
        const uint32_t dbase = tbase + (uint32_t)warp * (a.n_warps > 1 ? 96u : 0u);
        // the whole period of descriptors and accumulator addresses lives in registers: the unrolled body is a few moves per MMA
        uint64_t ad[P], bd[P];
        uint32_t ta[P], dc[NACC];
#pragma unroll
        for (int j = 0; j < P; ++j) { ad[j] = a0 + (uint64_t)(j * a.a_step16); bd[j] = b0 + (uint64_t)(j * a.b_step16); ta[j] = tbase + 448u + (uint32_t)(j * a.a_step16); }
#pragma unroll
        for (int j = 0; j < NACC; ++j) dc[j] = dbase + a.d_stride * j;
        constexpr int BODY = P * NACC;                 // one full period of (operand, accumulator) pairs
        const int reps = a.n_mma / BODY;
        const uint32_t go = a.uniform ? elect_one() : 1u;
        if (go) {
#pragma unroll
            for (int j = 0; j < BODY; ++j) {           // first period: the accumulators start fresh
                const uint32_t acc = j >= NACC ? 1u : 0u;
                if (TS) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(dc[j % NACC]), "r"(ta[j % P]), "l"(bd[j % P]), "r"(a.idesc), "r"(acc) : "memory");
                else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(dc[j % NACC]), "l"(ad[j % P]), "l"(bd[j % P]), "r"(a.idesc), "r"(acc) : "memory");
            }
#pragma unroll 1
            for (int r = 1; r < reps; ++r) {
#pragma unroll
                for (int j = 0; j < BODY; ++j) {
                    if (TS) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(dc[j % NACC]), "r"(ta[j % P]), "l"(bd[j % P]), "r"(a.idesc) : "memory");
                    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(dc[j % NACC]), "l"(ad[j % P]), "l"(bd[j % P]), "r"(a.idesc) : "memory");
                }
            }
        }
        const long long t1 = clock64();
        if (go) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[warp])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&mbar[warp])), "r"(0) : "memory");
        if (lane == 0) { t_end[warp] = clock64() - t0; t_issue[warp] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        long long e = 0, is = 0;
        for (int w = 0; w < a.n_warps; ++w) { e = t_end[w] > e ? t_end[w] : e; is = t_issue[w] > is ? t_issue[w] : is; }
        a.cycles[0] = e; a.cycles[1] = is;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int TS, int P, int NACC>
static int launch_timing(const TimingArgs& a) {
    auto k = umma_timing_kernel<TS, P, NACC>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 1;
    k<<<1, 128, a.smem_bytes + 1024, 0>>>(a);
    return 0;
}

// returns 0 on success; cycles_out[0] = first issue -> all MMAs complete, cycles_out[1] = issue loop alone (max over the issuing warps);
// n_mma is per issuing warp
extern "C" int umma_timing(int smem_bytes, int n_mma, int ts, uint32_t idesc, uint64_t a_desc0, uint32_t a_step16, int a_period,
                           uint64_t b_desc0, uint32_t b_step16, int b_period, uint32_t d_stride, int n_acc, int n_warps, int uniform, long long* cycles_out) {
    TimingArgs a{};
    a.smem_bytes = smem_bytes; a.n_mma = n_mma; a.ts = ts; a.idesc = idesc; a.a_desc0 = a_desc0; a.a_step16 = a_step16; a.a_period = a_period;
    a.b_desc0 = b_desc0; a.b_step16 = b_step16; a.b_period = b_period; a.d_stride = d_stride; a.n_acc = n_acc; a.n_warps = n_warps; a.uniform = uniform;
    void* d_cyc = nullptr;
    CK(cudaMalloc(&d_cyc, 16));
    a.cycles = (long long*)d_cyc;
    // n_mma is rounded down to whole periods of (a_period x n_acc) MMAs; a_period must equal b_period (4 or 8), n_acc 1, 2 or 5
    int rc = 2;
    if (a_period == 4 && n_acc == 1) rc = ts ? launch_timing<1, 4, 1>(a) : launch_timing<0, 4, 1>(a);
    else if (a_period == 4 && n_acc == 2) rc = ts ? launch_timing<1, 4, 2>(a) : launch_timing<0, 4, 2>(a);
    else if (a_period == 4 && n_acc == 5) rc = ts ? launch_timing<1, 4, 5>(a) : launch_timing<0, 4, 5>(a);
    else if (a_period == 8 && n_acc == 1) rc = launch_timing<0, 8, 1>(a);
    else if (a_period == 8 && n_acc == 2) rc = launch_timing<0, 8, 2>(a);
    if (rc) { snprintf(err, 256, "umma_timing: unsupported period / accumulator combination (%d, %d)", a_period, n_acc); return 1; }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cycles_out, d_cyc, 16, cudaMemcpyDeviceToHost));
    cudaFree(d_cyc);
    return 0;
}
