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
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
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
    if (tid == 0) {
        const uint64_t base16 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
        for (int i = 0; i < a.n_mma; ++i) {
            uint64_t da = a.a_desc[i], db = a.b_desc[i];
            db = (db & ~0x3FFFull) | (((db & 0x3FFF) + base16) & 0x3FFF);
            uint32_t d = tbase + a.d_col[i], id = a.idesc[i], acc = a.accum[i], kind = a.kind[i];
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
                uint64_t dc = (a.a_desc[i] & ~0x3FFFull) | (((a.a_desc[i] & 0x3FFF) + base16) & 0x3FFF);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(d), "l"(dc) : "memory");
            } else {                   // 128 lanes x 128 bit
                uint64_t dc = (a.a_desc[i] & ~0x3FFFull) | (((a.a_desc[i] & 0x3FFF) + base16) & 0x3FFF);
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

extern "C" const char* umma_probe_error(void) { return err; }

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
    cudaFree(d_s); cudaFree(d_t); cudaFree(d_a); cudaFree(d_b); cudaFree(d_i); cudaFree(d_c); cudaFree(d_acc); cudaFree(d_k); cudaFree(d_o);
    return 0;
}
