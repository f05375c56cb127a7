// Slot reduction (deterministic, fixed order) and the TF1-form Adam update.
// Replaces tf.train.AdamOptimizer(...).minimize's ApplyAdam nodes (PlateHoleQuarter/train/train.py:249-250)
// and the tf.reduce_mean scalar reductions (train.py:187-217).  See SURVEY.md A.3 for the epsilon convention.
#include "pe_common.cuh"

namespace {

#define RED_TX 32      // float4 columns per block
#define RED_TY 16      // slot groups per block

// blockDim = (RED_TX, RED_TY).  Thread (tx, ty) sums slots ty, ty+RED_TY, ... of float4 column i4 (4 loads in
// flight), the RED_TY partials are then added in fixed order by ty == 0: deterministic for a given n_slots.
__device__ __forceinline__ float4 sum_slots4(const float* __restrict__ partials, int n_slots, int stride, int i4, bool in_range) {
    __shared__ float4 red[RED_TY][RED_TX];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
        const float4* p = reinterpret_cast<const float4*>(partials) + i4;
        const size_t stride4 = (size_t)(stride >> 2);
        int k = threadIdx.y;
        for (; k + 3 * RED_TY < n_slots; k += 4 * RED_TY) {
            float4 a = __ldcg(p + (size_t)(k) * stride4);
            float4 b = __ldcg(p + (size_t)(k + RED_TY) * stride4);
            float4 c = __ldcg(p + (size_t)(k + 2 * RED_TY) * stride4);
            float4 d = __ldcg(p + (size_t)(k + 3 * RED_TY) * stride4);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
            s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        }
        for (; k < n_slots; k += RED_TY) {
            float4 a = __ldcg(p + (size_t)k * stride4);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int r = 1; r < RED_TY; ++r) {
            float4 a = red[r][threadIdx.x];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    return s;      // valid for ty == 0
}

__device__ __forceinline__ void reduce_terms(const float* __restrict__ term_partials, int n_slots, float* out_terms, float* copy) {
    // one warp: lane t < PE_MAX_TERMS sums its term over slots in fixed order
    int t = threadIdx.x;
    if (threadIdx.y == 1 && t < PE_MAX_TERMS) {
        float s = 0.f;
        for (int k = 0; k < n_slots; ++k) s += __ldcg(term_partials + (size_t)k * PE_MAX_TERMS + t);
        out_terms[t] = s;
        if (copy) copy[t] = s;
    }
}

__device__ __forceinline__ float adam_lr_t(int step, float lr, float b1, float b2) {
    double p1 = pow((double)b1, (double)step), p2 = pow((double)b2, (double)step);
    return (float)((double)lr * sqrt(1.0 - p2) / (1.0 - p1));
}

__device__ __forceinline__ void adam4(float4& p, float4 g, float4& m, float4& v, float lr_t, float b1, float b2, float eps) {
    m.x = b1 * m.x + (1.f - b1) * g.x; m.y = b1 * m.y + (1.f - b1) * g.y;
    m.z = b1 * m.z + (1.f - b1) * g.z; m.w = b1 * m.w + (1.f - b1) * g.w;
    v.x = b2 * v.x + (1.f - b2) * g.x * g.x; v.y = b2 * v.y + (1.f - b2) * g.y * g.y;
    v.z = b2 * v.z + (1.f - b2) * g.z * g.z; v.w = b2 * v.w + (1.f - b2) * g.w * g.w;
    p.x -= lr_t * m.x / (sqrtf(v.x) + eps); p.y -= lr_t * m.y / (sqrtf(v.y) + eps);
    p.z -= lr_t * m.z / (sqrtf(v.z) + eps); p.w -= lr_t * m.w / (sqrtf(v.w) + eps);
}

__global__ void reduce_kernel(const float* __restrict__ gp, const float* __restrict__ tp, int n_slots, int total, float* __restrict__ out, float* __restrict__ tcopy) {
    int i4 = blockIdx.x * RED_TX + threadIdx.x;
    const bool in_range = i4 < (total >> 2);
    float4 s = sum_slots4(gp, n_slots, total, i4, in_range);
    if (in_range && threadIdx.y == 0) reinterpret_cast<float4*>(out)[i4] = s;
    if (blockIdx.x == 0) reduce_terms(tp, n_slots, out + total, tcopy);
}

// step counter protocol: every thread reads *d_step (value before this update); the LAST block to finish
// increments it (threadfence + atomic ticket), so the kernel can be replayed from a CUDA graph.
__global__ void adam_kernel(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                            int* __restrict__ d_step, unsigned int* __restrict__ ticket, int total, float lr, float b1, float b2, float eps) {
    const int step = *reinterpret_cast<volatile int*>(d_step) + 1;
    const float lr_t = adam_lr_t(step, lr, b1, b2);
    int i4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 < (total >> 2)) {
        float4 p = reinterpret_cast<float4*>(params)[i4];
        float4 g = reinterpret_cast<const float4*>(grad)[i4];
        float4 mm = reinterpret_cast<float4*>(m)[i4];
        float4 vv = reinterpret_cast<float4*>(v)[i4];
        adam4(p, g, mm, vv, lr_t, b1, b2, eps);
        reinterpret_cast<float4*>(params)[i4] = p;
        reinterpret_cast<float4*>(m)[i4] = mm;
        reinterpret_cast<float4*>(v)[i4] = vv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) { *d_step = step; *ticket = 0u; __threadfence(); }
    }
}

__global__ void reduce_adam_kernel(const float* __restrict__ gp, const float* __restrict__ tp, int n_slots, int total, float* __restrict__ out, float* __restrict__ tcopy,
                                   float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                   int* __restrict__ d_step, unsigned int* __restrict__ ticket, float lr, float b1, float b2, float eps) {
    const int step = *reinterpret_cast<volatile int*>(d_step) + 1;
    const float lr_t = adam_lr_t(step, lr, b1, b2);
    int i4 = blockIdx.x * RED_TX + threadIdx.x;
    const bool in_range = i4 < (total >> 2);
    float4 g = sum_slots4(gp, n_slots, total, i4, in_range);
    if (in_range && threadIdx.y == 0) {
        reinterpret_cast<float4*>(out)[i4] = g;
        float4 p = reinterpret_cast<float4*>(params)[i4];
        float4 mm = reinterpret_cast<float4*>(m)[i4];
        float4 vv = reinterpret_cast<float4*>(v)[i4];
        adam4(p, g, mm, vv, lr_t, b1, b2, eps);
        reinterpret_cast<float4*>(params)[i4] = p;
        reinterpret_cast<float4*>(m)[i4] = mm;
        reinterpret_cast<float4*>(v)[i4] = vv;
    }
    if (blockIdx.x == 0) reduce_terms(tp, n_slots, out + total, tcopy);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) { *d_step = step; *ticket = 0u; __threadfence(); }
    }
}

}  // namespace

extern "C" int pe_reduce_partials(const pe_plan* plan, const float* d_grad_partials, const float* d_term_partials,
                                  int n_slots, float* d_out, float* d_terms_copy, void* stream) {
    if (!plan) { pe_set_error("null plan"); return 1; }
    int total = plan->lay.total;
    int n4 = total / 4;
    int grid = (n4 + RED_TX - 1) / RED_TX;
    dim3 bs(RED_TX, RED_TY);
    reduce_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(d_grad_partials, d_term_partials, n_slots, total, d_out, d_terms_copy);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("reduce_kernel: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

// the ticket word lives right after the step counter: d_step must point to at least 2 ints.
extern "C" int pe_adam_step(const pe_plan* plan, float* d_params, const float* d_grad, float* d_m, float* d_v,
                            int* d_step, float lr, float beta1, float beta2, float eps, void* stream) {
    if (!plan) { pe_set_error("null plan"); return 1; }
    int total = plan->lay.total;
    int n4 = total / 4;
    int bs = 128, grid = (n4 + bs - 1) / bs;
    adam_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(d_params, d_grad, d_m, d_v, d_step, reinterpret_cast<unsigned int*>(d_step + 1),
                                                       total, lr, beta1, beta2, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("adam_kernel: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

extern "C" int pe_reduce_adam(const pe_plan* plan, const float* d_grad_partials, const float* d_term_partials,
                              int n_slots, float* d_out, float* d_terms_copy, float* d_params, float* d_m, float* d_v, int* d_step,
                              float lr, float beta1, float beta2, float eps, void* stream) {
    if (!plan) { pe_set_error("null plan"); return 1; }
    int total = plan->lay.total;
    int n4 = total / 4;
    int grid = (n4 + RED_TX - 1) / RED_TX;
    dim3 bs(RED_TX, RED_TY);
    reduce_adam_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(d_grad_partials, d_term_partials, n_slots, total, d_out, d_terms_copy,
                                                             d_params, d_m, d_v, d_step, reinterpret_cast<unsigned int*>(d_step + 1),
                                                             lr, beta1, beta2, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("reduce_adam_kernel: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}
