// Device-resident L-BFGS vector algebra: two-loop recursion, curvature-pair storage, axpy, dot / max-norm.
// The reference drives SciPy's L-BFGS-B through tf.contrib.opt.ScipyOptimizerInterface (PlateHoleQuarter/train/train.py:240-247,
// 522-525): parameters and gradient cross the host boundary on every function evaluation and the O(m n) recursion runs on one
// CPU core (20 ms per evaluation at n = 10,655, m = 50 -- ten times the GPU evaluation itself).  Here the parameter vector, the
// gradient and the (S, Y) history never leave the GPU; the host only sees the scalars the line search branches on.
// One CTA of 1024 threads: n <= ~1e5 parameters, all dot products accumulate in double.
#include "pe_common.cuh"

namespace {

constexpr int LB_T = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = (l < LB_T / 32) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;          // every thread holds the total
}

// d = -H g with the L-BFGS two-loop recursion over the `count` newest pairs of the ring buffers S, Y ([m][n]); head = slot of the
// newest pair.  state[0] = gamma (initial Hessian scaling y.s / y.y), state[1 + i] = rho_i = 1 / (y_i . s_i).
__global__ void __launch_bounds__(LB_T, 1) lbfgs_direction_kernel(int n, int m, int count, int head, const float* __restrict__ g,
                                                                 const float* __restrict__ S, const float* __restrict__ Y,
                                                                 const float* __restrict__ state, float* __restrict__ d) {
    __shared__ double red[LB_T / 32];
    __shared__ double alpha[64];
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += LB_T) d[i] = g[i];
    __syncthreads();
    for (int j = 0; j < count; ++j) {                    // newest -> oldest
        const int slot = (head - j + m) % m;
        const float* s = S + (size_t)slot * n;
        const float* y = Y + (size_t)slot * n;
        double p = 0.0;
        for (int i = tid; i < n; i += LB_T) p += (double)s[i] * (double)d[i];
        const double a = (double)state[1 + slot] * block_sum(p, red);
        if (tid == 0) alpha[j] = a;
        for (int i = tid; i < n; i += LB_T) d[i] = (float)((double)d[i] - a * (double)y[i]);
        __syncthreads();
    }
    const float gamma = count > 0 ? state[0] : 1.0f;
    for (int i = tid; i < n; i += LB_T) d[i] *= gamma;
    __syncthreads();
    for (int j = count - 1; j >= 0; --j) {               // oldest -> newest
        const int slot = (head - j + m) % m;
        const float* s = S + (size_t)slot * n;
        const float* y = Y + (size_t)slot * n;
        double p = 0.0;
        for (int i = tid; i < n; i += LB_T) p += (double)y[i] * (double)d[i];
        const double b = (double)state[1 + slot] * block_sum(p, red);
        const double c = alpha[j] - b;
        for (int i = tid; i < n; i += LB_T) d[i] = (float)((double)d[i] + c * (double)s[i]);
        __syncthreads();
    }
    for (int i = tid; i < n; i += LB_T) d[i] = -d[i];
}

// Same recursion with the direction vector held in REGISTERS (n <= EPT * LB_T = 12,288: the 5x50 nets; 1,024 threads leave 64 registers each,
// wider nets use the kernel above): thread t owns elements t + LB_T * e.  Per pair the two
// history rows are fetched up front (coalesced, all loads in flight), so a step costs one L2 round trip + one block reduction instead of
// three passes over d in global memory.  Same summation order per thread and the same block reduction as the kernel above.
template <int EPT>
__global__ void __launch_bounds__(LB_T, 1) lbfgs_direction_reg_kernel(int n, int m, int count, int head, const float* __restrict__ g,
                                                                     const float* __restrict__ S, const float* __restrict__ Y,
                                                                     const float* __restrict__ state, float* __restrict__ d) {
    __shared__ double red[LB_T / 32];
    __shared__ double alpha[64];
    const int tid = threadIdx.x;
    float dv[EPT], sv[EPT], yv[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) { const int i = tid + LB_T * e; dv[e] = i < n ? g[i] : 0.f; }
    for (int j = 0; j < count; ++j) {                    // newest -> oldest
        const int slot = (head - j + m) % m;
        const float* s = S + (size_t)slot * n;
        const float* y = Y + (size_t)slot * n;
#pragma unroll
        for (int e = 0; e < EPT; ++e) { const int i = tid + LB_T * e; sv[e] = i < n ? __ldcg(s + i) : 0.f; yv[e] = i < n ? __ldcg(y + i) : 0.f; }
        double p = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) p += (double)sv[e] * (double)dv[e];
        const double a = (double)state[1 + slot] * block_sum(p, red);
        if (tid == 0) alpha[j] = a;
#pragma unroll
        for (int e = 0; e < EPT; ++e) dv[e] = (float)((double)dv[e] - a * (double)yv[e]);
    }
    const float gamma = count > 0 ? state[0] : 1.0f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) dv[e] *= gamma;
    __syncthreads();
    for (int j = count - 1; j >= 0; --j) {               // oldest -> newest
        const int slot = (head - j + m) % m;
        const float* s = S + (size_t)slot * n;
        const float* y = Y + (size_t)slot * n;
#pragma unroll
        for (int e = 0; e < EPT; ++e) { const int i = tid + LB_T * e; sv[e] = i < n ? __ldcg(s + i) : 0.f; yv[e] = i < n ? __ldcg(y + i) : 0.f; }
        double p = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) p += (double)yv[e] * (double)dv[e];
        const double b = (double)state[1 + slot] * block_sum(p, red);
        const double c = alpha[j] - b;
#pragma unroll
        for (int e = 0; e < EPT; ++e) dv[e] = (float)((double)dv[e] + c * (double)sv[e]);
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) { const int i = tid + LB_T * e; if (i < n) d[i] = -dv[e]; }
}

// s = x - x_prev, y = g - g_prev into slot `head`; rho, gamma and y.s into state.  state[1 + m] = y.s, state[2 + m] = y.y.
__global__ void __launch_bounds__(LB_T, 1) lbfgs_store_pair_kernel(int n, int m, int head, const float* __restrict__ x, const float* __restrict__ xp,
                                                                  const float* __restrict__ g, const float* __restrict__ gp,
                                                                  float* __restrict__ S, float* __restrict__ Y, float* __restrict__ state) {
    __shared__ double red[LB_T / 32];
    const int tid = threadIdx.x;
    float* s = S + (size_t)head * n;
    float* y = Y + (size_t)head * n;
    double ys = 0.0, yy = 0.0;
    for (int i = tid; i < n; i += LB_T) {
        const float si = x[i] - xp[i], yi = g[i] - gp[i];
        s[i] = si; y[i] = yi;
        ys += (double)yi * (double)si; yy += (double)yi * (double)yi;
    }
    ys = block_sum(ys, red);
    yy = block_sum(yy, red);
    if (tid == 0) {
        state[1 + m] = (float)ys; state[2 + m] = (float)yy;
        if (ys > 0.0) { state[1 + head] = (float)(1.0 / ys); state[0] = (float)(ys / yy); }
        else state[1 + head] = 0.f;                 // curvature condition lost in fp32: the pair contributes nothing
    }
}

__global__ void axpy_kernel(int n, float* __restrict__ out, const float* __restrict__ x, float alpha, const float* __restrict__ d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmaf(alpha, d[i], x[i]);
}

// res[0] = a . b, res[1] = max |a|
__global__ void __launch_bounds__(LB_T, 1) dot_max_kernel(int n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ res) {
    __shared__ double red[LB_T / 32];
    __shared__ float redm[LB_T / 32];
    const int tid = threadIdx.x;
    double p = 0.0; float mx = 0.f;
    for (int i = tid; i < n; i += LB_T) { p += (double)a[i] * (double)b[i]; mx = fmaxf(mx, fabsf(a[i])); }
    p = block_sum(p, red);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) redm[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        float m2 = 0.f;
        for (int w = 0; w < LB_T / 32; ++w) m2 = fmaxf(m2, redm[w]);
        res[0] = (float)p; res[1] = m2;
    }
}

}  // namespace

#define LB_CHECK(name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { pe_set_error(name ": %s", cudaGetErrorString(e_)); return 3; } } while (0)

extern "C" int pe_lbfgs_direction(int n, int m, int count, int head, const float* d_g, const float* d_S, const float* d_Y,
                                  const float* d_state, float* d_dir, void* stream) {
    if (n < 1 || m < 1 || m > 64 || count < 0 || count > m) { pe_set_error("pe_lbfgs_direction: bad sizes (m <= 64)"); return 1; }
    if (n <= 12 * LB_T) lbfgs_direction_reg_kernel<12><<<1, LB_T, 0, (cudaStream_t)stream>>>(n, m, count, head, d_g, d_S, d_Y, d_state, d_dir);
    else lbfgs_direction_kernel<<<1, LB_T, 0, (cudaStream_t)stream>>>(n, m, count, head, d_g, d_S, d_Y, d_state, d_dir);
    LB_CHECK("lbfgs_direction_kernel");
    return 0;
}

extern "C" int pe_lbfgs_store_pair(int n, int m, int head, const float* d_x, const float* d_xprev, const float* d_g, const float* d_gprev,
                                   float* d_S, float* d_Y, float* d_state, void* stream) {
    if (n < 1 || m < 1 || m > 64 || head < 0 || head >= m) { pe_set_error("pe_lbfgs_store_pair: bad sizes"); return 1; }
    lbfgs_store_pair_kernel<<<1, LB_T, 0, (cudaStream_t)stream>>>(n, m, head, d_x, d_xprev, d_g, d_gprev, d_S, d_Y, d_state);
    LB_CHECK("lbfgs_store_pair_kernel");
    return 0;
}

extern "C" int pe_vec_axpy(int n, float* d_out, const float* d_x, float alpha, const float* d_d, void* stream) {
    axpy_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_out, d_x, alpha, d_d);
    LB_CHECK("axpy_kernel");
    return 0;
}

extern "C" int pe_vec_dot_max(int n, const float* d_a, const float* d_b, float* d_res, void* stream) {
    dot_max_kernel<<<1, LB_T, 0, (cudaStream_t)stream>>>(n, d_a, d_b, d_res);
    LB_CHECK("dot_max_kernel");
    return 0;
}
