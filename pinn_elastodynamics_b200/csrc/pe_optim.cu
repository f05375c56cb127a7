// Slot reduction (deterministic, fixed order) and the TF1-form Adam update.
// Replaces tf.train.AdamOptimizer(...).minimize's ApplyAdam nodes (PlateHoleQuarter/train/train.py:249-250)
// and the tf.reduce_mean scalar reductions (train.py:187-217).  See SURVEY.md A.3 for the epsilon convention.
#include "pe_common.cuh"
#include <cstring>

namespace {

#define RED_TX 32      // peer-memory kernel: float4 columns per block (one arrival flag per block and peer: few, wide blocks)
#define RED_TY 16      //                     slot-group rows per block
#define RA_TX 8        // single-GPU kernels: 8 float4 columns (one 128-byte line per slot) x 32 rows per block -> four times the blocks
#define RA_TY 32       // (360 for the 5x50 net): the reduction is a latency chain of L2 loads, more CTAs in flight shorten it
#define RED_NP 32      // canonical partial sums per column: partial j = slots j, j + 32, j + 64, ... added in increasing order

// blockDim = (TX, TY), TY divides RED_NP.  The summation ORDER is independent of the block shape: 32 partial sums per column (slot index
// mod 32), then groups of 8 partials, then the 4 group sums -- so the single-GPU kernels, the NCCL path and the peer-memory kernel produce
// the same bits for the same slots.  Thread (tx, ty) owns partials ty, ty + TY, ...; all its loads are in flight together.
template <int TX, int TY>
__device__ __forceinline__ float4 sum_slots4(const float* __restrict__ partials, int n_slots, int stride, int i4, bool in_range) {
    constexpr int PER = RED_NP / TY;
    __shared__ float4 red[RED_NP][TX];
    const float4* p = reinterpret_cast<const float4*>(partials) + i4;
    const size_t stride4 = (size_t)(stride >> 2);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int j = threadIdx.y + q * TY;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_range) {
            int k = j;
            for (; k + 3 * RED_NP < n_slots; k += 4 * RED_NP) {
                float4 a = __ldcg(p + (size_t)(k) * stride4);
                float4 b = __ldcg(p + (size_t)(k + RED_NP) * stride4);
                float4 c = __ldcg(p + (size_t)(k + 2 * RED_NP) * stride4);
                float4 d = __ldcg(p + (size_t)(k + 3 * RED_NP) * stride4);
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
                s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
                s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
                s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
            }
            for (; k < n_slots; k += RED_NP) {
                float4 a = __ldcg(p + (size_t)k * stride4);
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            }
        }
        red[j][threadIdx.x] = s;
    }
    __syncthreads();
    constexpr int G = 8;                        // rows ty < RED_NP / G each add the G partials ty * G .. ty * G + G - 1
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.y < RED_NP / G) {
        s = red[threadIdx.y * G][threadIdx.x];
#pragma unroll
        for (int r = 1; r < G; ++r) {
            float4 a = red[threadIdx.y * G + r][threadIdx.x];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    __syncthreads();
    if (threadIdx.y < RED_NP / G) red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int r = 1; r < RED_NP / G; ++r) {
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
    int i4 = blockIdx.x * RA_TX + threadIdx.x;
    const bool in_range = i4 < (total >> 2);
    float4 s = sum_slots4<RA_TX, RA_TY>(gp, n_slots, total, i4, in_range);
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
    pe_grid_dep_wait();          // gradient partials of the residual kernel(s)
    pe_grid_dep_trigger();       // the next step's first kernel may be scheduled (it waits for this grid before it reads the parameters)
    const int step = *reinterpret_cast<volatile int*>(d_step) + 1;
    const float lr_t = adam_lr_t(step, lr, b1, b2);
    int i4 = blockIdx.x * RA_TX + threadIdx.x;
    const bool in_range = i4 < (total >> 2);
    float4 g = sum_slots4<RA_TX, RA_TY>(gp, n_slots, total, i4, in_range);
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


// ---------------------------------------------------------------------------------------------------------------------------
// Slot reduction + all-reduce over NVLink peer memory + Adam in ONE kernel (multi-GPU step).
// Every rank owns a region [2 buffers][world sources][n = total + 8 floats] + arrival flags, mapped into all peers (CUDA IPC).
// Block b of rank r: (1) sums its 128-float slice over the gradient-partial slots, (2) PUSHES the slice into source row r of every
// peer's region (fire-and-forget NVLink stores), fences, and raises flag[r][b] = seq in every peer, (3) polls its LOCAL flags
// [q][b] for all q, (4) adds the `world` rows in rank order -- the same order on every rank, so all ranks hold bit-identical
// gradients -- and applies Adam to its slice.  A block only ever waits for the same-index block of its peers, so no grid-wide
// co-residency is needed.  Buffers alternate with seq: a rank can be at most one step ahead of a peer (it needs the peer's
// flags of step s to finish step s), so buffer s&1 is never overwritten while still being read.
struct PeerArgs {
    float* data[PE_MAX_PEERS];
    unsigned int* flags[PE_MAX_PEERS];
    int* err;
    int rank, world, n, n_cta;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded wait (about 20 s of SM clock): a dead peer must not hang the GPU for good; the error word is read by pe_comm_error
__device__ __forceinline__ void wait_flag(const unsigned int* p, unsigned int seq, int* err) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - seq) < 0) {
        if (*reinterpret_cast<volatile int*>(err)) break;              // an earlier wait already failed: do not stall every later step
        if (clock64() - t0 > 40000000000LL) { atomicExch(err, 1); break; }
    }
}

__global__ void __launch_bounds__(RED_TX * RED_TY) reduce_peer_adam_kernel(const float* __restrict__ gp, const float* __restrict__ tp, int n_slots, int total, float* __restrict__ out,
                                        float* __restrict__ tcopy, float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                        int* __restrict__ d_step, unsigned int* __restrict__ ticket, float lr, float b1, float b2, float eps,
                                        PeerArgs pa, unsigned int seq, int do_adam) {
    pe_grid_dep_wait();
    pe_grid_dep_trigger();
    int step = 0;
    float lr_t = 0.f;
    if (do_adam) { step = *reinterpret_cast<volatile int*>(d_step) + 1; lr_t = adam_lr_t(step, lr, b1, b2); }
    const int buf = (int)(seq & 1u);
    const int lane = threadIdx.x;
    const size_t row = (size_t)pa.n;
    // the grid is capped at the number of SMs (every block of every rank is resident, so waiting for a peer's slice cannot starve the block
    // that produces it); a block walks over its 128-float slices sl = blockIdx.x, + gridDim.x, ..., one arrival flag per slice and peer
    for (int sl = blockIdx.x; sl < pa.n_cta; sl += gridDim.x) {
    const int i4 = sl * RED_TX + threadIdx.x;
    const bool in_range = i4 < (total >> 2);
    const float4 s = sum_slots4<RED_TX, RED_TY>(gp, n_slots, total, i4, in_range);
    if (threadIdx.y == 0) {
        if (in_range)
            for (int p = 0; p < pa.world; ++p)
                reinterpret_cast<float4*>(pa.data[p] + ((size_t)buf * pa.world + pa.rank) * row)[i4] = s;
        __threadfence_system();
        __syncwarp();
        if (lane < pa.world) st_release_sys(pa.flags[lane] + (size_t)pa.rank * (pa.n_cta + 1) + sl, seq);
        if (lane < pa.world) wait_flag(pa.flags[pa.rank] + (size_t)lane * (pa.n_cta + 1) + sl, seq, pa.err);
        __syncwarp();
        if (in_range) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < pa.world; ++q) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(pa.data[pa.rank] + ((size_t)buf * pa.world + q) * row) + i4);
                g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
            }
            reinterpret_cast<float4*>(out)[i4] = g;
            if (do_adam) {
                float4 p = reinterpret_cast<float4*>(params)[i4];
                float4 mm = reinterpret_cast<float4*>(m)[i4];
                float4 vv = reinterpret_cast<float4*>(v)[i4];
                adam4(p, g, mm, vv, lr_t, b1, b2, eps);
                reinterpret_cast<float4*>(params)[i4] = p;
                reinterpret_cast<float4*>(m)[i4] = mm;
                reinterpret_cast<float4*>(v)[i4] = vv;
            }
        }
    } else if (sl == 0 && threadIdx.y == 1) {          // the 8 loss terms ride the same exchange (flag column n_cta)
        float st = 0.f;
        if (lane < PE_MAX_TERMS) {
            for (int k = 0; k < n_slots; ++k) st += __ldcg(tp + (size_t)k * PE_MAX_TERMS + lane);
            for (int p = 0; p < pa.world; ++p) pa.data[p][((size_t)buf * pa.world + pa.rank) * row + total + lane] = st;
        }
        __threadfence_system();
        __syncwarp();
        if (lane < pa.world) st_release_sys(pa.flags[lane] + (size_t)pa.rank * (pa.n_cta + 1) + pa.n_cta, seq);
        if (lane < pa.world) wait_flag(pa.flags[pa.rank] + (size_t)lane * (pa.n_cta + 1) + pa.n_cta, seq, pa.err);
        __syncwarp();
        if (lane < PE_MAX_TERMS) {
            float g = 0.f;
            for (int q = 0; q < pa.world; ++q) g += __ldcg(pa.data[pa.rank] + ((size_t)buf * pa.world + q) * row + total + lane);
            out[total + lane] = g;
            if (tcopy) tcopy[lane] = g;
        }
    }
    __syncthreads();                 // the shared partial-sum array of sum_slots4 is reused by the next slice
    }
    if (do_adam && threadIdx.x == 0 && threadIdx.y == 0) {
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
    int grid = (n4 + RA_TX - 1) / RA_TX;
    dim3 bs(RA_TX, RA_TY);
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
    int grid = (n4 + RA_TX - 1) / RA_TX;
    dim3 bs(RA_TX, RA_TY);
    cudaError_t e = pe_launch_pdl(reduce_adam_kernel, dim3(grid), bs, 0, (cudaStream_t)stream, d_grad_partials, d_term_partials, n_slots, total, d_out, d_terms_copy,
                                  d_params, d_m, d_v, d_step, reinterpret_cast<unsigned int*>(d_step + 1), lr, beta1, beta2, eps);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("reduce_adam_kernel: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

// ---------------------------------------------------------------- peer-memory communicator (single node, CUDA IPC)
struct pe_comm {
    int rank, world, n, n_cta;
    size_t bytes, flags_off, err_off;
    unsigned char* local;
    unsigned char* peer[PE_MAX_PEERS];
    unsigned int seq;
    int connected;
};

extern "C" pe_comm* pe_comm_create(const pe_plan* plan, int rank, int world, unsigned char* handle_out) {
    if (!plan || !handle_out || world < 2 || world > PE_MAX_PEERS || rank < 0 || rank >= world) {
        pe_set_error("pe_comm_create: bad arguments (2 <= world <= %d)", PE_MAX_PEERS);
        return nullptr;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == PE_IPC_HANDLE_BYTES, "IPC handle size");
    pe_comm* c = new pe_comm();
    c->rank = rank; c->world = world;
    c->n = plan->lay.total + PE_MAX_TERMS;
    c->n_cta = (plan->lay.total / 4 + RED_TX - 1) / RED_TX;
    c->flags_off = sizeof(float) * 2 * (size_t)world * c->n;
    c->err_off = c->flags_off + sizeof(unsigned int) * (size_t)world * (c->n_cta + 1);
    c->bytes = ((c->err_off + 64 + (2u << 20) - 1) / (2u << 20)) * (2u << 20);          // whole 2 MiB pages: nothing else shares the IPC mapping
    cudaError_t e = cudaMalloc(&c->local, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) {
        pe_set_error("pe_comm_create: %s", cudaGetErrorString(e));
        if (c->local) cudaFree(c->local);
        delete c;
        return nullptr;
    }
    memcpy(handle_out, &h, sizeof(h));
    return c;
}

extern "C" int pe_comm_connect(pe_comm* c, const unsigned char* all_handles) {
    if (!c || !all_handles) { pe_set_error("pe_comm_connect: null argument"); return 1; }
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) { c->peer[p] = c->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)p * PE_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            pe_set_error("pe_comm_connect: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
            (void)cudaGetLastError();
            for (int q = 0; q < p; ++q) if (q != c->rank && c->peer[q]) { cudaIpcCloseMemHandle(c->peer[q]); c->peer[q] = nullptr; }
            return 3;
        }
        c->peer[p] = static_cast<unsigned char*>(ptr);
    }
    c->connected = 1;
    return 0;
}

extern "C" int pe_comm_error(pe_comm* c) {
    if (!c) return 1;
    int err = 0;
    if (cudaMemcpy(&err, c->local + c->err_off, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    if (err) pe_set_error("peer all-reduce timed out waiting for a peer rank");
    return err;
}

extern "C" void pe_comm_disconnect(pe_comm* c) {
    if (!c || !c->connected) return;
    for (int p = 0; p < c->world; ++p) if (p != c->rank && c->peer[p]) { cudaIpcCloseMemHandle(c->peer[p]); c->peer[p] = nullptr; }
    c->connected = 0;
}

extern "C" void pe_comm_destroy(pe_comm* c) {
    if (!c) return;
    pe_comm_disconnect(c);
    if (c->local) cudaFree(c->local);
    delete c;
}

extern "C" int pe_reduce_peer(const pe_plan* plan, pe_comm* c, const float* d_grad_partials, const float* d_term_partials, int n_slots,
                              float* d_out, float* d_terms_copy, float* d_params, float* d_m, float* d_v, int* d_step,
                              float lr, float beta1, float beta2, float eps, void* stream) {
    if (!plan || !c || !c->connected) { pe_set_error("pe_reduce_peer: communicator not connected"); return 1; }
    if (c->n != plan->lay.total + PE_MAX_TERMS) { pe_set_error("pe_reduce_peer: communicator was created for another plan"); return 1; }
    PeerArgs pa;
    for (int p = 0; p < c->world; ++p) {
        pa.data[p] = reinterpret_cast<float*>(c->peer[p]);
        pa.flags[p] = reinterpret_cast<unsigned int*>(c->peer[p] + c->flags_off);
    }
    pa.err = reinterpret_cast<int*>(c->local + c->err_off);
    pa.rank = c->rank; pa.world = c->world; pa.n = c->n; pa.n_cta = c->n_cta;
    c->seq += 1;
    const int total = plan->lay.total;
    dim3 bs(RED_TX, RED_TY);
    const int do_adam = d_params != nullptr;
    const int grid = c->n_cta < plan->sms ? c->n_cta : plan->sms;      // co-resident by construction (at least one 512-thread block per SM)
    cudaError_t e = pe_launch_pdl(reduce_peer_adam_kernel, dim3(grid), bs, 0, (cudaStream_t)stream, d_grad_partials, d_term_partials, n_slots, total, d_out, d_terms_copy,
                                  d_params, d_m, d_v, d_step, do_adam ? reinterpret_cast<unsigned int*>(d_step + 1) : nullptr,
                                  lr, beta1, beta2, eps, pa, c->seq, do_adam);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { pe_set_error("reduce_peer_adam_kernel: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}
