// Shared definitions of the sm_100a PINN-elastodynamics kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/pinn_elasto.h"

#define PE_P 32          // points per tile (lane = point)
#define PE_UJ 10         // hidden units owned by one thread in the forward / adjoint GEMMs

// Padded device layout of one network (see include/pinn_elasto.h).
struct PeLayout {
    int L;                          // number of weight matrices
    int d[PE_MAX_LAYERS + 1];       // widths d[0]=3 ... d[L]=O
    int ldw[PE_MAX_LAYERS];         // row stride of W_l = round_up(d[l+1], 4)
    int woff[PE_MAX_LAYERS];        // float offset of W_l (multiple of 4)
    int boff[PE_MAX_LAYERS];        // float offset of b_l (multiple of 4)
    int lda[PE_MAX_LAYERS + 1];     // shared-memory row stride for width d[l]: multiple of 4, (lda/4) odd
    int soff[PE_MAX_LAYERS + 1];    // per-stream-set float offset of layer l's activation plane in the stash, in units of (K*PE_P) rows: sum of lda over hidden layers before l
    int stash_rows;                 // sum of lda[l] over hidden layers l = 1..L-1
    int total;                      // padded parameter count (multiple of 4)
    int compact;                    // compact parameter count
    int maxw;                       // max over d[1..L-1] (hidden widths); d[L] <= PE_UJ required
    int groups;                     // warps per CTA = ceil(max(maxw, d[L]) / PE_UJ)
    int max_lda;                    // max lda over all layers
    int wmat_floats;                // max over layers of d[l]*ldw[l] (one staged weight matrix)
    int wstage_floats;              // one weight staging buffer: wmat_floats + max ldw (bias), multiple of 4
};

struct pe_plan {
    PeLayout lay;
    int device;
    int sms;
    int smem_optin;
    void* d_tc_images = nullptr;   // operand images of the tensor-core predict path (csrc/pe_tcf.cu pe_launch_fields_tcf), allocated at first use
};

struct PeResidArgs {
    PeLayout lay;
    pe_term_desc term;
    const float* points;
    const float* aux;
    const float* params;
    float* grad_partials;     // [slots][lay.total]
    float* term_partials;     // [slots][PE_MAX_TERMS]
    float* stash;             // [slots][stash_floats]
    int n;                    // local points
    int slot_base;
    int stash_floats;         // per slot
    float inv_n;              // 1 / n_global
};

struct PeFieldsArgs {
    PeLayout lay;
    const float* points;
    const float* aux;
    const float* params;
    float* out;
    int n, ld, aux_k, mode, formulation;   // mode 0 = fields[n][8], 1 = jets[n][K][O]
    float in_scale[3], in_shift[3];
};

static __host__ __device__ inline int pe_round4(int x) { return (x + 3) & ~3; }
static __host__ __device__ inline int pe_lda(int d) {
    int r = pe_round4(d);
    if (((r >> 2) & 1) == 0) r += 4;
    return r;
}

void pe_set_error(const char* fmt, ...);

// ---- programmatic dependent launch (PDL).  The Adam step is three back-to-back kernels on one stream (operand images -> residual ->
// slot reduction + Adam); launched with the programmatic-stream-serialization attribute the next kernel's CTAs are scheduled while the
// previous kernel drains, run their private prologue (shared-memory set-up, TMEM allocation) and block in pe_grid_dep_wait() until the
// previous grid has completed and its memory is visible.  A kernel launched without the attribute sees both calls as no-ops.
// PE_PDL=0 in the environment launches everything fully serialised.
bool pe_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pe_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pe_grid_dep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t pe_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pe_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// launchers (defined in the .cu files)
int pe_launch_resid_simt(const pe_plan* plan, const PeResidArgs& a, int K, int slots, cudaStream_t st);
int pe_launch_fields(const pe_plan* plan, const PeFieldsArgs& a, int K, cudaStream_t st);
int pe_launch_fields_tcf(const pe_plan* plan, const PeFieldsArgs& a, cudaStream_t st);
int pe_simt_smem_bytes(const PeLayout& lay, int K);
bool pe_simt_stage_weights(const pe_plan* plan, int K);
int pe_simt_ctas_per_sm(const pe_plan* plan, int K);
