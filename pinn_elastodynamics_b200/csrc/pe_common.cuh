// Shared definitions of the sm_100a PINN-elastodynamics kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
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

// launchers (defined in the .cu files)
int pe_launch_resid_simt(const pe_plan* plan, const PeResidArgs& a, int K, int slots, cudaStream_t st);
int pe_launch_fields(const pe_plan* plan, const PeFieldsArgs& a, int K, cudaStream_t st);
int pe_simt_smem_bytes(const PeLayout& lay, int K);
bool pe_simt_stage_weights(const pe_plan* plan, int K);
int pe_simt_ctas_per_sm(const pe_plan* plan, int K);
