// C ABI of libpinn_elasto.so (see include/pinn_elasto.h for the contract and the reference lines each
// entry point replaces).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "pe_common.cuh"

static thread_local char g_err[512] = "";

void pe_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pe_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PE_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int pe_launch_resid_tcs(const pe_plan* plan, const PeResidArgs& a, int K, int fast, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2);
int pe_launch_resid_tcf(const pe_plan* plan, const PeResidArgs& a, int K, int fast, int slots, cudaStream_t st,
                        const pe_term_desc* term2, const float* points2, int n2, const float* aux2);

// ---- sizing shared by the tcgen05 engines: 128-point tiles, one persistent CTA per SM (= one gradient-partial slot each); per slot a stash
// of (L-1) layers x 5 streams x 28,672 B (fp32 [chunk of 4][128][4] planes in pe_tcs.cu, fp16-pair planes in pe_tcf.cu: same bytes), then the
// operand images of every matrix (the larger of the two engines' formats: 2 x 36,864 B per matrix)
static const int TCG_STASH_LAYER_BYTES = 5 * 28672, TCG_IMG_LAYER_BYTES = 2 * 36864;
int pe_tc_slots(const pe_plan* plan, int n_points) {      // for a fused launch pass 128 * (tiles of set 1 + tiles of set 2)
    int ntiles = (n_points + PE_TC_TILE - 1) / PE_TC_TILE;
    int s = ntiles < plan->sms ? ntiles : plan->sms;
    return s < 1 ? 1 : s;
}
size_t pe_tc_stash_floats_per_slot(const pe_plan* plan) { return (size_t)(plan->lay.L - 1) * (TCG_STASH_LAYER_BYTES / 4) + 64; }
size_t pe_tc_image_floats(const pe_plan* plan) { return (size_t)plan->lay.L * (TCG_IMG_LAYER_BYTES / 4); }
// F5 (K = 5, 5 outputs) and F7 (K = 4, 7 outputs) collocation terms of networks with hidden widths <= 56 (7 chunks of 8 units per operand plane)
static int tcgen_supported(const pe_plan* plan, int K) {
    const PeLayout& lay = plan->lay;
    const int O = lay.d[lay.L];
    if (!((K == 5 && O == 5) || (K == 4 && O == 7)) || lay.L < 2) return 0;
    for (int l = 1; l < lay.L; ++l)
        if (lay.d[l] > 56) return 0;
    return 1;
}

extern "C" int pe_version(void) { return 100; }
extern "C" int pe_abi_sizeof_term_desc(void) { return (int)sizeof(pe_term_desc); }
extern "C" const char* pe_last_error(void) { return g_err; }

static int build_layout(PeLayout& lay, const int* dims, int n_dims) {
    if (n_dims < 3 || n_dims > PE_MAX_LAYERS + 1) { pe_set_error("need 3..%d layer sizes, got %d", PE_MAX_LAYERS + 1, n_dims); return 1; }
    if (dims[0] != 3) { pe_set_error("input width must be 3 (x, y, t), got %d", dims[0]); return 1; }
    memset(&lay, 0, sizeof(lay));
    lay.L = n_dims - 1;
    for (int i = 0; i < n_dims; ++i) {
        if (dims[i] < 1 || dims[i] > 512) { pe_set_error("layer width %d out of range", dims[i]); return 1; }
        lay.d[i] = dims[i];
        lay.lda[i] = pe_lda(dims[i]);
    }
    if (lay.d[lay.L] > PE_UJ) { pe_set_error("output width %d > %d not supported", lay.d[lay.L], PE_UJ); return 1; }
    int off = 0, compact = 0;
    for (int l = 0; l < lay.L; ++l) {
        lay.ldw[l] = pe_round4(lay.d[l + 1]);
        lay.woff[l] = off;
        off += lay.d[l] * lay.ldw[l];
        compact += lay.d[l] * lay.d[l + 1] + lay.d[l + 1];
    }
    for (int l = 0; l < lay.L; ++l) {
        lay.boff[l] = off;
        off += lay.ldw[l];
    }
    lay.total = off + 128;       // slack: vector loads of partial unit groups may run past the last row (never used)
    lay.compact = compact;
    lay.maxw = 0;
    lay.max_lda = lay.lda[0];
    int rows = 0;
    for (int l = 1; l <= lay.L; ++l) {
        if (l < lay.L) {
            if (lay.d[l] > lay.maxw) lay.maxw = lay.d[l];
            lay.soff[l] = rows;
            rows += lay.lda[l];
        }
        if (lay.lda[l] > lay.max_lda) lay.max_lda = lay.lda[l];
    }
    lay.stash_rows = rows;
    int wm = 0, wl = 0;
    for (int l = 0; l < lay.L; ++l) {
        // rows rounded up to whole 10-unit groups: the adjoint GEMM of a partial group reads (and discards) rows past the matrix
        const int rows = (lay.d[l] + PE_UJ - 1) / PE_UJ * PE_UJ;
        if (rows * lay.ldw[l] + 16 > wm) wm = rows * lay.ldw[l] + 16;
        if (lay.ldw[l] > wl) wl = lay.ldw[l];
    }
    lay.wmat_floats = wm;
    lay.wstage_floats = wm + wl;
    int mw = lay.maxw > lay.d[lay.L] ? lay.maxw : lay.d[lay.L];
    lay.groups = (mw + PE_UJ - 1) / PE_UJ;
    if (lay.groups > 16) { pe_set_error("hidden width %d too large (max %d)", lay.maxw, 16 * PE_UJ); return 1; }
    return 0;
}

extern "C" pe_plan* pe_plan_create(const int* dims, int n_dims, int device) {
    pe_plan* p = new (std::nothrow) pe_plan();
    if (!p) { pe_set_error("out of memory"); return nullptr; }
    if (build_layout(p->lay, dims, n_dims)) { delete p; return nullptr; }
    p->device = device;
    p->sms = 148;
    p->smem_optin = 227 * 1024;
    if (device >= 0) {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) { pe_set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e)); delete p; return nullptr; }
        p->sms = prop.multiProcessorCount;
        p->smem_optin = (int)prop.sharedMemPerBlockOptin;
        if (prop.major != 10) { pe_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); delete p; return nullptr; }
    }
    return p;
}

extern "C" void pe_plan_destroy(pe_plan* plan) {
    if (plan && plan->d_tc_images) cudaFree(plan->d_tc_images);
    delete plan;
}
extern "C" int pe_plan_param_count(const pe_plan* plan) { return plan ? plan->lay.compact : -1; }
extern "C" int pe_plan_param_count_padded(const pe_plan* plan) { return plan ? plan->lay.total : -1; }
extern "C" int pe_plan_weight_offset(const pe_plan* plan, int l) { return (plan && l >= 0 && l < plan->lay.L) ? plan->lay.woff[l] : -1; }
extern "C" int pe_plan_bias_offset(const pe_plan* plan, int l) { return (plan && l >= 0 && l < plan->lay.L) ? plan->lay.boff[l] : -1; }
extern "C" int pe_plan_weight_ld(const pe_plan* plan, int l) { return (plan && l >= 0 && l < plan->lay.L) ? plan->lay.ldw[l] : -1; }

static bool is_tcs(int engine) { return engine == PE_ENGINE_TCS_TF32X3 || engine == PE_ENGINE_TCS_TF32; }
static bool is_tcf(int engine) { return engine == PE_ENGINE_TCF || engine == PE_ENGINE_TCF_F16FWD; }
static bool is_tc(int engine) { return is_tcs(engine) || is_tcf(engine); }

extern "C" int pe_engine_supported(const pe_plan* plan, int kind, int K, int engine) {
    if (!plan) return 0;
    if (engine == PE_ENGINE_SIMT_FP32) return 1;
    if (is_tcf(engine) && plan->lay.L < 3) return 0;
    if (is_tc(engine)) return (kind == PE_RES_F5 || kind == PE_RES_F7) && tcgen_supported(plan, K);
    return 0;
}

extern "C" int pe_plan_slots(const pe_plan* plan, int n_points, int K, int engine) {
    if (!plan) return -1;
    if (is_tc(engine)) return pe_tc_slots(plan, n_points);
    int ntiles = (n_points + PE_P - 1) / PE_P;
    int cap = plan->sms * pe_simt_ctas_per_sm(plan, K);
    int s = ntiles < cap ? ntiles : cap;
    return s < 1 ? 1 : s;
}

static size_t simt_stash_floats_per_slot(const pe_plan* plan, int K) { return (size_t)K * PE_P * plan->lay.stash_rows + 64; }

extern "C" size_t pe_plan_scratch_floats(const pe_plan* plan, int n_points, int K, int engine) {
    if (!plan) return 0;
    size_t slots = (size_t)pe_plan_slots(plan, n_points, K, engine);
    if (is_tc(engine)) return slots * pe_tc_stash_floats_per_slot(plan) + pe_tc_image_floats(plan) + 64;
    return slots * simt_stash_floats_per_slot(plan, K);
}

extern "C" int pe_pack_params(const pe_plan* plan, const float* h_compact, float* h_padded) {
    if (!plan || !h_compact || !h_padded) { pe_set_error("null argument"); return 1; }
    const PeLayout& lay = plan->lay;
    memset(h_padded, 0, sizeof(float) * lay.total);
    int o = 0;
    for (int l = 0; l < lay.L; ++l)
        for (int i = 0; i < lay.d[l]; ++i)
            for (int j = 0; j < lay.d[l + 1]; ++j) h_padded[lay.woff[l] + i * lay.ldw[l] + j] = h_compact[o++];
    for (int l = 0; l < lay.L; ++l)
        for (int j = 0; j < lay.d[l + 1]; ++j) h_padded[lay.boff[l] + j] = h_compact[o++];
    return 0;
}

extern "C" int pe_unpack_params(const pe_plan* plan, const float* h_padded, float* h_compact) {
    if (!plan || !h_compact || !h_padded) { pe_set_error("null argument"); return 1; }
    const PeLayout& lay = plan->lay;
    int o = 0;
    for (int l = 0; l < lay.L; ++l)
        for (int i = 0; i < lay.d[l]; ++i)
            for (int j = 0; j < lay.d[l + 1]; ++j) h_compact[o++] = h_padded[lay.woff[l] + i * lay.ldw[l] + j];
    for (int l = 0; l < lay.L; ++l)
        for (int j = 0; j < lay.d[l + 1]; ++j) h_compact[o++] = h_padded[lay.boff[l] + j];
    return 0;
}

static int check_term(const pe_plan* plan, const pe_term_desc* t, int K) {
    const int O = plan->lay.d[plan->lay.L];
    if (t->ld < 3) { pe_set_error("point row stride %d < 3", t->ld); return 1; }
    if (t->n_global < 1) { pe_set_error("n_global must be >= 1"); return 1; }
    switch (t->kind) {
        case PE_RES_F5: if (K != 5 || O != 5) { pe_set_error("PE_RES_F5 needs K=5 and 5 outputs (K=%d, O=%d)", K, O); return 1; } break;
        case PE_RES_F7: if (K != 4 || O != 7) { pe_set_error("PE_RES_F7 needs K=4 and 7 outputs (K=%d, O=%d)", K, O); return 1; } break;
        case PE_RES_TRACTION: if (K != 1 || O != 5) { pe_set_error("PE_RES_TRACTION needs K=1 and 5 outputs"); return 1; } break;
        case PE_RES_COLS: if (K != 1) { pe_set_error("PE_RES_COLS needs K=1"); return 1; } break;
        case PE_RES_DT: if (K != 2) { pe_set_error("PE_RES_DT needs K=2"); return 1; } break;
        default: pe_set_error("unknown residual kind %d", t->kind); return 1;
    }
    if (t->kind == PE_RES_COLS || t->kind == PE_RES_DT) {
        if (t->ncols < 1 || t->ncols > PE_MAX_COLS) { pe_set_error("ncols %d out of range", t->ncols); return 1; }
        for (int c = 0; c < t->ncols; ++c) {
            if (t->col[c] < 0 || t->col[c] >= O) { pe_set_error("column %d out of range", t->col[c]); return 1; }
            if (t->tgt[c] >= t->ld) { pe_set_error("target column %d >= row stride %d", t->tgt[c], t->ld); return 1; }
            if (t->term[c] < 0 || t->term[c] >= PE_MAX_TERMS) { pe_set_error("term slot %d out of range", t->term[c]); return 1; }
        }
    } else {
        int nt = (t->kind == PE_RES_TRACTION) ? 1 : 2;
        for (int c = 0; c < nt; ++c)
            if (t->term[c] < 0 || t->term[c] >= PE_MAX_TERMS) { pe_set_error("term slot %d out of range", t->term[c]); return 1; }
    }
    if (t->aux_k != 0) {
        if (!(t->kind == PE_RES_F5 && t->aux_k == 5) && !(t->kind == PE_RES_TRACTION && t->aux_k == 1)) {
            pe_set_error("composite aux_k=%d not valid for kind %d", t->aux_k, t->kind); return 1;
        }
    }
    return 0;
}

static int residual_common(const pe_plan* plan, const pe_term_desc* term, int K, int engine,
                           const float* d_points, int n_local, const float* d_aux,
                           const pe_term_desc* term2, const float* d_points2, int n2_local, const float* d_aux2,
                           const float* d_params, float* d_grad_partials, float* d_term_partials, float* d_stash,
                           int slot_base, void* stream) {
    if (!plan || !term) { pe_set_error("null plan/term"); return 1; }
    if (plan->device < 0) { pe_set_error("plan was created without a device"); return 1; }
    if (check_term(plan, term, K)) return 1;
    if (n_local < 0) { pe_set_error("negative point count"); return 1; }
    if (term->aux_k && !d_aux) { pe_set_error("composite requested but d_aux is null"); return 1; }
    PeResidArgs a;
    a.lay = plan->lay;
    a.term = *term;
    a.points = d_points;
    a.aux = term->aux_k ? d_aux : nullptr;
    a.params = d_params;
    a.grad_partials = d_grad_partials;
    a.term_partials = d_term_partials;
    a.stash = d_stash;
    a.n = n_local;
    a.slot_base = slot_base;
    a.stash_floats = (int)simt_stash_floats_per_slot(plan, K);
    a.inv_n = 1.0f / (float)term->n_global;
    if (engine == PE_ENGINE_SIMT_FP32) {
        if (term2) { pe_set_error("fused point sets need a tensor-core engine"); return 1; }
        return pe_launch_resid_simt(plan, a, K, pe_plan_slots(plan, n_local, K, engine), (cudaStream_t)stream);
    }
    if (is_tc(engine)) {
        if (!pe_engine_supported(plan, term->kind, K, engine)) { pe_set_error("tensor-core engine %d does not support this residual kind / network (F5 K=5 or F7 K=4, hidden widths <= 56)", engine); return 1; }
        int n_eff = n_local;
        if (term2) {
            if (check_term(plan, term2, 1)) return 1;
            if (term2->kind != PE_RES_TRACTION && term2->kind != PE_RES_COLS) { pe_set_error("fused set must be PE_RES_TRACTION or PE_RES_COLS"); return 1; }
            if (term2->aux_k && !d_aux2) { pe_set_error("fused composite set needs d_aux2"); return 1; }
            n_eff = PE_TC_TILE * ((n_local + PE_TC_TILE - 1) / PE_TC_TILE + (n2_local + PE_TC_TILE - 1) / PE_TC_TILE);
        }
        if (is_tcf(engine))
            return pe_launch_resid_tcf(plan, a, K, engine == PE_ENGINE_TCF_F16FWD ? 1 : 0, pe_plan_slots(plan, n_eff, K, engine), (cudaStream_t)stream, term2, d_points2, n2_local, d_aux2);
        return pe_launch_resid_tcs(plan, a, K, engine == PE_ENGINE_TCS_TF32 ? 1 : 0, pe_plan_slots(plan, n_eff, K, engine), (cudaStream_t)stream, term2, d_points2, n2_local, d_aux2);
    }
    pe_set_error("unknown engine %d", engine);
    return 1;
}

extern "C" int pe_residual_loss_grad(const pe_plan* plan, const pe_term_desc* term, int K, int engine,
                                     const float* d_points, int n_local, const float* d_aux,
                                     const float* d_params,
                                     float* d_grad_partials, float* d_term_partials, float* d_stash,
                                     int slot_base, void* stream) {
    return residual_common(plan, term, K, engine, d_points, n_local, d_aux, nullptr, nullptr, 0, nullptr,
                           d_params, d_grad_partials, d_term_partials, d_stash, slot_base, stream);
}

extern "C" int pe_residual_loss_grad_fused(const pe_plan* plan, const pe_term_desc* term, int K, int engine,
                                           const float* d_points, int n_local, const float* d_aux,
                                           const pe_term_desc* term2, const float* d_points2, int n2_local, const float* d_aux2,
                                           const float* d_params,
                                           float* d_grad_partials, float* d_term_partials, float* d_stash,
                                           int slot_base, void* stream) {
    if (!term2) { pe_set_error("null term2"); return 1; }
    return residual_common(plan, term, K, engine, d_points, n_local, d_aux, term2, d_points2, n2_local, d_aux2,
                           d_params, d_grad_partials, d_term_partials, d_stash, slot_base, stream);
}

// engine of pe_forward_fields: PE_ENGINE_TCF (default; batches of >= 512 points of networks the tensor-core engine supports) or PE_ENGINE_SIMT_FP32;
// -1 = take $PE_FIELDS_ENGINE at the next call
static int g_fields_engine = -1;
extern "C" void pe_debug_set_fields_engine(int engine) { g_fields_engine = engine; }

extern "C" int pe_forward_fields(const pe_plan* plan, int formulation, const float* d_points, int ld, int n,
                                 const float* in_scale, const float* in_shift, const float* d_aux, int aux_k,
                                 const float* d_params, float* d_out, void* stream) {
    if (!plan) { pe_set_error("null plan"); return 1; }
    if (plan->device < 0) { pe_set_error("plan was created without a device"); return 1; }
    const int O = plan->lay.d[plan->lay.L];
    if ((formulation == PE_RES_F5 && O != 5) || (formulation == PE_RES_F7 && O != 7) ||
        (formulation != PE_RES_F5 && formulation != PE_RES_F7)) { pe_set_error("formulation %d does not match %d outputs", formulation, O); return 1; }
    if (aux_k && (aux_k != 4 || formulation != PE_RES_F5 || !d_aux)) { pe_set_error("composite fields need aux_k=4, F5 and d_aux"); return 1; }
    if (n <= 0) return 0;
    PeFieldsArgs a;
    a.lay = plan->lay;
    a.points = d_points; a.aux = d_aux; a.params = d_params; a.out = d_out;
    a.n = n; a.ld = ld; a.aux_k = aux_k; a.mode = 0; a.formulation = formulation;
    for (int i = 0; i < 3; ++i) { a.in_scale[i] = in_scale ? in_scale[i] : 1.f; a.in_shift[i] = in_shift ? in_shift[i] : 0.f; }
    // tensor-core forward sweep for frame batches (the reference predicts 251 x 251 x 81 points, plate:980-998); small batches, wide nets and
    // PE_FIELDS_ENGINE=simt take the fp32 SIMT kernel
    if (g_fields_engine < 0) { const char* e = getenv("PE_FIELDS_ENGINE"); g_fields_engine = (e && strcmp(e, "simt") == 0) ? PE_ENGINE_SIMT_FP32 : PE_ENGINE_TCF; }
    if (g_fields_engine == PE_ENGINE_TCF && n >= 4 * PE_TC_TILE && plan->lay.L >= 3 && tcgen_supported(plan, formulation == PE_RES_F5 ? 5 : 4))
        return pe_launch_fields_tcf(plan, a, (cudaStream_t)stream);
    return pe_launch_fields(plan, a, 4, (cudaStream_t)stream);
}

extern "C" int pe_forward_jets(const pe_plan* plan, int K, const float* d_points, int ld, int n,
                               const float* in_scale, const float* in_shift,
                               const float* d_params, float* d_out, void* stream) {
    if (!plan) { pe_set_error("null plan"); return 1; }
    if (plan->device < 0) { pe_set_error("plan was created without a device"); return 1; }
    if (n <= 0) return 0;
    PeFieldsArgs a;
    a.lay = plan->lay;
    a.points = d_points; a.aux = nullptr; a.params = d_params; a.out = d_out;
    a.n = n; a.ld = ld; a.aux_k = 0; a.mode = 1; a.formulation = 0;
    for (int i = 0; i < 3; ++i) { a.in_scale[i] = in_scale ? in_scale[i] : 1.f; a.in_shift[i] = in_shift ? in_shift[i] : 0.f; }
    return pe_launch_fields(plan, a, K, (cudaStream_t)stream);
}
