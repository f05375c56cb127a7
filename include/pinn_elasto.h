/*
 * pinn_elasto.h -- C ABI of the B200-native PINN-elastodynamics residual/training engine.
 *
 * The reference (Raocp/PINN-elastodynamics) has no FFI: its boundary is the Python class surface
 * (PINN / DeepHPM / DeepElasticWave) whose hot path is a TensorFlow-1 graph.
 * Each entry point below replaces the `sess.run` of one group of reference graph nodes; the reference
 * lines it replaces are cited per function (paths relative to the reference root):
 *     plate = PlateHoleQuarter/train/train.py        semi = ElasticWaveSemiInfinite/ElasticWave.py
 *     inf   = ElasticWaveInfinite/ElasticWave.py     conf = ElasticWaveConfined/ElasticWave.py
 *
 * Conventions
 *   - plain C: pointers, ints, floats.  No torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - every `d_*` pointer is DEVICE memory owned by the caller (the Python host allocates it with torch);
 *     every `h_*` pointer is host memory.  All device work is asynchronous on `stream`.
 *   - return value: 0 = ok, non-zero = error; pe_last_error() gives the message (thread-local).
 *   - network parameters live on the device in a PADDED layout (rows of W_l padded to a multiple of 4
 *     floats so that 128-bit loads are aligned): [W_0 | W_1 | ... | W_L | b_0 | ... | b_L], W_l stored
 *     row-major (in, out) like the reference (`plate:263`), pad entries are zero and stay zero.
 *     pe_pack_params / pe_unpack_params convert from/to the reference's compact order, which is also the
 *     ScipyOptimizerInterface packing (`var_list = weights + biases`, plate:241).
 *   - a "point set" is a row-major float array [n, ld]: columns 0..2 are (x, y, t) as in the reference's
 *     Collo/IC/... arrays (plate:45-88), further columns are per-point targets (SRC u,v: semi:52-53).
 *   - jet streams carried per point: K=5 (value, d/dx, d/dy, d/dt, d2/dt2) for the plate formulation,
 *     K=4 (value, d/dx, d/dy, d/dt) for the wave formulation, K=1 for primal-only boundary/data terms,
 *     K=2 (value, d/dt) for the d/dt data terms of the plate pre-training losses.
 */
#ifndef PINN_ELASTO_H
#define PINN_ELASTO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_MAX_LAYERS 16        /* weight matrices per network */
#define PE_MAX_TERMS 8          /* loss-term accumulators per model (loss_f_uv, loss_f_s, loss_IC, ...) */
#define PE_MAX_COLS 8           /* residual columns of one data term */
#define PE_TILE_POINTS 32       /* points per CTA tile of the SIMT engine (one warp lane per point) */
#define PE_TC_TILE 128           /* points per CTA tile of the tensor-core engines (UMMA M) */

/* residual kinds (what the point set contributes to the loss) */
enum {
    PE_RES_F5 = 0,       /* plane-stress momentum + constitutive residuals, 5 outputs  (plate:404-439) */
    PE_RES_F7 = 1,       /* plane-strain first-order system, 7 outputs              (semi:228-272, inf:221-265, conf:304-348) */
    PE_RES_COLS = 2,     /* selected output columns minus optional targets          (semi:103-107,119-126; conf:131-148; plate:194-198,201-215) */
    PE_RES_TRACTION = 3, /* hole traction tx, ty with n = -(x,y)/r                   (plate:452-461) */
    PE_RES_DT = 4        /* d/dt of selected output columns                          (plate:331-345 net_dist_dt, plate:354-355 net_part) */
};

/* precision / engine selection for the residual kernels */
enum {
    PE_ENGINE_SIMT_FP32 = 0,   /* fp32 FFMA kernels: parity anchor, any width */
    PE_ENGINE_TCS_TF32X3 = 5,  /* TF32x3 tcgen05 engine (csrc/pe_tcs.cu): warp-specialised (12 epilogue warps + MMA/TMA issuer warp), jet-stream
                                  groups pipelined through the forward pass, TMA-fed double-buffered weight images; PE_RES_F5 (K = 5) and
                                  PE_RES_F7 (K = 4), hidden widths <= 56.  Measured 5e-6 .. 8e-6 against the reference goldens: kept as the A/B
                                  partner of PE_ENGINE_TCF, not selected by 'auto' */
    PE_ENGINE_TCS_TF32 = 6,    /* same, single-pass TF32 (1e-3 class) */
    PE_ENGINE_TCF_F16FWD = 9,  /* PE_ENGINE_TCF with the FORWARD layer GEMMs as single fp16 products (11-bit significands, fp32 accumulation): the
                                  "16-bit forward / fp32 gradient" mode of BASELINE config 3.  (The config names bf16; mixed-format MMAs are illegal on
                                  B200 and the gradient GEMMs need the fp16 pair planes anyway, so the forward uses their hi halves: fp16, which is
                                  at least as precise as bf16.)  Adjoint and weight-gradient GEMMs stay fp32-grade.  ~1e-3 class on the loss. */
    PE_ENGINE_TCF = 8,         /* fp16-pair tcgen05 engine (csrc/pe_tcf.cu): every GEMM operand as an fp16 pair (hi, lo scaled by 2^11), products
                                  Ah Bh + 2^-11 (Ah Bl + Al Bh) on kind::f16 MMAs with the scale-input-d form, per-tile power-of-two seed
                                  scaling, TMA-fed weight gradient without a conversion pass; PE_RES_F5 (K = 5) and PE_RES_F7 (K = 4), hidden
                                  widths <= 56, at least two hidden layers */
};

typedef struct pe_plan pe_plan; /* host-side description of one network: dims, padded layout, launch config */

/* One loss contribution of one point set.  Replaces the graph nodes named by `kind`. */
typedef struct pe_term_desc {
    int kind;                    /* PE_RES_* */
    int n_global;                /* rows of the GLOBAL point set: the mean's denominator (tf.reduce_mean, plate:187);
                                    with index sharding each rank passes its shard but the global count */
    int ld;                      /* row stride of the point array in floats (>= 3) */
    /* material (plate:39-42, semi:35-37) */
    float E, mu, rho, hole_r;
    /* input normalisation H = 2(X-lb)/(ub-lb)-1 (inf:191): a0 = x*in_scale + in_shift; identity = {1,1,1},{0,0,0} */
    float in_scale[3], in_shift[3];
    /* F5 / F7: accumulate loss_f_uv into term slot term[0] with weight w[0], loss_f_s into term[1] with w[1]
       (plate:187-191,217; semi:112-118,127).  TRACTION: tx^2+ty^2 into term[0] (plate:192-193).
       COLS / DT: column c: residual = out[col[c]] (or its d/dt) - (tgt[c] >= 0 ? row[tgt[c]] : 0), into term[c], weight w[c]. */
    int ncols;
    int col[PE_MAX_COLS];
    int tgt[PE_MAX_COLS];
    int term[PE_MAX_COLS];
    float w[PE_MAX_COLS];
    /* hard-BC composite u = P + D*N (plate:382-387): d_aux holds per point the jets of the frozen
       dist/part nets, layout [n][2][aux_k][5] (D first, then P); aux_k = streams stored (K of this term). 0 = no composite */
    int aux_k;
} pe_term_desc;

/* ---------------------------------------------------------------- library */
int pe_version(void);
const char *pe_last_error(void);

/* sizeof(pe_term_desc) as compiled into the library: a binding (ctypes, cgo, ...) must assert that its mirror matches */
int pe_abi_sizeof_term_desc(void);

/* ---------------------------------------------------------------- plan (host only) */
/* dims = [3, w1, ..., wL-1, O] as the reference's `uv_layers` (plate:885); device = CUDA ordinal or -1 for
   "no device" (layout queries only; lets the CPU test-suite exercise the host logic).  NULL on error. */
pe_plan *pe_plan_create(const int *dims, int n_dims, int device);
void pe_plan_destroy(pe_plan *plan);
int pe_plan_param_count(const pe_plan *plan);          /* compact count P = sum d_l*d_{l+1} + sum d_{l+1} */
int pe_plan_param_count_padded(const pe_plan *plan);   /* padded device length (floats) */
int pe_plan_weight_offset(const pe_plan *plan, int layer);  /* offset of W_l in the padded vector */
int pe_plan_bias_offset(const pe_plan *plan, int layer);
int pe_plan_weight_ld(const pe_plan *plan, int layer);      /* padded row stride of W_l */
/* does `engine` (PE_ENGINE_*) implement residual `kind` with K streams for this network?  The tensor-core engines cover the
   PE_RES_F5 and PE_RES_F7 collocation terms of networks with hidden widths <= 56; the SIMT engine covers everything. */
int pe_engine_supported(const pe_plan *plan, int kind, int K, int engine);
/* number of CTAs (= gradient-partial slots) a launch over n points uses, and its scratch size in floats */
int pe_plan_slots(const pe_plan *plan, int n_points, int K, int engine);
size_t pe_plan_scratch_floats(const pe_plan *plan, int n_points, int K, int engine);
/* compact (reference order, weights then biases; W_l row-major (in,out)) <-> padded, host memory */
int pe_pack_params(const pe_plan *plan, const float *h_compact, float *h_padded);
int pe_unpack_params(const pe_plan *plan, const float *h_padded, float *h_compact);

/* ---------------------------------------------------------------- hot path */
/* Fused forward-jet MLP + residual + MSE partial sums + reverse sweep (weight/bias gradients) over one
 * point set.  Replaces, per Adam step / L-BFGS evaluation, the reference's
 *   net_f_sig / net_uv / net_e graph and its tf.gradients replay      (plate:404-439, 358-396)
 *   data / boundary terms                                              (plate:452-461, semi:103-107)
 *   reduce_mean(square(.)) partials                                    (plate:187-193)
 *   and the reverse-mode of all of the above w.r.t. uv weights+biases  (plate:249-250 minimize()).
 * Slot s in [slot_base, slot_base + pe_plan_slots()) of d_grad_partials ([slots][padded P]) and
 * d_term_partials ([slots][PE_MAX_TERMS]) is fully overwritten; nothing is accumulated across calls.
 * d_stash: scratch of pe_plan_scratch_floats() floats, private to this launch while it runs (launches on one
 * stream may share it): the per-CTA stash of hidden activations (kept L2-resident) + tensor-core operand images.
 * engine: PE_ENGINE_*.  */
int pe_residual_loss_grad(const pe_plan *plan, const pe_term_desc *term, int K, int engine,
                          const float *d_points, int n_local, const float *d_aux,
                          const float *d_params,
                          float *d_grad_partials, float *d_term_partials, float *d_stash,
                          int slot_base, void *stream);

/* Tensor-core engines only.  pe_residual_loss_grad for the F5 collocation set plus a SECOND, primal-only point set
 * (PE_RES_TRACTION, plate:452-461, or PE_RES_COLS; K = 1) processed by the same launch as extra 128-point tiles: its value
 * stream rides the K = 5 tiles and the derivative streams get zero seeds.  Saves one latency-bound launch per step (the
 * extra tiles usually disappear in the tail of the persistent grid).  Slots / scratch: query pe_plan_slots /
 * pe_plan_scratch_floats with n_points = PE_TC_TILE * (ceil(n/PE_TC_TILE) + ceil(n2/PE_TC_TILE)). */
int pe_residual_loss_grad_fused(const pe_plan *plan, const pe_term_desc *term, int K, int engine,
                                const float *d_points, int n_local, const float *d_aux,
                                const pe_term_desc *term2, const float *d_points2, int n2_local, const float *d_aux2,
                                const float *d_params,
                                float *d_grad_partials, float *d_term_partials, float *d_stash,
                                int slot_base, void *stream);

/* Deterministic fixed-order sum over slots: d_out[0..Pp) = sum_s partials[s], d_out[Pp..Pp+PE_MAX_TERMS) =
 * sum_s term partials.  d_out is the buffer a multi-GPU caller all-reduces.  If d_terms_copy is
 * non-NULL the PE_MAX_TERMS reduced term values are also written there (per-step loss history row). */
int pe_reduce_partials(const pe_plan *plan, const float *d_grad_partials, const float *d_term_partials,
                       int n_slots, float *d_out, float *d_terms_copy, void *stream);

/* tf.train.AdamOptimizer update in TF1 form (plate:249-250):
 *   lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps).
 * d_step points to TWO device ints {step, ticket}: step counts completed updates (the kernel uses step+1 and
 * increments it when its last block retires, so the launch can be replayed from a CUDA graph); ticket must be
 * zero-initialised.  d_grad = first Pp floats of the reduced buffer. */
int pe_adam_step(const pe_plan *plan, float *d_params, const float *d_grad, float *d_m, float *d_v,
                 int *d_step, float lr, float beta1, float beta2, float eps, void *stream);

/* pe_reduce_partials + pe_adam_step in one kernel (single-GPU path: no collective in between); also
 * writes the reduced gradient and terms to d_out. */
int pe_reduce_adam(const pe_plan *plan, const float *d_grad_partials, const float *d_term_partials,
                   int n_slots, float *d_out, float *d_terms_copy, float *d_params, float *d_m, float *d_v, int *d_step,
                   float lr, float beta1, float beta2, float eps, void *stream);

/* Forward-only fields for `predict` (plate:561-570; semi:348-358): d_out[n][8] =
 * (u, v, s11, s22, s12, e11, e22, e12); formulation PE_RES_F5 or PE_RES_F7 selects the output columns
 * (F7 nets return cols 0,1,4,5,6).  Composite as in pe_term_desc (aux_k = 4 streams: value,x,y,t).
 * Batches of >= 512 points of networks the tensor-core engine supports (hidden widths <= 56, at least two hidden
 * layers) run its forward sweep (operand-image kernel + resid_tcf_kernel<4, FWD>; the images live in a buffer the plan
 * allocates at the first such call, so one plan must not be used from two streams at once); everything else, and
 * everything under PE_FIELDS_ENGINE=simt, runs the fp32 SIMT fields kernel.  Both meet the fields bar (2e-5). */
int pe_forward_fields(const pe_plan *plan, int formulation, const float *d_points, int ld, int n,
                      const float *in_scale, const float *in_shift, const float *d_aux, int aux_k,
                      const float *d_params, float *d_out, void *stream);

/* Forward-only jets of all outputs: d_out[n][K][O].  Used to pre-compute the frozen dist/part jets of the
 * plate composite (plate:361-362: the two extra neural_net calls inside net_uv) and by predict_D/predict_P. */
int pe_forward_jets(const pe_plan *plan, int K, const float *d_points, int ld, int n,
                    const float *in_scale, const float *in_shift,
                    const float *d_params, float *d_out, void *stream);

/* ---------------------------------------------------------------- multi-GPU step over NVLink peer memory
 * The reference is single-device; data parallelism here shards every point set by index and needs ONE sum of
 * [grad | terms] per step.  Instead of reduce -> ncclAllReduce -> Adam (three launches and NCCL's latency) the ranks of one
 * node exchange their slices through IPC-mapped peer memory inside one kernel: slot reduction, push to all peers,
 * flag wait, rank-ordered sum (bit-identical on every rank), Adam.  Host protocol: every rank calls pe_comm_create,
 * all-gathers the 64-byte handles (any transport), calls pe_comm_connect, then pe_reduce_peer once per step in lockstep.
 * Teardown: every rank calls pe_comm_disconnect (unmaps the peers), then a barrier, then pe_comm_destroy (frees its region). */
#define PE_MAX_PEERS 8
#define PE_IPC_HANDLE_BYTES 64
typedef struct pe_comm pe_comm;
pe_comm *pe_comm_create(const pe_plan *plan, int rank, int world, unsigned char *handle_out /* 64 bytes */);
int pe_comm_connect(pe_comm *comm, const unsigned char *all_handles /* world x 64 bytes, rank order */);
/* non-zero if a kernel gave up waiting for a peer (about 20 s); synchronises the device. */
int pe_comm_error(pe_comm *comm);
void pe_comm_disconnect(pe_comm *comm);
void pe_comm_destroy(pe_comm *comm);
/* pe_reduce_partials + all-reduce (+ pe_adam_step when d_params != NULL) in one launch; arguments as pe_reduce_adam. */
int pe_reduce_peer(const pe_plan *plan, pe_comm *comm, const float *d_grad_partials, const float *d_term_partials,
                   int n_slots, float *d_out, float *d_terms_copy, float *d_params, float *d_m, float *d_v, int *d_step,
                   float lr, float beta1, float beta2, float eps, void *stream);

/* ---------------------------------------------------------------- device-resident L-BFGS
 * Vector algebra of the limited-memory BFGS driver that replaces the SciPy round trip of
 * ScipyOptimizerInterface.minimize (plate:240-247, 522-525; semi:151-156; conf:263-268).  All vectors are
 * padded parameter vectors of length n = pe_plan_param_count_padded (pads stay zero); S, Y are [m][n] ring
 * buffers (m = maxcor <= 64); d_state = m + 3 floats: [gamma | rho_0..rho_{m-1} | y.s | y.y of the last pair]. */
/* d_dir = -H d_g by the two-loop recursion over the `count` newest pairs; `head` = slot of the newest pair. */
int pe_lbfgs_direction(int n, int m, int count, int head, const float *d_g, const float *d_S, const float *d_Y,
                       const float *d_state, float *d_dir, void *stream);
/* slot `head` <- (s, y) = (x - x_prev, g - g_prev); updates rho[head] and gamma in d_state (rho = 0 when y.s <= 0). */
int pe_lbfgs_store_pair(int n, int m, int head, const float *d_x, const float *d_xprev, const float *d_g,
                        const float *d_gprev, float *d_S, float *d_Y, float *d_state, void *stream);
/* d_out = d_x + alpha * d_d (the line-search trial point, written straight into the parameter buffer). */
int pe_vec_axpy(int n, float *d_out, const float *d_x, float alpha, const float *d_d, void *stream);
/* d_res[0] = a . b (double accumulation), d_res[1] = max |a_i| (the projected-gradient norm of the pgtol test). */
int pe_vec_dot_max(int n, const float *d_a, const float *d_b, float *d_res, void *stream);

/* ---------------------------------------------------------------- debug / profiling */
/* Per-phase cycle counters of the tensor-core residual kernels: d_counters32 = 32 device uint64 (or NULL to switch off): 0..15 phases of
 * epilogue thread 0, 16..31 phases of the MMA/TMA issuer of CTA 0 (tests/tcf_gpu_check.py prof; selects the profiling instantiation). */
void pe_debug_set_tcs_profile(unsigned long long *d_counters32);     /* PE_ENGINE_TCS_* */
void pe_debug_set_tcf_profile(unsigned long long *d_counters32);     /* PE_ENGINE_TCF */
/* engine of pe_forward_fields: PE_ENGINE_TCF (default: batches of >= 512 points of networks the tensor-core engine supports run its forward
 * sweep, everything else the SIMT kernel), PE_ENGINE_SIMT_FP32, or -1 = re-read $PE_FIELDS_ENGINE ("simt" | "tcf") at the next call */
void pe_debug_set_fields_engine(int engine);

#ifdef __cplusplus
}
#endif
#endif /* PINN_ELASTO_H */
