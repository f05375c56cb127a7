// tcgen05 / TMEM tensor-core engine (placeholder until the UMMA path lands; reports "unsupported").
#include "pe_common.cuh"

int pe_tc_supported(const pe_plan* plan, int K, int engine) { (void)plan; (void)K; (void)engine; return 0; }
int pe_launch_resid_tc(const pe_plan* plan, const PeResidArgs& a, int K, int engine, int slots, cudaStream_t st) {
    (void)plan; (void)a; (void)K; (void)engine; (void)slots; (void)st;
    pe_set_error("tensor-core engine not built");
    return 1;
}
