// tcgen05 / TMEM / TMA bf16 GEMM (placeholder until the kernel lands; the dispatcher falls back
// to the SIMT kernel when gemm_tc_why_not() returns a reason).
#include "gemm_common.cuh"

namespace decaf {

const char *gemm_tc_why_not(const GemmArgs &a, int dtype) {
    (void)a; (void)dtype;
    return "tcgen05 kernel not built";
}

int gemm_tc_launch(const GemmArgs &a, int n_group, cudaStream_t st) {
    (void)a; (void)n_group; (void)st;
    set_error("tcgen05 kernel not built");
    return 1;
}

}  // namespace decaf
