// Fused persistent engine (placeholder until the kernels land): reports "unsupported" so the
// API layer routes to the generic engine.
#include "ising.cuh"
namespace dq {
int fused_supported(const dq_ising*) { return 0; }
int fused_grad_run(dq_ising*) { set_error("fused engine not built"); return DQ_ERR_UNSUPPORTED; }
int fused_evolve(dq_ising*, c128*, int, const double*, int, double*, bool) {
    set_error("fused engine not built");
    return DQ_ERR_UNSUPPORTED;
}
}  // namespace dq
