// Dense path entry points (placeholder until the DMMA kernels land).
#include "common.cuh"
extern "C" {
int dq_dense_set_H(dq_context*, int, const double*, int, const double*, const int32_t*, const double*, double, int) {
    dq::set_error("dense path not built yet"); return DQ_ERR_UNSUPPORTED; }
int dq_dense_trotter(dq_context*, const double*, double, double, int, const double*, int, int, double*, double*) {
    dq::set_error("dense path not built yet"); return DQ_ERR_UNSUPPORTED; }
int dq_dense_evolve(dq_context*, int, const double*, int, const double*, const double*, int, double, int, int,
                    const double*, double*) {
    dq::set_error("dense path not built yet"); return DQ_ERR_UNSUPPORTED; }
int dq_dense_grad(dq_context*, int, const double*, int, const double*, const double*, const double*, double, int,
                  const int32_t*, const double*, const double*, const int32_t*, const double*, const double*, int,
                  double*) {
    dq::set_error("dense path not built yet"); return DQ_ERR_UNSUPPORTED; }
}
