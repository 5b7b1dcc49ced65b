// Device-resident training loop of the dense path (dim <= 16): SURVEY 8(f) rank 1.
//
// SimulatorPlain.train_energy (sim_plain.py:245-305) does, per epoch: one full evolution and its energy (:276-281), one
// stochastic gradient sample (:290 -> compute_energy_grad_MC, :156-231), an Adam step (:266,:291-292) and a dense
// eigendecomposition of the observable for the log line (:294).  Here the whole loop is enqueued on the context's stream
// without a single host round trip between epochs:
//     k_train_rows    pulse rows of the epoch's trajectories from the coefficients AS THEY ARE ON THE DEVICE (generate_u,
//                     sim_plain.py:73-99, same operation order as k_pulse_rows) + the norm bound of the epoch
//     k_train_scale   bound -> squarings of the Taylor series (device int read by the engine kernels)
//     k_small         prefix kets phi_k and the full-evolution energy; k_small_mma: every shifted ket's energy
//     k_train_update  grad[i][j] = mean_k ps_k[i] dDdv_k[i][j]  (sim_plain.py:169-184,220-227), then torch's Adam update
//                     (single-tensor form: lerp, addcmul, bias corrections as lr / bc1 and sqrt(v) / sqrt(bc2) + eps)
// The sample times are drawn by the HOST beforehand from the reference's own stream (one np.random.uniform() per epoch,
// sim_plain.py:167); the step counts of every trajectory follow from them, so all descriptors are built and uploaded once.
// The observable's lowest eigenvalue (:294 recomputes it every epoch) is a constant passed in.
#include <algorithm>
#include <cmath>
#include <string.h>
#include <vector>
#include "dense.cuh"
#include "train_update.cuh"

namespace dq {
namespace dense {
namespace {

struct RowTraj {                // one trajectory's pulse rows: t_k = t0 (+= dt) k times, rows row0 .. row0 + n - 1
    double t0, dt;
    long long row0;
    int n, pad_;
};

__device__ __forceinline__ double bump(double x, double l, double r, double norm_factor) {
    if (x >= r || x <= l) return 0.0;                       // open support, sim_plain.py:62
    return __dmul_rn(__dadd_rn(x, -l), __dadd_rn(x, -r)) / norm_factor;
}

__global__ void k_train_rows(const RowTraj* __restrict__ tr, double T, const double* __restrict__ coeff,
                             const double* __restrict__ omegas, int n_H, int n_basis, const double* __restrict__ bl,
                             const double* __restrict__ br, double norm_factor, double* __restrict__ u,
                             const double* __restrict__ norm1, int sum_mode, unsigned long long* __restrict__ bound_bits) {
    const RowTraj t = tr[blockIdx.x];
    for (int k = threadIdx.x; k < t.n; k += blockDim.x) {
        double tt = t.t0;
        for (int i = 0; i < k; ++i) tt = __dadd_rn(tt, t.dt);            // the reference's repeated t += dt
        const double x = tt / T;
        double* __restrict__ out = u + (t.row0 + k) * n_H;
        double nb = norm1[0];
        for (int i = 0; i < n_H; ++i) {
            double a = 0.0;
            for (int j = 0; j < n_basis; ++j) a = __dadd_rn(a, __dmul_rn(bump(x, bl[j], br[j], norm_factor), coeff[i * n_basis + j]));
            const double sg = 1.0 / (1.0 + exp(-a));
            const double ui = __dmul_rn(__dadd_rn(__dmul_rn(sg, 2.0), -1.0), omegas[i]);
            out[i] = ui;
            nb = sum_mode ? nb + fabs(ui) * norm1[i + 1] : fmax(nb, fabs(ui) * norm1[i + 1]);
        }
        atomicMax(bound_bits, (unsigned long long)__double_as_longlong(fabs(t.dt) * nb));
    }
}

__global__ void k_train_scale(unsigned long long* __restrict__ bound_bits, int* __restrict__ s_out) {
    const double bound = __longlong_as_double((long long)*bound_bits);
    int s = 0;
    while (ldexp(bound, -s) > 1.0 && s < 20) ++s;
    *s_out = s;
    *bound_bits = 0ull;                                      // ready for the next epoch
}

}  // namespace
}  // namespace dense
}  // namespace dq

using namespace dq::dense;
using dq::k_train_update;

extern "C" int dq_dense_train(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                              const double* psi0, const double* omegas, double T, int per_step, int n_basis, double* coeff_inout,
                              int n_epoch, int K, const double* s_all, double lr, double beta1, double beta2, double eps, double r,
                              double e0, int mode, double* losses_out, double* final_state_out) {
    DQ_REQUIRE(ctx && H0 && M && psi0 && omegas && coeff_inout && s_all && losses_out, "dq_dense_train: NULL argument");
    DQ_REQUIRE(n_H >= 1 && n_basis >= 3 && per_step >= 1 && n_epoch >= 1 && K >= 1, "dq_dense_train: n_H=%d n_basis=%d per_step=%d n_epoch=%d K=%d",
               n_H, n_basis, per_step, n_epoch, K);
    DQ_REQUIRE(n_H * n_basis <= 1024, "dq_dense_train: %d coefficients exceed one update block (1024)", n_H * n_basis);
    DQ_REQUIRE(mode == 0 || mode == 1, "dq_dense_train: mode must be 0 (exact) or 1 (split)");
    DQ_REQUIRE(T > 0 && std::isfinite(T) && r > 0 && std::isfinite(r), "dq_dense_train: T and r must be positive");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->scratch_H;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    DQ_REQUIRE(small_fits(P), "dq_dense_train: the device-resident loop runs on the resident engine (dim <= 16), got %d", dim);
    DQ_TRY(small_upload(ctx, P, M));
    cudaStream_t st = ctx->stream;
    const int m = 18;
    const bool mma = mode == 0 && ctx->dense_small_mma != 0;
    const int nk = mma ? (2 * n_H > 8 ? 16 : 8) : ((n_H % 2 == 0) ? 4 : 2);
    const int warps_per_sample = (n_H + nk / 2 - 1) / (nk / 2);

    // ---- descriptors of every epoch (step counts follow from the sample times: sim_plain.py:123,133) --------------------------
    const int n_full = (int)(per_step * ((T - 0.0) + 1));
    DQ_REQUIRE(n_full >= 1, "dq_dense_train: per_step=%d gives no steps", per_step);
    const int rt_per = 1 + 2 * K, kt_per = 1 + K, ft_per = 1, st_per = K * warps_per_sample;
    std::vector<RowTraj> rows((size_t)n_epoch * rt_per);
    std::vector<SmallTraj> kets((size_t)n_epoch * kt_per), full((size_t)n_epoch * ft_per), suf((size_t)n_epoch * st_per);
    long long max_rows = 0;
    for (int e = 0; e < n_epoch; ++e) {
        long long row = 0;
        RowTraj* R = rows.data() + (size_t)e * rt_per;
        R[0] = RowTraj{0.0, T / n_full, row, n_full, 0};
        // ket slots: 0 = psi0, 1 = full-evolution state, 2 .. 1 + K = phi_k
        kets[(size_t)e * kt_per] = SmallTraj{row, T / n_full, 0.0, n_full, 0, 0, 1};
        full[(size_t)e] = SmallTraj{row, T / n_full, 0.0, n_full, 0, 0, 0};
        row += n_full;
        for (int k = 0; k < K; ++k) {
            const double sk = s_all[(size_t)e * K + k];
            DQ_REQUIRE(std::isfinite(sk) && sk >= 0.0 && sk <= T, "dq_dense_train: sample time %g outside [0, T]", sk);
            const int np = (int)(per_step * ((sk - 0.0) + 1)), ns = (int)(per_step * ((T - sk) + 1));
            const double dtp = np > 0 ? sk / np : 0.0, dts = ns > 0 ? (T - sk) / ns : 0.0;
            R[1 + k] = RowTraj{0.0, dtp, row, np, 0};
            kets[(size_t)e * kt_per + 1 + k] = SmallTraj{row, dtp, 0.0, np, 0, 0, 2 + k};
            row += np;
            R[1 + K + k] = RowTraj{sk, dts, row, ns, 0};
            for (int w = 0; w < warps_per_sample; ++w)
                suf[(size_t)e * st_per + (size_t)k * warps_per_sample + w] =
                    SmallTraj{row, dts, r, ns, 2 + k, w * (nk / 2), (k * n_H + w * (nk / 2)) * 2};
            row += ns;
        }
        max_rows = std::max(max_rows, row);
    }
    // ---- one device arena -------------------------------------------------------------------------------------------------------
    std::vector<double> bl(n_basis), br(n_basis);
    const double tau = 1. / (n_basis - 2);
    for (int b = 0; b < n_basis; ++b) {
        const double tau_b = tau * (b - 1.5);
        bl[b] = tau_b - 1.5 * tau;
        br[b] = tau_b + 1.5 * tau;
    }
    const double norm_factor = -((1.5 * tau) * (1.5 * tau));
    const size_t nc = (size_t)n_H * n_basis;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_coeff = take(nc * 8), o_m1 = take(nc * 8), o_m2 = take(nc * 8), o_om = take(n_H * 8), o_bl = take(n_basis * 8),
                 o_br = take(n_basis * 8), o_n1 = take((1 + n_H) * 8), o_s = take((size_t)n_epoch * K * 8), o_loss = take((size_t)n_epoch * 8),
                 o_bound = take(8), o_sint = take(4), o_efull = take(8), o_en = take((size_t)K * n_H * 2 * 8),
                 o_u = take((size_t)std::max<long long>(1, max_rows) * n_H * 8), o_kets = take((size_t)(2 + K) * 32 * 8),
                 o_rows = take(rows.size() * sizeof(RowTraj)), o_kt = take(kets.size() * sizeof(SmallTraj)),
                 o_ft = take(full.size() * sizeof(SmallTraj)), o_st = take(suf.size() * sizeof(SmallTraj));
    DQ_TRY(S->train.reserve(off));
    char* base = S->train.as<char>();
    DQ_CUDA(cudaMemsetAsync(base, 0, o_s, st));              // Adam state, bound
    DQ_CUDA(cudaMemsetAsync(base + o_bound, 0, 16, st));
    auto up = [&](size_t o, const void* src, size_t bytes) { return cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, st); };
    DQ_CUDA(up(o_coeff, coeff_inout, nc * 8));
    DQ_CUDA(up(o_om, omegas, n_H * 8));
    DQ_CUDA(up(o_bl, bl.data(), n_basis * 8));
    DQ_CUDA(up(o_br, br.data(), n_basis * 8));
    DQ_CUDA(up(o_n1, P.norm1.data(), (1 + n_H) * 8));
    DQ_CUDA(up(o_s, s_all, (size_t)n_epoch * K * 8));
    DQ_CUDA(up(o_rows, rows.data(), rows.size() * sizeof(RowTraj)));
    DQ_CUDA(up(o_kt, kets.data(), kets.size() * sizeof(SmallTraj)));
    DQ_CUDA(up(o_ft, full.data(), full.size() * sizeof(SmallTraj)));
    DQ_CUDA(up(o_st, suf.data(), suf.size() * sizeof(SmallTraj)));
    std::vector<double> k0(32, 0.0);
    memcpy(k0.data(), psi0, sizeof(double) * 2 * dim);
    DQ_CUDA(up(o_kets, k0.data(), 32 * 8));
    double* d_coeff = (double*)(base + o_coeff);
    double* d_u = (double*)(base + o_u);
    double* d_kets = (double*)(base + o_kets);
    int* d_s = (int*)(base + o_sint);
    if (!S->ev0) {
        DQ_CUDA(cudaEventCreate(&S->ev0));
        DQ_CUDA(cudaEventCreate(&S->ev1));
    }
    DQ_CUDA(cudaEventRecord(S->ev0, st));
    // ---- the loop: enqueue only ---------------------------------------------------------------------------------------------------
    for (int e = 0; e < n_epoch; ++e) {
        k_train_rows<<<rt_per, 64, 0, st>>>((const RowTraj*)(base + o_rows) + (size_t)e * rt_per, T, d_coeff, (const double*)(base + o_om),
                                            n_H, n_basis, (const double*)(base + o_bl), (const double*)(base + o_br), norm_factor, d_u,
                                            (const double*)(base + o_n1), mode == 0 ? 1 : 0, (unsigned long long*)(base + o_bound));
        k_train_scale<<<1, 1, 0, st>>>((unsigned long long*)(base + o_bound), d_s);
        ctx->launches += 2;
        DQ_TRY(small_enqueue(ctx, P, mode, 0, m, 1, (const SmallTraj*)(base + o_kt) + (size_t)e * kt_per, kt_per, d_u, d_kets, d_kets, nullptr, 1.0, d_s));
        DQ_TRY(small_enqueue(ctx, P, mode, 0, m, 1, (const SmallTraj*)(base + o_ft) + (size_t)e, 1, d_u, d_kets, nullptr, (double*)(base + o_efull), 1.0, d_s));
        DQ_TRY(small_enqueue(ctx, P, mode, 0, m, nk, (const SmallTraj*)(base + o_st) + (size_t)e * st_per, st_per, d_u, d_kets, nullptr,
                             (double*)(base + o_en), 1.0 / sqrt(1.0 + r * r), d_s));
        const double bc1 = 1.0 - pow(beta1, e + 1), bc2 = 1.0 - pow(beta2, e + 1);
        k_train_update<<<1, (unsigned)((nc + 31) / 32 * 32), 0, st>>>(
            (const double*)(base + o_en), (const double*)(base + o_s) + (size_t)e * K, K, d_coeff, (double*)(base + o_m1), (double*)(base + o_m2),
            (const double*)(base + o_om), T, n_H, n_basis, (const double*)(base + o_bl), (const double*)(base + o_br), norm_factor, r, beta1,
            beta2, eps, lr / bc1, sqrt(bc2), (const double*)(base + o_efull), e0, (double*)(base + o_loss), e);
        ctx->launches++;
    }
    DQ_CUDA(cudaGetLastError());
    DQ_CUDA(cudaEventRecord(S->ev1, st));
    DQ_CUDA(cudaMemcpyAsync(losses_out, base + o_loss, (size_t)n_epoch * 8, cudaMemcpyDeviceToHost, st));
    DQ_CUDA(cudaMemcpyAsync(coeff_inout, d_coeff, nc * 8, cudaMemcpyDeviceToHost, st));
    std::vector<double> fin(32, 0.0);
    if (final_state_out) DQ_CUDA(cudaMemcpyAsync(fin.data(), d_kets + 32, 32 * 8, cudaMemcpyDeviceToHost, st));   // state of the LAST epoch's evolution (:303)
    DQ_CUDA(cudaStreamSynchronize(st));
    if (final_state_out) memcpy(final_state_out, fin.data(), sizeof(double) * 2 * dim);
    float ms = 0.f;
    DQ_CUDA(cudaEventElapsedTime(&ms, S->ev0, S->ev1));
    S->last_kernel_ms = ms;
    S->last_strategy = 3;
    S->last_degree = m;
    return DQ_OK;
}
