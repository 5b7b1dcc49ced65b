// Device-resident training loop of the structured (Pauli-term) path: SURVEY 8(f) rank 1.
//
// SimulatorPlain.train_energy (sim_plain.py:245-305) per epoch: one full evolution and its energy (:276-281), one stochastic
// gradient sample (:290 -> compute_energy_grad_MC, :156-231), an Adam step (:266,:291-292), the log line against the
// observable's lowest eigenvalue (:294; here min(m_diag), a constant the caller passes).  On the structured path every
// piece already runs on the device except the pulse arithmetic; this file moves that over too, so the whole loop is
// enqueued on the context's stream with no host round trip between epochs:
//     k_ising_train_rows   angle rows of the epoch's 1 + 2K trajectories from the coefficients AS THEY ARE ON THE DEVICE
//                          (generate_u, sim_plain.py:73-99, then IsingProblem.angle_rows: base + pulses in term order, * dt)
//     fused_evolve / gen_* the full evolution, energy only (final state of the LAST epoch kept, :304)
//     fused_grad_run / ... the K samples' 2 n_H shifted trajectories each (the staged-batch driver of dq_ising_grad)
//     k_train_update       gradient mean + torch-style Adam + the epoch's log value (train_update.cuh)
// The sample times are drawn by the HOST beforehand from the reference's own stream (one np.random.uniform() per sample,
// sim_plain.py:167); step counts follow from them, so the per-epoch descriptors need nothing from the device.
#include <algorithm>
#include <cmath>
#include <vector>
#include "ising.cuh"
#include "train_update.cuh"

namespace dq {
namespace {

struct IsingRowTraj {           // one trajectory's rows: t_k = t0 (+= dt) k times -> out[which] rows row0 .. row0 + n - 1
    double t0, dt;
    long long row0;
    int n, which;               // which: 0 full-evolution table, 1 prefix table (rows_a), 2 suffix table (rows_b)
};

__global__ void k_ising_train_rows(const IsingRowTraj* __restrict__ tr, double T, const double* __restrict__ coeff,
                                   const double* __restrict__ omegas, int n_terms, int n_basis, const double* __restrict__ bl,
                                   const double* __restrict__ br, double norm_factor, const int* __restrict__ term_col,
                                   const double* __restrict__ row_base, int row_len, double* __restrict__ out_full,
                                   double* __restrict__ out_a, double* __restrict__ out_b) {
    const IsingRowTraj t = tr[blockIdx.x];
    double* __restrict__ table = t.which == 0 ? out_full : (t.which == 1 ? out_a : out_b);
    for (int k = threadIdx.x; k < t.n; k += blockDim.x) {
        double tt = t.t0;
        for (int i = 0; i < k; ++i) tt = __dadd_rn(tt, t.dt);            // the reference's repeated t += dt
        const double x = tt / T;
        double* __restrict__ row = table + (t.row0 + k) * row_len;
        for (int c = 0; c < row_len; ++c) row[c] = row_base[c];          // [c0 | w_e | 0]
        for (int i = 0; i < n_terms; ++i) {                              // pulses added in term order (angle_rows)
            double a = 0.0;
            for (int j = 0; j < n_basis; ++j) a = __dadd_rn(a, __dmul_rn(train_bump(x, bl[j], br[j], norm_factor), coeff[i * n_basis + j]));
            const double sg = 1.0 / (1.0 + exp(-a));
            const double ui = __dmul_rn(__dadd_rn(__dmul_rn(sg, 2.0), -1.0), omegas[i]);
            row[term_col[i]] = __dadd_rn(row[term_col[i]], ui);
        }
        for (int c = 0; c < row_len; ++c) row[c] = __dmul_rn(row[c], t.dt);
    }
}

__global__ void k_copy_state(const c128* __restrict__ src, c128* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace
}  // namespace dq

using dq::c128;

extern "C" int dq_ising_train(dq_ising* p, int n_terms, const int32_t* term_kind, const int32_t* term_index, const double* omegas,
                              const double* h0_zz, double h0_const, double T, int per_step, int n_basis, double* coeff_inout,
                              int n_epoch, int K, const double* s_all, double lr, double beta1, double beta2, double eps, double r,
                              double e0, const double* psi0, double* losses_out, double* final_state_out, double* device_ms_out) {
    using namespace dq;
    DQ_REQUIRE(p && term_kind && term_index && omegas && coeff_inout && s_all && losses_out, "dq_ising_train: NULL argument");
    DQ_REQUIRE(n_terms >= 1 && n_basis >= 3 && per_step >= 1 && n_epoch >= 1 && K >= 1,
               "dq_ising_train: n_terms=%d n_basis=%d per_step=%d n_epoch=%d K=%d", n_terms, n_basis, per_step, n_epoch, K);
    DQ_REQUIRE(n_terms * n_basis <= 1024, "dq_ising_train: %d coefficients exceed one update block (1024)", n_terms * n_basis);
    DQ_REQUIRE(T > 0 && std::isfinite(T) && r > 0 && std::isfinite(r), "dq_ising_train: T and r must be positive");
    DQ_TRY(p->ctx->set_device());
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim();
    const int row_len = p->row_len;

    // ---- per-epoch trajectories (step counts follow from the sample times: sim_plain.py:123,133) ---------------------------
    const int n_full = (int)(per_step * ((T - 0.0) + 1));
    DQ_REQUIRE(n_full >= 1, "dq_ising_train: per_step=%d gives no steps", per_step);
    const int rt_per = 1 + 2 * K;
    std::vector<IsingRowTraj> trajs((size_t)n_epoch * rt_per);
    std::vector<int32_t> pre_steps((size_t)n_epoch * K), suf_steps((size_t)n_epoch * K);
    long long max_a = 1, max_b = 1;
    double dt_max = T / n_full;
    for (int e = 0; e < n_epoch; ++e) {
        IsingRowTraj* R = trajs.data() + (size_t)e * rt_per;
        R[0] = IsingRowTraj{0.0, T / n_full, 0, n_full, 0};
        long long ra = 0, rb = 0;
        for (int k = 0; k < K; ++k) {
            const double sk = s_all[(size_t)e * K + k];
            DQ_REQUIRE(std::isfinite(sk) && sk >= 0.0 && sk <= T, "dq_ising_train: sample time %g outside [0, T]", sk);
            const int np = (int)(per_step * ((sk - 0.0) + 1)), ns = (int)(per_step * ((T - sk) + 1));
            const double dtp = np > 0 ? sk / np : 0.0, dts = ns > 0 ? (T - sk) / ns : 0.0;
            dt_max = std::max(dt_max, std::max(dtp, dts));
            R[1 + k] = IsingRowTraj{0.0, dtp, ra, np, 1};
            R[1 + K + k] = IsingRowTraj{sk, dts, rb, ns, 2};
            pre_steps[(size_t)e * K + k] = np;
            suf_steps[(size_t)e * K + k] = ns;
            ra += np;
            rb += ns;
        }
        max_a = std::max(max_a, ra);
        max_b = std::max(max_b, rb);
    }
    // ---- row template, term columns, angle bounds ------------------------------------------------------------------------------
    std::vector<double> base(row_len, 0.0), x_sum(p->n, 0.0);
    std::vector<int> col(n_terms);
    base[0] = h0_const;
    if (h0_zz) for (int e = 0; e < p->n_zz; ++e) base[1 + e] = h0_zz[e];
    double bound_all = std::fabs(h0_const);
    for (int e = 0; e < p->n_zz; ++e) bound_all += std::fabs(base[1 + e]);
    for (int i = 0; i < n_terms; ++i) {
        if (term_kind[i] == 0) {
            DQ_REQUIRE(term_index[i] >= 0 && term_index[i] < p->n_zz, "dq_ising_train: term %d: ZZ pair %d of %d", i, term_index[i], p->n_zz);
            col[i] = 1 + term_index[i];
        } else {
            DQ_REQUIRE(term_kind[i] == 1 && term_index[i] >= 0 && term_index[i] < p->n, "dq_ising_train: term %d: bad X control", i);
            col[i] = 1 + p->n_zz + term_index[i];
            x_sum[term_index[i]] += std::fabs(omegas[i]);
        }
        bound_all += std::fabs(omegas[i]);
    }
    const double x_max = *std::max_element(x_sum.begin(), x_sum.end());
    const bool scaled = std::max(dt_max * x_max, std::fabs(std::atan(r))) <= 1.0;      // |u_i| <= |omega_i| (sigmoid range)
    const double exact_bound = dt_max * bound_all;

    // ---- one device arena --------------------------------------------------------------------------------------------------------
    std::vector<double> bl(n_basis), br(n_basis);
    const double tau = 1. / (n_basis - 2);
    for (int b = 0; b < n_basis; ++b) {
        const double tau_b = tau * (b - 1.5);
        bl[b] = tau_b - 1.5 * tau;
        br[b] = tau_b + 1.5 * tau;
    }
    const double norm_factor = -((1.5 * tau) * (1.5 * tau));
    const size_t nc = (size_t)n_terms * n_basis;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_coeff = take(nc * 8), o_m1 = take(nc * 8), o_m2 = take(nc * 8), o_om = take(n_terms * 8), o_bl = take(n_basis * 8),
                 o_br = take(n_basis * 8), o_base = take(row_len * 8), o_col = take(n_terms * 4), o_s = take((size_t)n_epoch * K * 8),
                 o_loss = take((size_t)n_epoch * 8), o_efull = take(8), o_full = take((size_t)n_full * row_len * 8),
                 o_tr = take(trajs.size() * sizeof(IsingRowTraj)), o_trig = take((size_t)n_full * p->n * sizeof(double2)),
                 o_state = take(N * sizeof(c128));
    DQ_TRY(p->train.reserve(off));
    DQ_TRY(p->rows_a.reserve((size_t)max_a * row_len * 8));
    DQ_TRY(p->rows_b.reserve((size_t)max_b * row_len * 8));
    char* basep = p->train.as<char>();
    DQ_CUDA(cudaMemsetAsync(basep + o_m1, 0, o_om - o_m1, st));           // Adam state
    auto up = [&](size_t o, const void* src, size_t bytes) { return cudaMemcpyAsync(basep + o, src, bytes, cudaMemcpyHostToDevice, st); };
    DQ_CUDA(up(o_coeff, coeff_inout, nc * 8));
    DQ_CUDA(up(o_om, omegas, n_terms * 8));
    DQ_CUDA(up(o_bl, bl.data(), n_basis * 8));
    DQ_CUDA(up(o_br, br.data(), n_basis * 8));
    DQ_CUDA(up(o_base, base.data(), row_len * 8));
    DQ_CUDA(up(o_col, col.data(), n_terms * 4));
    DQ_CUDA(up(o_s, s_all, (size_t)n_epoch * K * 8));
    DQ_CUDA(up(o_tr, trajs.data(), trajs.size() * sizeof(IsingRowTraj)));
    double* d_coeff = (double*)(basep + o_coeff);
    double* d_full = (double*)(basep + o_full);
    double* d_efull = (double*)(basep + o_efull);
    c128* d_state = (c128*)(basep + o_state);
    const bool fused = engine_is_fused(p);
    // psi0 -> device (physical order) once, through the staging helper; shift descriptors once
    DQ_TRY(stage_meta(p, K, pre_steps.data(), suf_steps.data(), n_terms, term_kind, term_index, r, psi0));
    const int copy_grid = (int)std::min<size_t>((N + 255) / 256, (size_t)p->ctx->prop.multiProcessorCount * 8);

    cudaEvent_t ev0, ev1;
    DQ_CUDA(cudaEventCreate(&ev0));
    DQ_CUDA(cudaEventCreate(&ev1));
    int rc = DQ_OK;
    cudaEventRecord(ev0, st);
    // ---- the loop: enqueue only ----------------------------------------------------------------------------------------------------
    for (int e = 0; e < n_epoch && rc == DQ_OK; ++e) {
        k_ising_train_rows<<<rt_per, 64, 0, st>>>((const IsingRowTraj*)(basep + o_tr) + (size_t)e * rt_per, T, d_coeff,
                                                  (const double*)(basep + o_om), n_terms, n_basis, (const double*)(basep + o_bl),
                                                  (const double*)(basep + o_br), norm_factor, (const int*)(basep + o_col),
                                                  (const double*)(basep + o_base), row_len, d_full, p->rows_a.as<double>(),
                                                  p->rows_b.as<double>());
        p->ctx->launches++;
        // full evolution (:276) from psi0; the state itself is only needed after the last epoch (:304)
        if (p->st.uniform_psi0) rc = gen_fill_uniform(p, d_state, 1);
        else { k_copy_state<<<copy_grid, 256, 0, st>>>(p->st.psi0.as<c128>(), d_state, N); p->ctx->launches++; }
        if (rc != DQ_OK) break;
        const bool want_state = final_state_out && e == n_epoch - 1;
        if (fused) {
            rc = fused_evolve(p, d_state, 1, nullptr, n_full, d_efull, want_state, d_full, scaled ? 1 : 0);
        } else {
            if (p->step_mode == 1) {
                rc = gen_evolve_exact(p, d_state, 1, d_full, nullptr, n_full, exact_bound);
            } else {
                rc = gen_trig(p, d_full, n_full, (double2*)(basep + o_trig));
                if (rc == DQ_OK) rc = gen_evolve(p, d_state, 1, d_full, (const double2*)(basep + o_trig), n_full);
            }
            if (rc == DQ_OK) rc = gen_energy(p, d_state, 1, d_efull);
        }
        if (rc != DQ_OK) break;
        // the K gradient samples of this epoch (:290): same staged-batch driver as dq_ising_grad, rows already in place
        auto& s = p->st;
        for (int k = 0; k < K; ++k) {
            s.prefix_steps[k] = pre_steps[(size_t)e * K + k];
            s.suffix_steps[k] = suf_steps[(size_t)e * K + k];
            s.prefix_off[k + 1] = s.prefix_off[k] + s.prefix_steps[k];
            s.suffix_off[k + 1] = s.suffix_off[k] + s.suffix_steps[k];
        }
        s.scaled_ok = scaled;
        s.exact_bound = exact_bound;
        rc = stage_trig(p);
        if (rc != DQ_OK) break;
        s.valid = true;
        rc = dq_ising_grad_run_staged(p);
        if (rc != DQ_OK) break;
        const double bc1 = 1.0 - pow(beta1, e + 1), bc2 = 1.0 - pow(beta2, e + 1);
        k_train_update<<<1, (unsigned)((nc + 31) / 32 * 32), 0, st>>>(
            p->energies.as<double>(), (const double*)(basep + o_s) + (size_t)e * K, K, d_coeff, (double*)(basep + o_m1),
            (double*)(basep + o_m2), (const double*)(basep + o_om), T, n_terms, n_basis, (const double*)(basep + o_bl),
            (const double*)(basep + o_br), norm_factor, r, beta1, beta2, eps, lr / bc1, sqrt(bc2), d_efull, e0,
            (double*)(basep + o_loss), e);
        p->ctx->launches++;
    }
    p->st.valid = false;                         // the staged tables belong to the loop
    p->st.exact_bound = -1.0;
    if (rc == DQ_OK && cudaGetLastError() != cudaSuccess) { set_error("dq_ising_train: launch failed"); rc = DQ_ERR_CUDA; }
    cudaEventRecord(ev1, st);
    if (rc == DQ_OK) {
        cudaMemcpyAsync(losses_out, basep + o_loss, (size_t)n_epoch * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(coeff_inout, d_coeff, nc * 8, cudaMemcpyDeviceToHost, st);
        if (final_state_out) {
            const c128* src = d_state;
            if (!p->identity_layout) {
                rc = p->io.reserve(N * sizeof(c128));
                if (rc == DQ_OK) rc = gen_permute_out(p, d_state, p->io.as<c128>(), 1);
                src = p->io.as<c128>();
            }
            if (rc == DQ_OK) cudaMemcpyAsync(final_state_out, src, N * sizeof(c128), cudaMemcpyDeviceToHost, st);
        }
    }
    const cudaError_t sync = cudaStreamSynchronize(st);
    if (rc == DQ_OK && sync != cudaSuccess) { set_error("dq_ising_train: %s", cudaGetErrorString(sync)); rc = DQ_ERR_CUDA; }
    float ms = 0.f;
    if (rc == DQ_OK) cudaEventElapsedTime(&ms, ev0, ev1);
    if (device_ms_out) *device_ms_out = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return rc;
}
