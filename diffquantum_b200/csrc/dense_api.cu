// C-ABI entry points and host orchestration of the dense path (dq_dense_*).
//
// exp(A) K for A = -i dt H(t_k), K a block of kets, is evaluated one of three ways, chosen per call
// from the flop counts (all three are sequences of the same batched DMMA complex GEMM):
//   0  block-Taylor : K <- T_m(A/2^s) applied 2^s times to the block          (2^s m GEMMs  Dp x Dp x Ncp)
//   1  propagator   : P = T_m(A/2^s), s squarings, K <- P K                   (m-1+s GEMMs  Dp^3, one Dp x Dp x Ncp)
//   2  chained      : as 1 but U <- P U per step and one K <- U K at the end  (more kets than dim)
//   3  resident     : dim <= 16 only - one warp per ket, whole trajectory in one launch (dense_small.cu); the automatic
//                     choice there, because 16 x 16 GEMM launches are latency, not arithmetic
// The polynomial is evaluated in Horner form P <- I + (A/j) P, so every term is one GEMM with a fused
// "+ I" / "+ K" epilogue.  A is skew-Hermitian, so squaring is norm-preserving and well conditioned.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <string.h>
#include "dense.cuh"

namespace dq {
namespace dense {

State* state_of(dq_context* ctx) {
    if (!ctx->dense) ctx->dense = new State();
    return ctx->dense;
}

void release(dq_context* ctx) {
    State* S = ctx->dense;
    if (!S) return;
    DevBuf* bufs[] = {&S->global_H.H, &S->global_H.M, &S->scratch_H.H, &S->scratch_H.M, &S->A, &S->P0, &S->P1, &S->U,
                      &S->K0, &S->K1, &S->K2, &S->u_dev, &S->meta, &S->phi, &S->out, &S->small_H, &S->small_traj, &S->train};
    for (auto* b : bufs) b->release();
    if (S->ev0) cudaEventDestroy(S->ev0);
    if (S->ev1) cudaEventDestroy(S->ev1);
    delete S;
    ctx->dense = nullptr;
}

namespace {

inline int round8(int x) { return (x + 7) & ~7; }

// interleaved c128 [dim][dim] -> planar zero-padded [2][Dp*Dp]
void to_planar(const double* src, int dim, int Dp, double* dst) {
    const size_t plane = (size_t)Dp * Dp;
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
            dst[(size_t)i * Dp + j] = src[2 * ((size_t)i * dim + j)];
            dst[plane + (size_t)i * Dp + j] = src[2 * ((size_t)i * dim + j) + 1];
        }
}

double norm1_of(const double* src, int dim) {
    double best = 0.0;
    for (int j = 0; j < dim; ++j) {
        double s = 0.0;
        for (int i = 0; i < dim; ++i) s += hypot(src[2 * ((size_t)i * dim + j)], src[2 * ((size_t)i * dim + j) + 1]);
        best = std::max(best, s);
    }
    return best;
}

}  // namespace

int upload_problem(dq_context* ctx, Problem& P, int dim, const double* H0, int n_H, const double* Hs) {
    DQ_REQUIRE(dim >= 1 && dim <= 1024, "dense path: dim=%d outside [1,1024]", dim);
    DQ_REQUIRE(n_H >= 0 && H0 && (n_H == 0 || Hs), "dense path: NULL Hamiltonian");
    const size_t per = (size_t)dim * dim * 2;
    const size_t total = per * (1 + n_H);
    for (size_t i = 0; i < per; ++i) DQ_REQUIRE(std::isfinite(H0[i]), "dense path: non-finite H0 entry");
    for (size_t i = 0; i < per * n_H; ++i) DQ_REQUIRE(std::isfinite(Hs[i]), "dense path: non-finite Hs entry");
    if (P.dim == dim && P.n_H == n_H && P.host_copy.size() == total &&
        !memcmp(P.host_copy.data(), H0, per * sizeof(double)) &&
        (n_H == 0 || !memcmp(P.host_copy.data() + per, Hs, per * n_H * sizeof(double))))
        return DQ_OK;                                   // same operators as last time: keep the device copy
    P.dim = dim;
    P.Dp = round8(dim);
    P.n_H = n_H;
    P.host_copy.assign(total, 0.0);
    memcpy(P.host_copy.data(), H0, per * sizeof(double));
    if (n_H) memcpy(P.host_copy.data() + per, Hs, per * n_H * sizeof(double));
    const size_t plane = P.plane();
    std::vector<double> planar((size_t)(1 + n_H) * 2 * plane, 0.0);
    P.norm1.assign(1 + n_H, 0.0);
    for (int h = 0; h <= n_H; ++h) {
        const double* src = P.host_copy.data() + per * h;
        to_planar(src, dim, P.Dp, planar.data() + (size_t)h * 2 * plane);
        P.norm1[h] = norm1_of(src, dim);
    }
    DQ_TRY(P.H.reserve(planar.size() * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(P.H.p, planar.data(), planar.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return DQ_OK;
}

namespace {


int upload_observable(dq_context* ctx, Problem& P, const double* M) {
    const size_t per = (size_t)P.dim * P.dim * 2;
    for (size_t i = 0; i < per; ++i) DQ_REQUIRE(std::isfinite(M[i]), "dense path: non-finite M entry");
    std::vector<double> planar(2 * P.plane(), 0.0);
    to_planar(M, P.dim, P.Dp, planar.data());
    DQ_TRY(P.M.reserve(planar.size() * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(P.M.p, planar.data(), planar.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return DQ_OK;
}

inline int log2_ceil_ratio(double x, double theta) {
    if (!(x > theta)) return 0;
    return (int)std::ceil(std::log2(x / theta));
}

// max over all steps of ||dt H(t_k)||_1 (mode 0) or of the largest single term (mode 1): decides the scaling 2^s
int norm_bound(const Problem& P, int mode, int B, const int* steps, const double* dts, const long long* row_off,
               const double* h_u, double* out) {
    const int n_H = P.n_H;
    double bound = 0.0;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < steps[b]; ++k) {
            const double* row = h_u + (row_off[b] + k) * n_H;
            double nb = mode == 0 ? P.norm1[0] : 0.0;
            for (int h = 0; h < n_H; ++h) {
                DQ_REQUIRE(std::isfinite(row[h]), "dense path: non-finite pulse value (sample %d step %d term %d)", b, k, h);
                if (mode == 0) nb += fabs(row[h]) * P.norm1[h + 1];
                else nb = std::max(nb, fabs(row[h]) * P.norm1[h + 1]);
            }
            if (mode != 0) nb = std::max(nb, P.norm1[0]);
            bound = std::max(bound, fabs(dts[b]) * nb);
        }
    *out = bound;
    return DQ_OK;
}

// strategy 3 applies: automatic for dim <= 16, or forced; forcing it on a larger problem is an error
int want_resident(dq_context* ctx, const Problem& P, bool* yes) {
    const int f = ctx->dense_force_strategy;
    DQ_REQUIRE(f != 3 || small_fits(P), "dense path: strategy 3 (resident) needs dim <= 16, got %d", P.dim);
    *yes = small_fits(P) && (f < 0 || f == 3);
    return DQ_OK;
}

// Evolve B ket blocks (sample order, planar [B][2][Dp*Ncp], in place in d_K) through their own step lists.
// steps/dts/row_off are per sample; h_u is the packed host table [total_rows][n_H].
int evolve_blocks(dq_context* ctx, Problem& P, int mode, int B, int Ncp, const int* steps, const double* dts,
                  const long long* row_off, const double* h_u, long long total_rows, double* d_K) {
    State* S = state_of(ctx);
    cudaStream_t st = ctx->stream;
    const int Dp = P.Dp, n_H = P.n_H;
    const size_t plane = P.plane(), mat = 2 * plane, blk = (size_t)2 * Dp * Ncp;
    std::vector<int> order(B);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return steps[a] > steps[b]; });
    const int max_steps = steps[order[0]];
    if (max_steps <= 0) return DQ_OK;

    // ---- norm bound -> scaling -------------------------------------------------------------------
    double bound = 0.0;
    DQ_TRY(norm_bound(P, mode, B, steps, dts, row_off, h_u, &bound));
    const int m_blk = 18, m_mat = 9;
    const int s_blk = log2_ceil_ratio(bound, 1.0), s_mat = log2_ceil_ratio(bound, 0.125);
    DQ_REQUIRE(s_blk <= 20, "dense path: ||dt H|| = %g is too large", bound);
    const double d3 = (double)Dp * Dp * Dp, d2n = (double)Dp * Dp * Ncp;
    // time model per launch: ~4 us of launch latency + tensor-pipe time (DMMA GEMM at ~12 TFLOP/s) or, for the
    // skinny mat-vec kernel, the time to stream A (16 Dp^2 bytes per batch member at ~3 TB/s)
    const double n_blk = std::ldexp((double)m_blk, s_blk), n_mat = m_mat - 1 + s_mat;
    const double per_blk = (Ncp == 8 && Dp >= 64) ? 16.0 * Dp * Dp / 3e12 : 8.0 * d2n / 4e12;
    const double cost_blk = n_blk * (4e-6 + B * per_blk);
    const double cost_mat = n_mat * (4e-6 + B * 8.0 * d3 / 12e12) + (4e-6 + B * 8.0 * d2n / 4e12);
    int strategy = cost_blk <= cost_mat ? 0 : 1;
    if (strategy == 1 && mode == 0 && Ncp > Dp) strategy = 2;
    if (ctx->dense_force_strategy >= 0 && ctx->dense_force_strategy <= 2 && (ctx->dense_force_strategy < 2 || mode == 0))
        strategy = ctx->dense_force_strategy;
    const int s = strategy == 0 ? s_blk : s_mat, m = strategy == 0 ? m_blk : m_mat;
    S->last_strategy = strategy;
    S->last_squarings = s;
    S->last_degree = m;

    // ---- device tables -----------------------------------------------------------------------------
    DQ_TRY(S->u_dev.reserve(std::max<size_t>(1, (size_t)total_rows * n_H) * sizeof(double)));
    if (total_rows > 0 && n_H > 0)
        DQ_CUDA(cudaMemcpyAsync(S->u_dev.p, h_u, (size_t)total_rows * n_H * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<long long> rows(B);
    std::vector<double> scale(B);
    for (int z = 0; z < B; ++z) {
        rows[z] = row_off[order[z]];
        scale[z] = std::ldexp(dts[order[z]], -s);
    }
    const size_t off_scale = (size_t)B * sizeof(long long), off_order = off_scale + (size_t)B * sizeof(double);
    DQ_TRY(S->meta.reserve(off_order + (size_t)B * sizeof(int)));
    char* meta = S->meta.as<char>();
    DQ_CUDA(cudaMemcpyAsync(meta, rows.data(), B * sizeof(long long), cudaMemcpyHostToDevice, st));
    DQ_CUDA(cudaMemcpyAsync(meta + off_scale, scale.data(), B * sizeof(double), cudaMemcpyHostToDevice, st));
    DQ_CUDA(cudaMemcpyAsync(meta + off_order, order.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
    const long long* d_rows = reinterpret_cast<const long long*>(meta);
    const double* d_scale = reinterpret_cast<const double*>(meta + off_scale);
    const int* d_order = reinterpret_cast<const int*>(meta + off_order);

    DQ_TRY(S->A.reserve((size_t)B * mat * sizeof(double)));
    DQ_TRY(S->K0.reserve((size_t)B * blk * sizeof(double)));
    DQ_TRY(S->K1.reserve((size_t)B * blk * sizeof(double)));
    if (strategy == 0) DQ_TRY(S->K2.reserve((size_t)B * blk * sizeof(double)));
    else {
        DQ_TRY(S->P0.reserve((size_t)B * mat * sizeof(double)));
        DQ_TRY(S->P1.reserve((size_t)B * mat * sizeof(double)));
    }
    if (strategy == 2) DQ_TRY(S->U.reserve((size_t)B * mat * sizeof(double)));
    double* A = S->A.as<double>();
    double *cur = S->K0.as<double>(), *ya = S->K1.as<double>(), *yb = S->K2.as<double>();
    double *p0 = S->P0.as<double>(), *p1 = S->P1.as<double>(), *U = S->U.as<double>();
    DQ_TRY(gather_blocks(ctx, d_K, cur, d_order, B, blk, 0));
    double flops = 0.0;

    auto gemm_mat = [&](const double* a, const double* b, double* c, int nb, double alpha, int ident) {
        Gemm g{a, (long long)mat, (long long)plane, Dp, b, (long long)mat, (long long)plane, Dp,
               c, (long long)mat, (long long)plane, Dp, nullptr, Dp, Dp, Dp, nb, alpha, ident};
        flops += 8.0 * d3 * nb;
        return zgemm(ctx, g);
    };
    auto gemm_blk = [&](const double* a, const double* b, double* c, const double* add, int nb, double alpha) {
        Gemm g{a, (long long)mat, (long long)plane, Dp, b, (long long)blk, (long long)Dp * Ncp, Ncp,
               c, (long long)blk, (long long)Dp * Ncp, Ncp, add, Dp, Ncp, Dp, nb, alpha, 0};
        flops += 8.0 * d2n * nb;
        return zgemm(ctx, g);
    };

    if (strategy == 2) {                                               // U = 0 * (0 0) + I
        DQ_CUDA(cudaMemsetAsync(A, 0, (size_t)B * mat * sizeof(double), st));
        DQ_TRY(gemm_mat(A, A, U, B, 0.0, 1));
        flops -= 8.0 * d3 * B;
    }

    int nb = B;
    for (int k = 0; k < max_steps; ++k) {
        while (nb > 0 && steps[order[nb - 1]] <= k) --nb;
        int nb_next = nb;
        while (nb_next > 0 && steps[order[nb_next - 1]] <= k + 1) --nb_next;
        const int t_lo = mode == 0 ? -1 : 0, t_hi = mode == 0 ? -1 : n_H;
        for (int term = t_lo; term <= t_hi; ++term) {
            if (strategy == 0) {
                DQ_TRY(build_generator(ctx, P, nb, S->u_dev.as<double>(), d_rows, d_scale, k, term, A, nullptr, 0.0));
                for (long long rep = 0; rep < (1LL << s); ++rep) {
                    DQ_TRY(gemm_blk(A, cur, ya, cur, nb, 1.0 / m));
                    for (int j = m - 1; j >= 1; --j) {
                        DQ_TRY(gemm_blk(A, ya, yb, cur, nb, 1.0 / j));
                        std::swap(ya, yb);
                    }
                    std::swap(cur, ya);          // finished slots keep their data: they were scattered out already
                }
            } else {
                DQ_TRY(build_generator(ctx, P, nb, S->u_dev.as<double>(), d_rows, d_scale, k, term, A, p0, 1.0 / m));
                for (int j = m - 1; j >= 1; --j) {
                    DQ_TRY(gemm_mat(A, p0, p1, nb, 1.0 / j, 1));
                    std::swap(p0, p1);
                }
                for (int q = 0; q < s; ++q) {
                    DQ_TRY(gemm_mat(p0, p0, p1, nb, 1.0, 0));
                    std::swap(p0, p1);
                }
                if (strategy == 1) {
                    DQ_TRY(gemm_blk(p0, cur, ya, nullptr, nb, 1.0));
                    std::swap(cur, ya);
                } else {
                    DQ_TRY(gemm_mat(p0, U, p1, nb, 1.0, 0));            // U <- P U, result lands in p1
                    std::swap(U, p1);
                }
            }
        }
        if (nb_next < nb) {                      // slots [nb_next, nb) took their last step
            const int cnt = nb - nb_next;
            if (strategy == 2) {
                Gemm g{U + (size_t)nb_next * mat, (long long)mat, (long long)plane, Dp,
                       cur + (size_t)nb_next * blk, (long long)blk, (long long)Dp * Ncp, Ncp,
                       ya + (size_t)nb_next * blk, (long long)blk, (long long)Dp * Ncp, Ncp, nullptr, Dp, Ncp, Dp, cnt, 1.0, 0};
                flops += 8.0 * d2n * cnt;
                DQ_TRY(zgemm(ctx, g));
                DQ_TRY(gather_blocks(ctx, ya + (size_t)nb_next * blk, d_K, d_order + nb_next, cnt, blk, 1));
            } else {
                DQ_TRY(gather_blocks(ctx, cur + (size_t)nb_next * blk, d_K, d_order + nb_next, cnt, blk, 1));
            }
        }
    }
    // the buffer pointers were rotated locally; the DevBufs still own the same three allocations
    S->last_gemm_flops += flops;
    return DQ_OK;
}

// host block helpers: interleaved c128 kets <-> planar block columns
void put_column(std::vector<double>& blk, int Dp, int Ncp, int dim, int col, const double* psi) {
    for (int x = 0; x < dim; ++x) {
        blk[(size_t)x * Ncp + col] = psi[2 * x];
        blk[(size_t)Dp * Ncp + (size_t)x * Ncp + col] = psi[2 * x + 1];
    }
}
void get_column(const std::vector<double>& blk, int Dp, int Ncp, int dim, int col, double* psi) {
    for (int x = 0; x < dim; ++x) {
        psi[2 * x] = blk[(size_t)x * Ncp + col];
        psi[2 * x + 1] = blk[(size_t)Dp * Ncp + (size_t)x * Ncp + col];
    }
}

// ---- diffqc pulse model on the host (diffqc.cc:75-135) ------------------------------------------------
double expit_cc(double x) {
    if (x > 32.) return 1.;
    if (x < -32.) return 0.;
    return 1 / (1 + std::exp(-x));
}
double bump(int b, int n_basis, double t) {
    const double tau = 1. / (n_basis - 2.);
    const double tau_b = tau * (b - 1.5);
    const double l = tau_b - 1.5 * tau, r = tau_b + 1.5 * tau;
    if (t < r && t > l) return (t - l) * (t - r) / (-(1.5 * tau) * (1.5 * tau));
    return 0.;
}
typedef std::vector<std::vector<std::array<double, 4>>> Channels;
int f_u_channels(const Channels& channels, double duration, int func_type, int h, double t, const double* vv, int n_param,
                 int n_basis, double* out) {
    double ans = 0;
    for (const auto& chan : channels[h]) {
        const double omega = chan[1], w = chan[2];
        const int idx = (int)std::round(chan[3]);
        DQ_REQUIRE(idx >= 0 && idx < n_param, "dq_dense_trotter: channel parameter index %d outside vv (n_param=%d)", idx, n_param);
        double A = 0, B = 0;
        for (int j = 0; j < n_basis; ++j) {
            const double fv = func_type == 0 ? std::legendre((unsigned)j, 2 * t / duration - 1)
                                             : bump(j, n_basis, t / duration);
            A += vv[((size_t)0 * n_param + idx) * n_basis + j] * fv;
            B += vv[((size_t)1 * n_param + idx) * n_basis + j] * fv;
        }
        const double N = std::sqrt(A * A + B * B);
        if (std::fabs(N - 0.0) < 0.000001) ans += 0.0;
        else ans += omega * (2 * expit_cc(N) - 1) / N * (std::cos(w * t) * A + std::sin(w * t) * B);
    }
    *out = ans;
    return DQ_OK;
}
int f_u(const Problem& P, int h, double t, const double* vv, int n_param, int n_basis, double* out) {
    return f_u_channels(P.channels, P.duration, P.func_type, h, t, vv, n_param, n_basis, out);
}


// ---- shot sampling support: outcome distributions of measurement bases (sim_plain.py:101-117) -------------------------------
// For every ket and every basis m: distr[j] = |<e_mj | ket>|^2, j < dim -- what stochastic_measure computes one inner product at
// a time (:108-109) before it hands the vector to np.random.choice (:112; the draws stay on the host, in the reference's order).
struct Measure {
    int n_meas;
    const double* bases;            // host [n_meas][dim (j)][dim (x)] c128: component x of eigenvector j of basis m
    double* probs_out;              // host [n_kets][n_meas][dim]
};

// ket k = kets[(k / group) * group_stride + (k % group) * ket_stride + ...], element x: re at + x * xs, im at + x * xs + ims
__global__ void k_outcome_probs(const double* __restrict__ kets, long long n_kets, int group, long long group_stride, long long ket_stride,
                                long long xs, long long ims, const double2* __restrict__ bases, int n_meas, int dim,
                                double* __restrict__ probs) {
    const long long total = n_kets * n_meas * dim;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % dim);
        const int mi = (int)((idx / dim) % n_meas);
        const long long k = idx / ((long long)dim * n_meas);
        const double* kp = kets + (k / group) * group_stride + (k % group) * ket_stride;
        const double2* e = bases + ((size_t)mi * dim + j) * dim;
        double ar = 0.0, ai = 0.0;                                  // <e_j|ket> = sum_x conj(e_j[x]) ket[x]
        for (int x = 0; x < dim; ++x) {
            const double2 v = __ldg(e + x);
            const double kr = kp[x * xs], ki = kp[x * xs + ims];
            ar = fma(v.x, kr, fma(v.y, ki, ar));
            ai = fma(v.x, ki, fma(-v.y, kr, ai));
        }
        probs[idx] = ar * ar + ai * ai;
    }
}

int outcome_probs(dq_context* ctx, const Measure& meas, int dim, const double* d_kets, long long n_kets, int group, long long group_stride,
                  long long ket_stride, long long xs, long long ims, double* probs_out) {
    State* S = state_of(ctx);
    const size_t nb = (size_t)meas.n_meas * dim * dim * 2 * sizeof(double), np = (size_t)n_kets * meas.n_meas * dim * sizeof(double);
    DQ_TRY(S->K1.reserve(nb));
    DQ_TRY(S->K2.reserve(np));
    DQ_CUDA(cudaMemcpyAsync(S->K1.p, meas.bases, nb, cudaMemcpyHostToDevice, ctx->stream));
    const long long total = n_kets * meas.n_meas * dim;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->prop.multiProcessorCount * 16);
    k_outcome_probs<<<grid, 256, 0, ctx->stream>>>(d_kets, n_kets, group, group_stride, ket_stride, xs, ims, S->K1.as<double2>(),
                                                   meas.n_meas, dim, S->K2.as<double>());
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    DQ_CUDA(cudaMemcpyAsync(probs_out, S->K2.p, np, cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return DQ_OK;
}


// Resident engine (dim <= 16) for a batch of estimator samples: prefix kets, then every shifted ket of every sample.
// u_prefix / u_suffix: packed HOST pulse tables, or both NULL when S->u_dev already holds [prefix rows | suffix rows]
// (filled on the device by k_pulse_rows).
int grad_resident(dq_context* ctx, Problem& P, const double* M, const double* psi0, double r, int n_samples,
                  const int32_t* prefix_steps, const double* prefix_dt, const int32_t* suffix_steps, const double* suffix_dt,
                  const std::vector<long long>& pre_off, const std::vector<long long>& suf_off, double bound,
                  const double* u_prefix, const double* u_suffix, int mode, double* energies_out,
                  const Measure* meas = nullptr) {
    State* S = state_of(ctx);
    const int n_H = P.n_H, dim = P.dim;
    const long long n_pre = pre_off[n_samples], n_suf = suf_off[n_samples];
    const int s = log2_ceil_ratio(bound, 1.0), m = 18;
    DQ_REQUIRE(s <= 20, "dense path: ||dt H|| = %g is too large", bound);
    S->last_kernel_ms = 0;
    DQ_TRY(small_upload(ctx, P, M));
    if (u_prefix || u_suffix) {          // pulse rows: prefix table, then suffix table
        DQ_TRY(S->u_dev.reserve(std::max<size_t>(1, (size_t)(n_pre + n_suf) * n_H) * sizeof(double)));
        if (n_pre) DQ_CUDA(cudaMemcpyAsync(S->u_dev.p, u_prefix, (size_t)n_pre * n_H * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (n_suf) DQ_CUDA(cudaMemcpyAsync(S->u_dev.as<double>() + (size_t)n_pre * n_H, u_suffix, (size_t)n_suf * n_H * sizeof(double),
                                           cudaMemcpyHostToDevice, ctx->stream));
    }
    // kets: [0] = psi0, [1 .. n_samples] = phi_b
    std::vector<double> k0(32, 0.0);
    memcpy(k0.data(), psi0, sizeof(double) * 2 * dim);
    DQ_TRY(S->phi.reserve((size_t)(1 + n_samples) * 32 * sizeof(double)));
    DQ_TRY(S->out.reserve((size_t)n_samples * 2 * n_H * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->phi.p, k0.data(), 32 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // longest trajectories first: a warp is busy for steps x terms, CTAs retire in launch order
    std::vector<int> order(n_samples);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return prefix_steps[a] > prefix_steps[b]; });
    std::vector<SmallTraj> traj;
    traj.reserve((size_t)n_samples * 2 * n_H);
    for (int b : order) traj.push_back(SmallTraj{pre_off[b], std::ldexp(prefix_dt[b], -s), 0.0, prefix_steps[b], 0, 0, b});
    DQ_TRY(small_run(ctx, P, mode, s, m, 1, traj, S->u_dev.as<double>(), S->phi.as<double>(), S->phi.as<double>() + 32, nullptr, 1.0));
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return suffix_steps[a] > suffix_steps[b]; });
    traj.clear();
    if (meas) {
        // shot sampling (sim_plain.py:101-117,202-203,212-213): the shifted kets themselves are needed, one warp each
        // (k_small<false, 1>), then |<e_j|ket>|^2 for every eigenvector of every measurement basis
        const size_t n_kets = (size_t)n_samples * 2 * n_H;
        DQ_TRY(S->K0.reserve(n_kets * 32 * sizeof(double)));
        for (int b : order)
            for (int i = 0; i < n_H; ++i)
                for (int sg = 0; sg < 2; ++sg)
                    traj.push_back(SmallTraj{n_pre + suf_off[b], std::ldexp(suffix_dt[b], -s), sg ? -r : r, suffix_steps[b], b, i,
                                             (b * n_H + i) * 2 + sg});
        DQ_TRY(small_run(ctx, P, mode, s, m, 1, traj, S->u_dev.as<double>(), S->phi.as<double>() + 32, S->K0.as<double>(), nullptr,
                         1.0 / sqrt(1.0 + r * r)));
        DQ_TRY(outcome_probs(ctx, *meas, dim, S->K0.as<double>(), (long long)n_kets, 1, 32, 0, 2, 1, meas->probs_out));
        S->last_gemm_flops = 0;
        S->last_strategy = 3;
        S->last_squarings = s;
        S->last_degree = m;
        return DQ_OK;
    }
    // Every shifted ket of a sample shares its generator.  Exact step: a warp takes up to 16 of them through the tensor-core
    // recurrence (k_small_mma; option "small_mma" 0 = the DFMA kernel below instead).  Per-term product (mode 1): the +/-
    // kets of a control (and of two neighbouring controls when n_H is even) per warp on the DFMA kernel.
    const bool mma = mode == 0 && ctx->dense_small_mma != 0;
    const int nk = mma ? (2 * n_H > 8 ? 16 : 8) : ((n_H % 2 == 0) ? 4 : 2);
    for (int b : order)
        for (int i = 0; i < n_H; i += nk / 2)
            traj.push_back(SmallTraj{n_pre + suf_off[b], std::ldexp(suffix_dt[b], -s), r, suffix_steps[b], b, i, (b * n_H + i) * 2});
    DQ_TRY(small_run(ctx, P, mode, s, m, nk, traj, S->u_dev.as<double>(), S->phi.as<double>() + 32, nullptr, S->out.as<double>(),
                     1.0 / sqrt(1.0 + r * r)));
    DQ_CUDA(cudaMemcpyAsync(energies_out, S->out.p, (size_t)n_samples * 2 * n_H * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    S->last_gemm_flops = 0;
    S->last_strategy = 3;
    S->last_squarings = s;
    S->last_degree = m;
    return DQ_OK;
}

// ---- pulse rows of the Python twin on the device ---------------------------------------------------------------------------
// u[row][i] = omega_i (2 sigma(sum_j c_ij phi_j(t/T)) - 1), phi = the open-support quadratic bumps (sim_plain.py:52-99), for every
// step of every prefix (0 -> s_b) and suffix (s_b -> T) trajectory of a batch: the table pulses.u_table builds on the host,
// in the same operation order -- t accumulated by repeated `t += dt` (sim_plain.py:134,150), the sum over j sequential from
// j = 0 with a separate multiply and add (no FMA contraction), 1 / (1 + exp(-a)) -- so that the two agree to the last bit of
// everything but exp() (libdevice vs NumPy: each within 1 ulp).  One thread per row.
__global__ void k_pulse_rows(const double* __restrict__ s_list, const long long* __restrict__ pre_off,
                             const long long* __restrict__ suf_off, const int* __restrict__ pre_n, const int* __restrict__ suf_n,
                             const double* __restrict__ pre_dt, const double* __restrict__ suf_dt, int n_samples, long long n_pre_total,
                             double T, const double* __restrict__ coeff, const double* __restrict__ omegas, int n_H, int n_basis,
                             const double* __restrict__ bl, const double* __restrict__ br, double norm_factor,
                             double* __restrict__ u, const double* __restrict__ norm1, int sum_mode,
                             unsigned long long* __restrict__ bound_bits) {
    const int b = blockIdx.x;
    const int which = blockIdx.y;                          // 0 prefix, 1 suffix
    const int n = which ? suf_n[b] : pre_n[b];
    const double dt = which ? suf_dt[b] : pre_dt[b];
    const double t0 = which ? s_list[b] : 0.0;
    const long long row0 = which ? n_pre_total + suf_off[b] : pre_off[b];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double t = t0;
        for (int i = 0; i < k; ++i) t = __dadd_rn(t, dt);  // the reference's repeated t += dt
        const double x = t / T;
        double* __restrict__ out = u + (row0 + k) * n_H;
        double nb = norm1[0];                              // ||dt H(t_k)||_1 bound of this row (scaling choice of the Taylor series)
        for (int i = 0; i < n_H; ++i) {
            double a = 0.0;
            for (int j = 0; j < n_basis; ++j) {
                double phi = 0.0;
                if (!(x >= br[j] || x <= bl[j])) phi = __dmul_rn(__dadd_rn(x, -bl[j]), __dadd_rn(x, -br[j])) / norm_factor;
                a = __dadd_rn(a, __dmul_rn(phi, coeff[i * n_basis + j]));
            }
            const double sg = 1.0 / (1.0 + exp(-a));
            const double ui = __dmul_rn(__dadd_rn(__dmul_rn(sg, 2.0), -1.0), omegas[i]);
            out[i] = ui;
            nb = sum_mode ? nb + fabs(ui) * norm1[i + 1] : fmax(nb, fabs(ui) * norm1[i + 1]);
        }
        atomicMax(bound_bits, (unsigned long long)__double_as_longlong(fabs(dt) * nb));   // non-negative doubles order like integers
    }
}

}  // namespace
}  // namespace dense
}  // namespace dq

using namespace dq::dense;

extern "C" {

int dq_pulse_f_u_table(int n_H, const int32_t* chan_counts, const double* channels, double duration, int func_type,
                       const double* vv, int n_param, int n_basis, int n_t, const double* ts, double* u_out) {
    DQ_REQUIRE(n_H >= 0 && n_t >= 0 && (n_H == 0 || chan_counts) && vv && (n_t == 0 || (ts && u_out)), "dq_pulse_f_u_table: bad argument");
    DQ_REQUIRE(n_param >= 1 && n_basis >= 1, "dq_pulse_f_u_table: vv must be [2][n_param>=1][n_basis>=1]");
    DQ_REQUIRE(std::isfinite(duration) && duration != 0.0, "dq_pulse_f_u_table: duration must be finite and non-zero");
    Channels ch((size_t)n_H);
    size_t k = 0;
    for (int h = 0; h < n_H; ++h)
        for (int c = 0; c < chan_counts[h]; ++c, ++k) ch[h].push_back({channels[4 * k], channels[4 * k + 1], channels[4 * k + 2], channels[4 * k + 3]});
    for (int i = 0; i < n_t; ++i)
        for (int h = 0; h < n_H; ++h) DQ_TRY(f_u_channels(ch, duration, func_type, h, ts[i], vv, n_param, n_basis, &u_out[(size_t)i * n_H + h]));
    return DQ_OK;
}

int dq_dense_set_H(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const int32_t* chan_counts,
                   const double* channels, double duration, int func_type) {
    DQ_REQUIRE(ctx, "NULL context");
    DQ_REQUIRE(n_H == 0 || chan_counts, "dq_dense_set_H: NULL channel counts");
    DQ_REQUIRE(std::isfinite(duration) && duration != 0.0, "dq_dense_set_H: duration must be finite and non-zero");
    DQ_TRY(ctx->set_device());
    Problem& P = state_of(ctx)->global_H;
    P.is_set = false;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    P.channels.assign(n_H, {});
    size_t k = 0;
    for (int h = 0; h < n_H; ++h) {
        DQ_REQUIRE(chan_counts[h] >= 0, "dq_dense_set_H: negative channel count for term %d", h);
        DQ_REQUIRE(chan_counts[h] == 0 || channels, "dq_dense_set_H: NULL channel table");
        for (int c = 0; c < chan_counts[h]; ++c, ++k) {
            std::array<double, 4> ch = {channels[4 * k], channels[4 * k + 1], channels[4 * k + 2], channels[4 * k + 3]};
            for (double v : ch) DQ_REQUIRE(std::isfinite(v), "dq_dense_set_H: non-finite channel entry (term %d)", h);
            P.channels[h].push_back(ch);
        }
    }
    P.duration = duration;
    P.func_type = func_type;
    P.is_set = true;
    return DQ_OK;
}

int dq_dense_trotter(dq_context* ctx, const double* psi0, double T0, double T, int per_step, const double* vv,
                     int n_param, int n_basis, double* psi_out, double* u_out) {
    DQ_REQUIRE(ctx && psi0 && psi_out && vv, "dq_dense_trotter: NULL argument");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->global_H;
    if (!P.is_set) {
        dq::set_error("dq_dense_trotter: set_H has not been called (diffqc.cc:21-25 globals are empty)");
        return DQ_ERR_STATE;
    }
    DQ_REQUIRE(n_param >= 1 && n_basis >= 1, "dq_dense_trotter: vv must be [2][n_param>=1][n_basis>=1]");
    DQ_REQUIRE(P.func_type == 0 || n_basis != 2, "dq_dense_trotter: the bump basis needs n_basis != 2 (tau = 1/(n_basis-2))");
    DQ_REQUIRE(std::isfinite(T0) && std::isfinite(T), "dq_dense_trotter: non-finite time span");
    for (size_t i = 0; i < (size_t)2 * n_param * n_basis; ++i) DQ_REQUIRE(std::isfinite(vv[i]), "dq_dense_trotter: non-finite vv entry");
    for (int i = 0; i < 2 * P.dim; ++i) DQ_REQUIRE(std::isfinite(psi0[i]), "dq_dense_trotter: non-finite psi0 entry");
    // diffqc.cc:182-184: n_steps = (int)(per_step * (|T - T0| + 1)), dt = (T - T0) / n_steps, t accumulates
    const int n_steps = (int)(per_step * (std::fabs(T - T0) + 1));
    DQ_REQUIRE(n_steps >= 1, "dq_dense_trotter: per_step=%d gives %d steps (the reference divides by zero here)", per_step, n_steps);
    const double dt = (T - T0) / n_steps;
    std::vector<double> u((size_t)n_steps * std::max(1, P.n_H));
    double t = T0;
    for (int k = 0; k < n_steps; ++k) {
        for (int h = 0; h < P.n_H; ++h) DQ_TRY(f_u(P, h, t, vv, n_param, n_basis, &u[(size_t)k * P.n_H + h]));
        t += dt;
    }
    if (u_out) memcpy(u_out, u.data(), (size_t)n_steps * P.n_H * sizeof(double));
    const int Ncp = 8;
    std::vector<double> blk((size_t)2 * P.Dp * Ncp, 0.0);
    put_column(blk, P.Dp, Ncp, P.dim, 0, psi0);
    DQ_TRY(S->phi.reserve(blk.size() * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->phi.p, blk.data(), blk.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long long off = 0;
    S->last_gemm_flops = 0;
    DQ_TRY(evolve_blocks(ctx, P, 0, 1, Ncp, &n_steps, &dt, &off, u.data(), n_steps, S->phi.as<double>()));
    DQ_CUDA(cudaMemcpyAsync(blk.data(), S->phi.p, blk.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    get_column(blk, P.Dp, Ncp, P.dim, 0, psi_out);
    return DQ_OK;
}

int dq_dense_evolve(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* u, int n_steps,
                    double dt, int mode, int batch, const double* psi_in, double* psi_out) {
    DQ_REQUIRE(ctx && psi_in && psi_out, "dq_dense_evolve: NULL argument");
    DQ_REQUIRE(mode == 0 || mode == 1, "dq_dense_evolve: mode must be 0 (exact) or 1 (split)");
    DQ_REQUIRE(batch >= 1 && n_steps >= 0, "dq_dense_evolve: batch=%d n_steps=%d", batch, n_steps);
    DQ_REQUIRE(n_steps == 0 || n_H == 0 || u, "dq_dense_evolve: NULL pulse table");
    DQ_REQUIRE(std::isfinite(dt), "dq_dense_evolve: non-finite dt");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->scratch_H;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    for (size_t i = 0; i < (size_t)2 * dim * batch; ++i) DQ_REQUIRE(std::isfinite(psi_in[i]), "dq_dense_evolve: non-finite psi entry");
    bool resident = false;
    DQ_TRY(want_resident(ctx, P, &resident));
    if (resident) {
        static const double no_u0 = 0.0;
        const long long off0 = 0;
        double bound = 0.0;
        DQ_TRY(norm_bound(P, mode, 1, &n_steps, &dt, &off0, u ? u : &no_u0, &bound));
        const int s = log2_ceil_ratio(bound, 1.0), m = 18;
        DQ_REQUIRE(s <= 20, "dense path: ||dt H|| = %g is too large", bound);
        S->last_kernel_ms = 0;
        DQ_TRY(small_upload(ctx, P, nullptr));
        std::vector<double> kets((size_t)batch * 32, 0.0);
        for (int c = 0; c < batch; ++c) memcpy(kets.data() + (size_t)c * 32, psi_in + (size_t)2 * dim * c, sizeof(double) * 2 * dim);
        DQ_TRY(S->phi.reserve(kets.size() * sizeof(double)));
        DQ_TRY(S->out.reserve(kets.size() * sizeof(double)));
        const size_t u_count = (size_t)n_steps * n_H;
        DQ_TRY(S->u_dev.reserve(std::max<size_t>(1, u_count) * sizeof(double)));
        if (u_count) DQ_CUDA(cudaMemcpyAsync(S->u_dev.p, u, u_count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        DQ_CUDA(cudaMemcpyAsync(S->phi.p, kets.data(), kets.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<SmallTraj> traj(batch);
        for (int c = 0; c < batch; ++c) traj[c] = SmallTraj{0, std::ldexp(dt, -s), 0.0, n_steps, c, 0, c};
        DQ_TRY(small_run(ctx, P, mode, s, m, 1, traj, S->u_dev.as<double>(), S->phi.as<double>(), S->out.as<double>(), nullptr, 1.0));
        DQ_CUDA(cudaMemcpyAsync(kets.data(), S->out.p, kets.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        DQ_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < batch; ++c) memcpy(psi_out + (size_t)2 * dim * c, kets.data() + (size_t)c * 32, sizeof(double) * 2 * dim);
        S->last_gemm_flops = 0;
        S->last_strategy = 3;
        S->last_squarings = s;
        S->last_degree = m;
        return DQ_OK;
    }
    const int Ncp = round8(batch);
    std::vector<double> blk((size_t)2 * P.Dp * Ncp, 0.0);
    for (int c = 0; c < batch; ++c) put_column(blk, P.Dp, Ncp, dim, c, psi_in + (size_t)2 * dim * c);
    DQ_TRY(S->phi.reserve(blk.size() * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->phi.p, blk.data(), blk.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long long off = 0;
    S->last_gemm_flops = 0;
    static const double no_u = 0.0;
    DQ_TRY(evolve_blocks(ctx, P, mode, 1, Ncp, &n_steps, &dt, &off, u ? u : &no_u, n_steps, S->phi.as<double>()));
    DQ_CUDA(cudaMemcpyAsync(blk.data(), S->phi.p, blk.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < batch; ++c) get_column(blk, P.Dp, Ncp, dim, c, psi_out + (size_t)2 * dim * c);
    return DQ_OK;
}

int dq_dense_evolve_many(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M, int n_traj,
                         const double* psi_in, const int32_t* steps, const double* dts, const double* u, int mode,
                         double* energies_out, double* psi_out) {
    DQ_REQUIRE(ctx && psi_in && steps && dts && (energies_out || psi_out), "dq_dense_evolve_many: NULL argument");
    DQ_REQUIRE(!energies_out || M, "dq_dense_evolve_many: energies need the observable");
    DQ_REQUIRE(mode == 0 || mode == 1, "dq_dense_evolve_many: mode must be 0 (exact) or 1 (split)");
    DQ_REQUIRE(n_traj >= 1, "dq_dense_evolve_many: n_traj=%d", n_traj);
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->scratch_H;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    if (M) DQ_TRY(upload_observable(ctx, P, M));
    for (size_t i = 0; i < (size_t)2 * dim * n_traj; ++i) DQ_REQUIRE(std::isfinite(psi_in[i]), "dq_dense_evolve_many: non-finite psi entry");
    std::vector<long long> off(n_traj + 1, 0);
    for (int b = 0; b < n_traj; ++b) {
        DQ_REQUIRE(steps[b] >= 0 && std::isfinite(dts[b]), "dq_dense_evolve_many: bad step count or dt (trajectory %d)", b);
        off[b + 1] = off[b] + steps[b];
    }
    DQ_REQUIRE(off[n_traj] == 0 || n_H == 0 || u, "dq_dense_evolve_many: NULL pulse table");
    static const double no_u = 0.0;
    const double* uu = (u && n_H) ? u : &no_u;
    bool resident = false;
    DQ_TRY(want_resident(ctx, P, &resident));
    S->last_gemm_flops = 0;
    if (resident) {
        double bound = 0.0;
        DQ_TRY(norm_bound(P, mode, n_traj, steps, dts, off.data(), uu, &bound));
        const int s = log2_ceil_ratio(bound, 1.0), m = 18;
        DQ_REQUIRE(s <= 20, "dense path: ||dt H|| = %g is too large", bound);
        S->last_kernel_ms = 0;
        DQ_TRY(small_upload(ctx, P, M));
        std::vector<double> kets((size_t)n_traj * 32, 0.0);
        for (int c = 0; c < n_traj; ++c) memcpy(kets.data() + (size_t)c * 32, psi_in + (size_t)2 * dim * c, sizeof(double) * 2 * dim);
        DQ_TRY(S->phi.reserve(kets.size() * sizeof(double)));
        DQ_TRY(S->out.reserve((kets.size() + n_traj) * sizeof(double)));
        const size_t u_count = (size_t)off[n_traj] * n_H;
        DQ_TRY(S->u_dev.reserve(std::max<size_t>(1, u_count) * sizeof(double)));
        if (u_count) DQ_CUDA(cudaMemcpyAsync(S->u_dev.p, uu, u_count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        DQ_CUDA(cudaMemcpyAsync(S->phi.p, kets.data(), kets.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<SmallTraj> traj(n_traj);
        for (int c = 0; c < n_traj; ++c) traj[c] = SmallTraj{off[c], std::ldexp(dts[c], -s), 0.0, steps[c], c, 0, c};
        if (psi_out) {
            DQ_TRY(small_run(ctx, P, mode, s, m, 1, traj, S->u_dev.as<double>(), S->phi.as<double>(), S->out.as<double>(), nullptr, 1.0));
            DQ_CUDA(cudaMemcpyAsync(kets.data(), S->out.p, kets.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (energies_out) {
            double* dE = S->out.as<double>() + kets.size();
            DQ_TRY(small_run(ctx, P, mode, s, m, 1, traj, S->u_dev.as<double>(), S->phi.as<double>(), nullptr, dE, 1.0));
            DQ_CUDA(cudaMemcpyAsync(energies_out, dE, (size_t)n_traj * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        DQ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (psi_out)
            for (int c = 0; c < n_traj; ++c) memcpy(psi_out + (size_t)2 * dim * c, kets.data() + (size_t)c * 32, sizeof(double) * 2 * dim);
        S->last_strategy = 3;
        S->last_squarings = s;
        S->last_degree = m;
        return DQ_OK;
    }
    // GEMM strategies: one 8-column ket block per trajectory (column 0 carries the state), every block its own step list
    const int Ncp = 8, Dp = P.Dp;
    const size_t blk = (size_t)2 * Dp * Ncp;
    std::vector<double> host((size_t)n_traj * blk, 0.0), one(blk);
    for (int c = 0; c < n_traj; ++c) {
        std::fill(one.begin(), one.end(), 0.0);
        put_column(one, Dp, Ncp, dim, 0, psi_in + (size_t)2 * dim * c);
        memcpy(host.data() + (size_t)c * blk, one.data(), blk * sizeof(double));
    }
    DQ_TRY(S->phi.reserve(host.size() * sizeof(double)));
    DQ_TRY(S->out.reserve((size_t)n_traj * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->phi.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DQ_TRY(evolve_blocks(ctx, P, mode, n_traj, Ncp, steps, dts, off.data(), uu, off[n_traj], S->phi.as<double>()));
    if (energies_out) {
        DQ_TRY(energies(ctx, P, n_traj, S->phi.as<double>(), Ncp, 1, S->out.as<double>()));
        DQ_CUDA(cudaMemcpyAsync(energies_out, S->out.p, (size_t)n_traj * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (psi_out) DQ_CUDA(cudaMemcpyAsync(host.data(), S->phi.p, host.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (psi_out)
        for (int c = 0; c < n_traj; ++c) {
            memcpy(one.data(), host.data() + (size_t)c * blk, blk * sizeof(double));
            get_column(one, Dp, Ncp, dim, 0, psi_out + (size_t)2 * dim * c);
        }
    return DQ_OK;
}

static int grad_impl(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                  const double* psi0, double r, int n_samples, const int32_t* prefix_steps, const double* prefix_dt,
                  const double* u_prefix, const int32_t* suffix_steps, const double* suffix_dt, const double* u_suffix,
                  int mode, double* energies_out, const Measure* meas) {
    DQ_REQUIRE(ctx && psi0 && ((M && energies_out) || meas), "dq_dense_grad: NULL argument");
    DQ_REQUIRE(prefix_steps && prefix_dt && suffix_steps && suffix_dt, "dq_dense_grad: NULL step table");
    DQ_REQUIRE(mode == 0 || mode == 1, "dq_dense_grad: mode must be 0 (exact) or 1 (split)");
    DQ_REQUIRE(n_samples >= 1 && n_H >= 1, "dq_dense_grad: n_samples=%d n_H=%d", n_samples, n_H);
    DQ_REQUIRE(r > 0 && std::isfinite(r), "dq_dense_grad: r must be positive");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->scratch_H;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    if (M) DQ_TRY(upload_observable(ctx, P, M));
    for (int i = 0; i < 2 * dim; ++i) DQ_REQUIRE(std::isfinite(psi0[i]), "dq_dense_grad: non-finite psi0 entry");
    std::vector<long long> pre_off(n_samples + 1, 0), suf_off(n_samples + 1, 0);
    for (int b = 0; b < n_samples; ++b) {
        DQ_REQUIRE(prefix_steps[b] >= 0 && suffix_steps[b] >= 0, "dq_dense_grad: negative step count (sample %d)", b);
        DQ_REQUIRE(std::isfinite(prefix_dt[b]) && std::isfinite(suffix_dt[b]), "dq_dense_grad: non-finite dt (sample %d)", b);
        pre_off[b + 1] = pre_off[b] + prefix_steps[b];
        suf_off[b + 1] = suf_off[b] + suffix_steps[b];
    }
    DQ_REQUIRE((pre_off[n_samples] == 0 || u_prefix) && (suf_off[n_samples] == 0 || u_suffix), "dq_dense_grad: NULL pulse table");
    bool resident = false;
    DQ_TRY(want_resident(ctx, P, &resident));
    if (resident) {
        static const double no_u0 = 0.0;
        double b_pre = 0.0, b_suf = 0.0;
        DQ_TRY(norm_bound(P, mode, n_samples, prefix_steps, prefix_dt, pre_off.data(), u_prefix ? u_prefix : &no_u0, &b_pre));
        DQ_TRY(norm_bound(P, mode, n_samples, suffix_steps, suffix_dt, suf_off.data(), u_suffix ? u_suffix : &no_u0, &b_suf));
        return grad_resident(ctx, P, M, psi0, r, n_samples, prefix_steps, prefix_dt, suffix_steps, suffix_dt, pre_off, suf_off,
                             std::max(b_pre, b_suf), u_prefix, u_suffix, mode, energies_out, meas);
    }
    const int Ncp = round8(2 * n_H), Dp = P.Dp;
    const size_t mat_bytes = 2 * P.plane() * sizeof(double);
    // samples per chunk: three Dp x Dp work matrices and three ket blocks per sample within ~4 GiB
    const size_t per_sample = 4 * mat_bytes + 4 * (size_t)2 * Dp * Ncp * sizeof(double);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_samples, ((size_t)4 << 30) / per_sample));
    std::vector<double> phi_blk((size_t)2 * Dp * 8, 0.0);
    put_column(phi_blk, Dp, 8, dim, 0, psi0);
    S->last_gemm_flops = 0;
    static const double no_u = 0.0;
    for (int b0 = 0; b0 < n_samples; b0 += chunk) {
        const int nb = std::min(chunk, n_samples - b0);
        DQ_TRY(S->phi.reserve((size_t)nb * phi_blk.size() * sizeof(double)));
        DQ_TRY(S->out.reserve(((size_t)nb * 2 * Dp * Ncp + (size_t)nb * 2 * n_H) * sizeof(double)));
        for (int b = 0; b < nb; ++b)
            DQ_CUDA(cudaMemcpyAsync(S->phi.as<double>() + (size_t)b * phi_blk.size(), b == 0 ? (const void*)phi_blk.data() : S->phi.p,
                                    phi_blk.size() * sizeof(double), b == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                    ctx->stream));
        std::vector<long long> po(nb), so(nb);
        for (int b = 0; b < nb; ++b) { po[b] = pre_off[b0 + b] - pre_off[b0]; so[b] = suf_off[b0 + b] - suf_off[b0]; }
        const long long n_pre = pre_off[b0 + nb] - pre_off[b0], n_suf = suf_off[b0 + nb] - suf_off[b0];
        DQ_TRY(evolve_blocks(ctx, P, mode, nb, 8, prefix_steps + b0, prefix_dt + b0, po.data(),
                             u_prefix ? u_prefix + pre_off[b0] * n_H : &no_u, n_pre, S->phi.as<double>()));
        double* K = S->out.as<double>();
        double* E = K + (size_t)nb * 2 * Dp * Ncp;
        DQ_CUDA(cudaMemsetAsync(K, 0, (size_t)nb * 2 * Dp * Ncp * sizeof(double), ctx->stream));
        DQ_TRY(fanout(ctx, P, nb, S->phi.as<double>(), 8, K, Ncp, r));
        DQ_TRY(evolve_blocks(ctx, P, mode, nb, Ncp, suffix_steps + b0, suffix_dt + b0, so.data(),
                             u_suffix ? u_suffix + suf_off[b0] * n_H : &no_u, n_suf, K));
        if (meas) {                  // kets of sample b: planar block [2][Dp][Ncp], column c = 2 i + sign
            DQ_TRY(outcome_probs(ctx, *meas, dim, K, (long long)nb * 2 * n_H, 2 * n_H, (long long)2 * Dp * Ncp, 1, Ncp, (long long)Dp * Ncp,
                                 meas->probs_out + (size_t)b0 * 2 * n_H * meas->n_meas * dim));
            continue;
        }
        DQ_TRY(energies(ctx, P, nb, K, Ncp, 2 * n_H, E));
        DQ_CUDA(cudaMemcpyAsync(energies_out + (size_t)b0 * 2 * n_H, E, (size_t)nb * 2 * n_H * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->stream));
        DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return DQ_OK;
}

int dq_dense_grad(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                  const double* psi0, double r, int n_samples, const int32_t* prefix_steps, const double* prefix_dt,
                  const double* u_prefix, const int32_t* suffix_steps, const double* suffix_dt, const double* u_suffix,
                  int mode, double* energies_out) {
    DQ_REQUIRE(M && energies_out, "dq_dense_grad: NULL argument");
    return grad_impl(ctx, dim, H0, n_H, Hs, M, psi0, r, n_samples, prefix_steps, prefix_dt, u_prefix, suffix_steps, suffix_dt, u_suffix,
                     mode, energies_out, nullptr);
}

int dq_dense_grad_probs(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* psi0, double r,
                        int n_samples, const int32_t* prefix_steps, const double* prefix_dt, const double* u_prefix,
                        const int32_t* suffix_steps, const double* suffix_dt, const double* u_suffix, int mode, int n_meas,
                        const double* bases, double* probs_out) {
    DQ_REQUIRE(n_meas >= 1 && bases && probs_out, "dq_dense_grad_probs: n_meas=%d or NULL basis / output", n_meas);
    for (size_t i = 0; i < (size_t)n_meas * dim * dim * 2; ++i) DQ_REQUIRE(std::isfinite(bases[i]), "dq_dense_grad_probs: non-finite basis entry");
    Measure meas{n_meas, bases, probs_out};
    return grad_impl(ctx, dim, H0, n_H, Hs, nullptr, psi0, r, n_samples, prefix_steps, prefix_dt, u_prefix, suffix_steps, suffix_dt,
                     u_suffix, mode, nullptr, &meas);
}

int dq_dense_outcome_probs(dq_context* ctx, int dim, int n_kets, const double* kets, int n_meas, const double* bases, double* probs_out) {
    DQ_REQUIRE(ctx && kets && bases && probs_out && dim >= 1 && n_kets >= 1 && n_meas >= 1, "dq_dense_outcome_probs: bad argument");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    DQ_TRY(S->K0.reserve((size_t)n_kets * dim * 2 * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->K0.p, kets, (size_t)n_kets * dim * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    Measure meas{n_meas, bases, probs_out};
    return outcome_probs(ctx, meas, dim, S->K0.as<double>(), n_kets, 1, (long long)2 * dim, 0, 2, 1, probs_out);
}

int dq_dense_grad_times(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                        const double* psi0, double r, int n_samples, const double* s_list, double T, int per_step,
                        const double* coeff, const double* omegas, int n_basis, int mode, double* energies_out, double* u_out) {
    DQ_REQUIRE(ctx && M && psi0 && energies_out && s_list && coeff && omegas, "dq_dense_grad_times: NULL argument");
    DQ_REQUIRE(mode == 0 || mode == 1, "dq_dense_grad_times: mode must be 0 (exact) or 1 (split)");
    DQ_REQUIRE(n_samples >= 1 && n_H >= 1 && n_basis >= 3 && per_step >= 1, "dq_dense_grad_times: n_samples=%d n_H=%d n_basis=%d per_step=%d",
               n_samples, n_H, n_basis, per_step);
    DQ_REQUIRE(r > 0 && std::isfinite(r) && T > 0 && std::isfinite(T), "dq_dense_grad_times: r and T must be positive");
    DQ_TRY(ctx->set_device());
    State* S = state_of(ctx);
    Problem& P = S->scratch_H;
    DQ_TRY(upload_problem(ctx, P, dim, H0, n_H, Hs));
    bool resident = false;
    DQ_TRY(want_resident(ctx, P, &resident));
    if (!resident) {
        dq::set_error("dq_dense_grad_times: device-side pulse tables exist for the resident engine (dim <= 16) only");
        return DQ_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < 2 * dim; ++i) DQ_REQUIRE(std::isfinite(psi0[i]), "dq_dense_grad_times: non-finite psi0 entry");
    for (int i = 0; i < n_H * n_basis; ++i) DQ_REQUIRE(std::isfinite(coeff[i]), "dq_dense_grad_times: non-finite coefficient");
    // step grids exactly as sim_plain.py:123,133: n = int(per_step ((T1 - T0) + 1)), dt = (T1 - T0) / n
    std::vector<int32_t> pre_n(n_samples), suf_n(n_samples);
    std::vector<double> pre_dt(n_samples), suf_dt(n_samples);
    std::vector<long long> pre_off(n_samples + 1, 0), suf_off(n_samples + 1, 0);
    double w_sum = P.norm1[0], w_max = P.norm1[0], dt_max = 0.0;
    for (int h = 0; h < n_H; ++h) {
        DQ_REQUIRE(std::isfinite(omegas[h]), "dq_dense_grad_times: non-finite omega");
        w_sum += fabs(omegas[h]) * P.norm1[h + 1];          // |u_h| <= |omega_h|
        w_max = std::max(w_max, fabs(omegas[h]) * P.norm1[h + 1]);
    }
    for (int b = 0; b < n_samples; ++b) {
        const double sb = s_list[b];
        DQ_REQUIRE(std::isfinite(sb) && sb >= 0.0 && sb <= T, "dq_dense_grad_times: sample time %g outside [0, T]", sb);
        pre_n[b] = (int)(per_step * ((sb - 0.0) + 1));
        suf_n[b] = (int)(per_step * ((T - sb) + 1));
        pre_dt[b] = pre_n[b] > 0 ? (sb - 0.0) / pre_n[b] : 0.0;
        suf_dt[b] = suf_n[b] > 0 ? (T - sb) / suf_n[b] : 0.0;
        pre_off[b + 1] = pre_off[b] + pre_n[b];
        suf_off[b + 1] = suf_off[b] + suf_n[b];
        dt_max = std::max(dt_max, std::max(fabs(pre_dt[b]), fabs(suf_dt[b])));
    }
    const long long n_pre = pre_off[n_samples], n_suf = suf_off[n_samples];
    // bump supports as get_func_bspline computes them (sim_plain.py:53-58)
    std::vector<double> bl(n_basis), br(n_basis);
    const double tau = 1. / (n_basis - 2);
    for (int b = 0; b < n_basis; ++b) {
        const double tau_b = tau * (b - 1.5);
        bl[b] = tau_b - 1.5 * tau;
        br[b] = tau_b + 1.5 * tau;
    }
    const double norm_factor = -((1.5 * tau) * (1.5 * tau));
    // one staging buffer: s | pre_dt | suf_dt | coeff | omegas | bl | br | pre_off | suf_off | pre_n | suf_n
    const size_t nd = (size_t)3 * n_samples + (size_t)n_H * n_basis + n_H + 2 * n_basis + (1 + n_H) + 1;
    std::vector<double> hd(nd, 0.0);
    double* q = hd.data();
    memcpy(q, s_list, n_samples * sizeof(double)); q += n_samples;
    memcpy(q, pre_dt.data(), n_samples * sizeof(double)); q += n_samples;
    memcpy(q, suf_dt.data(), n_samples * sizeof(double)); q += n_samples;
    memcpy(q, coeff, (size_t)n_H * n_basis * sizeof(double)); q += (size_t)n_H * n_basis;
    memcpy(q, omegas, n_H * sizeof(double)); q += n_H;
    memcpy(q, bl.data(), n_basis * sizeof(double)); q += n_basis;
    memcpy(q, br.data(), n_basis * sizeof(double)); q += n_basis;
    memcpy(q, P.norm1.data(), (1 + n_H) * sizeof(double));            // then one zeroed slot: the bound (as integer bits)
    const size_t bytes_d = nd * sizeof(double), bytes_o = (size_t)2 * n_samples * sizeof(long long), bytes_n = (size_t)2 * n_samples * sizeof(int);
    DQ_TRY(S->meta.reserve(bytes_d + bytes_o + bytes_n));
    char* dm = S->meta.as<char>();
    DQ_CUDA(cudaMemcpyAsync(dm, hd.data(), bytes_d, cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(dm + bytes_d, pre_off.data(), n_samples * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(dm + bytes_d + n_samples * sizeof(long long), suf_off.data(), n_samples * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(dm + bytes_d + bytes_o, pre_n.data(), n_samples * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaMemcpyAsync(dm + bytes_d + bytes_o + n_samples * sizeof(int), suf_n.data(), n_samples * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    DQ_TRY(S->u_dev.reserve(std::max<size_t>(1, (size_t)(n_pre + n_suf) * n_H) * sizeof(double)));
    const double* dd = reinterpret_cast<const double*>(dm);
    const long long* dof = reinterpret_cast<const long long*>(dm + bytes_d);
    const int* dn = reinterpret_cast<const int*>(dm + bytes_d + bytes_o);
    k_pulse_rows<<<dim3((unsigned)n_samples, 2), 64, 0, ctx->stream>>>(
        dd, dof, dof + n_samples, dn, dn + n_samples, dd + n_samples, dd + 2 * n_samples, n_samples, n_pre, T,
        dd + 3 * n_samples, dd + 3 * n_samples + (size_t)n_H * n_basis, n_H, n_basis,
        dd + 3 * n_samples + (size_t)n_H * n_basis + n_H, dd + 3 * n_samples + (size_t)n_H * n_basis + n_H + n_basis, norm_factor,
        S->u_dev.as<double>(), dd + nd - (2 + n_H), mode == 0 ? 1 : 0,
        reinterpret_cast<unsigned long long*>(dm) + (nd - 1));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    double bound = 0.0;                                     // 8 bytes back: the scaling of the series is chosen from the actual pulses
    DQ_CUDA(cudaMemcpyAsync(&bound, dm + (nd - 1) * sizeof(double), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    DQ_REQUIRE(std::isfinite(bound), "dq_dense_grad_times: non-finite pulse value");
    if (u_out) {
        DQ_CUDA(cudaMemcpyAsync(u_out, S->u_dev.p, (size_t)(n_pre + n_suf) * n_H * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    DQ_TRY(upload_observable(ctx, P, M));
    (void)w_sum; (void)w_max; (void)dt_max;
    return grad_resident(ctx, P, M, psi0, r, n_samples, pre_n.data(), pre_dt.data(), suf_n.data(), suf_dt.data(), pre_off, suf_off, bound,
                         nullptr, nullptr, mode, energies_out);
}

int dq_dense_last_stat(dq_context* ctx, const char* name, double* value) {
    DQ_REQUIRE(ctx && name && value, "NULL argument");
    State* S = state_of(ctx);
    if (!strcmp(name, "gemm_flops")) *value = S->last_gemm_flops;
    else if (!strcmp(name, "strategy")) *value = S->last_strategy;
    else if (!strcmp(name, "squarings")) *value = S->last_squarings;
    else if (!strcmp(name, "degree")) *value = S->last_degree;
    else if (!strcmp(name, "kernel_ms")) *value = S->last_kernel_ms;
    else { dq::set_error("dq_dense_last_stat: unknown name '%s'", name); return DQ_ERR_INVALID; }
    return DQ_OK;
}

int dq_dense_set_option(dq_context* ctx, const char* name, int64_t value) {
    DQ_REQUIRE(ctx && name, "NULL argument");
    if (!strcmp(name, "strategy")) {
        DQ_REQUIRE(value >= -1 && value <= 3, "strategy must be -1 (auto), 0, 1, 2 or 3");
        ctx->dense_force_strategy = (int)value;
    } else if (!strcmp(name, "small_mma")) {
        ctx->dense_small_mma = value != 0;
    } else { dq::set_error("dq_dense_set_option: unknown option '%s'", name); return DQ_ERR_INVALID; }
    return DQ_OK;
}

}  // extern "C"
