// Resident engine of the dense path for dim <= 16 (strategy 3): one WARP owns one ket for its whole trajectory.
//
// The DMMA GEMM strategies (dense_api.cu) issue ~m 2^s launches per step; at D = 16 (demo_maxcut.py as shipped,
// H2 VQE: BASELINE configs[0], [1]) every one of them is a 16 x 16 problem and the path is launch-bound.  Here a
// trajectory  psi <- expm(-i dt (H0 + sum_h u_h(t_k) H_h)) psi, k < steps  (sim_plain.py:135-150, diffqc.cc:190-200)
// runs start to finish inside one warp: the generator A = -i dt/2^s H(t_k) is assembled in registers (lane = row r,
// half h: eight entries of row r), the ket lives in shared memory, and exp(A) psi is the same block-Taylor Horner
// recurrence as strategy 0,  y <- psi + (A/j) y for j = m..1, applied 2^s times.  The estimator
// (sim_plain.py:186-220) is two launches: the prefix states phi_b, then all 2 n_H shifted kets of every sample
// with the shift gate (I +/- i r H_i)/sqrt(1+r^2) fused in front and Re <ket|M|ket> fused behind.
// mode 1 is the per-term product of diffqc.cc:155-164 (one generator per term, list order).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>
#include "dense.cuh"

namespace dq {
namespace dense {
namespace {

constexpr int kD = 16;          // padded dimension
constexpr int kWarps = 4;       // trajectories per CTA
constexpr int kMaxDegree = 32;

__device__ __forceinline__ double2 cfma(const double2 a, const double2 b, double2 acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
    return acc;
}

// sum over this lane's eight columns, then over the two halves of the row: every lane of the pair gets (Mat x)_r
__device__ __forceinline__ double2 pair_sum(double2 acc) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
    return acc;
}

// Shared-memory slot of ket entry i: the upper half is shifted by one entry, so that the two halves of a row pair
// (lanes h = 0 / 1 read entries c and 8 + c in the same LDS.128) fall into different banks.
__device__ __forceinline__ int xslot(int i) { return i + (i >> 3); }
constexpr int kXs = kD + 1;

__device__ __forceinline__ double2 matvec_global(const double2* __restrict__ mat, const double2* x, int r, int h) {
    double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        acc0 = cfma(__ldg(mat + r * kD + 8 * h + c), x[9 * h + c], acc0);
        acc1 = cfma(__ldg(mat + r * kD + 8 * h + c + 1), x[9 * h + c + 1], acc1);
    }
    return pair_sum(make_double2(acc0.x + acc1.x, acc0.y + acc1.y));
}

// NK kets per warp share one trajectory (same pulse rows, same generator): ket g starts from the same source ket with the
// shift gate of control term + g/2 and sign (+, -, +, -); its result goes to out + g.  NK = 1: a plain trajectory.
template <bool ENERGY, int NK>
__global__ void __launch_bounds__(32 * kWarps, NK == 1 ? 6 : 4)
k_small(const double2* __restrict__ H, const double2* __restrict__ M, int n_H, int mode, int reps, int m,
        const double* __restrict__ u, const SmallTraj* __restrict__ traj, int n_traj, const double2* __restrict__ src,
        double2* __restrict__ dst_kets, double* __restrict__ dst_energy, double inv_norm, const int* __restrict__ d_s) {
    __shared__ double2 xs[kWarps][3][NK * kXs];
    __shared__ double inv_j[kMaxDegree + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x <= kMaxDegree) inv_j[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    const int w = blockIdx.x * kWarps + warp;
    if (w >= n_traj) return;
    SmallTraj t = traj[w];
    if (d_s) {                              // device-resident training: the scaling of the series was chosen on the device
        const int sd = *d_s;
        reps = 1 << sd;
        t.scale = ldexp(t.scale, -sd);
    }
    const int r = lane >> 1, h = lane & 1;
    double2* cur = xs[warp][0];
    double2* o1 = xs[warp][1];
    double2* o2 = xs[warp][2];
    if (lane < kD) {
        const double2 v = src[(size_t)t.src * kD + lane];
#pragma unroll
        for (int g = 0; g < NK; ++g) cur[g * kXs + xslot(lane)] = v;
    }
    __syncwarp();
    if (t.shift != 0.0) {                   // (phi + i shift H_i phi) / sqrt(1 + r^2)     (sim_plain.py:197-199)
#pragma unroll
        for (int g = 0; g < NK; ++g) {
            const double sh = (g & 1) ? -t.shift : t.shift;
            const double2 hp = matvec_global(H + (size_t)(t.term + (g >> 1) + 1) * kD * kD, cur + g * kXs, r, h);
            const double2 c = cur[g * kXs + xslot(r)];
            if (h == 0) o1[g * kXs + xslot(r)] = make_double2((c.x - sh * hp.y) * inv_norm, (c.y + sh * hp.x) * inv_norm);
        }
        __syncwarp();
        double2* tmp = cur; cur = o1; o1 = tmp;
    }
    const int t_lo = mode == 0 ? -1 : 0, t_hi = mode == 0 ? -1 : n_H;
    for (int k = 0; k < t.steps; ++k) {
        const double* __restrict__ ur = u + (t.row + k) * n_H;
        for (int term = t_lo; term <= t_hi; ++term) {
            // generator entries of this lane: A[r][8h .. 8h+7] = -i scale (H0 + sum_h u_h H_h)  or one term of it
            double2 a[8];
            const double2* __restrict__ hrow = H + r * kD + 8 * h;
            if (term < 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = __ldg(hrow + c);
                for (int q = 0; q < n_H; ++q) {
                    const double cq = __ldg(ur + q);
                    const double2* __restrict__ hq = hrow + (size_t)(q + 1) * kD * kD;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double2 e = __ldg(hq + c);
                        a[c].x = fma(cq, e.x, a[c].x);
                        a[c].y = fma(cq, e.y, a[c].y);
                    }
                }
            } else {
                const double cq = term == 0 ? 1.0 : __ldg(ur + term - 1);
                const double2* __restrict__ hq = hrow + (size_t)term * kD * kD;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double2 e = __ldg(hq + c);
                    a[c] = make_double2(cq * e.x, cq * e.y);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = make_double2(t.scale * a[c].y, -t.scale * a[c].x);   // -i (x + i y) = y - i x
            for (int rep = 0; rep < reps; ++rep) {
                double2 c0[NK];
#pragma unroll
                for (int g = 0; g < NK; ++g) c0[g] = cur[g * kXs + xslot(r)];
                const double2* in = cur;
                for (int j = m; j >= 1; --j) {
                    // 2 NK independent accumulation chains (NK = 1: the row is split in two): the recurrence is latency-bound
                    double2 acc[NK], acc1[NK];
#pragma unroll
                    for (int g = 0; g < NK; ++g) acc[g] = acc1[g] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
#pragma unroll
                        for (int g = 0; g < NK; ++g) {
                            acc[g] = cfma(a[c], in[g * kXs + 9 * h + c], acc[g]);
                            acc1[g] = cfma(a[c + 1], in[g * kXs + 9 * h + c + 1], acc1[g]);
                        }
                    }
                    const double f = inv_j[j];
#pragma unroll
                    for (int g = 0; g < NK; ++g) {
                        const double2 sum = pair_sum(make_double2(acc[g].x + acc1[g].x, acc[g].y + acc1[g].y));
                        if (h == 0) o1[g * kXs + xslot(r)] = make_double2(fma(f, sum.x, c0[g].x), fma(f, sum.y, c0[g].y));
                    }
                    __syncwarp();
                    in = o1;
                    double2* tmp = o1; o1 = o2; o2 = tmp;
                }
                // result sits in `in` (= o2 after the last swap); the old `cur` becomes a work buffer
                double2* res = o2;
                o2 = cur;
                cur = res;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < NK; ++g) {
        if (ENERGY) {                       // Re <ket| M |ket>     (sim_plain.py:205,215)
            const double2 mk = matvec_global(M, cur + g * kXs, r, h);
            const double2 c = cur[g * kXs + xslot(r)];
            double e = h == 0 ? fma(c.x, mk.x, c.y * mk.y) : 0.0;
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if (lane == 0) dst_energy[t.out + g] = e;
        } else {
            if (lane < kD) dst_kets[(size_t)(t.out + g) * kD + lane] = cur[g * kXs + xslot(lane)];
        }
    }
}

// ---- the same recurrence on the FP64 tensor cores ---------------------------------------------------------------------------
// All 2 n_H shifted kets of a sample share their generator (sim_plain.py:196-215: same H list, same step grid), so the
// Horner step  Y <- X + (A/j) Y  on the 16 x K block of kets is a complex 16x16 x 16xK matrix product: four real products,
// each a chain of mma.sync.m8n8k4.f64 (DMMA).  One warp owns up to 16 kets (KT ket tiles of 8) of ONE sample for the whole
// suffix trajectory; nothing but the per-step pulse row and the operator stack (shared memory) is read between launch and
// energy.  The block is kept TRANSPOSED, Z = Y^T (kets x states), so that the product is Z G with G = A^T the B operand:
//   * the accumulator layout of m8n8k4 (thread (g, q) holds columns 2q, 2q+1 of row g) read as an A operand (thread holds
//     column q of row g) is the same tile with its contraction index permuted to (0,2,4,6 | 1,3,5,7); G's B fragments are
//     loaded with the same permutation (it is fixed per step), so the result of one Horner step feeds the next one
//     without a single shuffle or shared-memory round trip;
//   * PTX has no negated operand for DMMA: the fragments of -Im G are kept next to Re G and Im G.
// Per Horner step and 16 kets: 64 DMMA (8 output tiles x 8-long accumulation chains) + 16 DFMA for the 1/j scaling.
__device__ __forceinline__ void dmma(double (&d)[2], const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// B fragments of G[c][n] = Mat[n][c] * (gr, gi) for this thread: index [ct][p][nn] <-> c = 8 ct + 2 q + p, n = 8 nn + g.
struct GFrag { double r[2][2][2], i[2][2][2], ni[2][2][2]; };

template <int KT>
__global__ void __launch_bounds__(32 * kWarps, KT == 1 ? 3 : 2)
k_small_mma(const double2* __restrict__ H, const double2* __restrict__ M, int n_H, int reps, int m,
            const double* __restrict__ u, const SmallTraj* __restrict__ traj, int n_traj, const double2* __restrict__ src,
            double* __restrict__ dst_energy, double inv_norm, const int* __restrict__ d_s) {
    extern __shared__ double2 smem_mma[];
    double2* Hs = smem_mma;                                     // [(2 + n_H)][16][16]: H0, H_1..H_nH, M
    double2* stage = Hs + (size_t)(2 + n_H) * kD * kD;           // [kWarps][8 KT][17] kets of a warp's chunk (+ phi)
    __shared__ double inv_j[kMaxDegree + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x <= kMaxDegree) inv_j[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    for (int i = threadIdx.x; i < (2 + n_H) * kD * kD; i += blockDim.x) Hs[i] = __ldg(H + i);   // M follows the H stack
    __syncthreads();
    const int w = blockIdx.x * kWarps + warp;
    if (w >= n_traj) return;
    SmallTraj t = traj[w];
    if (d_s) {
        const int sd = *d_s;
        reps = 1 << sd;
        t.scale = ldexp(t.scale, -sd);
    }
    const int g = lane >> 2, q = lane & 3;
    double2* st = stage + (size_t)warp * (8 * KT + 1) * kXs;
    double2* phi = st + (size_t)8 * KT * kXs;

    // ---- shift gates: ket k = (phi + sigma_k i r H_{term + k/2} phi) / sqrt(1 + r^2), sigma = +, -, +, - ...  (sim_plain.py:197-199)
    if (lane < kD) phi[lane] = src[(size_t)t.src * kD + lane];
    __syncwarp();
    const int n_valid = min(8 * KT, 2 * (n_H - t.term));
    for (int k = 0; k < 8 * KT; ++k) {
        if (lane < kD) {
            double2 v = make_double2(0.0, 0.0);
            if (k < n_valid) {
                const double2* __restrict__ hrow = Hs + (size_t)(t.term + (k >> 1) + 1) * kD * kD + lane * kD;
                double2 hp = make_double2(0.0, 0.0);
#pragma unroll
                for (int c = 0; c < kD; ++c) hp = cfma(hrow[c], phi[c], hp);
                const double sh = (k & 1) ? -t.shift : t.shift;
                const double2 c0 = phi[lane];
                v = make_double2((c0.x - sh * hp.y) * inv_norm, (c0.y + sh * hp.x) * inv_norm);
            }
            st[k * kXs + lane] = v;
        }
    }
    __syncwarp();
    // Z fragments: [rt][ct][e] <-> ket 8 rt + g, state 8 ct + 2 q + e
    double zr[KT][2][2], zi[KT][2][2];
#pragma unroll
    for (int rt = 0; rt < KT; ++rt)
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double2 v = st[(8 * rt + g) * kXs + 8 * ct + 2 * q + e];
                zr[rt][ct][e] = v.x;
                zi[rt][ct][e] = v.y;
            }

    auto load_g = [&](GFrag& G, const double* __restrict__ ur, const double scale) {
        // G = A^T with A = -i scale (H0 + sum_h u_h H_h):  Re A = scale Im H,  Im A = -scale Re H
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const int idx = (8 * nn + g) * kD + 8 * ct + 2 * q + p;          // Mat[n][c]
                    double2 a = Hs[idx];
                    for (int h = 0; h < n_H; ++h) {
                        const double cq = __ldg(ur + h);
                        const double2 e = Hs[(size_t)(h + 1) * kD * kD + idx];
                        a.x = fma(cq, e.x, a.x);
                        a.y = fma(cq, e.y, a.y);
                    }
                    G.r[ct][p][nn] = scale * a.y;
                    G.i[ct][p][nn] = -scale * a.x;
                    G.ni[ct][p][nn] = scale * a.x;
                }
    };
    // D = Y G (complex): dr = yr Gr + yi (-Gi), di = yr Gi + yi Gr; output tile (rt, nn), contraction over (ct, p)
    auto product = [&](const double (&yr)[KT][2][2], const double (&yi)[KT][2][2], const GFrag& G, double (&dr)[KT][2][2],
                       double (&di)[KT][2][2]) {
#pragma unroll
        for (int rt = 0; rt < KT; ++rt)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) dr[rt][nn][0] = dr[rt][nn][1] = di[rt][nn][0] = di[rt][nn][1] = 0.0;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int rt = 0; rt < KT; ++rt)
#pragma unroll
                    for (int nn = 0; nn < 2; ++nn) {
                        dmma(dr[rt][nn], yr[rt][ct][p], G.r[ct][p][nn]);
                        dmma(dr[rt][nn], yi[rt][ct][p], G.ni[ct][p][nn]);
                        dmma(di[rt][nn], yr[rt][ct][p], G.i[ct][p][nn]);
                        dmma(di[rt][nn], yi[rt][ct][p], G.r[ct][p][nn]);
                    }
    };

    for (int k = 0; k < t.steps; ++k) {
        GFrag G;
        load_g(G, u + (t.row + k) * n_H, t.scale);
        for (int rep = 0; rep < reps; ++rep) {
            double yr[KT][2][2], yi[KT][2][2];
#pragma unroll
            for (int rt = 0; rt < KT; ++rt)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { yr[rt][ct][e] = zr[rt][ct][e]; yi[rt][ct][e] = zi[rt][ct][e]; }
            for (int j = m; j >= 1; --j) {          // y <- x + (A / j) y
                double dr[KT][2][2], di[KT][2][2];
                product(yr, yi, G, dr, di);
                const double f = inv_j[j];
#pragma unroll
                for (int rt = 0; rt < KT; ++rt)
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            yr[rt][ct][e] = fma(f, dr[rt][ct][e], zr[rt][ct][e]);
                            yi[rt][ct][e] = fma(f, di[rt][ct][e], zi[rt][ct][e]);
                        }
            }
#pragma unroll
            for (int rt = 0; rt < KT; ++rt)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { zr[rt][ct][e] = yr[rt][ct][e]; zi[rt][ct][e] = yi[rt][ct][e]; }
        }
    }
    // ---- energies: W = Z M^T, E_k = sum_c Re conj(Z[k][c]) W[k][c]      (sim_plain.py:205,215) ---------------------------------
    {
        GFrag G;
        const double2* Ms = Hs + (size_t)(1 + n_H) * kD * kD;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const double2 a = Ms[(8 * nn + g) * kD + 8 * ct + 2 * q + p];
                    G.r[ct][p][nn] = a.x;
                    G.i[ct][p][nn] = a.y;
                    G.ni[ct][p][nn] = -a.y;
                }
        double wr[KT][2][2], wi[KT][2][2];
        product(zr, zi, G, wr, wi);
#pragma unroll
        for (int rt = 0; rt < KT; ++rt) {
            double e = 0.0;
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int c = 0; c < 2; ++c) e += zr[rt][ct][c] * wr[rt][ct][c] + zi[rt][ct][c] * wi[rt][ct][c];
            e += __shfl_xor_sync(0xffffffffu, e, 1);
            e += __shfl_xor_sync(0xffffffffu, e, 2);
            if (q == 0 && 8 * rt + g < n_valid) dst_energy[t.out + 8 * rt + g] = e;
        }
    }
}

// interleaved c128 [dim][dim] -> zero-padded [16][16]
void pad16(const double* src, int dim, double* dst) {
    std::fill(dst, dst + 2 * kD * kD, 0.0);
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
            dst[2 * (i * kD + j)] = src[2 * ((size_t)i * dim + j)];
            dst[2 * (i * kD + j) + 1] = src[2 * ((size_t)i * dim + j) + 1];
        }
}

}  // namespace

bool small_fits(const Problem& P) { return P.dim >= 1 && P.dim <= kD; }

int small_upload(dq_context* ctx, const Problem& P, const double* M) {
    State* S = state_of(ctx);
    const size_t per = (size_t)2 * kD * kD, src_per = (size_t)P.dim * P.dim * 2;
    std::vector<double> host((size_t)(2 + P.n_H) * per);
    for (int h = 0; h <= P.n_H; ++h) pad16(P.host_copy.data() + src_per * h, P.dim, host.data() + per * h);
    if (M) pad16(M, P.dim, host.data() + per * (1 + P.n_H));
    else std::fill(host.begin() + per * (1 + P.n_H), host.end(), 0.0);
    DQ_TRY(S->small_H.reserve(host.size() * sizeof(double)));
    DQ_CUDA(cudaMemcpyAsync(S->small_H.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));        // `host` goes out of scope
    return DQ_OK;
}

// Enqueue one launch of the resident engine on the context's stream; d_traj are DEVICE descriptors.  d_s: NULL, or a device int
// holding the squarings chosen on the device (then `s` is ignored and the descriptors carry the unscaled dt).
int small_enqueue(dq_context* ctx, const Problem& P, int mode, int s, int m, int kets_per_traj, const SmallTraj* d_traj, int n,
                  const double* d_u, const double* d_src, double* d_dst_kets, double* d_dst_energy, double inv_norm, const int* d_s) {
    State* S = state_of(ctx);
    DQ_REQUIRE(m >= 1 && m <= kMaxDegree && s >= 0 && s <= 20, "dense resident engine: degree %d / squarings %d out of range", m, s);
    DQ_REQUIRE(kets_per_traj == 1 || ((kets_per_traj == 2 || kets_per_traj == 4 || kets_per_traj == 8 || kets_per_traj == 16) && d_dst_energy),
               "dense resident engine: %d kets per trajectory", kets_per_traj);
    if (n <= 0) return DQ_OK;
    const double2* H = S->small_H.as<double2>();
    const double2* M = H + (size_t)(1 + P.n_H) * kD * kD;
    const double2* src = reinterpret_cast<const double2*>(d_src);
    const unsigned grid = (unsigned)((n + kWarps - 1) / kWarps);
    const int reps = 1 << s;
    if (d_dst_energy && kets_per_traj >= 8) {                 // suffix trajectories of the estimator, exact step: tensor cores
        DQ_REQUIRE(mode == 0 && (kets_per_traj == 8 || kets_per_traj == 16), "dense resident engine: DMMA path is mode 0, 8 or 16 kets per warp");
        const size_t smem = ((size_t)(2 + P.n_H) * kD * kD + (size_t)kWarps * (kets_per_traj + 1) * kXs) * sizeof(double2);
        DQ_REQUIRE(smem <= 200 * 1024, "dense resident engine: %d controls do not fit the shared-memory operator stack", P.n_H);
        if (kets_per_traj == 8) {
            DQ_CUDA(cudaFuncSetAttribute(k_small_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_small_mma<1><<<grid, 32 * kWarps, smem, ctx->stream>>>(H, M, P.n_H, reps, m, d_u, d_traj, n, src, d_dst_energy, inv_norm, d_s);
        } else {
            DQ_CUDA(cudaFuncSetAttribute(k_small_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_small_mma<2><<<grid, 32 * kWarps, smem, ctx->stream>>>(H, M, P.n_H, reps, m, d_u, d_traj, n, src, d_dst_energy, inv_norm, d_s);
        }
    } else if (!d_dst_energy)
        k_small<false, 1><<<grid, 32 * kWarps, 0, ctx->stream>>>(H, M, P.n_H, mode, reps, m, d_u, d_traj, n, src,
                                                                 reinterpret_cast<double2*>(d_dst_kets), nullptr, inv_norm, d_s);
    else if (kets_per_traj == 1)
        k_small<true, 1><<<grid, 32 * kWarps, 0, ctx->stream>>>(H, M, P.n_H, mode, reps, m, d_u, d_traj, n, src, nullptr, d_dst_energy, inv_norm, d_s);
    else if (kets_per_traj == 2)
        k_small<true, 2><<<grid, 32 * kWarps, 0, ctx->stream>>>(H, M, P.n_H, mode, reps, m, d_u, d_traj, n, src, nullptr, d_dst_energy, inv_norm, d_s);
    else
        k_small<true, 4><<<grid, 32 * kWarps, 0, ctx->stream>>>(H, M, P.n_H, mode, reps, m, d_u, d_traj, n, src, nullptr, d_dst_energy, inv_norm, d_s);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

// d_src / d_dst_kets: [..][16] c128; d_u: packed pulse rows [..][n_H]; traj: host descriptors (copied here)
int small_run(dq_context* ctx, const Problem& P, int mode, int s, int m, int kets_per_traj, const std::vector<SmallTraj>& traj,
              const double* d_u, const double* d_src, double* d_dst_kets, double* d_dst_energy, double inv_norm) {
    State* S = state_of(ctx);
    if (traj.empty()) return DQ_OK;
    DQ_TRY(S->small_traj.reserve(traj.size() * sizeof(SmallTraj)));
    DQ_CUDA(cudaMemcpyAsync(S->small_traj.p, traj.data(), traj.size() * sizeof(SmallTraj), cudaMemcpyHostToDevice, ctx->stream));
    if (!S->ev0) {
        DQ_CUDA(cudaEventCreate(&S->ev0));
        DQ_CUDA(cudaEventCreate(&S->ev1));
    }
    DQ_CUDA(cudaEventRecord(S->ev0, ctx->stream));
    DQ_TRY(small_enqueue(ctx, P, mode, s, m, kets_per_traj, S->small_traj.as<SmallTraj>(), (int)traj.size(), d_u, d_src, d_dst_kets,
                         d_dst_energy, inv_norm, nullptr));
    DQ_CUDA(cudaEventRecord(S->ev1, ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));        // the descriptor buffer is reused by the next call
    float ms = 0.f;
    DQ_CUDA(cudaEventElapsedTime(&ms, S->ev0, S->ev1));
    S->last_kernel_ms += ms;
    return DQ_OK;
}

}  // namespace dense
}  // namespace dq
