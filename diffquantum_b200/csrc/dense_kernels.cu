// Kernels of the dense path: batched complex FP64 GEMM on the tensor cores (DMMA, mma.sync.m8n8k4.f64),
// generator assembly, dense shift-gate fan-out and dense-observable energies.
#include "dense.cuh"

namespace dq {
namespace dense {
namespace {

constexpr int BK = 16;                 // k extent of a shared-memory stage (4 DMMA k-steps)

__device__ __forceinline__ void dmma(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CTA tile BM x BN, WM x WN warps, each warp (BM/WM) x (BN/WN) as 8x8 DMMA tiles.
// Fragment ownership of m8n8k4 (lane = 4*g + t):  A[g][t], B[t][g], C[g][2t], C[g][2t+1].
// Shared tiles are padded (+4 doubles per row) so both fragment loads are bank-conflict free.  Two stages: the k-slab k+1 is
// on its way (cp.async, 16-byte pieces, zero fill outside the matrices) while the DMMAs run on slab k.
template <int BM, int BN>
struct ZgemmSmem {
    static constexpr int SA = BK + 4, SB = BN + 4;
    double As[2][2][BM][SA];           // [stage][plane][row][k]
    double Bs[2][2][BK][SB];           // [stage][plane][k][col]
};

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32) k_zgemm(const Gemm g) {
    constexpr int NT = WM * WN * 32;
    constexpr int TM = BM / WM / 8, TN = BN / WN / 8;
    extern __shared__ __align__(16) unsigned char zgemm_smem[];
    ZgemmSmem<BM, BN>& S = *reinterpret_cast<ZgemmSmem<BM, BN>*>(zgemm_smem);

    const int z = blockIdx.z;
    const double* __restrict__ A = g.A + (long long)z * g.strideA;
    const double* __restrict__ B = g.B + (long long)z * g.strideB;
    double* __restrict__ C = g.C + (long long)z * g.strideC;
    const double* __restrict__ Add = g.Add ? g.Add + (long long)z * g.strideC : nullptr;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WN) * (BM / WM), wn0 = (warp % WN) * (BN / WN);
    const int gq = lane >> 2, tq = lane & 3;

    double cr[TM][TN][2], ci[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

    // stage A[m0:m0+BM, k0:k0+BK] and B[k0:k0+BK, n0:n0+BN], both planes
    auto issue = [&](int st, int k0) {
        for (int v = tid; v < BM * (BK / 2) * 2; v += NT) {
            const int pl = v / (BM * (BK / 2)), w = v % (BM * (BK / 2));
            const int r = w / (BK / 2), c = (w % (BK / 2)) * 2;
            const bool ok = m0 + r < g.M && k0 + c < g.K;
            cp_async16_zfill(&S.As[st][pl][r][c], ok ? A + pl * g.planeA + (long long)(m0 + r) * g.lda + k0 + c : A, ok);
        }
        for (int v = tid; v < BK * (BN / 2) * 2; v += NT) {
            const int pl = v / (BK * (BN / 2)), w = v % (BK * (BN / 2));
            const int r = w / (BN / 2), c = (w % (BN / 2)) * 2;
            const bool ok = k0 + r < g.K && n0 + c < g.N;
            cp_async16_zfill(&S.Bs[st][pl][r][c], ok ? B + pl * g.planeB + (long long)(k0 + r) * g.ldb + n0 + c : B, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nk = (g.K + BK - 1) / BK;
    issue(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nk) {
            issue(st ^ 1, (kt + 1) * BK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double ar[TM], ai[TM], nai[TM], br[TN], bi[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                ar[i] = S.As[st][0][wm0 + i * 8 + gq][kk + tq];
                ai[i] = S.As[st][1][wm0 + i * 8 + gq][kk + tq];
                nai[i] = -ai[i];
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                br[j] = S.Bs[st][0][kk + tq][wn0 + j * 8 + gq];
                bi[j] = S.Bs[st][1][kk + tq][wn0 + j * 8 + gq];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    dmma(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                    dmma(cr[i][j][0], cr[i][j][1], nai[i], bi[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
                    dmma(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
                }
        }
        __syncthreads();               // everybody is done with stage st before the next iteration refills it
    }
    // epilogue: C = alpha * acc (+ Add) (+ I)
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int row = m0 + wm0 + i * 8 + gq, col = n0 + wn0 + j * 8 + 2 * tq;
            if (row >= g.M || col >= g.N) continue;
            const long long o = (long long)row * g.ldc + col;
            double2 vr = make_double2(g.alpha * cr[i][j][0], g.alpha * cr[i][j][1]);
            double2 vi = make_double2(g.alpha * ci[i][j][0], g.alpha * ci[i][j][1]);
            if (Add) {
                const double2 xr = *reinterpret_cast<const double2*>(Add + o);
                const double2 xi = *reinterpret_cast<const double2*>(Add + g.planeC + o);
                vr.x += xr.x; vr.y += xr.y; vi.x += xi.x; vi.y += xi.y;
            }
            if (g.add_identity) {
                if (row == col) vr.x += 1.0;
                if (row == col + 1) vr.y += 1.0;
            }
            *reinterpret_cast<double2*>(C + o) = vr;
            *reinterpret_cast<double2*>(C + g.planeC + o) = vi;
        }
}

// Skinny case (N == 8, i.e. a block of <= 8 kets: diffqc.trotter, the solver hook, the estimator's prefix): the product
// is a batched mat-vec, bound by streaming A once, so tensor-core tiles only waste parallelism (a 64-row tile leaves
// 16 CTAs at dim 1024).  One warp per row: lanes stride over k with coalesced loads of A, the 8 ket columns of B come
// from L1, 16 accumulators per lane, butterfly reduction, fused "+ Add" epilogue.
__global__ void __launch_bounds__(256) k_zgemm_skinny(const Gemm g) {
    const int z = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= g.M) return;
    const double* __restrict__ Ar = g.A + (long long)z * g.strideA + (long long)row * g.lda;
    const double* __restrict__ Ai = Ar + g.planeA;
    const double* __restrict__ Br = g.B + (long long)z * g.strideB;
    const double* __restrict__ Bi = Br + g.planeB;
    double cr[8], ci[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cr[c] = ci[c] = 0.0;
    for (int k = lane; k < g.K; k += 32) {
        const double ar = Ar[k], ai = Ai[k];
        const double2* br = reinterpret_cast<const double2*>(Br + (long long)k * g.ldb);
        const double2* bi = reinterpret_cast<const double2*>(Bi + (long long)k * g.ldb);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 xr = __ldg(br + c), xi = __ldg(bi + c);
            cr[2 * c] = fma(ar, xr.x, fma(-ai, xi.x, cr[2 * c]));
            ci[2 * c] = fma(ar, xi.x, fma(ai, xr.x, ci[2 * c]));
            cr[2 * c + 1] = fma(ar, xr.y, fma(-ai, xi.y, cr[2 * c + 1]));
            ci[2 * c + 1] = fma(ar, xi.y, fma(ai, xr.y, ci[2 * c + 1]));
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        for (int o = 16; o > 0; o >>= 1) {
            cr[c] += __shfl_xor_sync(0xffffffffu, cr[c], o);
            ci[c] += __shfl_xor_sync(0xffffffffu, ci[c], o);
        }
    if (lane < 8) {
        double vr = 0.0, vi = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c == lane) { vr = cr[c]; vi = ci[c]; }
        const long long o = (long long)z * g.strideC + (long long)row * g.ldc + lane;
        vr *= g.alpha;
        vi *= g.alpha;
        if (g.Add) { vr += g.Add[o]; vi += g.Add[o + g.planeC]; }
        if (g.add_identity && row == lane) vr += 1.0;
        g.C[o] = vr;
        g.C[o + g.planeC] = vi;
    }
}

template <int BM, int BN, int WM, int WN>
int launch_zgemm(dq_context* ctx, const Gemm& g) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch);
    const size_t smem = sizeof(ZgemmSmem<BM, BN>);
    if (smem > (48u << 10))            // per device and cheap to set
        DQ_CUDA(cudaFuncSetAttribute(k_zgemm<BM, BN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_zgemm<BM, BN, WM, WN><<<grid, WM * WN * 32, smem, ctx->stream>>>(g);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

// A[z] = -i scale[z] (H0 + sum_h u[rows[z] + k][h] H_h)      (term < 0)
//      = -i scale[z] c H_term, c = 1 for term 0 and u[.][term-1] otherwise   (per-term product, diffqc.cc:155-164)
// optionally P[z] = I + inv_m * A[z]  (innermost Horner factor of the Taylor polynomial)
__global__ void k_build(const double* __restrict__ H, int n_H, int Dp, const double* __restrict__ u,
                        const long long* __restrict__ rows, const double* __restrict__ scale, int k, int term,
                        double* __restrict__ A, double* __restrict__ P, double inv_m) {
    const size_t plane = (size_t)Dp * Dp;
    const int z = blockIdx.y;
    const double* ur = u + (rows[z] + k) * n_H;
    const double sc = scale[z];
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
        double hr, hi;
        if (term < 0) {
            hr = H[e];
            hi = H[plane + e];
            for (int h = 0; h < n_H; ++h) {
                const double c = ur[h];
                hr = fma(c, H[(size_t)(h + 1) * 2 * plane + e], hr);
                hi = fma(c, H[(size_t)(h + 1) * 2 * plane + plane + e], hi);
            }
        } else {
            const double c = term == 0 ? 1.0 : ur[term - 1];
            hr = c * H[(size_t)term * 2 * plane + e];
            hi = c * H[(size_t)term * 2 * plane + plane + e];
        }
        const double ar = sc * hi, ai = -sc * hr;                 // -i (hr + i hi) = hi - i hr
        A[(size_t)z * 2 * plane + e] = ar;
        A[(size_t)z * 2 * plane + plane + e] = ai;
        if (P) {
            const bool diag = (e / Dp) == (e % Dp);
            P[(size_t)z * 2 * plane + e] = inv_m * ar + (diag ? 1.0 : 0.0);
            P[(size_t)z * 2 * plane + plane + e] = inv_m * ai;
        }
    }
}

// dst[z] = src[order[z]] (gather) or dst[order[z]] = src[z] (scatter); blocks of `n` doubles
__global__ void k_move_blocks(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ order,
                              size_t n, int scatter) {
    const int z = blockIdx.y;
    const double* s = src + (size_t)(scatter ? z : order[z]) * n;
    double* d = dst + (size_t)(scatter ? order[z] : z) * n;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) d[e] = s[e];
}

// K[b][:, 2i + sg] = (phi_b + sg' i r H_i phi_b) / sqrt(1 + r^2), sg' = +1, -1   (sim_plain.py:197-199)
// one CTA per (term i, sample b); phi_b is column 0 of a block with leading dimension ncp_phi.
__global__ void k_fanout(const double* __restrict__ H, int Dp, int dim, const double* __restrict__ phi, int ncp_phi,
                         double* __restrict__ K, int ncp, double r, double inv_norm) {
    const size_t plane = (size_t)Dp * Dp;
    const int i = blockIdx.x, b = blockIdx.y;
    const double* Hr = H + (size_t)(i + 1) * 2 * plane;
    const double* Hi = Hr + plane;
    const double* pr = phi + (size_t)b * 2 * Dp * ncp_phi;
    const double* pi = pr + (size_t)Dp * ncp_phi;
    double* Kr = K + (size_t)b * 2 * Dp * ncp;
    double* Ki = Kr + (size_t)Dp * ncp;
    for (int x = threadIdx.x; x < Dp; x += blockDim.x) {
        double hr = 0.0, hi = 0.0;
        if (x < dim)
            for (int y = 0; y < dim; ++y) {
                const double a = Hr[(size_t)x * Dp + y], c = Hi[(size_t)x * Dp + y];
                const double vr = pr[(size_t)y * ncp_phi], vi = pi[(size_t)y * ncp_phi];
                hr += a * vr - c * vi;
                hi += a * vi + c * vr;
            }
        const double fr = pr[(size_t)x * ncp_phi], fi = pi[(size_t)x * ncp_phi];
        // phi + i r hp = (fr - r hi) + i (fi + r hr)
        Kr[(size_t)x * ncp + 2 * i] = (fr - r * hi) * inv_norm;
        Ki[(size_t)x * ncp + 2 * i] = (fi + r * hr) * inv_norm;
        Kr[(size_t)x * ncp + 2 * i + 1] = (fr + r * hi) * inv_norm;
        Ki[(size_t)x * ncp + 2 * i + 1] = (fi - r * hr) * inv_norm;
    }
}

// out[b][c] = Re <k_c| M |k_c>   (sim_plain.py:205,215); one CTA per (column c, sample b)
__global__ void k_energy(const double* __restrict__ M, int Dp, int dim, const double* __restrict__ K, int ncp,
                         int n_cols, double* __restrict__ out) {
    const size_t plane = (size_t)Dp * Dp;
    const int c = blockIdx.x, b = blockIdx.y;
    const double* Kr = K + (size_t)b * 2 * Dp * ncp + c;
    const double* Ki = Kr + (size_t)Dp * ncp;
    double acc = 0.0;
    for (int x = threadIdx.x; x < dim; x += blockDim.x) {
        double mr = 0.0, mi = 0.0;
        for (int y = 0; y < dim; ++y) {
            const double a = M[(size_t)x * Dp + y], d = M[plane + (size_t)x * Dp + y];
            const double vr = Kr[(size_t)y * ncp], vi = Ki[(size_t)y * ncp];
            mr += a * vr - d * vi;
            mi += a * vi + d * vr;
        }
        acc += Kr[(size_t)x * ncp] * mr + Ki[(size_t)x * ncp] * mi;      // Re conj(k) (M k)
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[(size_t)b * n_cols + c] = v;
    }
}

}  // namespace

int zgemm(dq_context* ctx, const Gemm& g) {
    DQ_REQUIRE(g.M % 8 == 0 && g.N % 8 == 0 && g.K % 8 == 0 && g.batch >= 1, "zgemm: dims must be multiples of 8");
    if (g.N == 8 && g.M >= 64) {
        dim3 grid(g.M / 8, g.batch);
        k_zgemm_skinny<<<grid, 256, 0, ctx->stream>>>(g);
        ctx->launches++;
        DQ_CUDA(cudaGetLastError());
        return DQ_OK;
    }
    if (g.M >= 64) {
        if (g.N >= 64) return launch_zgemm<64, 64, 2, 2>(ctx, g);
        if (g.N >= 16) return launch_zgemm<64, 16, 4, 1>(ctx, g);
        return launch_zgemm<64, 8, 4, 1>(ctx, g);
    }
    if (g.N >= 16) return launch_zgemm<16, 16, 1, 1>(ctx, g);
    return launch_zgemm<16, 8, 1, 1>(ctx, g);
}

int build_generator(dq_context* ctx, const Problem& P, int nb, const double* d_u, const long long* d_rows,
                    const double* d_scale, int k, int term, double* d_A, double* d_P, double inv_m) {
    const size_t plane = P.plane();
    dim3 grid((unsigned)std::min<size_t>((plane + 255) / 256, 1024), nb);
    k_build<<<grid, 256, 0, ctx->stream>>>(P.H.as<double>(), P.n_H, P.Dp, d_u, d_rows, d_scale, k, term, d_A, d_P, inv_m);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gather_blocks(dq_context* ctx, const double* src, double* dst, const int* d_order, int nb, size_t n, int scatter) {
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 256), nb);
    k_move_blocks<<<grid, 256, 0, ctx->stream>>>(src, dst, d_order, n, scatter);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int fanout(dq_context* ctx, const Problem& P, int B, const double* d_phi, int ncp_phi, double* d_K, int ncp, double r) {
    dim3 grid(P.n_H, B);
    k_fanout<<<grid, 128, 0, ctx->stream>>>(P.H.as<double>(), P.Dp, P.dim, d_phi, ncp_phi, d_K, ncp, r,
                                            1.0 / sqrt(1.0 + r * r));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int energies(dq_context* ctx, const Problem& P, int B, const double* d_K, int ncp, int n_cols, double* d_out) {
    dim3 grid(n_cols, B);
    k_energy<<<grid, 128, 0, ctx->stream>>>(P.M.as<double>(), P.Dp, P.dim, d_K, ncp, n_cols, d_out);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

}  // namespace dense
}  // namespace dq
