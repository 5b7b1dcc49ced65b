// Generic engine for the structured path: one streaming kernel per diagonal phase and per X
// rotation.  Works for any n >= 1; it is the small-n path (n < 12) and the on-device cross-check
// for the fused engine.  Step semantics: diffqc.cc:155-164 (H0 / ZZ phases, then X rotations).
#include "ising.cuh"

namespace dq {
namespace {

constexpr int kThreads = 256;

__global__ void k_fill_uniform(c128* psi, size_t total, double amp) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) psi[i] = make_double2(amp, 0.0);
}

// dst[phys(x)] = src[x] (to_phys) or dst[x] = src[phys(x)]; bit k of x (reference order) moves to
// physical bit map[k].
struct BitMap { int8_t map[40]; };
template <typename T>
__global__ void k_permute(const T* __restrict__ src, T* __restrict__ dst, int n, BitMap bm,
                          int to_phys) {
    size_t N = (size_t)1 << n;
    size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const T* s = src + blockIdx.y * N;
    T* d = dst + blockIdx.y * N;
    for (; x < N; x += stride) {
        size_t y = 0;
        for (int k = 0; k < n; ++k) y |= ((x >> k) & 1) << bm.map[k];
        if (to_phys) d[y] = s[x]; else d[x] = s[y];
    }
}

__global__ void k_trig(const double* __restrict__ rows, int64_t n_rows, int row_len, int off_x,
                       int n, double2* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_rows * n) return;
    int64_t r = i / n;
    int q = (int)(i - r * n);
    double s, c;
    sincos(rows[r * row_len + off_x + q], &s, &c);
    out[i] = make_double2(c, s);
}

// psi[x] *= exp(-i (a_c + sum_e a_e z_a z_b)); all `batch` states (grid.y) share the row.
__global__ void k_phase(c128* __restrict__ psi, int n, int n_zz, const int2* __restrict__ pairs,
                        const double* __restrict__ row) {
    extern __shared__ double sh[];
    double* ang = sh;                      // n_zz angles
    int2* pr = (int2*)(sh + n_zz);
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { ang[e] = row[1 + e]; pr[e] = pairs[e]; }
    __syncthreads();
    const double a_c = row[0];
    size_t N = (size_t)1 << n;
    c128* s = psi + blockIdx.y * N;
    size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; x < N; x += stride) {
        double a = a_c;
        for (int e = 0; e < n_zz; ++e) {
            int par = (int)(((x >> pr[e].x) ^ (x >> pr[e].y)) & 1);
            a += par ? -ang[e] : ang[e];
        }
        double sn, cs;
        sincos(a, &sn, &cs);
        c128 v = s[x];
        s[x] = make_double2(v.x * cs + v.y * sn, v.y * cs - v.x * sn);   // v * (cs - i sn)
    }
}

// exp(-i theta X) on physical bit `bit`: a' = c a - i s b ; b' = c b - i s a
__global__ void k_rx(c128* __restrict__ psi, int n, int bit, const double2* __restrict__ cs_ptr) {
    const double2 cs = *cs_ptr;
    size_t N = (size_t)1 << n;
    size_t half = N >> 1;
    c128* s = psi + blockIdx.y * N;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t low = ((size_t)1 << bit) - 1;
    for (; i < half; i += stride) {
        size_t x0 = ((i & ~low) << 1) | (i & low);
        size_t x1 = x0 | ((size_t)1 << bit);
        c128 a = s[x0], b = s[x1];
        s[x0] = make_double2(cs.x * a.x + cs.y * b.y, cs.x * a.y - cs.y * b.x);
        s[x1] = make_double2(cs.x * b.x + cs.y * a.y, cs.x * b.y - cs.y * a.x);
    }
}

// kets[k] = exp(sign_k i alpha P_k) phi, alpha = atan(r)   (sim_plain.py:197-199)
__global__ void k_fanout(const c128* __restrict__ phi, c128* __restrict__ kets, int n,
                         const ShiftDesc* __restrict__ desc, double ca, double sa) {
    const ShiftDesc d = desc[blockIdx.y];
    size_t N = (size_t)1 << n;
    c128* out = kets + blockIdx.y * N;
    size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const double s = d.sign * sa;
    for (; x < N; x += stride) {
        c128 v = phi[x];
        if (d.kind == 0) {
            double z = (((x >> d.b0) ^ (x >> d.b1)) & 1) ? -s : s;     // i * z * s
            out[x] = make_double2(ca * v.x - z * v.y, ca * v.y + z * v.x);
        } else {
            c128 w = phi[x ^ ((size_t)1 << d.b0)];
            out[x] = make_double2(ca * v.x - s * w.y, ca * v.y + s * w.x);
        }
    }
}

__global__ void k_energy(const c128* __restrict__ psi, const double* __restrict__ mdiag, int n,
                         double* __restrict__ partial) {
    size_t N = (size_t)1 << n;
    const c128* s = psi + blockIdx.y * N;
    size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (; x < N; x += stride) {
        c128 v = s[x];
        acc += mdiag[x] * (v.x * v.x + v.y * v.y);
    }
    __shared__ double red[kThreads / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
}

// Shot sampling (stochastic_measure, sim_plain.py:101-117) on Z-string observables: the outcome distribution of Z_a Z_b in a
// state collapses to P(-1) = (1 - <Z_a Z_b>) / 2, so per state and pair  zz[e] = sum_x |psi_x|^2 (-1)^(x_a xor x_b).
// One pass over the state serves kPairChunk pairs (register accumulators); partial[(state * gridDim.x + block) * n_zz + e].
constexpr int kPairChunk = 16;
__global__ void k_pair_expect(const c128* __restrict__ psi, int n, int n_zz, const int2* __restrict__ pairs, int e0,
                              double* __restrict__ partial) {
    const size_t N = (size_t)1 << n;
    const c128* s = psi + blockIdx.y * N;
    unsigned mask[kPairChunk];
    double acc[kPairChunk];
#pragma unroll
    for (int k = 0; k < kPairChunk; ++k) {
        const int e = e0 + k;
        mask[k] = e < n_zz ? ((1u << pairs[e].x) | (1u << pairs[e].y)) : 0u;
        acc[k] = 0.0;
    }
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const c128 v = s[x];
        const double pr = v.x * v.x + v.y * v.y;
#pragma unroll
        for (int k = 0; k < kPairChunk; ++k) acc[k] += (__popc((unsigned)x & mask[k]) & 1) ? -pr : pr;
    }
    __shared__ double red[kThreads / 32][kPairChunk];
#pragma unroll
    for (int k = 0; k < kPairChunk; ++k) {
        double a = acc[k];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = a;
    }
    __syncthreads();
    if (threadIdx.x < kPairChunk && e0 + threadIdx.x < n_zz) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * n_zz + e0 + threadIdx.x] = v;
    }
}

// out[state][e] = sum over blocks, fixed order
__global__ void k_sum_pair_partials(const double* __restrict__ partial, int blocks, int n_zz, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_zz) return;
    double acc = 0.0;
    for (int b = 0; b < blocks; ++b) acc += partial[((size_t)blockIdx.y * blocks + b) * n_zz + e];
    out[(size_t)blockIdx.y * n_zz + e] = acc;
}

__global__ void k_sum_partials(const double* __restrict__ partial, int per, double* __restrict__ out) {
    // deterministic: one warp per state sums its partials in a fixed order
    double acc = 0.0;
    for (int i = threadIdx.x; i < per; i += 32) acc += partial[blockIdx.x * per + i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

__global__ void k_build_mdiag(double* __restrict__ m, int n, int n_zz, const int2* __restrict__ pairs,
                              const double* __restrict__ w, double c0) {
    size_t N = (size_t)1 << n;
    size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; x < N; x += stride) {
        double a = c0;
        for (int e = 0; e < n_zz; ++e) {
            int par = (int)(((x >> pairs[e].x) ^ (x >> pairs[e].y)) & 1);
            a += par ? -w[e] : w[e];
        }
        m[x] = a;
    }
}

// ---- exact step (live reference semantics, sim_plain.py:135-150) without the dense matrix -----------------
// psi <- exp(-i (Diag + sum_q x_q X_q)) psi evaluated as a scaled Taylor series on the vector, the matrix-free
// variant the reference leaves commented at sim_plain.py:147 (expm_multiply).  diag[x] = c + sum_e g_e z_a z_b.
__global__ void k_build_diag(double* __restrict__ diag, int n, int n_zz, const int2* __restrict__ pairs,
                             const double* __restrict__ row) {
    extern __shared__ double sh[];
    double* ang = sh;
    int2* pr = (int2*)(sh + n_zz);
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { ang[e] = row[1 + e]; pr[e] = pairs[e]; }
    __syncthreads();
    const double c = row[0];
    const size_t N = (size_t)1 << n;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        double a = c;
        for (int e = 0; e < n_zz; ++e) a += (((x >> pr[e].x) ^ (x >> pr[e].y)) & 1) ? -ang[e] : ang[e];
        diag[x] = a;
    }
}

struct XTerms { int n; int bit[40]; double ang[40]; const double* rows; };   // rows != NULL: X angles read from the device row

// term_out = (scale / k) * (-i) * (diag .* term_in + sum_q ang_q term_in[x ^ bit_q]);  acc += term_out
__global__ void k_taylor_term(const c128* __restrict__ tin, c128* __restrict__ tout, c128* __restrict__ acc,
                              const double* __restrict__ diag, int n, XTerms xt, double f) {
    const size_t N = (size_t)1 << n;
    const c128* ti = tin + blockIdx.y * N;
    c128* to = tout + blockIdx.y * N;
    c128* ac = acc + blockIdx.y * N;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const c128 v = ti[x];
        const double d = diag[x];
        double hr = d * v.x, hi = d * v.y;
        for (int q = 0; q < xt.n; ++q) {
            const c128 w = ti[x ^ ((size_t)1 << xt.bit[q])];
            const double aq = xt.rows ? __ldg(xt.rows + q) : xt.ang[q];
            hr = fma(aq, w.x, hr);
            hi = fma(aq, w.y, hi);
        }
        const c128 t = make_double2(f * hi, -f * hr);          // -i (hr + i hi) = hi - i hr
        to[x] = t;
        const c128 a = ac[x];
        ac[x] = make_double2(a.x + t.x, a.y + t.y);
    }
}

inline int grid_for(size_t work, int sms) {
    size_t b = (work + kThreads - 1) / kThreads;
    size_t cap = (size_t)sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int gen_fill_uniform(dq_ising* p, c128* psi, int batch) {
    size_t total = p->dim() * batch;
    k_fill_uniform<<<grid_for(total, p->ctx->prop.multiProcessorCount), kThreads, 0, p->ctx->stream>>>(
        psi, total, 1.0 / sqrt((double)p->dim()));
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

template <typename T>
static int permute(dq_ising* p, const T* src, T* dst, int batch, int to_phys) {
    BitMap bm;
    for (int k = 0; k < p->n; ++k) bm.map[k] = (int8_t)p->bitpos[p->n - 1 - k];  // ref bit k = qubit n-1-k
    dim3 grid(grid_for(p->dim(), p->ctx->prop.multiProcessorCount), batch);
    k_permute<T><<<grid, kThreads, 0, p->ctx->stream>>>(src, dst, p->n, bm, to_phys);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}
int gen_permute_in(dq_ising* p, const c128* s, c128* d, int batch) { return permute(p, s, d, batch, 1); }
int gen_permute_out(dq_ising* p, const c128* s, c128* d, int batch) { return permute(p, s, d, batch, 0); }
int gen_permute_real_in(dq_ising* p, const double* s, double* d) { return permute(p, s, d, 1, 1); }

int gen_trig(dq_ising* p, const double* d_rows, int64_t n_rows, double2* d_trig) {
    if (n_rows == 0) return DQ_OK;
    int64_t total = n_rows * p->n;
    k_trig<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, p->ctx->stream>>>(
        d_rows, n_rows, p->row_len, 1 + p->n_zz, p->n, d_trig);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_evolve(dq_ising* p, c128* d_states, int batch, const double* d_rows, const double2* d_trig,
               int n_steps) {
    const int sms = p->ctx->prop.multiProcessorCount;
    cudaStream_t st = p->ctx->stream;
    dim3 gfull(grid_for(p->dim(), sms), batch), ghalf(grid_for(p->dim() / 2 ? p->dim() / 2 : 1, sms), batch);
    size_t sh = p->n_zz * (sizeof(double) + sizeof(int2));
    for (int k = 0; k < n_steps; ++k) {
        k_phase<<<gfull, kThreads, sh, st>>>(d_states, p->n, p->n_zz, p->pairs_dev.as<int2>(),
                                             d_rows + (size_t)k * p->row_len);
        for (int q = 0; q < p->n; ++q)
            k_rx<<<ghalf, kThreads, 0, st>>>(d_states, p->n, p->bitpos[q], d_trig + (size_t)k * p->n + q);
        p->ctx->launches += 1 + p->n;
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_evolve_exact(dq_ising* p, c128* d_states, int batch, const double* d_rows, const double* h_rows, int n_steps,
                     double uniform_bound) {
    const int sms = p->ctx->prop.multiProcessorCount;
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim();
    DQ_TRY(p->exact_diag.reserve(N * sizeof(double)));
    DQ_TRY(p->exact_t0.reserve(N * batch * sizeof(c128)));
    DQ_TRY(p->exact_t1.reserve(N * batch * sizeof(c128)));
    dim3 grid(grid_for(N, sms), batch);
    const size_t sh = p->n_zz * (sizeof(double) + sizeof(int2));
    const int m = 18;                                           // theta <= 1: truncation < 1e-17
    DQ_REQUIRE(h_rows || uniform_bound >= 0.0, "exact step: neither host rows nor a norm bound");
    for (int k = 0; k < n_steps; ++k) {
        // h_rows == NULL (device-resident training): the rows exist on the device only; the caller's bound on every
        // ||dt H(t_k)|| picks the number of squarings (any s with bound / 2^s <= 1 gives the same propagator)
        double bound = h_rows ? 0.0 : uniform_bound;
        XTerms xt;
        xt.n = p->n;
        xt.rows = h_rows ? nullptr : d_rows + (size_t)k * p->row_len + 1 + p->n_zz;
        for (int q = 0; q < p->n; ++q) xt.bit[q] = p->bitpos[q];
        if (h_rows) {
            const double* row = h_rows + (size_t)k * p->row_len;
            bound = fabs(row[0]);
            for (int e = 0; e < p->n_zz; ++e) bound += fabs(row[1 + e]);
            for (int q = 0; q < p->n; ++q) {
                xt.ang[q] = row[1 + p->n_zz + q];
                bound += fabs(xt.ang[q]);
            }
        }
        int s = 0;
        while (ldexp(bound, -s) > 1.0) ++s;
        DQ_REQUIRE(s <= 24, "exact step: ||dt H|| = %g is too large", bound);
        k_build_diag<<<grid_for(N, sms), kThreads, sh, st>>>(p->exact_diag.as<double>(), p->n, p->n_zz, p->pairs_dev.as<int2>(),
                                                            d_rows + (size_t)k * p->row_len);
        p->ctx->launches++;
        for (long long rep = 0; rep < (1LL << s); ++rep) {
            c128* bufs[2] = {p->exact_t0.as<c128>(), p->exact_t1.as<c128>()};
            // term 0 = a copy of the state (the kernel reads neighbours of `tin` while it updates the state in place)
            DQ_CUDA(cudaMemcpyAsync(bufs[0], d_states, N * batch * sizeof(c128), cudaMemcpyDeviceToDevice, st));
            const c128* tin = bufs[0];
            for (int j = 1; j <= m; ++j) {
                c128* tout = bufs[j & 1];
                k_taylor_term<<<grid, kThreads, 0, st>>>(tin, tout, d_states, p->exact_diag.as<double>(), p->n, xt,
                                                        ldexp(1.0, -s) / j);
                p->ctx->launches++;
                tin = tout;
            }
        }
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_fanout(dq_ising* p, const c128* d_phi, c128* d_kets, int n_kets, const ShiftDesc* d_desc,
               double r) {
    const double alpha = atan(r);
    dim3 grid(grid_for(p->dim(), p->ctx->prop.multiProcessorCount), n_kets);
    k_fanout<<<grid, kThreads, 0, p->ctx->stream>>>(d_phi, d_kets, p->n, d_desc, cos(alpha), sin(alpha));
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_energy(dq_ising* p, const c128* d_states, int batch, double* d_out) {
    int gx = grid_for(p->dim(), p->ctx->prop.multiProcessorCount);
    if (gx > 256) gx = 256;
    DQ_TRY(p->scratch.reserve((size_t)gx * batch * sizeof(double)));
    dim3 grid(gx, batch);
    k_energy<<<grid, kThreads, 0, p->ctx->stream>>>(d_states, p->mdiag.as<double>(), p->n,
                                                    p->scratch.as<double>());
    k_sum_partials<<<batch, 32, 0, p->ctx->stream>>>(p->scratch.as<double>(), gx, d_out);
    p->ctx->launches += 2;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_pair_expect(dq_ising* p, const c128* d_states, int batch, double* d_out) {
    if (p->n_zz == 0) return DQ_OK;
    int gx = grid_for(p->dim(), p->ctx->prop.multiProcessorCount);
    if (gx > 128) gx = 128;
    DQ_REQUIRE(batch <= 65535, "pair expectations: batch=%d exceeds 65535 states per call", batch);
    DQ_TRY(p->scratch.reserve((size_t)gx * batch * p->n_zz * sizeof(double)));
    dim3 grid(gx, batch);
    for (int e0 = 0; e0 < p->n_zz; e0 += kPairChunk) {
        k_pair_expect<<<grid, kThreads, 0, p->ctx->stream>>>(d_states, p->n, p->n_zz, p->pairs_dev.as<int2>(), e0, p->scratch.as<double>());
        p->ctx->launches++;
    }
    dim3 g2((p->n_zz + 63) / 64, batch);
    k_sum_pair_partials<<<g2, 64, 0, p->ctx->stream>>>(p->scratch.as<double>(), gx, p->n_zz, d_out);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int gen_build_mdiag(dq_ising* p, const double* m_zz, double m_const) {
    DQ_TRY(p->scratch.reserve((size_t)(p->n_zz + 1) * sizeof(double)));
    if (p->n_zz)
        DQ_CUDA(cudaMemcpyAsync(p->scratch.p, m_zz, p->n_zz * sizeof(double), cudaMemcpyHostToDevice,
                                p->ctx->stream));
    k_build_mdiag<<<grid_for(p->dim(), p->ctx->prop.multiProcessorCount), kThreads, 0, p->ctx->stream>>>(
        p->mdiag.as<double>(), p->n, p->n_zz, p->pairs_dev.as<int2>(), p->scratch.as<double>(), m_const);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return DQ_OK;
}

}  // namespace dq
