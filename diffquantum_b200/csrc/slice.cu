// Slice kernels for a single state distributed over ranks on its high-order index bits (SURVEY 8e,
// BASELINE configs[4]).  A rank owns 2^L consecutive amplitudes; the global basis index of local x is
// (high_bits << L) | x.  The diagonal phase needs no communication in ANY layout — the host passes the
// current physical bit position of every ZZ endpoint — and an X rotation is local whenever its qubit
// currently sits on a bit below L.  Qubits on the top g bits are made local by an all-to-all that swaps
// bits [L-g, L) with the rank bits (done by the host through NCCL, diffquantum_b200/distributed.py).
// Step semantics: the per-term product of diffqc.cc:155-164.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPairs = 256;

struct PhaseArgs {
    int n_zz;
    unsigned char a[kMaxPairs], b[kMaxPairs];
    double ang[kMaxPairs];
    double c0;
};

__global__ void __launch_bounds__(kThreads) k_slice_phase(double2* __restrict__ psi, int L, unsigned long long high,
                                                          const PhaseArgs* __restrict__ pa) {
    __shared__ double ang[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { ang[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    __syncthreads();
    const double c0 = pa->c0;
    const size_t N = (size_t)1 << L;
    const unsigned long long hi = high << L;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long g = hi | x;
        double a = c0;
        for (int e = 0; e < n_zz; ++e) a += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -ang[e] : ang[e];
        double sn, cs;
        sincos(a, &sn, &cs);
        const double2 v = psi[x];
        psi[x] = make_double2(v.x * cs + v.y * sn, v.y * cs - v.x * sn);
    }
}

__global__ void __launch_bounds__(kThreads) k_slice_rx(double2* __restrict__ psi, int L, int bit, double c, double s) {
    const size_t half = (size_t)1 << (L - 1);
    const size_t low = ((size_t)1 << bit) - 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        const size_t x0 = ((i & ~low) << 1) | (i & low), x1 = x0 | ((size_t)1 << bit);
        const double2 a = psi[x0], b = psi[x1];
        psi[x0] = make_double2(c * a.x + s * b.y, c * a.y - s * b.x);
        psi[x1] = make_double2(c * b.x + s * a.y, c * b.y - s * a.x);
    }
}

__global__ void __launch_bounds__(kThreads) k_slice_fill(double2* __restrict__ psi, size_t N, double amp) {
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x)
        psi[x] = make_double2(amp, 0.0);
}

// partial <psi|M|psi>, M = c0 + sum_e w_e z_a z_b evaluated on the fly (a 2^n table would not fit at n = 32)
__global__ void __launch_bounds__(kThreads) k_slice_energy(const double2* __restrict__ psi, int L, unsigned long long high,
                                                           const PhaseArgs* __restrict__ pa, double* __restrict__ partial) {
    __shared__ double w[kMaxPairs];
    __shared__ unsigned char ba[kMaxPairs], bb[kMaxPairs];
    __shared__ double red[kThreads / 32];
    const int n_zz = pa->n_zz;
    for (int e = threadIdx.x; e < n_zz; e += blockDim.x) { w[e] = pa->ang[e]; ba[e] = pa->a[e]; bb[e] = pa->b[e]; }
    __syncthreads();
    const double c0 = pa->c0;
    const size_t N = (size_t)1 << L;
    const unsigned long long hi = high << L;
    double acc = 0.0;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long g = hi | x;
        double m = c0;
        for (int e = 0; e < n_zz; ++e) m += (((g >> ba[e]) ^ (g >> bb[e])) & 1ull) ? -w[e] : w[e];
        const double2 v = psi[x];
        acc = fma(m, v.x * v.x + v.y * v.y, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

int grid_for(const dq_context* ctx, size_t work) {
    const size_t b = (work + kThreads - 1) / kThreads, cap = (size_t)ctx->prop.multiProcessorCount * 16;
    return (int)std::max<size_t>(1, std::min(b, cap));
}

int fill_args(PhaseArgs& h, int n_total, int n_zz, const int32_t* pair_bits, const double* vals, double c0) {
    DQ_REQUIRE(n_zz >= 0 && n_zz <= kMaxPairs, "slice: n_zz=%d outside [0,%d]", n_zz, kMaxPairs);
    DQ_REQUIRE(n_zz == 0 || (pair_bits && vals), "slice: NULL pair table");
    h.n_zz = n_zz;
    h.c0 = c0;
    for (int e = 0; e < n_zz; ++e) {
        const int a = pair_bits[2 * e], b = pair_bits[2 * e + 1];
        DQ_REQUIRE(a >= 0 && b >= 0 && a < n_total && b < n_total && a != b, "slice: pair %d = bits (%d,%d) invalid for %d qubits", e, a, b, n_total);
        DQ_REQUIRE(isfinite(vals[e]), "slice: non-finite value for pair %d", e);
        h.a[e] = (unsigned char)a;
        h.b[e] = (unsigned char)b;
        h.ang[e] = vals[e];
    }
    DQ_REQUIRE(isfinite(c0), "slice: non-finite constant");
    return DQ_OK;
}

int upload_args(dq_context* ctx, const PhaseArgs& h, PhaseArgs** d_out) {
    // a small ring so that consecutive asynchronous launches never share a table
    if (!ctx->slice_ring) DQ_CUDA(cudaMalloc(&ctx->slice_ring, sizeof(PhaseArgs) * 64));
    PhaseArgs* d = reinterpret_cast<PhaseArgs*>(ctx->slice_ring) + (ctx->slice_cursor++ & 63);
    if ((ctx->slice_cursor & 63) == 0) DQ_CUDA(cudaStreamSynchronize(ctx->stream));      // ring wrapped: drain
    DQ_CUDA(cudaMemcpyAsync(d, &h, sizeof(PhaseArgs), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = d;
    return DQ_OK;
}

}  // namespace

extern "C" {

int dq_slice_fill_uniform(dq_context* ctx, void* psi_dev, int L, int n_total) {
    DQ_REQUIRE(ctx && psi_dev, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_fill_uniform: L=%d n=%d", L, n_total);
    DQ_TRY(ctx->set_device());
    const size_t N = (size_t)1 << L;
    k_slice_fill<<<grid_for(ctx, N), kThreads, 0, ctx->stream>>>((double2*)psi_dev, N, exp2(-0.5 * n_total));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int dq_slice_phase(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                   const int32_t* pair_bits, const double* angles) {
    DQ_REQUIRE(ctx && psi_dev && angles, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_phase: L=%d n=%d", L, n_total);
    DQ_REQUIRE((high_bits >> (n_total - L)) == 0, "dq_slice_phase: high_bits do not fit %d global bits", n_total - L);
    DQ_TRY(ctx->set_device());
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, angles + 1, angles[0]));
    PhaseArgs* d;
    DQ_TRY(upload_args(ctx, h, &d));
    k_slice_phase<<<grid_for(ctx, (size_t)1 << L), kThreads, 0, ctx->stream>>>((double2*)psi_dev, L, high_bits, d);
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int dq_slice_rx(dq_context* ctx, void* psi_dev, int L, int bit, double theta) {
    DQ_REQUIRE(ctx && psi_dev, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && bit >= 0 && bit < L, "dq_slice_rx: bit %d is not local (L=%d)", bit, L);
    DQ_REQUIRE(isfinite(theta), "dq_slice_rx: non-finite angle");
    DQ_TRY(ctx->set_device());
    k_slice_rx<<<grid_for(ctx, (size_t)1 << (L - 1)), kThreads, 0, ctx->stream>>>((double2*)psi_dev, L, bit, cos(theta), sin(theta));
    ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int dq_slice_energy(dq_context* ctx, const void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                    const int32_t* pair_bits, const double* m_zz, double m_const, double* partial_out) {
    DQ_REQUIRE(ctx && psi_dev && partial_out, "NULL argument");
    DQ_REQUIRE(L >= 1 && L <= 33 && n_total >= L && n_total <= 40, "dq_slice_energy: L=%d n=%d", L, n_total);
    DQ_TRY(ctx->set_device());
    PhaseArgs h;
    DQ_TRY(fill_args(h, n_total, n_zz, pair_bits, m_zz, m_const));
    PhaseArgs* d;
    DQ_TRY(upload_args(ctx, h, &d));
    const int blocks = std::min(grid_for(ctx, (size_t)1 << L), 1024);
    double* d_part = nullptr;
    DQ_CUDA(cudaMalloc(&d_part, blocks * sizeof(double)));
    k_slice_energy<<<blocks, kThreads, 0, ctx->stream>>>((const double2*)psi_dev, L, high_bits, d, d_part);
    ctx->launches++;
    std::vector<double> part(blocks);
    cudaError_t e = cudaMemcpyAsync(part.data(), d_part, blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_part);
    DQ_CUDA(e);
    double acc = 0.0;
    for (double v : part) acc += v;            // fixed order: deterministic
    *partial_out = acc;
    return DQ_OK;
}

}  // extern "C"
