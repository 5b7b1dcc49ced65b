// Context, error reporting and calibration micro-benchmarks.
#include <stdarg.h>
#include <memory>
#include "common.cuh"
#include "dense.cuh"

namespace dq {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dq

int dq_context::set_device() const {
    DQ_CUDA(cudaSetDevice(device));
    return DQ_OK;
}

extern "C" {

const char* dq_version(void) { return "dev"; }
const char* dq_last_error(void) { return dq::g_err; }

int dq_device_count(int* count) {
    DQ_REQUIRE(count != nullptr, "dq_device_count: NULL output");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        dq::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return DQ_ERR_CUDA;
    }
    return DQ_OK;
}

int dq_context_create(int device, dq_context** out) {
    DQ_REQUIRE(out != nullptr, "dq_context_create: NULL output");
    *out = nullptr;
    int count = 0;
    DQ_TRY(dq_device_count(&count));
    DQ_REQUIRE(device >= 0 && device < count, "dq_context_create: device %d of %d", device, count);
    std::unique_ptr<dq_context> owner(new dq_context());      // freed on every early return below
    dq_context* c = owner.get();
    c->device = device;
    DQ_CUDA(cudaSetDevice(device));
    DQ_CUDA(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major != 10) {
        dq::set_error("dq_context_create: device %d is sm_%d%d; this library is built for sm_100a only",
                      device, c->prop.major, c->prop.minor);
        return DQ_ERR_UNSUPPORTED;
    }
    DQ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = owner.release();
    return DQ_OK;
}

int dq_context_destroy(dq_context* ctx) {
    if (!ctx) return DQ_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dq::dense::release(ctx);
    if (ctx->slice_ring) cudaFree(ctx->slice_ring);
    if (ctx->slice_partials) cudaFree(ctx->slice_partials);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DQ_OK;
}

int dq_context_synchronize(dq_context* ctx) {
    DQ_REQUIRE(ctx != nullptr, "NULL context");
    DQ_TRY(ctx->set_device());
    DQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return DQ_OK;
}

int dq_context_stream(dq_context* ctx, uint64_t* stream_out) {
    DQ_REQUIRE(ctx != nullptr && stream_out != nullptr, "NULL argument");
    *stream_out = (uint64_t)(uintptr_t)ctx->stream;
    return DQ_OK;
}

int dq_context_launch_count(dq_context* ctx, uint64_t* count_out) {
    DQ_REQUIRE(ctx != nullptr && count_out != nullptr, "NULL argument");
    *count_out = ctx->launches;
    return DQ_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// micro-benchmarks (calibration only; not on the product path)
// ------------------------------------------------------------------------------------------
namespace {

__global__ void mb_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __ldcg(src + i);
}

__global__ void mb_rw_inplace(double2* __restrict__ buf, size_t n, int reps) {
    // read-modify-write the same window repeatedly: L2-resident when the window fits
    for (int r = 0; r < reps; ++r) {
        size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
        size_t stride = (size_t)gridDim.x * blockDim.x;
        for (; i < n; i += stride) {
            double2 v = __ldcg(buf + i);
            v.x += 1.0;
            __stcg(buf + i, v);
        }
    }
}

__global__ void mb_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void mb_smem(double* out, int iters) {
    extern __shared__ double2 sm[];
    int t = threadIdx.x;
    double2 acc = make_double2(0, 0);
    for (int i = t; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            double2 v = sm[(t + k * 128) & 4095];
            acc.x += v.x;
            acc.y += v.y;
        }
    }
    out[blockIdx.x * blockDim.x + t] = acc.x + acc.y;
}

// released on every exit path of dq_microbench (the DQ_CUDA / DQ_REQUIRE macros return early)
struct ScopedDev { void* p = nullptr; ~ScopedDev() { if (p) cudaFree(p); } };
struct ScopedEvent { cudaEvent_t e = nullptr; ~ScopedEvent() { if (e) cudaEventDestroy(e); } };

}  // namespace

extern "C" int dq_microbench(dq_context* ctx, int kind, int64_t bytes, int iters, double* result) {
    DQ_REQUIRE(ctx && result, "NULL argument");
    DQ_REQUIRE(iters > 0, "iters must be positive");
    DQ_TRY(ctx->set_device());
    ScopedEvent g0, g1;
    ScopedDev ga, gb;
    DQ_CUDA(cudaEventCreate(&g0.e));
    DQ_CUDA(cudaEventCreate(&g1.e));
    cudaEvent_t e0 = g0.e, e1 = g1.e;
    float ms = 0;
    const int sms = ctx->prop.multiProcessorCount;
    if (kind == 0 || kind == 2) {
        DQ_REQUIRE(bytes >= 1 << 20, "bytes too small");
        size_t n = (size_t)bytes / sizeof(double2);
        double2 *a = nullptr, *b = nullptr;
        DQ_CUDA(cudaMalloc(&a, n * sizeof(double2)));
        ga.p = a;
        DQ_CUDA(cudaMemsetAsync(a, 0, n * sizeof(double2), ctx->stream));
        if (kind == 0) {
            DQ_CUDA(cudaMalloc(&b, n * sizeof(double2)));
            gb.p = b;
            mb_copy<<<sms * 8, 256, 0, ctx->stream>>>(a, b, n);
            DQ_CUDA(cudaEventRecord(e0, ctx->stream));
            for (int i = 0; i < iters; ++i) mb_copy<<<sms * 8, 256, 0, ctx->stream>>>(a, b, n);
            DQ_CUDA(cudaEventRecord(e1, ctx->stream));
        } else {
            mb_rw_inplace<<<sms * 8, 256, 0, ctx->stream>>>(a, n, 2);
            DQ_CUDA(cudaEventRecord(e0, ctx->stream));
            mb_rw_inplace<<<sms * 8, 256, 0, ctx->stream>>>(a, n, iters);
            DQ_CUDA(cudaEventRecord(e1, ctx->stream));
        }
        DQ_CUDA(cudaEventSynchronize(e1));
        DQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *result = 2.0 * n * sizeof(double2) * iters / (ms * 1e-3) / 1e9;
    } else if (kind == 1 || kind == 3) {
        double* out = nullptr;
        const int blocks = sms * 8, threads = 256;
        DQ_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
        ga.p = out;
        if (kind == 1) {
            mb_dfma<<<blocks, threads, 0, ctx->stream>>>(out, 16);
            DQ_CUDA(cudaEventRecord(e0, ctx->stream));
            mb_dfma<<<blocks, threads, 0, ctx->stream>>>(out, iters);
            DQ_CUDA(cudaEventRecord(e1, ctx->stream));
            DQ_CUDA(cudaEventSynchronize(e1));
            DQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *result = 2.0 * 64.0 * iters * blocks * threads / (ms * 1e-3) / 1e12;   // TFLOP/s
        } else {
            DQ_CUDA(cudaFuncSetAttribute(mb_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            mb_smem<<<sms * 2, 128, 65536, ctx->stream>>>(out, 4);
            DQ_CUDA(cudaEventRecord(e0, ctx->stream));
            mb_smem<<<sms * 2, 128, 65536, ctx->stream>>>(out, iters);
            DQ_CUDA(cudaEventRecord(e1, ctx->stream));
            DQ_CUDA(cudaEventSynchronize(e1));
            DQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *result = 16.0 * 32.0 * iters * (sms * 2) * 128 / (ms * 1e-3) / 1e9;   // GB/s smem read
        }
    } else {
        dq::set_error("dq_microbench: unknown kind %d", kind);
        return DQ_ERR_INVALID;
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}
