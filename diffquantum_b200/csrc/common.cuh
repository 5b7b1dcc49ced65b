// Shared helpers for the diffqc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/diffqc_b200.h"

namespace dq {

void set_error(const char* fmt, ...);

#define DQ_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            dq::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DQ_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define DQ_REQUIRE(cond, ...)                                                             \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            dq::set_error(__VA_ARGS__);                                                   \
            return DQ_ERR_INVALID;                                                        \
        }                                                                                 \
    } while (0)

#define DQ_TRY(expr)                                                                      \
    do {                                                                                  \
        int _s = (expr);                                                                  \
        if (_s != DQ_OK) return _s;                                                       \
    } while (0)

typedef double2 c128;

__host__ __device__ __forceinline__ c128 cmul(c128 a, c128 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ c128 cconj(c128 a) { return make_double2(a.x, -a.y); }

// growable device buffer owned by a context/problem
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return DQ_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return DQ_ERR_NOMEM;
        }
        cap = bytes;
        return DQ_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace dq

namespace dq { namespace dense { struct State; } }

struct dq_context {
    dq::dense::State* dense = nullptr;     // dense-path state (diffqc.set_H globals + workspaces), lazily created
    int dense_force_strategy = -1;         // tests: -1 auto, 0 block-Taylor, 1 per-step propagator, 2 chained
    void* slice_ring = nullptr;            // slice.cu: ring of small device argument tables
    int dense_small_mma = 1;               // dense resident engine: shifted kets of a sample on the FP64 tensor cores (k_small_mma)
    void* slice_partials = nullptr;        // slice.cu: per-block energy partials (1024 doubles)
    unsigned slice_cursor = 0;
    bool slice_tma_attr = false;           // slice.cu: dynamic shared-memory attribute of the TMA tile kernels set on this device
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop;
    uint64_t launches = 0;
    int set_device() const;
};
