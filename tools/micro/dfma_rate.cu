// Micro-benchmark (GPU box): issue rate of the lifting butterfly run (320 in-place DFMA per 32 complex registers) for ONE warp
// per SM sub-partition and for two, in the pair-by-pair order and in the "all a-updates, then all b-updates" order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_rate tools/micro/dfma_rate.cu && /tmp/dfma_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 c128;
constexpr int kRegs = 32;
template <int B, int ORDER>
__device__ __forceinline__ void lift_bit(c128 (&v)[kRegs], const double2 rc) {
    if (ORDER == 0) {
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            if (j & (1 << B)) continue;
            const int k = j | (1 << B);
            v[j].x = fma(rc.x, v[k].y, v[j].x);
            v[j].y = fma(-rc.x, v[k].x, v[j].y);
            v[k].x = fma(rc.y, v[j].y, v[k].x);
            v[k].y = fma(-rc.y, v[j].x, v[k].y);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            if (j & (1 << B)) continue;
            const int k = j | (1 << B);
            v[j].x = fma(rc.x, v[k].y, v[j].x);
            v[j].y = fma(-rc.x, v[k].x, v[j].y);
        }
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            if (j & (1 << B)) continue;
            const int k = j | (1 << B);
            v[k].x = fma(rc.y, v[j].y, v[k].x);
            v[k].y = fma(-rc.y, v[j].x, v[k].y);
        }
    }
}
template <int ORDER>
__global__ void __launch_bounds__(256, 1) k(const double2* __restrict__ rc, c128* out, int iters, long long* cyc) {
    c128 v[kRegs];
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = make_double2(threadIdx.x + j, 1.0 / (j + 1));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const double2* r = rc + 5 * (it & 3);
        lift_bit<0, ORDER>(v, r[0]); lift_bit<1, ORDER>(v, r[1]); lift_bit<2, ORDER>(v, r[2]); lift_bit<3, ORDER>(v, r[3]); lift_bit<4, ORDER>(v, r[4]);
    }
    const long long t1 = clock64();
    c128 s = make_double2(0, 0);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) { s.x += v[j].x; s.y += v[j].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double2 h[20];
    for (int i = 0; i < 20; ++i) h[i] = make_double2(1e-3 * (i + 1), 9e-4 * (i + 1));
    double2* rc; c128* out; long long* cyc;
    cudaMalloc(&rc, sizeof(h)); cudaMemcpy(rc, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 148 * 256 * sizeof(c128)); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int order = 0; order < 2; ++order)
        for (int threads : {128, 256}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (order == 0) k<0><<<148, threads>>>(rc, out, iters, cyc); else k<1><<<148, threads>>>(rc, out, iters, cyc);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("order %d  warps/SMSP %d: %.2f cycles per DFMA per warp  (pipe: %.2f cycles per DFMA)\n", order, threads / 128,
                   (double)c / (iters * 320.0), (double)c / (iters * 320.0) / (threads / 128));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
