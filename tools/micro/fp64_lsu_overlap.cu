// Micro-benchmark (GPU box): do the FP64 pipe and the shared-memory (LSU) pipe of an SM overlap when different warps of the same
// sub-partition use them?  One CTA per SM.  Roles per warp: M = the lifting butterfly run of k_fused_ws (320 in-place DFMA per 32
// complex registers), X = its register <-> shared-memory exchange (32 STS.128 + 32 LDS.128 per thread, conflict-free).
//   case 0: 4 warps M (one per sub-partition)          case 1: 4 warps X
//   case 2: 4 warps M + 4 warps X (one of each per sub-partition)
//   case 3: 8 warps M                                  case 4: 8 warps X
//   case 5: 8 warps, each alternating M and X (the kernel's own pattern, two independent warps per sub-partition)
// Printed: cycles per iteration of each role.  Overlap in hardware <=> case 2 runs each role at (nearly) its case 0 / case 1 rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_lsu tools/micro/fp64_lsu_overlap.cu && /tmp/fp64_lsu
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 c128;
constexpr int kRegs = 32;
template <int B>
__device__ __forceinline__ void lift_bit(c128 (&v)[kRegs], const double2 rc) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
        if (j & (1 << B)) continue;
        const int k = j | (1 << B);
        v[j].x = fma(rc.x, v[k].y, v[j].x);
        v[j].y = fma(-rc.x, v[k].x, v[j].y);
        v[k].x = fma(rc.y, v[j].y, v[k].x);
        v[k].y = fma(-rc.y, v[j].x, v[k].y);
    }
}
__device__ __forceinline__ void math(c128 (&v)[kRegs], const double2* r) {
    lift_bit<0>(v, r[0]); lift_bit<1>(v, r[1]); lift_bit<2>(v, r[2]); lift_bit<3>(v, r[3]); lift_bit<4>(v, r[4]);
}
// a warp's private 1024-amplitude region: store by rows, load by a lane-rotated pattern (both conflict-free: 16-byte units,
// consecutive lanes)
__device__ __forceinline__ void exchange(c128 (&v)[kRegs], c128* region, int lane) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) region[j * 32 + lane] = v[j];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = region[((j + 1) & 31) * 32 + ((lane + j) & 31)];
    __syncwarp();
}
__global__ void __launch_bounds__(256, 1) k(const double2* __restrict__ rc, c128* out, int iters, int mode, long long* cyc) {
    extern __shared__ c128 smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    c128* region = smem + warp * 1024;
    c128 v[kRegs];
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = make_double2(threadIdx.x + j, 1.0 / (j + 1));
    // role: 0 = M, 1 = X, 2 = alternate, 3 = idle
    int role;
    switch (mode) {
        case 0: role = warp < 4 ? 0 : 3; break;
        case 1: role = warp < 4 ? 1 : 3; break;
        case 2: role = warp < 4 ? 0 : 1; break;
        case 3: role = 0; break;
        case 4: role = 1; break;
        default: role = 2; break;
    }
    __syncthreads();
    const long long t0 = clock64();
    if (role == 0) {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) math(v, rc + 5 * (it & 3));
    } else if (role == 1) {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) exchange(v, region, lane);
    } else if (role == 2) {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            if (it == 0 && (warp & 4)) exchange(v, region, lane);      // the two warps of a sub-partition start out of phase
            math(v, rc + 5 * (it & 3));
            exchange(v, region, lane);
        }
    }
    const long long t1 = clock64();
    c128 s = make_double2(0, 0);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) { s.x += v[j].x; s.y += v[j].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}
int main() {
    double2 h[20];
    for (int i = 0; i < 20; ++i) h[i] = make_double2(1e-3 * (i + 1), 9e-4 * (i + 1));
    double2* rc; c128* out; long long* cyc;
    cudaMalloc(&rc, sizeof(h)); cudaMemcpy(rc, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 148 * 256 * sizeof(c128)); cudaMalloc(&cyc, 64);
    const int iters = 2000, smem = 8 * 1024 * sizeof(c128);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"4 warps M", "4 warps X", "4 M + 4 X", "8 warps M", "8 warps X", "8 warps alternating M, X"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 256, smem>>>(rc, out, iters, mode, cyc); cudaDeviceSynchronize(); }
        long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
        printf("case %d  %-26s  warp 0: %8.1f cycles/iter   warp 4: %8.1f cycles/iter\n", mode, names[mode], (double)c[0] / iters, (double)c[4] / iters);
    }
    printf("per iteration: M = 320 DFMA per thread (ideal 640 cycles per warp alone on its FP64 pipe), X = 16 KiB written + 16 KiB read per warp "
           "(128 B/clk/SM: 256 cycles per warp alone, 1024 when 4 warps share the pipe, 2048 when 8 do)\n%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
