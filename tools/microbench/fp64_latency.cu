// Dependent-issue latency of the FP64 instructions on the fit's critical path (DFMA, DMUL, MUFU.RSQ64H), a shared-memory
// round trip and a 256-thread barrier, in SM clocks.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/fp64_latency tools/microbench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(double* out, long long* cyc, double a, double b) {
    __shared__ double sh[64];
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    double y = x;
#pragma unroll
    for (int i = 0; i < 256; ++i) y = y * b;
    long long t2 = clock64();
    double z = y + 2.0;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
        z = r + 1.5;
    }
    long long t3 = clock64();
    double w = z;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
        sh[threadIdx.x & 63] = w;
        __syncthreads();
        w = sh[(threadIdx.x + 1) & 63] + 1.0;
        __syncthreads();
    }
    long long t4 = clock64();
    double v = w;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
        volatile double* p = sh;
        p[threadIdx.x & 63] = v;
        v = p[threadIdx.x & 63] + 1.0;
    }
    long long t5 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    }
}

int main() {
    double* d; long long* c; long long h[5];
    cudaMalloc(&d, 1024 * 8); cudaMalloc(&c, 5 * 8);
    for (int threads : {32, 256}) {
        chain<<<1, threads>>>(d, c, 1.0, 0.999);
        chain<<<1, threads>>>(d, c, 1.0, 0.999);
        cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%3d threads: DFMA %.1f clk  DMUL %.1f clk  rsqrt.approx.f64 + DADD %.1f clk  (store, barrier, load, add, barrier) %.1f clk  "
               "(store, load, add) %.1f clk  per dependent step\n", threads, h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0);
    }
    return 0;
}
