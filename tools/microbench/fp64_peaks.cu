// Microbenchmark: FP64 issue rates on sm_100a (DFMA, DMMA shapes, exp/sqrt, mixed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 2048;

__global__ void k_dfma(double* out, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>  // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void k_dmma(double* out, double a, double b) {
    constexpr int NACC = 8;
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x; }
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av[0]), "d"(bv[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(av[0]), "d"(av[1]), "d"(bv[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(bv[0]), "d"(bv[1]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]),
                               "d"(bv[0]), "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the warps DFMA, half DMMA m8n8k4: do the pipes overlap?
__global__ void k_mixed(double* out, double a, double b) {
    int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) s += acc[i];
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = threadIdx.x; }
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, double a) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = -1e-3 * (threadIdx.x + i);
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = exp(acc[i]) * a - 1.0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}
__global__ void k_sqrt(double* out, double a) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = sqrt(acc[i]) + a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <typename F>
float time_it(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    for (int wpb : {4, 8, 16, 32}) {
        for (int bps : {1, 2}) {
            if (wpb * bps > 64) continue;
            int threads = wpb * 32, blocks = sms * bps;
            double nthreads = (double)threads * blocks;
            float ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("DFMA       warps/blk=%2d blk/sm=%d : %.3f ms  %.2f TFLOP/s\n", wpb, bps, ms, nthreads * ITERS * 16 * 2 / ms / 1e9);
            double nwarps = nthreads / 32;
            ms = time_it([&] { k_dmma<0><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("DMMA 8x8x4   warps/blk=%2d blk/sm=%d : %.3f ms  %.2f TFLOP/s\n", wpb, bps, ms, nwarps * ITERS * 8 * (8 * 8 * 4 * 2.0) / ms / 1e9);
            ms = time_it([&] { k_dmma<1><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("DMMA 16x8x4  warps/blk=%2d blk/sm=%d : %.3f ms  %.2f TFLOP/s\n", wpb, bps, ms, nwarps * ITERS * 8 * (16 * 8 * 4 * 2.0) / ms / 1e9);
            ms = time_it([&] { k_dmma<2><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("DMMA 16x8x8  warps/blk=%2d blk/sm=%d : %.3f ms  %.2f TFLOP/s\n", wpb, bps, ms, nwarps * ITERS * 8 * (16 * 8 * 8 * 2.0) / ms / 1e9);
            ms = time_it([&] { k_dmma<3><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("DMMA 16x8x16 warps/blk=%2d blk/sm=%d : %.3f ms  %.2f TFLOP/s\n", wpb, bps, ms, nwarps * ITERS * 8 * (16 * 8 * 16 * 2.0) / ms / 1e9);
            ms = time_it([&] { k_mixed<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
            printf("MIXED        warps/blk=%2d blk/sm=%d : %.3f ms  dfma-part %.2f + dmma-part %.2f TFLOP/s\n", wpb, bps, ms,
                   nthreads / 2 * ITERS * 16 * 2 / ms / 1e9, nwarps / 2 * ITERS * 8 * 512.0 / ms / 1e9);
        }
    }
    {
        int threads = 256, blocks = sms * 8;
        double nthreads = (double)threads * blocks;
        float ms = time_it([&] { k_exp<<<blocks, threads>>>(out, 0.5); });
        printf("DP exp : %.3f ms  %.2f Gexp/s  (%.1f SM-cycles per warp-exp at %d kHz)\n", ms, nthreads * (ITERS / 8) * 4 / ms / 1e6,
               (double)ms * 1e-3 * p.clockRate * 1e3 * sms / (nthreads / 32 * (ITERS / 8) * 4), p.clockRate);
        ms = time_it([&] { k_sqrt<<<blocks, threads>>>(out, 0.5); });
        printf("DP sqrt: %.3f ms  %.2f Gsqrt/s\n", ms, nthreads * (ITERS / 8) * 4 / ms / 1e6);
    }
    return 0;
}
