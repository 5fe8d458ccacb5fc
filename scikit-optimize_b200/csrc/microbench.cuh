// Roofline denominators measured on the device the bench runs on (skb_measure_peaks): MEASURED_PEAKS.json carries HBM and
// bf16 numbers only, the two pipes this path is bound by are the tcgen05 kind::i8 tensor pipe and the FP64 (DMMA) pipe.
//   umma_i8_rate_kernel : one CTA per SM issues back-to-back 128 x 256 x 32 s8 MMAs on operands that stay in shared
//                         memory (no data movement): MACs per clock per SM from clock64, TOP/s from CUDA events.
//   dmma_rate_kernel    : 8 warps x 2 CTAs per SM of independent mma.sync.m8n8k4.f64 chains: TFLOP/s from CUDA events.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

namespace skb {

constexpr int MB_M = 128, MB_N = 256, MB_KC = 32, MB_CHUNKS = 4;
constexpr size_t MB_SMEM = (size_t)MB_CHUNKS * (MB_M + MB_N) * MB_KC + 1024;

__global__ void __launch_bounds__(128, 1) umma_i8_rate_kernel(int repeat, long long* cycles) {
    extern __shared__ __align__(128) unsigned char mb_smem[];
    unsigned char* sA = mb_smem;
    unsigned char* sB = sA + (size_t)MB_CHUNKS * MB_M * MB_KC;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < MB_CHUNKS * (MB_M + MB_N) * MB_KC / 4; i += 128)
        reinterpret_cast<uint32_t*>(mb_smem)[i] = 0x01010101u * (i & 1);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MB_N >> 3) << 17) | ((uint32_t)(MB_M >> 4) << 24);
    const long long t0 = clock64();
    if (tid == 0) {
        for (int rep = 0; rep < repeat; ++rep)
            for (int kc = 0; kc < MB_CHUNKS; ++kc) {
                const uint32_t a = smem_u32(sA + (size_t)kc * MB_M * MB_KC), b = smem_u32(sB + (size_t)kc * MB_N * MB_KC);
                const uint64_t da = (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(((MB_M / 8) * 128 >> 4) & 0x3FFF) << 16) |
                                    ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
                const uint64_t db = (uint64_t)((b >> 4) & 0x3FFF) | ((uint64_t)(((MB_N / 8) * 128 >> 4) & 0x3FFF) << 16) |
                                    ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

constexpr int MB_DMMA_ITERS = 4096;
__global__ void __launch_bounds__(256, 2) dmma_rate_kernel(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < MB_DMMA_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace skb
