// What does one tcgen05.mma kind::i8 (M = 128, K = 32) cost as a function of N and of the accumulator it targets?
// K2o (csrc/ozaki.cuh) issues, per 32 training points, one instruction per digit plane of L_inv whose N shrinks from 384
// to 64 columns; this probe times candidate instruction schedules (compile-time lists of (N, accumulator column, A plane,
// first B row)) issued back to back by one thread on operands resident in shared memory, one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_shapes umma_i8_shapes.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int M = 128, KC = 32, MAXI = 10;
struct Sched { int n; int N[MAXI]; int col[MAXI]; int aplane[MAXI]; int brow[MAXI]; };

__host__ __device__ constexpr Sched sched(int id) {
    switch (id) {
        case 0: return {1, {16}, {0}, {0}, {0}};
        case 1: return {1, {64}, {0}, {0}, {0}};
        case 2: return {1, {96}, {0}, {0}, {0}};
        case 3: return {1, {128}, {0}, {0}, {0}};
        case 4: return {1, {160}, {0}, {0}, {0}};
        case 5: return {1, {176}, {0}, {0}, {0}};
        case 6: return {1, {192}, {0}, {0}, {0}};
        case 7: return {1, {224}, {0}, {0}, {0}};
        case 8: return {1, {256}, {0}, {0}, {0}};
        case 9: return {4, {64, 64, 64, 64}, {0, 64, 128, 192}, {0, 1, 2, 3}, {0, 0, 0, 0}};
        case 10: return {4, {128, 128, 128, 128}, {0, 128, 256, 384}, {0, 1, 2, 3}, {0, 0, 0, 0}};
        case 11: return {8, {256, 128, 160, 160, 256, 192, 128, 64}, {0, 256, 64, 224, 128, 192, 256, 320}, {0, 0, 1, 1, 2, 3, 4, 5}, {0, 256, 0, 160, 0, 0, 0, 0}};
        case 12: return {8, {192, 192, 160, 160, 256, 192, 128, 64}, {0, 192, 64, 224, 128, 192, 256, 320}, {0, 0, 1, 1, 2, 3, 4, 5}, {0, 192, 0, 160, 0, 0, 0, 0}};
        case 13: return {8, {64, 128, 192, 256, 160, 160, 192, 192}, {320, 256, 192, 128, 64, 224, 0, 192}, {5, 4, 3, 2, 1, 1, 0, 0}, {0, 0, 0, 0, 0, 160, 0, 192}};
        case 14: return {8, {256, 128, 64, 160, 160, 128, 256, 192}, {0, 256, 320, 64, 224, 256, 128, 192}, {0, 0, 5, 1, 1, 4, 2, 3}, {0, 256, 0, 0, 160, 0, 0, 0}};
        case 15: return {10, {256, 192, 192, 192, 160, 160, 256, 192, 128, 64}, {0, 256, 64, 256, 128, 288, 192, 256, 320, 384}, {0, 0, 1, 1, 2, 2, 3, 4, 5, 6}, {0, 256, 0, 192, 0, 160, 0, 0, 0, 0}};
        case 16: return {9, {240, 240, 208, 192, 160, 160, 240, 160, 80}, {0, 240, 80, 288, 160, 320, 240, 320, 400}, {0, 0, 1, 1, 2, 2, 3, 4, 5}, {0, 240, 0, 208, 0, 160, 0, 0, 0}};
        // six instructions of N = 224: the same 1344 columns if every instruction could be made equally wide
        case 17: return {6, {224, 224, 224, 224, 224, 224}, {0, 32, 64, 96, 128, 160}, {0, 1, 2, 3, 4, 5}, {0, 0, 0, 0, 0, 0}};
        // 5 x 256 + 64
        case 18: return {6, {256, 256, 256, 256, 256, 64}, {0, 32, 64, 96, 128, 320}, {0, 1, 2, 3, 4, 5}, {0, 0, 0, 0, 0, 0}};
        default: return {1, {256}, {0}, {0}, {0}};
    }
}
const char* names[] = {"N=16 same accumulator", "N=64 same accumulator", "N=96 same accumulator", "N=128 same accumulator", "N=160 same accumulator",
                       "N=176 same accumulator", "N=192 same accumulator", "N=224 same accumulator", "N=256 same accumulator",
                       "N=64 x4 disjoint accumulators", "N=128 x4 disjoint accumulators",
                       "K2o now: 256+128|160+160|256|192|128|64", "K2o alt: 192+192|160+160|256|192|128|64",
                       "K2o small first: 64|128|192|256|160+160|192+192", "K2o interleaved: 256 128|64|160 160|128|256|192",
                       "K2o P=7: 256+192|192+192|160+160|256|192|128|64", "80 cand: 240+240|208+192|160+160|240|160|80",
                       "6 x N=224 (hypothetical even split)", "5 x 256 + 64 (hypothetical)"};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t idesc(int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A: 8 planes of 128 x 32 (4 KB each); B: 512 rows x 32 (16 KB), LBO = 512 / 8 * 128
template <int ID, int A_TMEM, int EXTRA = 0>   // EXTRA bit 0: tcgen05.commit to a scratch mbarrier after every pass, bit 1: tcgen05.fence before every pass
__global__ void __launch_bounds__(128, 1) sched_kernel(int repeat, long long* cycles) {
    constexpr Sched s = sched(ID);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = sA + 8 * M * KC;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (8 * M * KC + 512 * KC) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    if (tid == 0) {
#pragma unroll 1
        for (int rep = 0; rep < repeat; ++rep) {
            if (EXTRA & 2) asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int i = 0; i < s.n; ++i) {
                const uint64_t db = make_desc(b0 + (s.brow[i] >> 3) * 128, (512 / 8) * 128, 128);
                if (A_TMEM)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem + s.col[i]), "r"(tmem + 448 + 8 * s.aplane[i]), "l"(db), "r"(idesc(s.N[i])), "r"(1u) : "memory");
                else {
                    const uint64_t da = make_desc(a0 + s.aplane[i] * M * KC, (M / 8) * 128, 128);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + s.col[i]), "l"(da), "l"(db), "r"(idesc(s.N[i])), "r"(1u) : "memory");
                }
            }
            if (EXTRA & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static long long* dcyc;
template <int ID, int A_TMEM = 0, int EXTRA = 0>
static void run(int blocks = 148) {
    constexpr Sched s = sched(ID);
    const int repeat = 1024;
    const size_t smem = 8 * M * KC + 512 * KC + 1024;
    CK(cudaFuncSetAttribute(sched_kernel<ID, A_TMEM, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long best = 1LL << 62;
    for (int r = 0; r < 3; ++r) {
        sched_kernel<ID, A_TMEM, EXTRA><<<blocks, 128, smem>>>(repeat, dcyc);
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        if (cyc < best) best = cyc;
    }
    long long cols = 0;
    for (int i = 0; i < s.n; ++i) cols += s.N[i];
    printf("%-52s %2d instr, %5lld cols: %8.1f clk per pass, ideal %7.1f, eff %.3f%s%s\n", names[ID], s.n, cols, (double)best / repeat,
           cols / 2.0, cols / 2.0 / ((double)best / repeat), A_TMEM ? "  [A in TMEM]" : "", blocks == 1 ? "  [1 CTA]" : (EXTRA == 1 ? "  [commit per pass]" : EXTRA == 2 ? "  [fence per pass]" : EXTRA == 3 ? "  [commit + fence per pass]" : ""));
}

int main() {
    CK(cudaMalloc(&dcyc, 8));
    run<0>(); run<1>(); run<2>(); run<3>(); run<4>(); run<5>(); run<6>(); run<7>(); run<8>();
    run<9>(); run<10>();
    run<11>(); run<12>(); run<13>(); run<14>(); run<15>(); run<16>(); run<17>(); run<18>();
    run<1, 1>(); run<3, 1>(); run<8, 1>(); run<11, 1>(); run<17, 1>();
    run<11>(1);
    run<11, 0, 1>(); run<11, 0, 2>(); run<11, 0, 3>();
    return 0;
}
