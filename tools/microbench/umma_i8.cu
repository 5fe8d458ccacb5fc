// Feasibility probe for a split-integer (Ozaki-style) FP64 contraction on the 5th-gen tensor cores:
// one CTA computes C[128 x 64] (int32) = A[128 x K] (s8) * B[64 x K]^T (s8) with tcgen05.mma kind::i8, operands in the
// canonical K-major no-swizzle shared-memory layout (8-row x 16-byte core matrices), accumulator in TMEM, result read
// back with tcgen05.ld and checked against the host.  Then a timing loop measures the int8 MMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8 umma_i8.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 64, KC = 32;       // one MMA: 128 x 64 x 32 (8-bit operands)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: element (row, k) of a (rows x 32) chunk at
//   (k/16) * LBO + (row/8) * SBO + (row%8) * 16 + (k%16),  LBO = rows/8 * 128, SBO = 128
__host__ __device__ inline int canon_off(int rows, int row, int k) { return (k / 16) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 16); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version for sm_100
    return d;                                    // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128, 1) umma_i8_kernel(const int8_t* __restrict__ Ap, const int8_t* __restrict__ Bp,
                                                         int32_t* __restrict__ C, int nchunks, int repeat, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    int8_t* sA = reinterpret_cast<int8_t*>(smem);                 // nchunks * M * KC
    int8_t* sB = sA + (size_t)nchunks * M * KC;                   // nchunks * N * KC
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < nchunks * M * KC / 16; i += 128) reinterpret_cast<int4*>(sA)[i] = reinterpret_cast<const int4*>(Ap)[i];
    for (int i = tid; i < nchunks * N * KC / 16; i += 128) reinterpret_cast<int4*>(sB)[i] = reinterpret_cast<const int4*>(Bp)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA unit
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    long long t0 = clock64();
    if (tid == 0) {
        for (int rep = 0; rep < repeat; ++rep) {
            for (int kc = 0; kc < nchunks; ++kc) {
                const uint64_t da = make_desc(smem_u32(sA + (size_t)kc * M * KC), (M / 8) * 128, 128);
                const uint64_t db = make_desc(smem_u32(sB + (size_t)kc * N * KC), (N / 8) * 128, 128);
                const uint32_t accumulate = (rep > 0 || kc > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs to retire
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0 && cycles) *cycles = t1 - t0;

    // warp w owns TMEM lanes [32w, 32w+32): thread = one row, 64 consecutive int32 columns
    uint32_t r[64];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = tid;
#pragma unroll
    for (int j = 0; j < 64; ++j) C[row * N + j] = (int32_t)r[j];

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

template <int NN, int KIND>   // KIND 0: i8 (K=32 bytes), 1: f16 (K=16 elements = 32 bytes)
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int nchunks, int repeat, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = sA + (size_t)nchunks * M * KC;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < nchunks * (M + NN) * KC / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
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
    constexpr uint32_t idesc = KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24))
                                         : ((1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
    long long t0 = clock64();
    if (tid == 0) {
        for (int rep = 0; rep < repeat; ++rep)
            for (int kc = 0; kc < nchunks; ++kc) {
                const uint64_t da = make_desc(smem_u32(sA + (size_t)kc * M * KC), (M / 8) * 128, 128);
                const uint64_t db = make_desc(smem_u32(sB + (size_t)kc * NN * KC), (NN / 8) * 128, 128);
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

template <int NN, int KIND>
void rate(long long* dcyc, int blocks) {
    const int nchunks = 4, repeat = 256;
    const size_t smem = (size_t)nchunks * (M + NN) * KC + 1024;
    CK(cudaFuncSetAttribute(umma_rate_kernel<NN, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_rate_kernel<NN, KIND><<<blocks, 128, smem>>>(nchunks, repeat, dcyc);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double mmas = (double)repeat * nchunks;
    const int kelems = KIND == 0 ? 32 : 16;
    printf("%s M=128 N=%3d blocks=%3d: %.1f cycles per MMA, %.0f MAC/clk/SM\n", KIND == 0 ? "i8 " : "f16", NN, blocks, cyc / mmas,
           mmas * M * NN * kelems / cyc);
}

int main() {
    const int nchunks = 8, K = nchunks * KC;
    std::vector<int8_t> A(M * K), B(N * K), Ap(M * K), Bp(N * K);
    srand(1);
    for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
    for (int kc = 0; kc < nchunks; ++kc)
        for (int k = 0; k < KC; ++k) {
            for (int i = 0; i < M; ++i) Ap[(size_t)kc * M * KC + canon_off(M, i, k)] = A[i * K + kc * KC + k];
            for (int j = 0; j < N; ++j) Bp[(size_t)kc * N * KC + canon_off(N, j, k)] = B[j * K + kc * KC + k];
        }
    int8_t *dA, *dB; int32_t* dC; long long* dcyc;
    CK(cudaMalloc(&dA, Ap.size())); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)nchunks * (M + N) * KC + 1024;
    CK(cudaFuncSetAttribute(umma_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_i8_kernel<<<1, 128, smem>>>(dA, dB, dC, nchunks, 1, dcyc);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> C(M * N);
    CK(cudaMemcpy(C.data(), dC, M * N * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int32_t ref = 0;
            for (int k = 0; k < K; ++k) ref += (int32_t)A[i * K + k] * (int32_t)B[j * K + k];
            if (ref != C[i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, C[i * N + j], ref); ++bad; }
        }
    printf("tcgen05 kind::i8 128x64x%d: %lld mismatches of %d\n", K, bad, M * N);
    // issue-rate probe: many MMAs back to back on one SM
    for (int repeat : {64, 512}) {
        umma_i8_kernel<<<1, 128, smem>>>(dA, dB, dC, nchunks, repeat, dcyc);
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        const double mmas = (double)repeat * nchunks;
        printf("repeat %4d: %lld cycles for %.0f MMAs -> %.1f cycles per 128x64x32 MMA, %.0f MAC/clk/SM\n", repeat, cyc, mmas,
               cyc / mmas, mmas * M * N * KC / cyc);
    }
    rate<64, 0>(dcyc, 1); rate<128, 0>(dcyc, 1); rate<256, 0>(dcyc, 1);
    rate<64, 1>(dcyc, 1); rate<128, 1>(dcyc, 1); rate<256, 1>(dcyc, 1);
    rate<64, 0>(dcyc, 148); rate<256, 0>(dcyc, 148); rate<256, 1>(dcyc, 148);
    return 0;
}
