// umma_mn_test.cu -- correctness probe: tcgen05.mma kind::f16 with MN-major (transposed) shared-memory operands in the
// no-swizzle layout the kernels of this repo keep their activation tiles in:
//     byte(row r, column c) = (c / 8) * 2048 + r * 16 + (c % 8) * 2          (128 rows, 16-bit elements)
// Read as an MN-major operand (MN = column, K = row) this is the canonical INTERLEAVE layout with SBO = 2048 (between
// 8-column groups) and LBO = 128 (between 8-row groups).  D[m][n] = sum_r A[r][m] * B[r][n]   (weight gradients).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_mn_test tools/umma_mn_test.cu && tools/umma_mn_test
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
constexpr int NCOL_A = 128, NCOL_B = 64;

// variant 0: LBO=128, SBO=2048;  variant 1: swapped
__global__ void __launch_bounds__(128, 1) k(const __half* A /*[128][128] row-major*/, const __half* B /*[128][64]*/, float* D /*[128][64]*/,
                                           int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* sa = smem;                    // 16 chunks x 2048
    uint8_t* sbm = smem + 16 * 2048;       // 8 chunks x 2048
    for (int i = threadIdx.x; i < 128 * NCOL_A; i += 128) {
        const int r = i / NCOL_A, c = i % NCOL_A;
        *reinterpret_cast<__half*>(sa + (c / 8) * 2048 + r * 16 + (c % 8) * 2) = A[i];
    }
    for (int i = threadIdx.x; i < 128 * NCOL_B; i += 128) {
        const int r = i / NCOL_B, c = i % NCOL_B;
        *reinterpret_cast<__half*>(sbm + (c / 8) * 2048 + r * 16 + (c % 8) * 2) = B[i];
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (threadIdx.x == 0) {
        // idesc: c_format f32 (bit 4), a/b f16, a_major = b_major = 1 (MN), N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NCOL_B >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lbo = variant == 0 ? 128u : 2048u, sbo = variant == 0 ? 2048u : 128u;
        for (int j = 0; j < 8; ++j)          // K-step j = rows 16j .. 16j+15 = two 8-row groups = +256 bytes
            mma(tmem, desc(smem_u32(sa) + j * 256, lbo, sbo), desc(smem_u32(sbm) + j * 256, lbo, sbo), idesc, j > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < NCOL_B; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[threadIdx.x * NCOL_B + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    std::vector<__half> A(128 * NCOL_A), B(128 * NCOL_B);
    std::vector<float> Af(A.size()), Bf(B.size());
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(Af[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (rand() % 13 - 6) / 4.0f; B[i] = __float2half(Bf[i]); }
    std::vector<float> ref(128 * NCOL_B, 0.f);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NCOL_B; ++n) {
            float s = 0.f;
            for (int r = 0; r < 128; ++r) s += Af[r * NCOL_A + m] * Bf[r * NCOL_B + n];
            ref[m * NCOL_B + n] = s;
        }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, ref.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0, ref.size() * 4);
        k<<<1, 128, 64 * 1024>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> out(ref.size());
        cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (size_t i = 0; i < out.size(); ++i) mx = fmax(mx, fabs(out[i] - ref[i]));
        printf("umma_mn_test variant %d (LBO=%d SBO=%d): %s max |err| = %g  (D[0][0]=%g ref %g; D[5][3]=%g ref %g)\n", variant,
               variant == 0 ? 128 : 2048, variant == 0 ? 2048 : 128, cudaGetErrorString(e), mx, out[0], ref[0], out[5 * NCOL_B + 3], ref[5 * NCOL_B + 3]);
    }
    return 0;
}
