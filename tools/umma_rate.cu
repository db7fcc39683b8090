// umma_rate.cu -- microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128 per CTA) for different
// shared-memory operand layouts.  Contents are garbage; only clock64 deltas matter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
    if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ long long gt1() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
struct P { int n, layout, a_lbo, a_sbo, b_lbo, b_sbo, kadv_a, kadv_b, nmma, ksteps; long long* out; int mode; int m; };

template <int CG>
__global__ void __launch_bounds__(128, 1) k(P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2s, bar3s;
    __shared__ uint32_t slot;
    const uint32_t sb = (smem_u32(smem) + 1023) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar2s)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3s)) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)((p.m ? p.m : (CG == 2 ? 256 : 128)) >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32 && rank == 0 && (p.mode & 8)) {
        // warp-converged issue: the whole warp runs the loop, one elected lane issues
        const uint32_t a0 = sb, b0 = sb + 64 * 1024;
        const uint64_t da = desc(a0, p.a_lbo, p.a_sbo, p.layout), db = desc(b0, p.b_lbo, p.b_sbo, p.layout);
        const uint64_t sa = (uint64_t)(p.kadv_a >> 4), sbb = (uint64_t)(p.kadv_b >> 4);
        const uint32_t bar2 = smem_u32(&bar2s), bar3 = smem_u32(&bar3s);
        if (elect_one()) {
            for (int i = 0; i < 8; ++i) mma<CG>(tm, da, db, idesc, i > 0);
            commit<CG>(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t0 = clock64();
        const long long g0 = gt1();
        for (int it = 0; it < p.nmma / 4; ++it) {
            const uint32_t d = tm + (it & 1) * 256;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (p.mode & 2) mbar_wait(bar3, 0);
                if (p.mode & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    mma<CG>(d, da + ks * sa, db + ks * sbb, idesc, ks > 0);
                    if (p.mode & 1) commit<CG>(bar2);
                }
                __syncwarp();
            }
        }
        if (elect_one()) commit<CG>(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 1);
        t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { p.out[0] = t1 - t0; p.out[1] = gt1() - g0; }
    } else if (threadIdx.x == 0 && rank == 0) {
        const uint32_t a0 = sb, b0 = sb + 64 * 1024;
        // warm-up
        for (int i = 0; i < 8; ++i) mma<CG>(tm, desc(a0, p.a_lbo, p.a_sbo, p.layout), desc(b0, p.b_lbo, p.b_sbo, p.layout), idesc, i > 0);
        commit<CG>(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t0 = clock64();
        const uint64_t da = desc(a0, p.a_lbo, p.a_sbo, p.layout), db = desc(b0, p.b_lbo, p.b_sbo, p.layout);
        const uint64_t sa = (uint64_t)(p.kadv_a >> 4), sbb = (uint64_t)(p.kadv_b >> 4);
        const uint32_t bar2 = smem_u32(&bar2s), bar3 = smem_u32(&bar3s);
        for (int it = 0; it < p.nmma / 4; ++it) {
            const uint32_t d = tm + (it & 1) * 256;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (p.mode & 2) mbar_wait(bar3, 0);          // already-complete barrier
                if (p.mode & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma<CG>(d, da + ks * sa, db + ks * sbb, idesc, ks > 0);
                if (p.mode & 1) commit<CG>(bar2);
            }
        }
        commit<CG>(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 1);
        t1 = clock64();
        if (blockIdx.x == 0) { p.out[0] = t1 - t0; }
    } else if (CG == 2 && threadIdx.x == 0) {
        mbar_wait(smem_u32(&bar), 0);
        mbar_wait(smem_u32(&bar), 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x < 32) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}

static void run(const char* name, int cg, P p, int grid) {
    long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    p.out = d;
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cg == 2 ? cudaLaunchKernelEx(&cfg, k<2>, p) : cudaLaunchKernelEx(&cfg, k<1>, p);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s cg=%d grid=%3d M=%3d N=%3d : %7.1f cycles/MMA %7.1f ns/MMA (%s %s)\n", name, cg, grid, p.m, p.n, (double)h[0] / p.nmma,
           (double)h[1] / p.nmma, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d);
}

int main() {
    const int NM = 1024;
    for (int grid : {1, 148}) {
        for (int m : {128, 64}) {
            for (int n : {256, 128, 64, 32, 16}) {
                run("noswz elect", 1, P{n, 0, 2048, 128, n * 16, 128, 4096, n * 32, NM, 16, nullptr, 8, m}, grid);
            }
        }
        run("noswz elect", 2, P{256, 0, 2048, 128, 128 * 16, 128, 4096, 128 * 32, NM, 16, nullptr, 8, 256}, grid);
        run("noswz elect", 2, P{64, 0, 2048, 128, 32 * 16, 128, 4096, 32 * 32, NM, 16, nullptr, 8, 256}, grid);
        run("noswz elect", 2, P{256, 0, 2048, 128, 128 * 16, 128, 4096, 128 * 32, NM, 16, nullptr, 8, 128}, grid);
    }
    return 0;
}
