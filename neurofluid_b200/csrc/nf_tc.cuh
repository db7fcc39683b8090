// nf_tc.cuh -- device helpers shared by the tcgen05 kernels of the NeRF MLP (forward: nf_mlp.cu, backward: nf_mlp_bwd.cu):
// mbarrier / bulk-copy / tcgen05 PTX wrappers, UMMA descriptors, and the positional-encoding row writer.
#pragma once
#include <type_traits>

#include "nf_common.cuh"

namespace nf {
namespace mlp {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// ---- cluster (CTA pair) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {  // shared::cta -> shared::cluster of CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Arrive on a barrier anywhere in the cluster window.  Default semantics (release at CTA scope), as in
// CUTLASS's ClusterBarrier::arrive(cta_id): a cluster-scope release costs ~1000 cycles per arrive (measured),
// and the data guarded here is this CTA's own shared memory, ordered for the tensor core by fence.proxy.async.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// bulk copy global -> the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <bool PAIR>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    if constexpr (PAIR) {  // executed by the same warp of both CTAs of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <bool PAIR>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (PAIR)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 or bf16 operands, fp32 accumulate)
// PAIR: one instruction drives both SMs of the pair -- M = 256 (128 rows of A and of D per CTA), each CTA
// supplies half of the N rows of B from its own shared memory (same offsets in both CTAs).
template <bool PAIR>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    if constexpr (PAIR)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// PAIR: the arrive is multicast to the barrier at the same offset in the CTAs of `mask` (cluster ranks; default: the
// two CTAs of a cluster of 2).
template <bool PAIR>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t mask = 3) {
    if constexpr (PAIR)
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
            "h"(mask)
            : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K-major, no-swizzle UMMA shared-memory descriptor.  Canonical layout (units of 16 B):
// ((8 rows, m groups), 2 k-chunks) : ((1, SBO), LBO)  -- 8 rows x 16 B core matrices.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__device__ __forceinline__ uint32_t umma_idesc(int n, bool bf16, int m) {
    const uint32_t fmt = bf16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Streams feature columns (compile-time column index) of one row into the A-operand layout:
// byte address of (row r, column k) = base + (k/8)*2048 + r*16 + (k%8)*2.
template <bool BF16>
struct RowWriter {
    uint32_t base;  // smem address of (k-chunk 0, this row)
    float lo;
    uint32_t pk[4];
    template <int COL>
    __device__ __forceinline__ void put(float v) {
        if constexpr ((COL & 1) == 0) {
            lo = v;
        } else {
            pk[(COL & 7) >> 1] = pack2<BF16>(lo, v);
            if constexpr ((COL & 7) == 7) st_shared_v4(base + (COL >> 3) * 2048, pk[0], pk[1], pk[2], pk[3]);
        }
    }
};

// [v, sin(2^0 v), cos(2^0 v), ..., sin(2^(L-1) v), cos(2^(L-1) v)] for a C-vector, reference column
// order (models/nerf.py:33-38).  sin/cos of 2^f v by double-angle recurrence, re-anchored with a
// direct sincosf at f = 5 so that the absolute error stays below ~2e-6 (operands are rounded to
// fp16/bf16 afterwards: 2.4e-4 / 2e-3 relative).
template <bool BF16, int BASE, int C, int L>
__device__ __forceinline__ void emit_encoding(RowWriter<BF16>& w, const float* v) {
    static_for<0, C>([&](auto ci) { w.template put<BASE + decltype(ci)::value>(v[decltype(ci)::value]); });
    float s[C], c[C];
#pragma unroll
    for (int i = 0; i < C; ++i) sincosf(v[i], &s[i], &c[i]);
    static_for<0, L>([&](auto fi) {
        constexpr int f = decltype(fi)::value;
        if constexpr (f == 5) {
#pragma unroll
            for (int i = 0; i < C; ++i) sincosf(32.0f * v[i], &s[i], &c[i]);
        } else if constexpr (f > 0) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const float s2 = 2.0f * s[i] * c[i];
                const float c2 = 1.0f - 2.0f * s[i] * s[i];
                s[i] = s2;
                c[i] = c2;
            }
        }
        static_for<0, C>([&](auto ci) {
            constexpr int i = decltype(ci)::value;
            w.template put<BASE + C + 2 * C * f + i>(s[i]);
        });
        static_for<0, C>([&](auto ci) {
            constexpr int i = decltype(ci)::value;
            w.template put<BASE + C + 2 * C * f + C + i>(c[i]);
        });
    });
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <class T>
__device__ __forceinline__ T uniform(T v) { return __shfl_sync(0xffffffffu, v, 0); }   // provably warp-uniform copy


}  // namespace mlp
}  // namespace nf
