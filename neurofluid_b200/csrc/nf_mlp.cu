// nf_mlp.cu -- persistent fused positional-encoding + NeRF MLP on tcgen05 tensor cores (sm_100a).
//
// replaces: Embedding.forward x6 (models/nerf.py:21-38 via models/renderer.py:125-179) and
//           NeRF.forward (models/nerf.py:83-124; two instances, models/renderer.py:43-44).
//
// Work unit: a tile of 128 compact "geometry records" (16 fp32 per evaluated ray sample, written by
// the ray-stage kernels in nf_render.cu).  One persistent CTA per SM walks tiles; inside a CTA
//
//   warps 6-9  (PE producers, thread = row)  record -> sin/cos features (fp32 math, double-angle
//              recurrences re-anchored at 2^5) -> fp16/bf16 A-operand tiles written straight into
//              shared memory in the UMMA K-major core-matrix layout.  The 252-wide feature row never
//              exists in HBM.
//   warp 5     (weight producer, one lane)   streams the pre-packed weight K-slabs (8 KB each, already
//              in UMMA layout) from L2 with cp.async.bulk into a 7-stage mbarrier ring.
//   warp 4     (MMA issuer, one lane)        issues tcgen05.mma M=128,N=256(128),K=16 per slab; the
//              fp32 accumulator of layer l lives in TMEM columns [256*(l&1), +256): two buffers, so
//              layer l+1's MMAs run while layer l's accumulator is being drained.
//   warps 0-3  (epilogue, thread = row)      tcgen05.ld -> +bias -> ReLU -> fp16 -> back into the
//              shared-memory activation tile (in place), signalling the MMA warp per 64-column chunk so
//              the next layer starts before the epilogue ends.  sigma (256->1) and rgb (128->3) heads
//              are register dot products in the epilogues of layers 8 and 10; (r,g,b,sigma) goes out as
//              one float4 per row.
//
// Per row the tensor pipe does 671,744 MAC (665,984 algorithmic + K padding 198->208, 54->64).
// HBM traffic per row: 64 B record in, 16 B out (+4 B row id); weights (1.34 MB/net) stay in L2.
#include <type_traits>

#include <stdlib.h>

#include "nf_common.cuh"
#include "nf_mlp.cuh"

namespace nf {
namespace mlp {

constexpr int TILE_M = 128;
constexpr int KX_STEPS = 13;   // xyz-like features 198 -> 208 = 13 K-steps of 16
constexpr int KD_STEPS = 4;    // dir-like features 54 -> 64
constexpr int KH_STEPS = 16;   // hidden width 256
constexpr int STAGE_BYTES = 8192;  // one K-step of a 256-row weight slab: 2 k-chunks x 256 rows x 16 B
constexpr int NSTAGE = 7;
constexpr int N256_STEPS = 154;
constexpr int N128_STEPS = 20;
constexpr int W_BYTES = N256_STEPS * 8192 + N128_STEPS * 4096;
// small fp32 params appended after the weight slabs
constexpr int SP_BIAS = 0;       // [10][256]
constexpr int SP_WSIG = 2560;    // [256]
constexpr int SP_BSIG = 2816;    // [1] (+3 pad)
constexpr int SP_WRGB = 2820;    // [3][128]
constexpr int SP_BRGB = 3204;    // [3] (+1 pad)
constexpr int SP_FLOATS = 3208;
static_assert(PACKED_BYTES == W_BYTES + SP_FLOATS * 4, "packed size");

// shared memory map
constexpr int SM_HIDDEN = 0;                          // 128 x 256 halves
constexpr int SM_PEXYZ = SM_HIDDEN + 65536;           // 128 x 208 halves
constexpr int SM_PEDIR = SM_PEXYZ + 26 * 2048;        // 2 x (128 x 64 halves)
constexpr int SM_WRING = SM_PEDIR + 2 * 8 * 2048;     // NSTAGE x 8 KB
constexpr int SM_SPARAM = SM_WRING + NSTAGE * STAGE_BYTES;
constexpr int SM_BAR = SM_SPARAM + SP_FLOATS * 4;
constexpr int NUM_BARS = 2 * NSTAGE + 12;
constexpr int SM_TMEM_SLOT = SM_BAR + NUM_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM_SLOT + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(SM_BAR % 8 == 0, "barrier alignment");

enum Bar {
    B_WFULL = 0,
    B_WEMPTY = NSTAGE,
    B_PEXYZ_READY = 2 * NSTAGE,
    B_PEXYZ_FREE,
    B_PEDIR_READY,  // 2
    B_PEDIR_FREE = B_PEDIR_READY + 2,  // 2
    B_ACT_READY = B_PEDIR_FREE + 2,    // 4
    B_ACC_FULL = B_ACT_READY + 4,      // 2
};
static_assert(B_ACC_FULL + 2 == NUM_BARS, "barrier count");

constexpr int NUM_THREADS = 320;

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 or bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
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
__device__ __forceinline__ uint32_t umma_idesc(int n, bool bf16) {
    const uint32_t fmt = bf16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
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

__device__ __forceinline__ int layer_pe_steps(int l) { return (l == 0 || l == 4) ? KX_STEPS : (l == 9 ? KD_STEPS : 0); }

#define NF_TRACE(slot) do { if (a.trace && blockIdx.x == 0 && ti == 2) a.trace[(slot)] = clock64(); } while (0)

template <bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_nerf_mlp(const KernelArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n_rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows_cap) : a.n_rows_host;
    const int ntiles = (n_rows + TILE_M - 1) / TILE_M;
    if ((int)blockIdx.x >= ntiles) return;
    const int nl = a.n_layers;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_hidden = s_base + SM_HIDDEN, s_pexyz = s_base + SM_PEXYZ, s_pedir = s_base + SM_PEDIR;
    const uint32_t s_wring = s_base + SM_WRING, s_bar = s_base + SM_BAR;
    float* sp = reinterpret_cast<float*>(smem + SM_SPARAM);
    auto bar = [&](int i) { return s_bar + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bar(B_WFULL + i), 1);
            mbar_init(bar(B_WEMPTY + i), 1);
        }
        mbar_init(bar(B_PEXYZ_READY), 128);
        mbar_init(bar(B_PEXYZ_FREE), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_PEDIR_READY + i), 128);
            mbar_init(bar(B_PEDIR_FREE + i), 1);
            mbar_init(bar(B_ACC_FULL + i), 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_ACT_READY + i), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // small params -> smem
        const float4* src = reinterpret_cast<const float4*>(a.packed + W_BYTES);
        float4* dst = reinterpret_cast<float4*>(sp);
        for (int i = threadIdx.x; i < SP_FLOATS / 4; i += NUM_THREADS) dst[i] = __ldg(src + i);
    }
    if (warp == 4) tmem_alloc(s_base + SM_TMEM_SLOT, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_SLOT);

    const uint32_t a_lbo = a.desc_swap ? 128u : 2048u, a_sbo = a.desc_swap ? 2048u : 128u;

    if (warp == 4) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            uint32_t ws = 0, wph = 0, hidw = 0, lc = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
                for (int l = 0; l < nl; ++l, ++lc) {
                    const uint32_t d_tmem = tmem_base + (lc & 1) * 256;
                    const int n = (l == 9) ? 128 : 256;
                    const uint32_t idesc = umma_idesc(n, BF16);
                    const uint32_t b_lbo = a.desc_swap ? 128u : (uint32_t)n * 16u;
                    const uint32_t b_sbo = a.desc_swap ? (uint32_t)n * 16u : 128u;
                    uint32_t acc = 0;
                    NF_TRACE(100 + l * 8);
                    const int npe = layer_pe_steps(l);
                    if (npe) {
                        uint32_t abase;
                        if (l == 9) {
                            mbar_wait(bar(B_PEDIR_READY + (ti & 1)), (ti >> 1) & 1);
                            abase = s_pedir + (ti & 1) * (8 * 2048);
                        } else {
                            mbar_wait(bar(B_PEXYZ_READY), ti & 1);
                            abase = s_pexyz;
                        }
                        tc_fence_after();
                        for (int j = 0; j < npe; ++j) {
                            mbar_wait(bar(B_WFULL + ws), wph);
                            tc_fence_after();
                            umma_f16(d_tmem, umma_desc(abase + j * 4096, a_lbo, a_sbo),
                                     umma_desc(s_wring + ws * STAGE_BYTES, b_lbo, b_sbo), idesc, acc);
                            acc = 1;
                            umma_commit(bar(B_WEMPTY + ws));
                            if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                        }
                        if (l == 4) umma_commit(bar(B_PEXYZ_FREE));
                    }
                    if (l > 0) {
                        for (int j = 0; j < KH_STEPS; ++j) {
                            if ((j & 3) == 0) {
                                mbar_wait(bar(B_ACT_READY + (j >> 2)), hidw & 1);
                                tc_fence_after();
                                NF_TRACE(100 + l * 8 + 1 + (j >> 2));
                            }
                            mbar_wait(bar(B_WFULL + ws), wph);
                            tc_fence_after();
                            umma_f16(d_tmem, umma_desc(s_hidden + j * 4096, a_lbo, a_sbo),
                                     umma_desc(s_wring + ws * STAGE_BYTES, b_lbo, b_sbo), idesc, acc);
                            acc = 1;
                            umma_commit(bar(B_WEMPTY + ws));
                            if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                        }
                        ++hidw;
                    }
                    umma_commit(bar(B_ACC_FULL + (lc & 1)));
                    NF_TRACE(100 + l * 8 + 5);
                    if (l == 9) umma_commit(bar(B_PEDIR_FREE + (ti & 1)));
                }
            }
        }
    } else if (warp == 5) {
        // ================================================================ weight producer
        if (lane == 0) {
            uint32_t ws = 0, wph = 0;
            const int nsteps = (nl == 10) ? (N256_STEPS + N128_STEPS) : 138;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t* src = a.packed;
                for (int s = 0; s < nsteps; ++s) {
                    const uint32_t bytes = (s < N256_STEPS) ? 8192u : 4096u;
                    mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                    mbar_arrive_expect_tx(bar(B_WFULL + ws), bytes);
                    bulk_g2s(s_wring + ws * STAGE_BYTES, src, bytes, bar(B_WFULL + ws));
                    src += bytes;
                    if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp < 4) {
        // ================================================================ epilogue (thread = row)
        const int tr = threadIdx.x;  // 0..127
        uint32_t lc = 0;
        int ti = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int row = tile * TILE_M + tr;
            float sigma = 0.f;
            for (int l = 0; l < nl; ++l, ++lc) {
                const uint32_t buf = lc & 1;
                mbar_wait(bar(B_ACC_FULL + buf), (lc >> 1) & 1);
                tc_fence_after();
                if (tr == 0) NF_TRACE(l * 8);
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256;
                if (l < 9) {
                    const bool writes = (l + 1 < nl);
                    const float* bias = sp + SP_BIAS + l * 256;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[64];
                        tmem_ld32(taddr + c * 64, v);
                        tmem_ld32(taddr + c * 64 + 32, v + 32);
                        tmem_ld_wait();
                        float f[64];
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 64 + i);
                            f[i] = __uint_as_float(v[i]) + b4.x;
                            f[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
                            f[i + 2] = __uint_as_float(v[i + 2]) + b4.z;
                            f[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
                        }
                        if (l != 8) {
#pragma unroll
                            for (int i = 0; i < 64; ++i) f[i] = fmaxf(f[i], 0.f);
                        }
                        if (l == 7) {
                            const float* wsig = sp + SP_WSIG + c * 64;
#pragma unroll
                            for (int i = 0; i < 64; i += 4) {
                                const float4 w4 = *reinterpret_cast<const float4*>(wsig + i);
                                sigma = fmaf(f[i], w4.x, sigma);
                                sigma = fmaf(f[i + 1], w4.y, sigma);
                                sigma = fmaf(f[i + 2], w4.z, sigma);
                                sigma = fmaf(f[i + 3], w4.w, sigma);
                            }
                        }
                        if (writes) {
                            const uint32_t dst = s_hidden + (uint32_t)(c * 8) * 2048 + (uint32_t)tr * 16;
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                st_shared_v4(dst + q * 2048, pack2<BF16>(f[8 * q], f[8 * q + 1]),
                                             pack2<BF16>(f[8 * q + 2], f[8 * q + 3]),
                                             pack2<BF16>(f[8 * q + 4], f[8 * q + 5]),
                                             pack2<BF16>(f[8 * q + 6], f[8 * q + 7]));
                            fence_proxy_async();
                            tc_fence_before();
                            mbar_arrive(bar(B_ACT_READY + c));
                        }
                        if (tr == 0) NF_TRACE(l * 8 + 1 + c);
                    }
                    if (l == 7) {
                        sigma += sp[SP_BSIG];
                        if (!writes && row < n_rows) {  // sigma-only network
                            const int dst = a.rowid ? a.rowid[row] : row;
                            if (dst >= 0) a.out4[dst] = make_float4(0.f, 0.f, 0.f, sigma);
                        }
                    }
                } else {
                    float rgb[3] = {0.f, 0.f, 0.f};
                    const float* bias = sp + SP_BIAS + 9 * 256;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t v[64];
                        tmem_ld32(taddr + c * 64, v);
                        tmem_ld32(taddr + c * 64 + 32, v + 32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 64; ++i) {
                            const float f = fmaxf(__uint_as_float(v[i]) + bias[c * 64 + i], 0.f);
                            rgb[0] = fmaf(f, sp[SP_WRGB + c * 64 + i], rgb[0]);
                            rgb[1] = fmaf(f, sp[SP_WRGB + 128 + c * 64 + i], rgb[1]);
                            rgb[2] = fmaf(f, sp[SP_WRGB + 256 + c * 64 + i], rgb[2]);
                        }
                    }
                    tc_fence_before();
                    if (row < n_rows) {
                        const int dst = a.rowid ? a.rowid[row] : row;
                        if (dst >= 0) {
                            float4 o;
                            o.x = 1.0f / (1.0f + expf(-(rgb[0] + sp[SP_BRGB])));
                            o.y = 1.0f / (1.0f + expf(-(rgb[1] + sp[SP_BRGB + 1])));
                            o.z = 1.0f / (1.0f + expf(-(rgb[2] + sp[SP_BRGB + 2])));
                            o.w = sigma;
                            a.out4[dst] = o;
                        }
                    }
                }
            }
        }
    } else {
        // ================================================================ PE producers (warps 6-9)
        const int tp = threadIdx.x - 192;  // 0..127
        int ti = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int row = tile * TILE_M + tp;
            float r[16];
            if (row < n_rows) {
                const float4* src = reinterpret_cast<const float4*>(a.records + (size_t)row * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t = __ldg(src + i);
                    r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0.f;
            }
            // xyz-like block: [PE10(x) 63 | PE4(density) 9 | PE10(smoothed) 63 | PE10(variance) 63 | 0 x10]
            mbar_wait(bar(B_PEXYZ_FREE), (ti & 1) ^ 1);
            if (tp == 0) NF_TRACE(300);
            {
                RowWriter<BF16> w;
                w.base = s_pexyz + (uint32_t)tp * 16;
                emit_encoding<BF16, 0, 3, 10>(w, r + 0);
                emit_encoding<BF16, 63, 1, 4>(w, r + 3);
                emit_encoding<BF16, 72, 3, 10>(w, r + 4);
                emit_encoding<BF16, 135, 3, 10>(w, r + 7);
                static_for<198, 208>([&](auto ci) { w.template put<decltype(ci)::value>(0.f); });
            }
            fence_proxy_async();
            mbar_arrive(bar(B_PEXYZ_READY));
            if (tp == 0) NF_TRACE(301);
            if (nl == 10) {
                // dir-like block: [PE4(ray dir) 27 | PE4(smoothed dir) 27 | 0 x10]
                mbar_wait(bar(B_PEDIR_FREE + (ti & 1)), ((ti >> 1) & 1) ^ 1);
                RowWriter<BF16> w;
                w.base = s_pedir + (ti & 1) * (8 * 2048) + (uint32_t)tp * 16;
                emit_encoding<BF16, 0, 3, 4>(w, r + 10);
                emit_encoding<BF16, 27, 3, 4>(w, r + 13);
                static_for<54, 64>([&](auto ci) { w.template put<decltype(ci)::value>(0.f); });
                fence_proxy_async();
                mbar_arrive(bar(B_PEDIR_READY + (ti & 1)));
                if (tp == 0) NF_TRACE(302);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// weight packer: fp32 nn.Linear tensors -> K-step slabs in UMMA core-matrix order + fp32 small params
// slab(step)[kc][n][e]  (kc: 8-column chunk 0/1, n: output row, e: 0..7) = W_layer[n][k_src(step,kc,e)]
// ------------------------------------------------------------------------------------------------
struct PackArgs {
    const float* w[12];
    const float* b[12];
};

__device__ __forceinline__ void step_source(int s, int& layer, int& k0, int& src_off, int& src_valid, int& ld) {
    // returns: layer (0..9 in the index space of PackArgs: 0-7 xyz_encoding, 8 final, 9 dir), first padded
    // column k0 of this step inside its segment, source column offset, number of valid source columns in the
    // segment, and the source leading dimension.
    if (s < 13) { layer = 0; k0 = s * 16; src_off = 0; src_valid = 198; ld = 198; }
    else if (s < 61) { layer = 1 + (s - 13) / 16; k0 = ((s - 13) % 16) * 16; src_off = 0; src_valid = 256; ld = 256; }
    else if (s < 74) { layer = 4; k0 = (s - 61) * 16; src_off = 0; src_valid = 198; ld = 454; }
    else if (s < 90) { layer = 4; k0 = (s - 74) * 16; src_off = 198; src_valid = 256; ld = 454; }
    else if (s < 138) { layer = 5 + (s - 90) / 16; k0 = ((s - 90) % 16) * 16; src_off = 0; src_valid = 256; ld = 256; }
    else if (s < 154) { layer = 8; k0 = (s - 138) * 16; src_off = 0; src_valid = 256; ld = 256; }
    else if (s < 158) { layer = 9; k0 = (s - 154) * 16; src_off = 256; src_valid = 54; ld = 310; }
    else { layer = 9; k0 = (s - 158) * 16; src_off = 0; src_valid = 256; ld = 310; }
}

template <bool BF16>
__global__ void k_pack_weights(PackArgs p, uint8_t* out) {
    // one thread per (step, kc, n): writes 8 halves (16 B)
    const int total = N256_STEPS * 2 * 256 + N128_STEPS * 2 * 128;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) {
        int s, kc, n, nrows;
        size_t byte_off;
        if (t < N256_STEPS * 512) {
            s = t / 512; kc = (t % 512) / 256; n = t % 256; nrows = 256;
            byte_off = (size_t)s * 8192;
        } else {
            const int u = t - N256_STEPS * 512;
            s = N256_STEPS + u / 256; kc = (u % 256) / 128; n = u % 128; nrows = 128;
            byte_off = (size_t)N256_STEPS * 8192 + (size_t)(s - N256_STEPS) * 4096;
        }
        int layer, k0, src_off, src_valid, ld;
        step_source(s, layer, k0, src_off, src_valid, ld);
        const float* W = p.w[layer];
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const int ka = k0 + kc * 8 + e, kb = ka + 1;
            const float va = ka < src_valid ? W[(size_t)n * ld + src_off + ka] : 0.f;
            const float vb = kb < src_valid ? W[(size_t)n * ld + src_off + kb] : 0.f;
            pk[e >> 1] = pack2<BF16>(va, vb);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + byte_off + ((size_t)kc * nrows + n) * 16);
        *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    // small params
    float* sp = reinterpret_cast<float*>(out + W_BYTES);
    if (t < SP_FLOATS) {
        float v = 0.f;
        if (t < 2560) {
            const int l = t / 256, i = t % 256;
            v = (l < 9 || i < 128) ? p.b[l][i] : 0.f;
        } else if (t < SP_BSIG) v = p.w[10][t - SP_WSIG];
        else if (t == SP_BSIG) v = p.b[10][0];
        else if (t >= SP_WRGB && t < SP_BRGB) v = p.w[11][t - SP_WRGB];
        else if (t >= SP_BRGB && t < SP_BRGB + 3) v = p.b[11][t - SP_BRGB];
        sp[t] = v;
    }
}

int launch(const KernelArgs& a, int dtype, cudaStream_t st) {
    static bool attr_set[2] = {false, false};
    const int di = dtype == NF_DTYPE_BF16 ? 1 : 0;
    if (!attr_set[di]) {
        if (di)
            NF_CUDA_OK(cudaFuncSetAttribute(k_nerf_mlp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        else
            NF_CUDA_OK(cudaFuncSetAttribute(k_nerf_mlp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set[di] = true;
    }
    const int grid = num_sms();
    if (di)
        k_nerf_mlp<true><<<grid, NUM_THREADS, SM_TOTAL, st>>>(a);
    else
        k_nerf_mlp<false><<<grid, NUM_THREADS, SM_TOTAL, st>>>(a);
    NF_LAUNCH_OK();
    return NF_OK;
}

}  // namespace mlp
}  // namespace nf

using namespace nf;

extern "C" size_t nf_render_packed_weights_bytes(void) { return (size_t)mlp::PACKED_BYTES; }

extern "C" int nf_render_pack_weights(const float* const* params, int dtype, void* packed_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(params && packed_out, NF_E_INVALID, "nf_render_pack_weights: null argument");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_render_pack_weights: dtype %d", dtype);
    mlp::PackArgs p;
    for (int i = 0; i < 12; ++i) {
        p.w[i] = params[2 * i];
        p.b[i] = params[2 * i + 1];
        NF_REQUIRE(p.w[i] && p.b[i], NF_E_INVALID, "nf_render_pack_weights: null parameter %d", i);
    }
    const int total = mlp::N256_STEPS * 512 + mlp::N128_STEPS * 256;
    if (dtype == NF_DTYPE_BF16)
        mlp::k_pack_weights<true><<<(total + 255) / 256, 256, 0, st>>>(p, (uint8_t*)packed_out);
    else
        mlp::k_pack_weights<false><<<(total + 255) / 256, 256, 0, st>>>(p, (uint8_t*)packed_out);
    NF_LAUNCH_OK();
    return NF_OK;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

extern "C" int nf_nerf_mlp_forward(const void* packed, int dtype, const float* records, int n_rows, int sigma_only,
                                   float* out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(packed && out && n_rows >= 0, NF_E_INVALID, "nf_nerf_mlp_forward: bad arguments");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_nerf_mlp_forward: dtype %d", dtype);
    if (n_rows == 0) return NF_OK;
    NF_REQUIRE(records != nullptr, NF_E_INVALID, "nf_nerf_mlp_forward: null records");
    mlp::KernelArgs a;
    a.packed = (const uint8_t*)packed;
    a.records = records;
    a.rowid = nullptr;
    a.n_rows_dev = nullptr;
    a.n_rows_host = n_rows;
    a.n_rows_cap = n_rows;
    a.n_layers = sigma_only ? 8 : 10;
    a.desc_swap = env_int("NF_MLP_DESC_SWAP", 0);
    a.out4 = (float4*)out;
    a.trace = (long long*)(uintptr_t)strtoull(getenv("NF_MLP_TRACE_PTR") ? getenv("NF_MLP_TRACE_PTR") : "0", nullptr, 0);
    return mlp::launch(a, dtype, st);
}
