// nf_cconv.cuh -- shared pieces of the transition-model kernels (nf_cconv.cu: forward + operator entry points,
// nf_cconv_bwd.cu: backward): neighbour pair / slab-list formats, the filter geometry, the tensor-core ContinuousConv
// kernel k_cconv_tc (forward layers 1-2, and their feature gradients with re-packed weights), weight packing and the
// workspace layouts.  Reference: models/transmodel.py:106-142 (open3d ContinuousConv + nn.Linear).
#pragma once
#include <stdlib.h>
#include <type_traits>

#include "nf_common.cuh"

namespace nf {
namespace cconv {

constexpr int MAXNBR = 128;           // neighbour slots per particle (fixed stride)
constexpr int FSIZE = 4;              // filter size per axis
constexpr int NCELL = 64;

struct __align__(16) Pair {
    int j;                // neighbour index
    unsigned char cell[8];  // (z*4+y)*4+x of the 8 trilinear corners
    int pad;
    float w[8];           // trilinear weight * window
};
static_assert(sizeof(Pair) == 48, "Pair layout");

// Slab lists: the fluid->fluid pairs regrouped by the (z,y) row of the 4x4x4 filter they touch.  A neighbour's 8
// trilinear corners lie in <= 4 of the 16 rows; for row s the entry is {j, weight per x cell of that row}.  Per
// particle: off[17] (prefix over rows, u16) and SLABCAP = 4*MAXNBR entries {j (int), wx (float4)}; entries keep the
// pair order, so every sum runs in the same order as a walk over the pair list.
constexpr int SLABCAP = 4 * MAXNBR;
constexpr int SLABOFF = 32;           // u16 per particle (17 used; 64-byte rows)

// ---------------------------------------------------------------- PTX wrappers (same as nf_mlp.cu)
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int m, int n, bool bf16) {
    const uint32_t fmt = bf16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// filter geometry (open3d ContinuousConv: ball_to_cube_volume_preserving + linear interpolation,
// align_corners=True)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

__device__ __forceinline__ void ball_to_cube(float& X, float& Y, float& Z) {
    const float sq = X * X + Y * Y + Z * Z;
    const float n = sqrtf(sq);
    if (sq < 1e-12f) { X = Y = Z = 0.f; return; }
    const float xy2 = X * X + Y * Y;
    if (1.25f * Z * Z > xy2) {
        const float s = sqrtf(3.0f * n / (n + fabsf(Z)));
        X *= s; Y *= s; Z = sgnf(Z) * n;
    } else {
        const float s = n / sqrtf(xy2);
        X *= s; Y *= s; Z *= 1.5f;
    }
    const float nxy2 = X * X + Y * Y;
    if (nxy2 < 1e-12f) {
        X = 0.f; Y = 0.f;
    } else {
        const float nxy = sqrtf(nxy2);
        const float four_over_pi = 1.2732395447351628f;
        if (fabsf(Y) <= fabsf(X)) {
            const float t = sgnf(X) * nxy;
            Y = t * four_over_pi * atanf(Y / X);
            X = t;
        } else {
            const float t = sgnf(Y) * nxy;
            X = t * four_over_pi * atanf(X / Y);
            Y = t;
        }
    }
}

__device__ __forceinline__ void filter_corners(float rx, float ry, float rz, float inv_radius, float window, Pair& p) {
    float x = rx * inv_radius, y = ry * inv_radius, z = rz * inv_radius;
    ball_to_cube(x, y, z);
    const float c[3] = {x * 0.5f, y * 0.5f, z * 0.5f};
    int i0[3], i1[3];
    float f[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t = (c[a] + 0.5f) * (float)(FSIZE - 1);
        const float fl = floorf(t);
        f[a] = t - fl;
        i0[a] = min(max((int)fl, 0), FSIZE - 1);
        i1[a] = min(max((int)fl + 1, 0), FSIZE - 1);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
        const int ix = bx ? i1[0] : i0[0], iy = by ? i1[1] : i0[1], iz = bz ? i1[2] : i0[2];
        p.cell[k] = (unsigned char)((iz * FSIZE + iy) * FSIZE + ix);
        p.w[k] = window * (bx ? f[0] : 1.f - f[0]) * (by ? f[1] : 1.f - f[1]) * (bz ? f[2] : 1.f - f[2]);
    }
}


// features in one of three storage types (kind 0: fp32, 1: fp16, 2: bf16), row stride ld elements
__device__ __forceinline__ float load_feat(const void* p, size_t idx, int kind) {
    if (kind == 0) return __ldg(reinterpret_cast<const float*>(p) + idx);
    if (kind == 1) return __half2float(reinterpret_cast<const __half*>(p)[idx]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}

// ------------------------------------------------------------------------------------------------
// layers 1..3 on tensor cores
// ------------------------------------------------------------------------------------------------
struct ConvArgs {
    const int* slab_j; const float4* slab_w; const unsigned short* slab_off;   // fluid->fluid slab lists
    const void* x_in;                      // (N,CIN) fp16/bf16, already ReLU'd
    const uint8_t* w_packed;               // slabs (16 conv slabs + dense slab) + fp32 bias[COUT] at the end
    const float* residual;                 // (N, ld_res) fp32 or NULL
    int ld_res;
    float* ans;                            // (N, COUT_PAD) fp32
    void* x_out;                           // (N, COUT_PAD) fp16/bf16 = relu(ans), or NULL
    int n;                                 // total particles (rows of x_in)
    int begin, end;                        // rows computed by this launch
    int cout;                              // real output channels (<= COUT_PAD)
    int dense;                             // 1: the 17th slab is the dense (nn.Linear) branch on the particle's own row;
                                           // 0: plain ContinuousConv (operator-level entry point; in / out sets may differ)
    const float* mask_src;                 // backward: (N, ld_mask) fp32 pre-activations; the result is zeroed where <= 0
    int ld_mask;                           //           (ReLU backward), BEFORE the residual is added.  NULL: no mask
    int relu_out;                          // 1: x_out = relu(ans) (forward);  0: x_out = ans (backward: next gradient)
    const float4* order;                   // NULL, or the fluid grid's cell-sorted copy (.w = particle index): tile row t of the
                                           // launch is particle order[begin + t] -- a tile then holds 128 spatial neighbours whose
                                           // neighbour rows overlap (~300 distinct rows per tile: the gathers hit L1, not L2)
    int tile_rows;                         // 0 / 128: full tiles.  16..64: a CTA takes only that many rows, spread over all 16 worker
                                           // warps (row = r * 16 + warp): a rank of the sharded step has ~3,700 rows = 30 full tiles
                                           // on 148 SMs, and a CTA's time is set by the rows per WARP, not by the CTAs in flight
};

template <int CIN, int COUT_PAD>
struct ConvCfg {
    static constexpr int CPL = CIN / 32;                 // channels per lane
    static constexpr int KSLAB = 4 * CIN;                // columns of a conv slab
    static constexpr int KSTEPS = KSLAB / 16;
    static constexpr int KSTEPS_DENSE = CIN / 16;
    static constexpr int STEP_BYTES = COUT_PAD * 32;     // one K-step of the B operand
    static constexpr int SLAB_BYTES = KSTEPS * STEP_BYTES;
    static constexpr int DENSE_BYTES = KSTEPS_DENSE * STEP_BYTES;
    static constexpr int W_BYTES = 16 * SLAB_BYTES + DENSE_BYTES;
    static constexpr int PACKED_BYTES = W_BYTES + COUT_PAD * 4;
    static constexpr int SM_A = 0;                                   // 128 x KSLAB halves
    static constexpr int SM_W = SM_A + 128 * KSLAB * 2;
    static constexpr int SM_BIAS = SM_W + SLAB_BYTES;
    static constexpr int SM_OFFS = SM_BIAS + COUT_PAD * 4;           // 16 warps x 8 rows x 32 u16: slab list row starts
    static constexpr int SM_ROWMAP = SM_OFFS + 16 * 8 * 32 * 2;       // 128 ints: tile row -> particle
    static constexpr int SM_BAR = SM_ROWMAP + 128 * 4;
    static constexpr int SM_TOTAL = SM_BAR + 64;
    static_assert(SM_TOTAL <= 232448, "smem budget");
};

constexpr int WORKER_WARPS = 16;
constexpr int ROWS_PER_WARP = 128 / WORKER_WARPS;
constexpr int CONV_THREADS = WORKER_WARPS * 32 + 32;   // worker warps + 1 issuer warp

template <int CIN, int COUT_PAD, bool BF16>
__global__ void __launch_bounds__(CONV_THREADS, 1) k_cconv_tc(const ConvArgs a) {
    using C = ConvCfg<CIN, COUT_PAD>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_rows = (a.tile_rows > 0 && a.tile_rows < 128) ? a.tile_rows : 128;
    const bool spread = tile_rows < 128;
    const int row0 = a.begin + blockIdx.x * tile_rows;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_a = s_base + C::SM_A, s_w = s_base + C::SM_W, s_bar = s_base + C::SM_BAR;
    float* sbias = reinterpret_cast<float*>(smem + C::SM_BIAS);
    const uint32_t bar_a_ready = s_bar, bar_w_full = s_bar + 8, bar_mma_done = s_bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::SM_BAR + 32);

    if (threadIdx.x == 0) {
        mbar_init(bar_a_ready, WORKER_WARPS * 32);
        mbar_init(bar_w_full, 1);
        mbar_init(bar_mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < COUT_PAD) sbias[threadIdx.x] = __ldg(reinterpret_cast<const float*>(a.w_packed + C::W_BYTES) + threadIdx.x);
    constexpr uint32_t TMEM_COLS = COUT_PAD <= 64 ? 64 : 128;
    if (warp == WORKER_WARPS) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == WORKER_WARPS) {
        // ============================================================ issuer: weight slabs + MMAs
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, COUT_PAD, BF16);
            const uint8_t* src = a.w_packed;
            uint32_t acc = 0;
            for (int s = 0; s <= 16; ++s) {
                const uint32_t bytes = (s < 16) ? C::SLAB_BYTES : C::DENSE_BYTES;
                const int ksteps = (s < 16) ? C::KSTEPS : C::KSTEPS_DENSE;
                if (s > 0) mbar_wait(bar_mma_done, (s - 1) & 1);      // W buffer free again
                mbar_arrive_expect_tx(bar_w_full, bytes);
                bulk_g2s(s_w, src, bytes, bar_w_full);
                src += bytes;
                mbar_wait(bar_a_ready, s & 1);
                mbar_wait(bar_w_full, s & 1);
                tc_fence_after();
                for (int j = 0; j < ksteps; ++j) {
                    umma_f16(tmem_base, umma_desc(s_a + j * 4096, 2048, 128),
                             umma_desc(s_w + j * C::STEP_BYTES, COUT_PAD * 16, 128), idesc, acc);
                    acc = 1;
                }
                umma_commit(bar_mma_done);
            }
        }
    } else {
        // ============================================================ workers: slab construction
        // A warp owns 8 particles (rows of the tile).  Its lanes split into NG groups of GL lanes; a group works on ONE
        // particle at a time and each of its lanes owns CPL consecutive input channels (CIN = 64: 4 groups x 8 lanes x 8
        // channels, one 16-byte feature load per entry; CIN = 96: 2 groups x 16 lanes x 6 channels, three 4-byte loads).
        // For filter row s the group walks the particle's slab-list entries {j, weight per x cell}: every lane of the
        // group reads the same entry (a broadcast load, no shuffles), gathers its channels of neighbour j and adds
        // w[x] * f into acc[x][channel].  NG particles advance per warp instruction: ~10 (CIN 64) / ~18 (CIN 96) warp
        // instructions per entry instead of the ~40 of the lane-per-channel-pair version it replaces.
        constexpr int GL = (CIN == 64) ? 8 : 16;
        constexpr int CPL = CIN / GL;                      // 8 or 6 channels per lane
        constexpr int NG = 32 / GL;
        static_assert(CIN == 64 || CIN == 96, "channel mapping");
        const int rbase = warp * ROWS_PER_WARP;
        const int gq = lane / GL, cl = lane % GL;
        // row starts of this warp's 8 slab lists: smem [r][32] u16 (17 used)
        unsigned short* offs = reinterpret_cast<unsigned short*>(smem + C::SM_OFFS) + warp * ROWS_PER_WARP * 32;
        int* rowmap = reinterpret_cast<int*>(smem + C::SM_ROWMAP);      // tile row -> particle index (or -1)
        // tile row of this warp's r-th particle: consecutive rows (full tiles), or rows r * 16 + warp (short tiles)
        auto tile_row = [&](int r) { return spread ? r * WORKER_WARPS + warp : rbase + r; };
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int rl = tile_row(r);
            const int tpos = row0 + rl;
            int row = -1;
            if (rl < tile_rows && tpos < a.end) row = a.order ? __float_as_int(__ldg(&a.order[tpos].w)) : tpos;
            if (lane == 0) rowmap[rl] = row;
            offs[r * 32 + lane] = (row >= 0 && lane < 17) ? __ldg(a.slab_off + (size_t)row * SLABOFF + lane) : (unsigned short)0;
        }
        __syncwarp();
        const uint8_t* xin = reinterpret_cast<const uint8_t*>(a.x_in);
        auto cvt2 = [](uint32_t v, float& lo, float& hi) {
            if (BF16) { lo = __uint_as_float(v << 16); hi = __uint_as_float(v & 0xffff0000u); }
            else { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&v)); lo = t.x; hi = t.y; }
        };
        auto pack = [](float lo, float hi) -> uint32_t {
            if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
            __half2 h = __floats2half2_rn(lo, hi);
            return *reinterpret_cast<uint32_t*>(&h);
        };
        // CPL halves of neighbour j's feature row, as CPL/2 packed words
        auto load_feat = [&](int j, uint32_t (&f)[CPL / 2]) {
            const uint8_t* p = xin + (size_t)j * (CIN * 2) + cl * (CPL * 2);
            if constexpr (CPL == 8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
                f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < CPL / 2; ++i) f[i] = __ldg(reinterpret_cast<const uint32_t*>(p) + i);
            }
        };
        auto fma_feat = [&](const float4& w, const uint32_t (&f)[CPL / 2], float (&acc)[4][CPL]) {
#pragma unroll
            for (int i = 0; i < CPL / 2; ++i) {
                float f0, f1;
                cvt2(f[i], f0, f1);
                acc[0][2 * i] += w.x * f0; acc[1][2 * i] += w.y * f0; acc[2][2 * i] += w.z * f0; acc[3][2 * i] += w.w * f0;
                acc[0][2 * i + 1] += w.x * f1; acc[1][2 * i + 1] += w.y * f1; acc[2][2 * i + 1] += w.z * f1; acc[3][2 * i + 1] += w.w * f1;
            }
        };
#pragma unroll 1
        for (int s = 0; s <= 16; ++s) {
#pragma unroll 1
            for (int R = 0; R < ROWS_PER_WARP / NG; ++R) {
                const int r = R * NG + gq;
                const int rl = tile_row(r), row = rowmap[rl];
                if (spread && R > 0 && !__any_sync(NF_FULL, row >= 0)) continue;     // short tile: nothing in this iteration
                float acc[4][CPL];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int i = 0; i < CPL; ++i) acc[x][i] = 0.f;
                if (s < 16) {
                    const int beg = offs[r * 32 + s];
                    const int n = (int)offs[r * 32 + s + 1] - beg;
                    const int nmax = __reduce_max_sync(NF_FULL, n);
                    const size_t ebase = (size_t)max(row, 0) * SLABCAP + beg;       // no particle: n = 0, never dereferenced
                    // EB entries per iteration, the next iteration's {j, w} already in flight while this one's feature rows
                    // are gathered (the lists stream from L2 / HBM: one exposed round trip per iteration, not two).
                    // Entries past the group's own list read {j = 0, w = 0}: row 0 is an L1 hit.
                    constexpr int EB = 4;
                    int jn[EB];
                    float4 wn[EB];
                    auto fetch = [&](int e) {
#pragma unroll
                        for (int u = 0; u < EB; ++u) {
                            jn[u] = 0; wn[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (e + u < n) { jn[u] = __ldg(a.slab_j + ebase + e + u); wn[u] = __ldg(a.slab_w + ebase + e + u); }
                        }
                    };
                    fetch(0);
#pragma unroll 1
                    for (int e = 0; e < nmax; e += EB) {
                        int jc[EB];
                        float4 wc[EB];
                        uint32_t f[EB][CPL / 2];
#pragma unroll
                        for (int u = 0; u < EB; ++u) { jc[u] = jn[u]; wc[u] = wn[u]; }
#pragma unroll
                        for (int u = 0; u < EB; ++u) load_feat(jc[u], f[u]);
                        fetch(e + EB);
#pragma unroll
                        for (int u = 0; u < EB; ++u) fma_feat(wc[u], f[u], acc);
                    }
                } else if (row >= 0 && a.dense) {
                    // dense branch: the particle's own (ReLU'd) features, K = CIN
                    uint32_t f[CPL / 2];
                    load_feat(row, f);
#pragma unroll
                    for (int i = 0; i < CPL / 2; ++i) cvt2(f[i], acc[0][2 * i], acc[0][2 * i + 1]);
                }
                if (R == 0 && s > 0) mbar_wait(bar_mma_done, (s - 1) & 1);   // previous slab consumed
                const int nx = (s < 16) ? 4 : 1;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (x < nx) {
                        const int k = x * CIN + cl * CPL;                     // first of this lane's CPL slab columns
                        if constexpr (CPL == 8) {
                            const uint32_t addr = s_a + (uint32_t)(k >> 3) * 2048 + (uint32_t)rl * 16;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack(acc[x][0], acc[x][1])),
                                         "r"(pack(acc[x][2], acc[x][3])), "r"(pack(acc[x][4], acc[x][5])), "r"(pack(acc[x][6], acc[x][7]))
                                         : "memory");
                        } else {
#pragma unroll
                            for (int i = 0; i < CPL / 2; ++i) {
                                const int kk = k + 2 * i;
                                const uint32_t addr = s_a + (uint32_t)(kk >> 3) * 2048 + (uint32_t)rl * 16 + (uint32_t)(kk & 7) * 2;
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack(acc[x][2 * i], acc[x][2 * i + 1])) : "memory");
                            }
                        }
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_a_ready);
        }
        // ============================================================ epilogue (warps 0-3, thread = row)
        if (warp < 4) {
            mbar_wait(bar_mma_done, 0);     // 17 commits: the last one completes phase index 16 -> parity 0
            tc_fence_after();
            const int rl = warp * 32 + lane;
            const int row = rowmap[rl];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c0 = 0; c0 < COUT_PAD; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (row >= 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = c0 + i;
                        float o = __uint_as_float(v[i]) + sbias[c];
                        if (a.mask_src && !(a.mask_src[(size_t)row * a.ld_mask + c] > 0.f)) o = 0.f;
                        if (a.residual && c < a.cout) o += a.residual[(size_t)row * a.ld_res + c];
                        if (c >= a.cout) o = 0.f;
                        a.ans[(size_t)row * COUT_PAD + c] = o;
                        if (a.x_out) {
                            const float xo = a.relu_out ? fmaxf(o, 0.f) : o;
                            if (BF16) reinterpret_cast<__nv_bfloat16*>(a.x_out)[(size_t)row * COUT_PAD + c] = __float2bfloat16(xo);
                            else reinterpret_cast<__half*>(a.x_out)[(size_t)row * COUT_PAD + c] = __float2half(xo);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WORKER_WARPS) tmem_dealloc(tmem_base, TMEM_COLS);
}


constexpr int C3_IN = 64, C3_OUT = 3, C3_G = NCELL * C3_OUT;      // 192 projected values per particle

// ------------------------------------------------------------------------------------------------
// weight packing: conv kernel (4,4,4,CIN,COUT) + dense (COUT,CIN) -> K-step slabs in UMMA order
//   conv slab s=(z*4+y): column k = x*CIN + ch   <-  kernel[z][y][x][ch][cout]
//   dense slab         : column k = ch           <-  dense_w[cout][ch]
//   K-step bytes: [kc(2)][cout(COUT_PAD)][e(8)] halves
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT_PAD, bool BF16>
__global__ void k_pack_conv(const float* __restrict__ kern, const float* __restrict__ bconv, const float* __restrict__ wd,
                            const float* __restrict__ bd, int cout, uint8_t* __restrict__ out) {
    using C = ConvCfg<CIN, COUT_PAD>;
    const int total_steps = 16 * C::KSTEPS + C::KSTEPS_DENSE;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;    // one thread per (step, kc, cout)
    if (t < total_steps * 2 * COUT_PAD) {
        const int step = t / (2 * COUT_PAD), kc = (t / COUT_PAD) % 2, n = t % COUT_PAD;
        unsigned short e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (n < cout) {
                if (step < 16 * C::KSTEPS) {
                    const int s = step / C::KSTEPS, k = (step % C::KSTEPS) * 16 + kc * 8 + i;
                    const int x = k / CIN, ch = k % CIN;
                    v = kern[((size_t)(s * 4 + x) * CIN + ch) * cout + n];
                } else {
                    const int k = (step - 16 * C::KSTEPS) * 16 + kc * 8 + i;
                    v = wd ? wd[(size_t)n * CIN + k] : 0.f;
                }
            }
            if (BF16) { __nv_bfloat16 h = __float2bfloat16(v); e[i] = *reinterpret_cast<unsigned short*>(&h); }
            else { __half h = __float2half(v); e[i] = *reinterpret_cast<unsigned short*>(&h); }
        }
        uint4 pk;
        pk.x = e[0] | ((unsigned)e[1] << 16); pk.y = e[2] | ((unsigned)e[3] << 16);
        pk.z = e[4] | ((unsigned)e[5] << 16); pk.w = e[6] | ((unsigned)e[7] << 16);
        *reinterpret_cast<uint4*>(out + (size_t)step * C::STEP_BYTES + ((size_t)kc * COUT_PAD + n) * 16) = pk;
    }
    if (t < COUT_PAD) reinterpret_cast<float*>(out + C::W_BYTES)[t] = t < cout ? (bconv ? bconv[t] : 0.f) + (bd ? bd[t] : 0.f) : 0.f;
}

// packed layout of the whole ParticleNet: fp32 layer-0 tensors, then the three tensor-core layers
struct PackedLayout {
    size_t k_fluid, b_fluid, k_obst, b_obst, w_dense0, b_dense0, l1, l2, k3, b3, w_dense3, b_dense3, total;
};
inline PackedLayout packed_layout() {
    PackedLayout L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    L.k_fluid = take(64 * 4 * 32 * 4); L.b_fluid = take(32 * 4);
    L.k_obst = take(64 * 3 * 32 * 4); L.b_obst = take(32 * 4);
    L.w_dense0 = take(32 * 4 * 4); L.b_dense0 = take(32 * 4);
    L.l1 = take(ConvCfg<96, 64>::PACKED_BYTES);
    L.l2 = take(ConvCfg<64, 64>::PACKED_BYTES);
    L.k3 = take(NCELL * 64 * 3 * 4); L.b3 = take(3 * 4);          // conv3 / dense3 stay fp32 (k_conv3_*)
    L.w_dense3 = take(3 * 64 * 4); L.b_dense3 = take(3 * 4);
    L.total = o;
    return L;
}

struct WsLayout {
    size_t pos_new, vel_new, grid_f, grid_b, pairs_ff, cnt_ff, pairs_fb, cnt_fb, slab_j, slab_w, slab_off, ans0, x0, ans1, x1, ans2, x2, ans3,
        g3, flags, out10, total;
};
constexpr int ROW_PAD = 8;      // row arrays that are all-gathered in place hold world * ceil(n / world) <= n + 7 rows (world <= 8)
inline WsLayout ws_layout(int n, int m) {
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1) + ROW_PAD;
    L.pos_new = take(N * 12); L.vel_new = take(N * 12);
    L.grid_f = take(grid_layout(n).total);
    L.grid_b = take(grid_layout(m).total);
    L.pairs_ff = take(N * MAXNBR * sizeof(Pair)); L.cnt_ff = take(N * 4);
    L.pairs_fb = take(N * MAXNBR * sizeof(Pair)); L.cnt_fb = take(N * 4);
    L.slab_j = take(N * SLABCAP * 4); L.slab_w = take(N * SLABCAP * 16); L.slab_off = take(N * SLABOFF * 2);
    L.ans0 = take(N * 96 * 4); L.x0 = take(N * 96 * 2);
    L.ans1 = take(N * 64 * 4); L.x1 = take(N * 64 * 2);
    L.ans2 = take(N * 64 * 4); L.x2 = take(N * 64 * 2);
    L.ans3 = take(N * 16 * 4);
    L.g3 = take(N * C3_G * 4);
    L.flags = take(256);
    L.out10 = take(N * 10 * 4);
    L.total = o;
    return L;
}

// sharded step: a rank's rows of (pos_out, vel_out, count, delta) packed 10 floats wide for ONE all-gather, then unpacked

template <int CIN, int COUT_PAD>
inline int launch_conv(const ConvArgs& a, int dtype, cudaStream_t st) {
    using C = ConvCfg<CIN, COUT_PAD>;
    if (a.end <= a.begin) return NF_OK;
    const int tr = (a.tile_rows > 0 && a.tile_rows < 128) ? a.tile_rows : 128;
    const int grid = (a.end - a.begin + tr - 1) / tr;
    // the attribute is per device, not per process: set it on every launch instead of caching a flag
    if (dtype == NF_DTYPE_BF16) {
        NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_tc<CIN, COUT_PAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SM_TOTAL));
        k_cconv_tc<CIN, COUT_PAD, true><<<grid, CONV_THREADS, C::SM_TOTAL, st>>>(a);
    } else {
        NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_tc<CIN, COUT_PAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SM_TOTAL));
        k_cconv_tc<CIN, COUT_PAD, false><<<grid, CONV_THREADS, C::SM_TOTAL, st>>>(a);
    }
    NF_LAUNCH_OK();
    return NF_OK;
}

}  // namespace cconv
}  // namespace nf
