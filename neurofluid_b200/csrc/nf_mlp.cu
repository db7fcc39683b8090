// nf_mlp.cu -- weight packing and the C entry points of the fused positional-encoding + NeRF MLP; the production kernel is
// k_nerf_mlp2 (nf_mlp2.cu, two tiles per CTA).
//
// replaces: Embedding.forward x6 (models/nerf.py:21-38 via models/renderer.py:125-179) and
//           NeRF.forward (models/nerf.py:83-124; two instances, models/renderer.py:43-44).
//
// This file also keeps the round-1/2 ONE-tile-per-CTA kernel k_nerf_mlp, compiled into the tuning build only
// (-DNF_TUNING, NF_MLP_IMPL=1; NF_MLP_CLUSTER=4 adds its 4-CTA weight-multicast flavour): it is the comparison the
// two-tile kernel is measured against (tests/gpu_mlp_power.py, profiles/r02_notes.md).  Its layout: 14 warps --
// 0-7 epilogue (two groups of four, pipelined tcgen05.ld one chunk ahead), 8 MMA issuer (rank 0) / weight relay (rank 1),
// 9 weight loader (4-stage x 16 KB ring of weight units), 10-13 encoding producers writing the A tiles straight into
// shared memory; accumulators double-buffered in TMEM by layer parity.
#include <type_traits>

#include <stdlib.h>

#include "nf_common.cuh"
#include "nf_mlp.cuh"
#include "nf_tc.cuh"

namespace nf {
namespace mlp {

#ifdef NF_TUNING
// Weight ring: one stage holds one weight UNIT (nf_mlp.cuh): up to 8 consecutive K-steps of one N-HALF of a layer
// (128 of its 256 output features; 64 of 128 for the dir layer), one CTA's 64 (32) rows of it = 16 KB.  The issuer
// waits once and commits once per unit, and runs each layer as [half 0, K low] [half 1, K low] [half 0, K high]
// [half 1, K high]: half 0's accumulator is complete -- and its epilogue running -- while the tensor pipe still works
// on half 1, and the next layer's low-K units only need the activations half 0's epilogue produces.
constexpr int NSTAGE = 4;
constexpr int STAGE = 16384;
constexpr int RING_BYTES = NSTAGE * STAGE;

// shared memory map
constexpr int SM_HIDDEN = 0;                          // 128 x 256 halves
constexpr int SM_PEXYZ = SM_HIDDEN + 65536;           // 128 x 208 halves
constexpr int SM_PEDIR = SM_PEXYZ + 26 * 2048;        // 2 x (128 x 64 halves)
constexpr int SM_WRING = SM_PEDIR + 2 * 8 * 2048;     // NSTAGE x 16 KB
constexpr int SM_SPARAM = SM_WRING + RING_BYTES;
constexpr int SM_PART = SM_SPARAM + SP_FLOATS * 4;    // 128 x float4: head partial sums of epilogue group B
constexpr int SM_BAR = SM_PART + 128 * 16;
constexpr int NUM_BARS = 2 * NSTAGE + 14;
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
    B_ACC_FULL = B_ACT_READY + 4,      // 4: [accumulator buffer (layer parity)][N-half]
};
static_assert(B_ACC_FULL + 4 == NUM_BARS, "barrier count");

// warp roles
constexpr int W_EPI = 0;        // warps 0-7: epilogue, two groups of four (TMEM lane quarter = warp & 3)
constexpr int W_ISSUE = 8;      // MMA issuer (rank 0) / weight relay (rank 1 of a pair)
constexpr int W_LOAD = 9;       // weight producer
constexpr int W_PE = 10;        // warps 10-13: positional-encoding producers
constexpr int NUM_THREADS = 14 * 32;

// number of weight units (ring stages) one tile consumes
__device__ __forceinline__ int weight_units(int nl) {
    int n = 0;
    for (int l = 0; l < nl; ++l) n += 2 * ((layer_pe_steps(l) + wu_ksteps(0) - 1) / wu_ksteps(0) + (l > 0 ? KH_STEPS / wu_ksteps(1) : 0));
    return n;
}

#define NF_TRACE(slot) do { if (tracing) a.trace[(slot)] = clock64(); } while (0)

// PAIR = two CTAs on the two SMs of a TPC run as one unit (cluster of 2, tcgen05 cta_group::2): each CTA owns a
// 128-row tile (A operand, TMEM accumulator, PE producers, epilogue) and streams only HALF of every weight slab
// (its half of the N output rows of B), so the L2 -> SM weight traffic per FLOP halves.  CTA rank 0 issues every
// MMA; "operand ready" barriers live in rank 0 and collect one arrive per producing warp of both CTAs (remote
// arrives through the cluster window); "accumulator full" / "slot free" commits are multicast to both CTAs.
//
// The issuer warp runs converged and elects one lane around the tcgen05 instructions only, so that descriptors
// and barrier addresses stay in uniform registers; it waits and commits once per GROUP of K-steps (one ring
// stage), not per MMA.  (Round-1 build: one divergent lane, wait + commit per MMA = ~70 SASS instructions
// and ~360 cycles per 178-cycle MMA -- the kernel was issue-bound, profiles/r01_notes.md.)
// CL = cluster size: 2 = one CTA pair; 4 = two pairs that share every weight unit: each CTA fetches a QUARTER of a unit's
// rows and multicasts it to the CTA of the other pair that needs the same rows, halving the L2 -> SM weight traffic per
// FLOP once more (the kernel sat at the ~6.9 TB/s L2 -> SM ceiling: every pair streamed the whole 1.34 MB network per
// 256 rows).  Both pairs then walk the ring in lockstep: a stage is free when BOTH pairs' MMAs have read it (their
// commits are multicast to all four CTAs), so every pair of the grid runs the same number of passes (a pair without a
// real tile computes an empty one).
template <bool BF16, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_nerf_mlp(const KernelArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    static_assert(CL == 2 || CL == 4, "cluster of one or two CTA pairs");
    constexpr bool PAIR = true;
    constexpr int TPU = 2;              // tiles per unit (CTA pair) per pass
    const int warp = uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
    const uint32_t crank = uniform(cluster_ctarank());
    const uint32_t prank = crank & 1u;                          // rank inside the pair
    const uint32_t pbase = crank & ~1u;                         // cluster rank of the pair's first CTA (its MMA issuer)
    const uint16_t pair_mask = (uint16_t)(3u << pbase);         // commits that concern this pair only
    const uint16_t all_mask = (uint16_t)((1u << CL) - 1u);      // "weight stage consumed": every CTA that shares the ring
    const int unit = (int)(blockIdx.x >> 1);
    const int nunits = (int)(gridDim.x >> 1);
    const int n_rows = uniform(a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows_cap) : a.n_rows_host);
    const int ntiles = (n_rows + TILE_M - 1) / TILE_M;
    // every pair runs the same number of passes (CL = 4: the two pairs of a cluster share the weight ring); a pass beyond
    // the last tile pair works on rows >= n_rows: zero records in, nothing written
    const int npass = ((ntiles + TPU - 1) / TPU + nunits - 1) / nunits * nunits;
    if (ntiles == 0) return;
    const int nl = a.n_layers;
#ifdef NF_TUNING
    const bool no_weights = (a.desc_swap & 16) != 0;   // tuning builds only: do not stream / wait for weights (timing only)
#else
    constexpr bool no_weights = false;
#endif

    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_hidden = s_base + SM_HIDDEN, s_pexyz = s_base + SM_PEXYZ, s_pedir = s_base + SM_PEDIR;
    const uint32_t s_wring = s_base + SM_WRING, s_bar = s_base + SM_BAR;
    float* sp = reinterpret_cast<float*>(smem + SM_SPARAM);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART);
    auto bar = [&](int i) { return s_bar + 8u * (uint32_t)i; };
    // "operand ready" barriers are consumed by the issuer in CTA rank 0
    auto ready_bar = [&](int i) { return mapa_rank(bar(i), pbase); };
    constexpr uint32_t READY_COUNT = PAIR ? 8 : 4;   // PE tiles: one arrive per producing warp (4 per CTA)
    constexpr uint32_t ACT_COUNT = PAIR ? 16 : 8;    // activation chunks: all 8 epilogue warps of a CTA

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bar(B_WFULL + i), prank == 0 ? 2 : 1);   // own expect_tx arrive (+ the pair peer's relay)
            mbar_init(bar(B_WEMPTY + i), CL / 2);              // one commit per pair sharing the ring
        }
        mbar_init(bar(B_PEXYZ_READY), READY_COUNT);
        mbar_init(bar(B_PEXYZ_FREE), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_PEDIR_READY + i), READY_COUNT);
            mbar_init(bar(B_PEDIR_FREE + i), 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_ACC_FULL + i), 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_ACT_READY + i), ACT_COUNT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // small params -> smem
        const float4* src = reinterpret_cast<const float4*>(a.packed + W_BYTES);
        float4* dst = reinterpret_cast<float4*>(sp);
        for (int i = threadIdx.x; i < SP_FLOATS / 4; i += NUM_THREADS) dst[i] = __ldg(src + i);
    }
    if (warp == W_ISSUE) tmem_alloc<PAIR>(s_base + SM_TMEM_SLOT, 512);
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = uniform(*reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_SLOT));

    if (warp == W_ISSUE) {
        if (prank == 0) {
            // ================================================================ MMA issuer (rank 0; converged warp)
            const uint64_t adesc_hidden = umma_desc(s_hidden, 2048u, 128u);
            const uint64_t adesc_pexyz = umma_desc(s_pexyz, 2048u, 128u);
            const uint64_t adesc_pedir = umma_desc(s_pedir, 2048u, 128u);
            const uint32_t idesc128 = umma_idesc(128, BF16, 2 * TILE_M);
            const uint32_t idesc64 = umma_idesc(64, BF16, 2 * TILE_M);
            uint32_t ws = 0, wph = 0, hidw = 0, lc = 0;
            int ti = 0;
            for (int pass = unit; pass < npass; pass += nunits, ++ti) {
                const bool tracing = a.trace && blockIdx.x == 0 && ti == 2 && lane == 0;
                for (int l = 0; l < nl; ++l, ++lc) {
                    const bool n128 = (l == 9);
                    const uint32_t nhalf = n128 ? 64u : 128u;                  // N of one MMA = one N-half of the layer
                    const uint32_t rpc = nhalf >> 1;                           // rows of B each CTA of the pair supplies
                    const uint32_t d_tmem = tmem_base + (lc & 1) * 256;
                    const uint32_t idesc = n128 ? idesc64 : idesc128;
                    const uint64_t bdesc0 = umma_desc(s_wring, rpc * 16u, 128u);
                    const uint32_t bstep = (2u * rpc * 16u) >> 4;              // descriptor units (16 B) per K-step of a unit
                    uint32_t acc0 = 0, acc1 = 0;
                    NF_TRACE(100 + l * 8);
                    const int npe = layer_pe_steps(l);
                    for (int seg = 0; seg < 2; ++seg) {
                        const int nsteps = seg == 0 ? npe : (l > 0 ? KH_STEPS : 0);
                        if (nsteps == 0) continue;
                        uint64_t adesc = adesc_hidden;
                        if (seg == 0) {
                            if (l == 9) {
                                mbar_wait(bar(B_PEDIR_READY + (ti & 1)), (ti >> 1) & 1);
                                adesc = adesc_pedir + (uint64_t)((ti & 1) * ((8 * 2048) >> 4));
                            } else {
                                mbar_wait(bar(B_PEXYZ_READY), ti & 1);
                                adesc = adesc_pexyz;
                            }
                        }
                        const bool last_seg = (seg == 1) || (l == 0);
                        const int kb = wu_ksteps(seg);
                        for (int k0 = 0; k0 < nsteps; k0 += kb) {
                            const int g = min(kb, nsteps - k0);
                            if (seg == 1) {      // the 8 K-steps of this block read activation chunks k0/4 and k0/4 + 1
                                mbar_wait(bar(B_ACT_READY + (k0 >> 2)), hidw & 1);
                                NF_TRACE(100 + l * 8 + 1 + (k0 >> 2));
                                mbar_wait(bar(B_ACT_READY + (k0 >> 2) + 1), hidw & 1);
                                NF_TRACE(100 + l * 8 + 2 + (k0 >> 2));
                            }
                            const bool last_blk = last_seg && (k0 + kb >= nsteps);
#pragma unroll
                            for (uint32_t nh = 0; nh < 2; ++nh) {
                                if (!no_weights) mbar_wait(bar(B_WFULL + ws), wph);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint64_t bd = bdesc0 + (uint64_t)(ws * (STAGE >> 4));
                                    uint32_t acc = nh ? acc1 : acc0;
#pragma unroll
                                    for (int j = 0; j < WU_KSTEPS; ++j) {
                                        if (j < g) {
                                            umma_f16<PAIR>(d_tmem + nh * nhalf, adesc + (uint64_t)((k0 + j) * (4096 >> 4)), bd + (uint64_t)(j * bstep),
                                                           idesc, acc);
                                            acc = 1;
                                        }
                                    }
                                    umma_commit<PAIR>(bar(B_WEMPTY + ws), all_mask);
                                    if (last_blk) umma_commit<PAIR>(bar(B_ACC_FULL + (lc & 1) * 2 + nh), pair_mask);
                                }
                                __syncwarp();
                                if (nh) acc1 = 1; else acc0 = 1;
                                if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                            }
                        }
                        if (seg == 0 && l == 4) {
                            if (elect_one()) umma_commit<PAIR>(bar(B_PEXYZ_FREE), pair_mask);
                            __syncwarp();
                        }
                    }
                    if (l > 0) ++hidw;
                    if (l == 9) {
                        if (elect_one()) umma_commit<PAIR>(bar(B_PEDIR_FREE + (ti & 1)), pair_mask);
                        __syncwarp();
                    }
                    NF_TRACE(100 + l * 8 + 5);
                }
            }
        } else if (PAIR && lane == 0 && !no_weights) {
            // ================================================================ relay (rank 1): tells the issuer
            // that this CTA's share of a weight group has landed
            uint32_t ws = 0, wph = 0;
            const int ngroups = weight_units(nl);
            const uint32_t remote0 = mapa_rank(bar(B_WFULL), pbase);
            for (int pass = unit; pass < npass; pass += nunits) {
                for (int s = 0; s < ngroups; ++s) {
                    mbar_wait(bar(B_WFULL + ws), wph);
                    mbar_arrive_cluster(remote0 + 8u * ws);
                    if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == W_LOAD) {
        // ================================================================ weight producer
        if (lane == 0 && !no_weights) {
            uint32_t ws = 0, wph = 0;
            for (int pass = unit; pass < npass; pass += nunits) {
                const uint8_t* src = a.packed;
                for (int l = 0; l < nl; ++l) {
                    const uint32_t rpc = (l == 9) ? 32u : 64u;
                    const int npe = layer_pe_steps(l);
                    for (int seg = 0; seg < 2; ++seg) {
                        const int nsteps = seg == 0 ? npe : (l > 0 ? KH_STEPS : 0);
                        const int kb = wu_ksteps(seg);
                        for (int k0 = 0; k0 < nsteps; k0 += kb) {
                            const uint32_t mine = (uint32_t)min(kb, nsteps - k0) * 2u * rpc * 16u;   // this CTA's rows of the unit
                            for (int nh = 0; nh < 2; ++nh) {
                                mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                                mbar_arrive_expect_tx(bar(B_WFULL + ws), mine);
                                if constexpr (CL == 4) {       // my quarter -> me and the CTA with my pair rank in the other pair
                                    const uint32_t q = mine >> 1, sub = crank >> 1;
                                    bulk_g2s_multicast(s_wring + ws * STAGE + sub * q, src + prank * mine + sub * q, q, bar(B_WFULL + ws),
                                                       (uint16_t)((1u << prank) | (1u << (prank + 2))));
                                } else {
                                    bulk_g2s(s_wring + ws * STAGE, src + prank * mine, mine, bar(B_WFULL + ws));
                                }
                                src += 2 * mine;
                                if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                            }
                        }
                    }
                }
            }
            if constexpr (PAIR) {   // every multicast "slot free" arrive has landed before this CTA may exit
                for (int i = 0; i < NSTAGE; ++i) {
                    mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                    if (++ws == NSTAGE) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp < W_ISSUE) {
        // ================================================================ epilogue: thread = (row, column half)
        // both groups drain every 64-column chunk together: group g (warps 4g..4g+3) owns columns [64c+32g, +32)
        // of chunk c; the TMEM load of chunk c+1 is in flight while chunk c is processed.
        const int grp = warp >> 2;
        const int tr = (warp & 3) * 32 + lane;   // row of the tile = TMEM lane
        uint32_t lc = 0;
        int ti = 0;
        const uint32_t act_ready0 = ready_bar(B_ACT_READY);
        for (int pass = unit; pass < npass; pass += nunits, ++ti) {
            const bool tracing = a.trace && blockIdx.x == 0 && ti == 2 && (tr == 0);
            const int row = (pass * TPU + (int)prank) * TILE_M + tr;
            float sigma = 0.f;
            for (int l = 0; l < nl; ++l, ++lc) {
                const uint32_t buf = lc & 1;
                mbar_wait(bar(B_ACC_FULL + buf * 2), (lc >> 1) & 1);          // N-half 0 complete (half 1 still accumulating)
                tc_fence_after();
                if (grp == 0) NF_TRACE(l * 8);
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 256 + grp * 32;
                if (l < 9) {
                    const bool writes = (l + 1 < nl);
                    const float* bias = sp + SP_BIAS + l * 256 + grp * 32;
                    uint32_t v[2][32];
                    tmem_ld32(taddr, v[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tmem_ld_wait();
                        if (c == 1) {       // columns 128.. belong to N-half 1
                            mbar_wait(bar(B_ACC_FULL + buf * 2 + 1), (lc >> 1) & 1);
                            tc_fence_after();
                        }
                        if (c < 3) tmem_ld32(taddr + (c + 1) * 64, v[(c + 1) & 1]);
                        const uint32_t* vc = v[c & 1];
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 64 + i);
                            f[i] = __uint_as_float(vc[i]) + b4.x;
                            f[i + 1] = __uint_as_float(vc[i + 1]) + b4.y;
                            f[i + 2] = __uint_as_float(vc[i + 2]) + b4.z;
                            f[i + 3] = __uint_as_float(vc[i + 3]) + b4.w;
                        }
                        if (l != 8) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                        }
                        if (l == 7) {
                            const float* wsig = sp + SP_WSIG + c * 64 + grp * 32;
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 w4 = *reinterpret_cast<const float4*>(wsig + i);
                                sigma = fmaf(f[i], w4.x, sigma);
                                sigma = fmaf(f[i + 1], w4.y, sigma);
                                sigma = fmaf(f[i + 2], w4.z, sigma);
                                sigma = fmaf(f[i + 3], w4.w, sigma);
                            }
                        }
                        if (writes) {
                            const uint32_t dst = s_hidden + (uint32_t)(c * 8 + grp * 4) * 2048 + (uint32_t)tr * 16;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                st_shared_v4(dst + q * 2048, pack2<BF16>(f[8 * q], f[8 * q + 1]),
                                             pack2<BF16>(f[8 * q + 2], f[8 * q + 3]),
                                             pack2<BF16>(f[8 * q + 4], f[8 * q + 5]),
                                             pack2<BF16>(f[8 * q + 6], f[8 * q + 7]));
                            fence_proxy_async();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(act_ready0 + 8u * c);
                        }
                        if (grp == 0) NF_TRACE(l * 8 + 1 + c);
                    }
                    if (l == 7 && !writes) {  // sigma-only network: combine the two column halves and emit
                        tc_fence_before();
                        if (grp == 1) part[tr] = make_float4(0.f, 0.f, 0.f, sigma);
                        named_bar_sync(1, 256);
                        if (grp == 0 && row < n_rows) {
                            const int dst = a.rowid ? a.rowid[row] : row;
                            if (dst >= 0) a.out4[dst] = make_float4(0.f, 0.f, 0.f, sigma + part[tr].w + sp[SP_BSIG]);
                        }
                        named_bar_sync(2, 256);   // part[] may be rewritten by the next tile only after it was read
                    }
                } else {
                    // rgb head on the 128-wide dir layer: group g owns columns [64c + 32g, +32), c = 0, 1
                    float rgb[3] = {0.f, 0.f, 0.f};
                    const float* bias = sp + SP_BIAS + 9 * 256 + grp * 32;
                    const float* wrgb = sp + SP_WRGB + grp * 32;
                    uint32_t v[2][32];
                    tmem_ld32(taddr, v[0]);
                    mbar_wait(bar(B_ACC_FULL + buf * 2 + 1), (lc >> 1) & 1);   // the dir layer's halves are 64 columns each
                    tc_fence_after();
                    tmem_ld32(taddr + 64, v[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float f = fmaxf(__uint_as_float(v[c][i]) + bias[c * 64 + i], 0.f);
                            rgb[0] = fmaf(f, wrgb[c * 64 + i], rgb[0]);
                            rgb[1] = fmaf(f, wrgb[128 + c * 64 + i], rgb[1]);
                            rgb[2] = fmaf(f, wrgb[256 + c * 64 + i], rgb[2]);
                        }
                    }
                    tc_fence_before();
                    if (grp == 1) part[tr] = make_float4(rgb[0], rgb[1], rgb[2], sigma);
                    named_bar_sync(1, 256);
                    if (grp == 0 && row < n_rows) {
                        const int dst = a.rowid ? a.rowid[row] : row;
                        if (dst >= 0) {
                            const float4 pb = part[tr];
                            float4 o;
                            o.x = 1.0f / (1.0f + expf(-(rgb[0] + pb.x + sp[SP_BRGB])));
                            o.y = 1.0f / (1.0f + expf(-(rgb[1] + pb.y + sp[SP_BRGB + 1])));
                            o.z = 1.0f / (1.0f + expf(-(rgb[2] + pb.z + sp[SP_BRGB + 2])));
                            o.w = sigma + pb.w + sp[SP_BSIG];
                            a.out4[dst] = o;
                        }
                    }
                    named_bar_sync(2, 256);
                }
            }
        }
    } else {
        // ================================================================ PE producers (warps 10-13)
        const int tp = threadIdx.x - W_PE * 32;  // 0..127
        int ti = 0;
        const uint32_t pexyz_ready = ready_bar(B_PEXYZ_READY);
        const uint32_t pedir_ready0 = ready_bar(B_PEDIR_READY), pedir_ready1 = ready_bar(B_PEDIR_READY + 1);
        for (int pass = unit; pass < npass; pass += nunits, ++ti) {
            const bool tracing = a.trace && blockIdx.x == 0 && ti == 2 && tp == 0;
            const int row = (pass * TPU + (int)prank) * TILE_M + tp;
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
            NF_TRACE(300);
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
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(pexyz_ready);
            NF_TRACE(301);
            if (nl == 10) {
                // dir-like block: [PE4(ray dir) 27 | PE4(smoothed dir) 27 | 0 x10]
                mbar_wait(bar(B_PEDIR_FREE + (ti & 1)), ((ti >> 1) & 1) ^ 1);
                RowWriter<BF16> w;
                w.base = s_pedir + (ti & 1) * (8 * 2048) + (uint32_t)tp * 16;
                emit_encoding<BF16, 0, 3, 4>(w, r + 10);
                emit_encoding<BF16, 27, 3, 4>(w, r + 13);
                static_for<54, 64>([&](auto ci) { w.template put<decltype(ci)::value>(0.f); });
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster((ti & 1) ? pedir_ready1 : pedir_ready0);
                NF_TRACE(302);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
    if (warp == W_ISSUE) tmem_dealloc<PAIR>(tmem_base, 512);
}

#endif  // NF_TUNING (one-tile kernel)

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

// byte offset of weight unit (layer, segment, K-block k0, N-half nh) in the packed stream (see nf_mlp.cuh)
__device__ __forceinline__ size_t unit_offset(int layer, int seg, int k0, int nh) {
    size_t off = 0;
    for (int l = 0; l <= layer; ++l) {
        const int rpc = (l == 9) ? 32 : 64;
        for (int sg = 0; sg < 2; ++sg) {
            const int nsteps = sg == 0 ? layer_pe_steps(l) : (l > 0 ? KH_STEPS : 0);
            for (int kb = 0; kb < nsteps; kb += wu_ksteps(sg)) {
                const int g = min(wu_ksteps(sg), nsteps - kb);
                for (int h = 0; h < 2; ++h) {
                    if (l == layer && sg == seg && kb == k0 && h == nh) return off;
                    off += (size_t)2 * g * 2 * rpc * 16;
                }
            }
        }
    }
    return off;
}

template <bool BF16>
__global__ void k_pack_weights(PackArgs p, uint8_t* out, int enc) {
    // one thread per (K-step, kc, n): writes 8 halves (16 B)
    const int total = N256_STEPS * 2 * 256 + N128_STEPS * 2 * 128;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) {
        int s, kc, n;
        if (t < N256_STEPS * 512) {
            s = t / 512; kc = (t % 512) / 256; n = t % 256;
        } else {
            const int u = t - N256_STEPS * 512;
            s = N256_STEPS + u / 256; kc = (u % 256) / 128; n = u % 128;
        }
        int layer, k0c, src_off, src_valid, ld;
        step_source(s, layer, k0c, src_off, src_valid, ld);
        const float* W = p.w[layer];
        // source column of fixed-layout column k of this segment (encoding ablations narrow the network's own inputs)
        const int in_xyz = enc_in_xyz(enc), in_dir = enc_in_dir(enc);
        const bool seg_xyz = (src_valid == 198), seg_dir = (src_valid == 54);
        const int ld_e = layer == 0 ? in_xyz : (layer == 4 ? in_xyz + 256 : (layer == 9 ? 256 + in_dir : ld));
        const int off_e = (layer == 4 && !seg_xyz) ? in_xyz : src_off;       // skip layer: [input_xyz | h]; dir layer: [final | input_dir]
        auto src_col = [&](int k) -> int {
            if (seg_xyz) return enc_col_xyz(k, enc);
            if (seg_dir) { const int c = enc_col_dir(k, enc); return c < 0 ? -1 : 256 + c; }
            return k < src_valid ? off_e + k : -1;
        };
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const int ca = src_col(k0c + kc * 8 + e), cb = src_col(k0c + kc * 8 + e + 1);
            const float va = ca >= 0 ? W[(size_t)n * ld_e + ca] : 0.f;
            const float vb = cb >= 0 ? W[(size_t)n * ld_e + cb] : 0.f;
            pk[e >> 1] = pack2<BF16>(va, vb);
        }
        // destination: the unit of (layer, segment, K-block, N-half) this (K-step, row) belongs to
        const int seg = (layer_pe_steps(layer) > 0 && src_valid != 256) ? 0 : 1;      // encoded-feature segments are 198 / 54 wide
        const int kseg = k0c / 16;
        const int rpc = (layer == 9) ? 32 : 64;
        const int kblk = (kseg / wu_ksteps(seg)) * wu_ksteps(seg);
        const int nsteps = seg == 0 ? layer_pe_steps(layer) : KH_STEPS;
        const int g = min(wu_ksteps(seg), nsteps - kblk);
        const int nh = n / (2 * rpc), r = (n / rpc) % 2, row = n % rpc;
        const size_t off = unit_offset(layer, seg, kblk, nh) + ((((size_t)r * g + (kseg - kblk)) * 2 + kc) * rpc + row) * 16;
        *reinterpret_cast<uint4*>(out + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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

#ifdef NF_TUNING
template <bool BF16, int CL>
static int launch_t(const KernelArgs& a, cudaStream_t st, int grid) {
    auto* kern = k_nerf_mlp<BF16, CL>;
    // per device, not per process: set it on every launch (a few hundred ns) instead of caching a flag
    NF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return NF_OK;
}

// How many clusters of four CTAs (one CTA per SM: 227 KB of shared memory each) the device can hold at once: a cluster
// lives inside one GPC, so GPCs whose SM count is not a multiple of four leave SMs idle.  Queried once per device.
template <bool BF16>
static int max_clusters4() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 0;
    if (cached[dev] == 0) {
        auto* kern = k_nerf_mlp<BF16, 4>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(num_sms() & ~3), 1, 1);
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = SM_TOTAL;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
        cached[dev] = n > 0 ? n : -1;
    }
    return cached[dev] > 0 ? cached[dev] : 0;
}
#endif  // NF_TUNING


// Production: the two-tile kernel (nf_mlp2.cu).  The tuning build can select the one-tile kernel (NF_MLP_IMPL=1), optionally
// as clusters of two pairs with multicast weight quarters (NF_MLP_CLUSTER=4), for back-to-back comparisons on one box.
int launch(const KernelArgs& a, int dtype, cudaStream_t st) {
#ifdef NF_TUNING
    const char* impl = getenv("NF_MLP_IMPL");
    if (impl && atoi(impl) == 1) {
        const bool bf = dtype == NF_DTYPE_BF16;
        const int pairs_grid = num_sms() & ~1;
        const char* e = getenv("NF_MLP_CLUSTER");
        if (e && atoi(e) == 4) {
            const int c4 = bf ? max_clusters4<true>() : max_clusters4<false>();
            if (c4 * 4 * 10 >= pairs_grid * 9) {       // at most 10 % of the SMs left without a cluster
                const int grid = 4 * (c4 < num_sms() / 4 ? c4 : num_sms() / 4);
                return bf ? launch_t<true, 4>(a, st, grid) : launch_t<false, 4>(a, st, grid);
            }
        }
        return bf ? launch_t<true, 2>(a, st, pairs_grid) : launch_t<false, 2>(a, st, pairs_grid);
    }
#endif
    return launch2(a, dtype, st);
}

}  // namespace mlp
}  // namespace nf

using namespace nf;

extern "C" size_t nf_render_packed_weights_bytes(void) { return (size_t)mlp::PACKED_BYTES; }

extern "C" int nf_render_pack_weights(const float* const* params, int dtype, void* packed_out, void* stream_) {
    return nf_render_pack_weights_ex(params, dtype, NF_ENC_ALL, packed_out, stream_);
}

extern "C" int nf_render_pack_weights_ex(const float* const* params, int dtype, int enc_flags, void* packed_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(params && packed_out, NF_E_INVALID, "nf_render_pack_weights: null argument");
    NF_REQUIRE(enc_flags >= 0 && enc_flags <= NF_ENC_ALL, NF_E_INVALID, "nf_render_pack_weights: enc_flags %d", enc_flags);
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_render_pack_weights: dtype %d", dtype);
    mlp::PackArgs p;
    for (int i = 0; i < 12; ++i) {
        p.w[i] = params[2 * i];
        p.b[i] = params[2 * i + 1];
        NF_REQUIRE(p.w[i] && p.b[i], NF_E_INVALID, "nf_render_pack_weights: null parameter %d", i);
    }
    const int total = mlp::N256_STEPS * 512 + mlp::N128_STEPS * 256;
    if (dtype == NF_DTYPE_BF16)
        mlp::k_pack_weights<true><<<(total + 255) / 256, 256, 0, st>>>(p, (uint8_t*)packed_out, enc_flags);
    else
        mlp::k_pack_weights<false><<<(total + 255) / 256, 256, 0, st>>>(p, (uint8_t*)packed_out, enc_flags);
    NF_LAUNCH_OK();
    return NF_OK;
}

#ifdef NF_TUNING
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
#endif

extern "C" size_t nf_nerf_mlp_workspace_bytes(void) { return mlp::PE_SCRATCH_BYTES; }

extern "C" int nf_nerf_mlp_forward(const void* packed, int dtype, const float* records, int n_rows, int sigma_only,
                                   float* out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(packed && out && n_rows >= 0, NF_E_INVALID, "nf_nerf_mlp_forward: bad arguments");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_nerf_mlp_forward: dtype %d", dtype);
    if (n_rows == 0) return NF_OK;
    NF_REQUIRE(records != nullptr, NF_E_INVALID, "nf_nerf_mlp_forward: null records");
    NF_REQUIRE(workspace && workspace_bytes >= mlp::PE_SCRATCH_BYTES, NF_E_WORKSPACE, "nf_nerf_mlp_forward: workspace %zu < %zu",
               workspace_bytes, (size_t)mlp::PE_SCRATCH_BYTES);
    mlp::KernelArgs a;
    a.pe_scratch = (uint8_t*)workspace;
    a.packed = (const uint8_t*)packed;
    a.records = records;
    a.rowid = nullptr;
    a.n_rows_dev = nullptr;
    a.n_rows_host = n_rows;
    a.n_rows_cap = n_rows;
    a.n_layers = sigma_only ? 8 : 10;
    a.out4 = (float4*)out;
#ifdef NF_TUNING      // tests/gpu_mlp_trace.py (clock64 timeline of one tile) runs against the tuning build only
    a.desc_swap = env_int("NF_MLP_DESC_SWAP", 0);
    a.trace = (long long*)(uintptr_t)strtoull(getenv("NF_MLP_TRACE_PTR") ? getenv("NF_MLP_TRACE_PTR") : "0", nullptr, 0);
#else
    a.desc_swap = 0;
    a.trace = nullptr;
#endif
    return mlp::launch(a, dtype, st);
}
