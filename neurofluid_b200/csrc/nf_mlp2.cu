// nf_mlp2.cu -- the fused positional-encoding + NeRF MLP forward (sm_100a, tcgen05): TWO 128-row tiles per CTA.
//
// replaces: Embedding.forward x6 (models/nerf.py:21-38 via models/renderer.py:125-179) and NeRF.forward
//           (models/nerf.py:83-124; two instances, models/renderer.py:43-44).
//
// Work unit: a tile of 128 compact "geometry records" (16 fp32 per evaluated ray sample, written by the ray-stage kernels
// in nf_render.cu).  Persistent CTAs, one per SM, run as CTA PAIRS (cluster of 2, tcgen05 cta_group::2, M = 256); a pair
// takes 4 tiles per pass: tile T (0, 1) of CTA r = rows of tile 4 * pass + 2 r + T.
//
// Why two tiles per CTA (profiles/r02_mlp_timeline.txt, profiles/r02_notes.md).  A layer needs 2,048 cycles of MMA per
// tile AND 2,048 cycles to drain its 128 x 256 fp32 accumulator out of TMEM (tcgen05.ld moves 64 B per clock per SM), and
// the next layer of the same tile needs that drain's result: with one tile per CTA the two never overlap (measured:
// ~3,750 cycles per layer).  Here the tensor pipe works on one tile while the other tile's accumulator is drained, and
// every weight unit that lands in shared memory is used by BOTH tiles (half the L2 -> SM weight traffic per row).
// Measured on the same box, back to back: fine network 37.1 -> 32.0 ms per 800 x 800 image, bit-identical outputs.
//
// Shared memory (222 KB): two in-place activation tiles (2 x 64 KB, UMMA K-major no-swizzle "tile image":
// byte(row r, column k) = (k / 8) * 2048 + r * 16 + (k % 8) * 2), a 5-stage x 16 KB mbarrier ring, the fp32 biases / heads.
// The ring carries, in consumption order (nf_mlp.cuh), the weight units AND -- for layers 4 (xyz-like skip input) and 9
// (dir-like) -- "A pieces": 4 K-steps of a tile's encoded features.  Layer 0's encodings are bulk-copied into the tile's own
// activation buffer (dead between the last layer's MMAs and layer 0's first epilogue) by a second loader lane.  The encodings are written by the 4
// producer warps to a per-CTA scratch in the caller's workspace (KernelArgs::pe_scratch, L2-resident) one pass ahead, as
// tile images, so that they occupy no shared memory between the layers that use them.  (Writing them straight into ring
// stages was tried: with 4 of the 5 stages held by one K-block there is no room to produce ahead, 32 -> 54 ms.)
// TMEM: accumulators of tile T in columns [256 T, 256 T + 256), single-buffered: the issuer waits for "accumulator half
// drained" before it overwrites one.
//
// Roles (14 warps): 0-7 epilogue (two groups of four; group g owns columns [64c + 32g, +32) of every 64-column chunk c;
// drain order per layer: (half 0, tile 0) (half 0, tile 1) (half 1, tile 0) (half 1, tile 1); tcgen05.ld -> +bias -> ReLU ->
// fp16 -> back into the tile's activation image; sigma (256 -> 1) and rgb (128 -> 3) heads are register dot products),
// 8 MMA issuer (rank 0; converged warp, one elected lane issues; one wait + one commit per weight unit) / weight relay
// (rank 1), 9 loader (cp.async.bulk), 10-13 encoding producers (record -> sin/cos by double-angle recurrence re-anchored
// at 2^5, fp32 -> fp16/bf16).
// The issuer takes a hidden layer's eight (weight unit, tile) steps in an order that lets tile 0's high-K units go before
// tile 1's (K low, half 1): the drain runs gap-free (6.0k -> 5.2k cycles per layer of two tiles, 33.2 -> 31.9 ms).  A variant
// that runs tile 1 a whole weight unit behind tile 0 everywhere measured 41 ms: not kept.  The tuning build (-DNF_TUNING) can still run the one-tile kernel of nf_mlp.cu for comparison.
//
// Per row the tensor pipe does 671,744 MAC (665,984 algorithmic + K padding 198->208, 54->64).
// HBM traffic per row: 64 B record in, 16 B out (+4 B row id); weights (1.34 MB/net) stay in L2.
#include "nf_common.cuh"
#include "nf_mlp.cuh"
#include "nf_tc.cuh"

namespace nf {
namespace mlp {
namespace v2 {

#ifdef NF_TUNING
#define NF_TRACE2(cond, slot) do { if (a.trace && (cond)) a.trace[(slot)] = clock64(); } while (0)
#else
#define NF_TRACE2(cond, slot) do { } while (0)
#endif

constexpr int NST = 5;
constexpr int STAGE = 16384;
constexpr int SM_HID = 0;                               // 2 x (128 x 256 halves)
constexpr int SM_RING = SM_HID + 2 * 65536;
constexpr int SM_SPARAM = SM_RING + NST * STAGE;
constexpr int SM_PART = SM_SPARAM + SP_FLOATS * 4;      // 128 x float4: head partial sums of epilogue group 1
constexpr int SM_BAR = SM_PART + 128 * 16;
enum Bar {
    B_WFULL = 0,
    B_WEMPTY = NST,
    B_ACT_READY = 2 * NST,          // [tile 2][chunk 4]
    B_ACC_FULL = B_ACT_READY + 8,   // [tile 2][half 2]
    B_ACC_FREE = B_ACC_FULL + 4,    // [tile 2][half 2]
    B_PE_READY = B_ACC_FREE + 4,    // [parity 2]
    B_PE_FREE = B_PE_READY + 2,     // [parity 2]
    B_L0_FULL = B_PE_FREE + 2,      // [tile 2]: layer 0's xyz-like encodings have landed in the tile's activation buffer
    B_HID_FREE = B_L0_FULL + 2,     // [tile 2]: the last layer's MMAs have read the tile's activation buffer
    NUM_BARS = B_HID_FREE + 2,
};
constexpr int SM_TMEM_SLOT = SM_BAR + NUM_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM_SLOT + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(SM_BAR % 8 == 0, "barrier alignment");

constexpr int W_ISSUE = 8, W_LOAD = 9, W_PE = 10;
constexpr int NUM_THREADS = 14 * 32;
constexpr int PE_TILE_BYTES = (26 + 8) * 2048;          // xyz-like image (26 chunks) + dir-like image (8 chunks)
constexpr int PE_CTA_BYTES = 2 /*parity*/ * 2 /*tile*/ * PE_TILE_BYTES;
static_assert((size_t)PE_SCRATCH_CTAS * PE_CTA_BYTES == PE_SCRATCH_BYTES, "encoding scratch size (nf_mlp.cuh)");
constexpr int TPP = 4;                                  // tiles per pair per pass

// encoded-feature row -> tile image in GLOBAL memory: byte(row r, column k) = (k / 8) * 2048 + r * 16 + (k % 8) * 2
template <bool BF16>
struct RowWriterG {
    uint8_t* base;  // image + row * 16
    float lo;
    uint32_t pk[4];
    template <int COL>
    __device__ __forceinline__ void put(float v) {
        if constexpr ((COL & 1) == 0) {
            lo = v;
        } else {
            pk[(COL & 7) >> 1] = pack2<BF16>(lo, v);
            if constexpr ((COL & 7) == 7) *reinterpret_cast<uint4*>(base + (COL >> 3) * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
};
template <bool BF16, int BASE, int C, int L>
__device__ __forceinline__ void emit_encoding_g(RowWriterG<BF16>& w, const float* v) {
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

__device__ __forceinline__ int stages_per_pass(int nl, bool l0_hid) {
    int n = 0;
    for (int l = 0; l < nl; ++l) {
        const int npe = layer_pe_steps(l);
        n += ((l == 0 && l0_hid) ? 2 : 4) * ((npe + wu_ksteps(0) - 1) / wu_ksteps(0)) + (l > 0 ? 2 * (KH_STEPS / wu_ksteps(1)) : 0);
    }
    return n;
}

template <bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_nerf_mlp2(const KernelArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr bool PAIR = true;
    const int warp = uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
    const uint32_t prank = uniform(cluster_ctarank());
    const int unit = (int)(blockIdx.x >> 1), nunits = (int)(gridDim.x >> 1);
    const int n_rows = uniform(a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows_cap) : a.n_rows_host);
    const int ntiles = (n_rows + TILE_M - 1) / TILE_M;
    const int npass = (ntiles + TPP - 1) / TPP;
    if (unit >= npass) return;          // both CTAs of a pair leave together
    const int nl = a.n_layers;
    // Layer 0 reads its A operand (the xyz-like encodings, 52 KB per tile) from the tile's own activation buffer, which is dead
    // between the last layer's MMAs and layer 0's first epilogue: one bulk copy per tile from the encoding scratch, issued as
    // soon as the buffer is free, i.e. under the last layer's drain, instead of 8 ring stages that compete with layer 0's
    // weights for the SM's ingest (the ring carries <= 64 KB in flight: ~24 B/clk, and a 4-K-step block wants 48 KB per 1,024).
#ifdef NF_TUNING
    const bool l0_hid = (a.desc_swap & 64) == 0;       // NF_MLP_DESC_SWAP=64: layer 0's A pieces through the ring, for comparison
#else
    constexpr bool l0_hid = true;
#endif

    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_ring = s_base + SM_RING, s_bar = s_base + SM_BAR;
    float* sp = reinterpret_cast<float*>(smem + SM_SPARAM);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART);
    auto bar = [&](int i) { return s_bar + 8u * (uint32_t)i; };
    auto bar0 = [&](int i) { return mapa_rank(bar(i), 0); };       // the same barrier in the issuer's CTA (rank 0)
    uint8_t* my_scratch = a.pe_scratch + (size_t)blockIdx.x * PE_CTA_BYTES;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(bar(B_WFULL + i), prank == 0 ? 2 : 1);    // own expect_tx arrive (+ the peer's relay)
            mbar_init(bar(B_WEMPTY + i), 1);
        }
        for (int i = 0; i < 8; ++i) mbar_init(bar(B_ACT_READY + i), 16);     // 8 epilogue warps x 2 CTAs
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar(B_ACC_FULL + i), 1);
            mbar_init(bar(B_ACC_FREE + i), 16);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_PE_READY + i), 4);                  // the 4 producer warps of this CTA
            mbar_init(bar(B_PE_FREE + i), 1);
            mbar_init(bar(B_L0_FULL + i), prank == 0 ? 2 : 1);  // own expect_tx arrive (+ the peer's relay)
            mbar_init(bar(B_HID_FREE + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a.packed + W_BYTES);
        float4* dst = reinterpret_cast<float4*>(sp);
        for (int i = threadIdx.x; i < SP_FLOATS / 4; i += NUM_THREADS) dst[i] = __ldg(src + i);
    }
    if (warp == W_ISSUE) tmem_alloc<PAIR>(s_base + SM_TMEM_SLOT, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = uniform(*reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_SLOT));

    if (warp == W_ISSUE) {
        if (prank == 0) {
            // ================================================================ MMA issuer (converged warp, one elected lane issues)
            const uint32_t idesc128 = umma_idesc(128, BF16, 2 * TILE_M), idesc64 = umma_idesc(64, BF16, 2 * TILE_M);
            const uint32_t pe_free_remote = mapa_rank(bar(B_PE_FREE), 1);
            uint32_t ws = 0, wph = 0, hidw = 0, lcount = 0;
            int ti = 0;
            auto advance = [&]() { if (++ws == NST) { ws = 0; wph ^= 1; } };
#ifdef NF_TUNING
            const bool t0_ahead = (a.desc_swap & 32) == 0;     // NF_MLP_DESC_SWAP=32: the plain lockstep order, for comparison
#else
            constexpr bool t0_ahead = true;
#endif
            for (int pass = unit; pass < npass; pass += nunits, ++ti) {
                for (int l = 0; l < nl; ++l, ++lcount) {
                    const bool n128 = (l == 9);
                    int tslot = 100 + l * 12;
                    NF_TRACE2(blockIdx.x == 0 && ti == 2 && lane == 0, tslot++);
                    const uint32_t nhalf = n128 ? 64u : 128u, rpc = nhalf >> 1;
                    const uint32_t idesc = n128 ? idesc64 : idesc128;
                    const uint32_t bstep = (2u * rpc * 16u) >> 4;
                    uint32_t accf = 0;                                  // bit (T * 2 + nh): that accumulator half has been started
                    const int npe = layer_pe_steps(l);
                    // ---- encoded-feature segment (layers 0, 4, 9): K-blocks of 4 K-steps, A pieces through the ring
                    for (int k0 = 0; k0 < npe; k0 += wu_ksteps(0)) {
                        const int g = min(wu_ksteps(0), npe - k0);
                        const bool last_blk = (l == 0) && (k0 + wu_ksteps(0) >= npe);
                        const bool from_hid = l0_hid && l == 0;
                        uint32_t sa[2] = {0u, 0u};
                        if (from_hid) {
                            if (k0 == 0) { mbar_wait(bar(B_L0_FULL), ti & 1); mbar_wait(bar(B_L0_FULL + 1), ti & 1); }
                        } else {
#pragma unroll
                            for (int T = 0; T < 2; ++T) {           // the two tiles' encoded-feature pieces of this K-block
                                mbar_wait(bar(B_WFULL + ws), wph);
                                sa[T] = ws;
                                advance();
                            }
                        }
#pragma unroll
                        for (uint32_t nh = 0; nh < 2; ++nh) {
                            mbar_wait(bar(B_WFULL + ws), wph);
                            const uint64_t bd = umma_desc(s_ring + ws * STAGE, rpc * 16u, 128u);
#pragma unroll
                            for (uint32_t T = 0; T < 2; ++T) {
                                const uint32_t abit = 1u << (T * 2 + nh);
                                if (!(accf & abit)) {      // the previous layer's epilogue has drained what these MMAs overwrite
                                    // the dir layer's halves are 64 columns wide: both lie inside half 0 of a 256-wide layer, so layer
                                    // 9 only needs layer 8's half-0 drains (waiting for half 1 held the block's four ring stages ~2k
                                    // cycles longer), and layer 0 of the next pass needs both of layer 9's
                                    mbar_wait(bar(B_ACC_FREE + T * 2 + (n128 ? 0u : nh)), (lcount & 1) ^ 1);
                                    if (l == 0 && nh == 0) mbar_wait(bar(B_ACC_FREE + T * 2 + 1), (lcount & 1) ^ 1);
                                }
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint64_t ad = from_hid ? umma_desc(s_base + SM_HID + T * 65536 + (uint32_t)k0 * 4096u, 2048u, 128u)
                                                                 : umma_desc(s_ring + sa[T] * STAGE, 2048u, 128u);
                                    uint32_t acc = (accf & abit) ? 1u : 0u;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (j < g) {
                                            umma_f16<PAIR>(tmem_base + T * 256 + nh * nhalf, ad + (uint64_t)(j * (4096 >> 4)), bd + (uint64_t)(j * bstep), idesc, acc);
                                            acc = 1;
                                        }
                                    }
                                    if (last_blk) umma_commit<PAIR>(bar(B_ACC_FULL + T * 2 + nh));
                                    if (nh == 1 && !from_hid) umma_commit<PAIR>(bar(B_WEMPTY + sa[T]));
                                    if (T == 1) umma_commit<PAIR>(bar(B_WEMPTY + ws));
                                }
                                __syncwarp();
                                accf |= abit;
                            }
                            advance();
                        }
                    }
                    // ---- hidden segment (layers 1..9): four weight units u = 2 blk + nh (K-block blk of 8 K-steps, N-half nh), each used
                    // by both tiles.  Order of the eight (unit, tile) steps: tile 0 takes (blk 1, nh 0) BEFORE tile 1 takes (blk 0, nh 1):
                    // tile 0's high-K units only need tile 0's own activation chunks (the previous layer's THIRD drain), tile 1's
                    // (blk 0, nh 1) needs the fourth drain to have started; so tile 0's first accumulator half is complete by the time
                    // the previous layer's last drain ends and the drain -- the busiest resource -- runs without a gap.
                    if (l > 0) {
                        uint32_t us[4], up[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { us[u] = ws; up[u] = wph; advance(); }
                        const uint32_t perm = t0_ahead ? 0x76534210u : 0x76543210u;
#pragma unroll
                        for (int step = 0; step < 8; ++step) {
                            const uint32_t code = (perm >> (4 * step)) & 7u;
                            const uint32_t blk = code >> 2, nh = (code >> 1) & 1u, T = code & 1u, u = code >> 1;
                            const uint32_t k0 = blk * 8u;
                            mbar_wait(bar(B_WFULL + us[u]), up[u]);
                            if (nh == 0) {          // this block's K-steps read activation chunks k0/4, k0/4 + 1
                                mbar_wait(bar(B_ACT_READY + T * 4 + (k0 >> 2)), hidw & 1);
                                mbar_wait(bar(B_ACT_READY + T * 4 + (k0 >> 2) + 1), hidw & 1);
                            }
                            const uint32_t abit = 1u << (T * 2 + nh);
                            if (!(accf & abit)) mbar_wait(bar(B_ACC_FREE + T * 2 + nh), (lcount & 1) ^ 1);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t bd = umma_desc(s_ring + us[u] * STAGE, rpc * 16u, 128u);
                                const uint64_t ad = umma_desc(s_base + SM_HID + T * 65536 + k0 * 4096u, 2048u, 128u);
                                uint32_t acc = (accf & abit) ? 1u : 0u;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    umma_f16<PAIR>(tmem_base + T * 256 + nh * nhalf, ad + (uint64_t)(j * (4096 >> 4)), bd + (uint64_t)(j * bstep), idesc, acc);
                                    acc = 1;
                                }
                                if (blk == 1) umma_commit<PAIR>(bar(B_ACC_FULL + T * 2 + nh));
                                if (T == 1) umma_commit<PAIR>(bar(B_WEMPTY + us[u]));
                                if (l == nl - 1 && code >= 6u) umma_commit<PAIR>(bar(B_HID_FREE + T));   // (blk 1, nh 1, T): the tile's last step
                            }
                            __syncwarp();
                            accf |= abit;
                            if (T == 1) NF_TRACE2(blockIdx.x == 0 && ti == 2 && lane == 0, tslot++);
                        }
                    }
                    if (l > 0) ++hidw;
                }
                // every stage of this pass has landed: the encoding scratch of this parity may be rewritten
                if (lane == 0) {
                    mbar_arrive(bar(B_PE_FREE + (ti & 1)));
                    mbar_arrive_cluster(pe_free_remote + 8u * (ti & 1));
                }
                __syncwarp();
            }
        } else if (lane == 0) {
            // ================================================================ relay (rank 1): "my share of stage s has landed"
            uint32_t ws = 0, wph = 0;
            const int nstages = stages_per_pass(nl, l0_hid);
            const uint32_t remote0 = bar0(B_WFULL);
            for (int pass = unit; pass < npass; pass += nunits) {
                for (int s = 0; s < nstages; ++s) {
                    mbar_wait(bar(B_WFULL + ws), wph);
                    mbar_arrive_cluster(remote0 + 8u * ws);
                    if (++ws == NST) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == W_LOAD) {
        // ================================================================ loader: weight units + encoded-feature pieces
        if (lane == 0) {
            uint32_t ws = 0, wph = 0;
            int ti = 0;
            auto fill = [&](const void* src, uint32_t bytes) {
                mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                mbar_arrive_expect_tx(bar(B_WFULL + ws), bytes);
                bulk_g2s(s_ring + ws * STAGE, src, bytes, bar(B_WFULL + ws));
                if (++ws == NST) { ws = 0; wph ^= 1; }
            };
            for (int pass = unit; pass < npass; pass += nunits, ++ti) {
                const uint8_t* src = a.packed;
                const uint8_t* pe = my_scratch + (size_t)(ti & 1) * 2 * PE_TILE_BYTES;
                mbar_wait(bar(B_PE_READY + (ti & 1)), (ti >> 1) & 1);
                for (int l = 0; l < nl; ++l) {
                    const uint32_t rpc = (l == 9) ? 32u : 64u;
                    const int npe = layer_pe_steps(l);
                    for (int seg = 0; seg < 2; ++seg) {
                        const int nsteps = seg == 0 ? npe : (l > 0 ? KH_STEPS : 0);
                        const int kb = wu_ksteps(seg);
                        for (int k0 = 0; k0 < nsteps; k0 += kb) {
                            const int g = min(kb, nsteps - k0);
                            if (seg == 0 && !(l0_hid && l == 0))
                                for (int T = 0; T < 2; ++T)
                                    fill(pe + (size_t)T * PE_TILE_BYTES + (l == 9 ? 26 * 2048 : 0) + (size_t)k0 * 4096, (uint32_t)g * 4096u);
                            const uint32_t mine = (uint32_t)g * 2u * rpc * 16u;
                            for (int nh = 0; nh < 2; ++nh) {
                                fill(src + prank * mine, mine);
                                src += 2 * mine;
                            }
                        }
                    }
                }
            }
            for (int i = 0; i < NST; ++i) {      // every multicast "slot free" arrive has landed before this CTA may exit
                mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                if (++ws == NST) { ws = 0; wph ^= 1; }
            }
        } else if (lane == 1 && l0_hid) {
            // second loader lane: layer 0's encodings, one bulk copy per tile, as soon as the tile's activation buffer is free
            constexpr uint32_t L0_BYTES = KX_STEPS * 4096u;
            const uint32_t l0_remote = bar0(B_L0_FULL);
            int ti = 0;
            for (int pass = unit; pass < npass; pass += nunits, ++ti) {
                const uint8_t* pe = my_scratch + (size_t)(ti & 1) * 2 * PE_TILE_BYTES;
                mbar_wait(bar(B_PE_READY + (ti & 1)), (ti >> 1) & 1);
                for (int T = 0; T < 2; ++T) {
                    mbar_wait(bar(B_HID_FREE + T), (ti & 1) ^ 1);
                    mbar_arrive_expect_tx(bar(B_L0_FULL + T), L0_BYTES);
                    bulk_g2s(s_base + SM_HID + T * 65536, pe + (size_t)T * PE_TILE_BYTES, L0_BYTES, bar(B_L0_FULL + T));
                }
                if (prank != 0) {           // tell the issuer (rank 0) that this CTA's tiles have landed too
                    for (int T = 0; T < 2; ++T) {
                        mbar_wait(bar(B_L0_FULL + T), ti & 1);
                        mbar_arrive_cluster(l0_remote + 8u * T);
                    }
                }
            }
            // the last pass's (multicast) "buffer free" arrives have landed before this CTA may exit
            for (int T = 0; T < 2; ++T) mbar_wait(bar(B_HID_FREE + T), (ti & 1) ^ 1);
        }
    } else if (warp < W_ISSUE) {
        // ================================================================ epilogue
        const int grp = warp >> 2;
        const int tr = (warp & 3) * 32 + lane;           // row of the tile = TMEM lane
        const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * 32;
        const uint32_t act_ready0 = bar0(B_ACT_READY), acc_free0 = bar0(B_ACC_FREE);
        uint32_t lcount = 0;
        int ti = 0;
        for (int pass = unit; pass < npass; pass += nunits, ++ti) {
            float sigma[2] = {0.f, 0.f};
            float rgb[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            for (int l = 0; l < nl; ++l, ++lcount) {
                const bool writes = (l + 1 < nl);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int T = 0; T < 2; ++T) {
                        const int row = ((pass * TPP) + (int)prank * 2 + T) * TILE_M + tr;
                        mbar_wait(bar(B_ACC_FULL + T * 2 + h), lcount & 1);
                        tc_fence_after();
                        NF_TRACE2(blockIdx.x == 0 && ti == 2 && threadIdx.x == 0, l * 8 + (h * 2 + T) * 2);
                        const uint32_t taddr = tlane + T * 256;
                        if (l < 9) {
                            const float* bias = sp + SP_BIAS + l * 256 + grp * 32;
                            uint32_t v[2][32];
                            tmem_ld32(taddr + (2 * h) * 64, v[0]);
                            tmem_ld32(taddr + (2 * h + 1) * 64, v[1]);
                            tmem_ld_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(acc_free0 + 8u * (T * 2 + h));       // this warp's part of the half is in registers
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc) {
                                const int c = 2 * h + cc;
                                float f[32];
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 64 + i);
                                    f[i] = __uint_as_float(v[cc][i]) + b4.x;
                                    f[i + 1] = __uint_as_float(v[cc][i + 1]) + b4.y;
                                    f[i + 2] = __uint_as_float(v[cc][i + 2]) + b4.z;
                                    f[i + 3] = __uint_as_float(v[cc][i + 3]) + b4.w;
                                }
                                if (l != 8) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                                }
                                if (l == 7) {
                                    const float* wsig = sp + SP_WSIG + c * 64 + grp * 32;
                                    float sg = sigma[T];
#pragma unroll
                                    for (int i = 0; i < 32; i += 4) {
                                        const float4 w4 = *reinterpret_cast<const float4*>(wsig + i);
                                        sg = fmaf(f[i], w4.x, sg);
                                        sg = fmaf(f[i + 1], w4.y, sg);
                                        sg = fmaf(f[i + 2], w4.z, sg);
                                        sg = fmaf(f[i + 3], w4.w, sg);
                                    }
                                    sigma[T] = sg;
                                }
                                if (writes) {
                                    const uint32_t dst = s_base + SM_HID + T * 65536 + (uint32_t)(c * 8 + grp * 4) * 2048 + (uint32_t)tr * 16;
#pragma unroll
                                    for (int q = 0; q < 4; ++q)
                                        st_shared_v4(dst + q * 2048, pack2<BF16>(f[8 * q], f[8 * q + 1]), pack2<BF16>(f[8 * q + 2], f[8 * q + 3]),
                                                     pack2<BF16>(f[8 * q + 4], f[8 * q + 5]), pack2<BF16>(f[8 * q + 6], f[8 * q + 7]));
                                    fence_proxy_async();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive_cluster(act_ready0 + 8u * (T * 4 + c));
                                }
                            }
                            NF_TRACE2(blockIdx.x == 0 && ti == 2 && threadIdx.x == 0, l * 8 + (h * 2 + T) * 2 + 1);
                            if (l == 7 && !writes && h == 1) {   // sigma-only network: combine the two column halves and emit
                                if (grp == 1) part[tr] = make_float4(0.f, 0.f, 0.f, sigma[T]);
                                named_bar_sync(1, 256);
                                if (grp == 0 && row < n_rows) {
                                    const int dst = a.rowid ? a.rowid[row] : row;
                                    if (dst >= 0) a.out4[dst] = make_float4(0.f, 0.f, 0.f, sigma[T] + part[tr].w + sp[SP_BSIG]);
                                }
                                named_bar_sync(2, 256);
                            }
                        } else {
                            // rgb head on the 128-wide dir layer: N-half h = columns [64h, 64h + 64); group g owns [64h + 32g, +32)
                            const float* bias = sp + SP_BIAS + 9 * 256 + grp * 32;
                            const float* wrgb = sp + SP_WRGB + grp * 32;
                            uint32_t v[32];
                            tmem_ld32(taddr + h * 64, v);
                            tmem_ld_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(acc_free0 + 8u * (T * 2 + h));
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {       // 16-byte shared-memory loads (all offsets are multiples of 4 floats)
                                const float4 b4 = *reinterpret_cast<const float4*>(bias + h * 64 + i);
                                const float4 wr = *reinterpret_cast<const float4*>(wrgb + h * 64 + i);
                                const float4 wg = *reinterpret_cast<const float4*>(wrgb + 128 + h * 64 + i);
                                const float4 wb = *reinterpret_cast<const float4*>(wrgb + 256 + h * 64 + i);
                                const float f0 = fmaxf(__uint_as_float(v[i]) + b4.x, 0.f), f1 = fmaxf(__uint_as_float(v[i + 1]) + b4.y, 0.f);
                                const float f2 = fmaxf(__uint_as_float(v[i + 2]) + b4.z, 0.f), f3 = fmaxf(__uint_as_float(v[i + 3]) + b4.w, 0.f);
                                rgb[T][0] = fmaf(f0, wr.x, rgb[T][0]); rgb[T][1] = fmaf(f0, wg.x, rgb[T][1]); rgb[T][2] = fmaf(f0, wb.x, rgb[T][2]);
                                rgb[T][0] = fmaf(f1, wr.y, rgb[T][0]); rgb[T][1] = fmaf(f1, wg.y, rgb[T][1]); rgb[T][2] = fmaf(f1, wb.y, rgb[T][2]);
                                rgb[T][0] = fmaf(f2, wr.z, rgb[T][0]); rgb[T][1] = fmaf(f2, wg.z, rgb[T][1]); rgb[T][2] = fmaf(f2, wb.z, rgb[T][2]);
                                rgb[T][0] = fmaf(f3, wr.w, rgb[T][0]); rgb[T][1] = fmaf(f3, wg.w, rgb[T][1]); rgb[T][2] = fmaf(f3, wb.w, rgb[T][2]);
                            }
                            if (h == 1) {
                                if (grp == 1) part[tr] = make_float4(rgb[T][0], rgb[T][1], rgb[T][2], sigma[T]);
                                named_bar_sync(1, 256);
                                if (grp == 0 && row < n_rows) {
                                    const int dst = a.rowid ? a.rowid[row] : row;
                                    if (dst >= 0) {
                                        const float4 pb = part[tr];
                                        float4 o;
                                        o.x = 1.0f / (1.0f + expf(-(rgb[T][0] + pb.x + sp[SP_BRGB])));
                                        o.y = 1.0f / (1.0f + expf(-(rgb[T][1] + pb.y + sp[SP_BRGB + 1])));
                                        o.z = 1.0f / (1.0f + expf(-(rgb[T][2] + pb.z + sp[SP_BRGB + 2])));
                                        o.w = sigma[T] + pb.w + sp[SP_BSIG];
                                        a.out4[dst] = o;
                                    }
                                }
                                named_bar_sync(2, 256);
                            }
                        }
                    }
                }
            }
        }
    } else {
        // ================================================================ encoding producers (warps 10-13): one pass ahead, to global scratch
        const int tp = threadIdx.x - W_PE * 32;  // 0..127
        int ti = 0;
        for (int pass = unit; pass < npass; pass += nunits, ++ti) {
            mbar_wait(bar(B_PE_FREE + (ti & 1)), ((ti >> 1) & 1) ^ 1);
            uint8_t* base = my_scratch + (size_t)(ti & 1) * 2 * PE_TILE_BYTES + (size_t)tp * 16;
#pragma unroll 1
            for (int T = 0; T < 2; ++T) {
                const int row = ((pass * TPP) + (int)prank * 2 + T) * TILE_M + tp;
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
                RowWriterG<BF16> w;
                w.base = base + (size_t)T * PE_TILE_BYTES;
                emit_encoding_g<BF16, 0, 3, 10>(w, r + 0);
                emit_encoding_g<BF16, 63, 1, 4>(w, r + 3);
                emit_encoding_g<BF16, 72, 3, 10>(w, r + 4);
                emit_encoding_g<BF16, 135, 3, 10>(w, r + 7);
                static_for<198, 208>([&](auto ci) { w.template put<decltype(ci)::value>(0.f); });
                if (nl == 10) {
                    // dir-like block: [PE4(ray dir) 27 | PE4(smoothed dir) 27 | 0 x10]
                    RowWriterG<BF16> wd;
                    wd.base = base + (size_t)T * PE_TILE_BYTES + 26 * 2048;
                    emit_encoding_g<BF16, 0, 3, 4>(wd, r + 10);
                    emit_encoding_g<BF16, 27, 3, 4>(wd, r + 13);
                    static_for<54, 64>([&](auto ci) { wd.template put<decltype(ci)::value>(0.f); });
                }
            }
            __threadfence();                                               // the images are read back by bulk copies (async proxy)
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_PE_READY + (ti & 1)));
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_ISSUE) tmem_dealloc<PAIR>(tmem_base, 512);
}

template <bool BF16>
static int launch_t(const KernelArgs& a, cudaStream_t st) {
    auto* kern = k_nerf_mlp2<BF16>;
    NF_REQUIRE(a.pe_scratch != nullptr, NF_E_WORKSPACE, "nf_mlp: no encoding scratch in the workspace");
    NF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)min(num_sms() & ~1, PE_SCRATCH_CTAS), 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return NF_OK;
}

}  // namespace v2

int launch2(const KernelArgs& a, int dtype, cudaStream_t st) {
    return dtype == NF_DTYPE_BF16 ? v2::launch_t<true>(a, st) : v2::launch_t<false>(a, st);
}

}  // namespace mlp
}  // namespace nf
