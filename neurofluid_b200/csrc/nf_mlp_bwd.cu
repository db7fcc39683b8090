// nf_mlp_bwd.cu -- backward pass of the fused positional-encoding + NeRF MLP on tcgen05 tensor cores (sm_100a).
//
// replaces: autograd through Embedding.forward x6 + NeRF.forward (models/nerf.py:21-38, 83-124), i.e. what
//           loss.backward() runs for the renderer at trainer/trainer_e2e.py:277 and trainer/trainer_renderer.py:96.
//
// Per 128-row tile of evaluated samples (same compact record rows as the forward kernel):
//
//   k_mlp_bwd_dgrad   one CTA per tile, three roles: 4 worker warps (thread = row), 1 MMA issuer lane, 1 weight loader lane.
//     forward recompute: the ten layers again (same operand dtype and the same packed weights as the forward kernel, so
//       the ReLU masks are the forward's), every layer input saved to HBM as a bf16 "tile image" -- the UMMA K-major
//       no-swizzle layout the tile has in shared memory: byte(row r, column c) = (c / 8) * 2048 + r * 16 + (c % 8) * 2;
//     data gradient: the chain dY_l -> dX_l = dY_l W_l runs as ten more GEMMs against transposed weight slabs
//       (nf_render_pack_weights_bwd), bf16 operands (gradients need the exponent range), fp32 accumulators in TMEM; the
//       epilogues apply the ReLU masks (read back from the saved images), add the sigma head's rank-1 term, and save
//       every dY_l as an image too.  The gradient w.r.t. the 252 encoded features leaves as fp32 rows (dfeat), the
//       sigma / rgb head weight gradients are column sums over the tile in shared memory.
//   k_mlp_bwd_wgrad   dW_l = dY_l^T X_l over all tiles: the saved images are bulk-copied back into shared memory and fed
//     to tcgen05.mma as MN-MAJOR operands (the K-major image of an activation tile is exactly the canonical MN-major
//     layout of its transpose: SBO = 2048, LBO = 128; tools/umma_mn_test.cu), K = the 128 rows of a tile, accumulating
//     over the tiles of a split in TMEM; 22 (layer, 128-row block, input segment) items x splits CTAs; fp32 atomics into
//     the flat parameter-gradient buffer (nn.Linear layout).  Bias gradients are column sums of the dY images.
#include "nf_common.cuh"
#include "nf_mlp.cuh"
#include "nf_tc.cuh"

namespace nf {
namespace mlp {
namespace bwd {

constexpr int CH = 2048;                         // one 8-column chunk of a 128-row tile image
constexpr int IMG_XPE = 0;                       // 26 chunks: encoded xyz-like features (208 columns)
constexpr int IMG_H1 = 26 * CH;                  // H(i), i = 1..8: input of layer i (output of layer i-1), 32 chunks each
constexpr int IMG_F = IMG_H1 + 8 * 65536;        // output of xyz_encoding_final
constexpr int IMG_PD = IMG_F + 65536;            // 8 chunks: encoded dir-like features (64 columns)
constexpr int IMG_D = IMG_PD + 8 * CH;           // 16 chunks: ReLU'd output of dir_encoding (128 columns)
constexpr int IMG_DP1 = IMG_D + 16 * CH;         // DP(i), i = 1..8: gradient w.r.t. the pre-activation of layer i-1
constexpr int IMG_DF = IMG_DP1 + 8 * 65536;      // gradient w.r.t. the output of xyz_encoding_final
constexpr int IMG_DD = IMG_DF + 65536;           // 16 chunks: gradient w.r.t. the pre-activation of dir_encoding
constexpr int TILE_SCRATCH = IMG_DD + 16 * CH;   // 1,314,816 bytes per tile
__host__ __device__ inline int img_h(int i) { return IMG_H1 + (i - 1) * 65536; }
__host__ __device__ inline int img_dp(int i) { return IMG_DP1 + (i - 1) * 65536; }

constexpr int DFEAT_W = 272;                     // 208 xyz-like + 64 dir-like columns (padding columns are zero gradients)

// ---- transposed weight slabs, in consumption order G9, G8, ..., G0 (G_l = data gradient through layer index l;
//      8 = xyz_encoding_final, 9 = dir_encoding).  One K-step = 16 output features o; slab[kc][row i][8 o's].
__host__ __device__ inline void bwd_gemm(int g /*0..9 in consumption order*/, int& nsteps, int& rows, int& layer) {
    layer = 9 - g;
    nsteps = (g == 0) ? 8 : 16;
    rows = (g == 0) ? 320 : (g == 5 ? 464 : (g == 9 ? 208 : 256));
}
// Every K-step is stored twice: the bf16 rounding of the weights ("hi") and the bf16 rounding of what the first
// rounding lost ("lo"); the data-gradient GEMMs run hi x hi + lo x hi + hi x lo, i.e. ~16 mantissa bits per operand, so
// that the ten-layer chain does not accumulate bf16 rounding (measured without the split: 1.0-1.6e-2 relative L2 on the
// final gradients, all of it rounding).
__host__ __device__ inline size_t bwd_pack_bytes() {
    size_t t = 0;
    for (int g = 0; g < 10; ++g) {
        int ns, rows, l;
        bwd_gemm(g, ns, rows, l);
        t += (size_t)ns * 2 * rows * 16 * 2;
    }
    return t;
}

// flat parameter(-gradient) layout: ordered_params order, nn.Linear (out, in) row-major
struct ParamOff {
    int w[12], b[12], total;
};
__host__ __device__ inline ParamOff param_offsets() {
    ParamOff P;
    const int out[12] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 128, 1, 3};
    const int in[12] = {198, 256, 256, 256, 454, 256, 256, 256, 256, 310, 256, 128};
    int o = 0;
    for (int i = 0; i < 12; ++i) {
        P.w[i] = o; o += out[i] * in[i];
        P.b[i] = o; o += out[i];
    }
    P.total = o;
    return P;
}

struct PackArgs {
    const float* w[12];
};

__global__ void k_pack_weights_bwd(PackArgs p, uint8_t* out, int enc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int base = 0;
    size_t byte0 = 0;
    for (int g = 0; g < 10; ++g) {
        int ns, rows, layer;
        bwd_gemm(g, ns, rows, layer);
        const int cnt = ns * 2 * rows;
        if (t < base + cnt) {
            const int u = t - base;
            const int step = u / (2 * rows), kc = (u / rows) % 2, i = u % rows;
            // source column of W_layer for slab row i (-1: padding)
            const int in_xyz = enc_in_xyz(enc), in_dir = enc_in_dir(enc);      // encoding ablations: narrower inputs (nf_mlp.cuh)
            int col = i, ld = 256;
            if (layer == 9) { ld = 256 + in_dir; col = i < 256 ? i : (enc_col_dir(i - 256, enc) < 0 ? -1 : 256 + enc_col_dir(i - 256, enc)); }
            else if (layer == 4) { ld = in_xyz + 256; col = i < 256 ? in_xyz + i : enc_col_xyz(i - 256, enc); }
            else if (layer == 0) { ld = in_xyz; col = enc_col_xyz(i, enc); }
            const float* W = p.w[layer];
            uint32_t pk[4], pl[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const int oa = step * 16 + kc * 8 + e;
                const float va = col >= 0 ? W[(size_t)oa * ld + col] : 0.f;
                const float vb = col >= 0 ? W[(size_t)(oa + 1) * ld + col] : 0.f;
                const float ha = __bfloat162float(__float2bfloat16(va)), hb = __bfloat162float(__float2bfloat16(vb));
                pk[e >> 1] = pack2<true>(va, vb);
                pl[e >> 1] = pack2<true>(va - ha, vb - hb);
            }
            const size_t half = (size_t)2 * rows * 16;
            uint8_t* dst = out + byte0 + (size_t)step * 2 * half + ((size_t)kc * rows + i) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + half) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            return;
        }
        base += cnt;
        byte0 += (size_t)ns * 2 * rows * 16 * 2;
    }
}

// ---------------------------------------------------------------- dgrad kernel
constexpr int NST = 4;
constexpr int WSTAGE = 16384;                     // >= the largest slab of one K-step (14,848 B)
constexpr int DS_HIDDEN = 0;
constexpr int DS_PEXYZ = DS_HIDDEN + 65536;
constexpr int DS_PEDIR = DS_PEXYZ + 26 * CH;
constexpr int DS_LO = DS_PEXYZ;                   // backward phase: the "lo" half of the gradient tile overlays the (dead) PE tiles
static_assert(26 * CH + 8 * CH >= 65536, "lo tile fits over the PE tiles");
constexpr int DS_WRING = DS_PEDIR + 8 * CH;
constexpr int DS_SPARAM = DS_WRING + NST * WSTAGE;
constexpr int DS_GS = DS_SPARAM + SP_FLOATS * 4;  // 128 x float4
constexpr int DS_BAR = DS_GS + 128 * 16;
constexpr int DS_TMEM = DS_BAR + (2 * NST + 2) * 8;
constexpr int DS_TOTAL = DS_TMEM + 16;
static_assert(DS_TOTAL <= 232448, "dgrad smem budget");
static_assert(DS_BAR % 8 == 0 && DS_GS % 16 == 0, "alignment");
constexpr int DG_THREADS = 6 * 32;
constexpr uint32_t TM_H = 0, TM_X = 256;

struct DgradArgs {
    const uint8_t* wf;       // forward pack (nf_render_pack_weights; CTA-pair slab layout) incl. the fp32 small params
    const uint8_t* wb;       // transposed bf16 slabs (nf_render_pack_weights_bwd)
    const float* records;    // (rows, 16), absolute row index
    const int* rowid;        // (rows) index into dout4, NULL = identity
    const float4* dout4;     // gradient w.r.t. (pre-sigmoid r, g, b, sigma) of every sample
    int row0, n_rows;        // this launch covers rows [row0, row0 + n_rows); tile t of the launch = rows row0 + 128 t ...
    uint8_t* scratch;        // ceil(n_rows / 128) * TILE_SCRATCH
    float* dfeat;            // (rows, 272) fp32, absolute row index
    float* dparams;          // flat parameter gradients (accumulated): only the sigma / rgb heads are touched here
};

__device__ __forceinline__ void step_shape(int k, int& npe, int& nh, int& slab) {
    if (k < 10) { npe = (k == 0 || k == 4) ? KX_STEPS : (k == 9 ? KD_STEPS : 0); nh = k > 0 ? 16 : 0; slab = k == 9 ? 4096 : 8192; }
    else {
        int rows, layer;
        bwd_gemm(k - 10, nh, rows, layer);
        npe = 0; slab = 2 * rows * 16;
    }
}

template <bool BF16>
__device__ __forceinline__ float half_bits_to_float(unsigned short v) {
    if (BF16) return __uint_as_float((uint32_t)v << 16);
    return __half2float(*reinterpret_cast<const __half*>(&v));
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
template <bool FROM_BF16>
__device__ __forceinline__ uint32_t to_bf16x2(uint32_t v) {      // two halves of the forward dtype -> two bf16
    if (FROM_BF16) return v;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    return pack2<true>(f.x, f.y);
}

template <bool FBF16>
__global__ void __launch_bounds__(DG_THREADS, 1) k_mlp_bwd_dgrad(const DgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (a.n_rows + TILE_M - 1) / TILE_M;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_hidden = s_base + DS_HIDDEN, s_pexyz = s_base + DS_PEXYZ, s_pedir = s_base + DS_PEDIR;
    const uint32_t s_wring = s_base + DS_WRING, s_bar = s_base + DS_BAR;
    float* sp = reinterpret_cast<float*>(smem + DS_SPARAM);
    float4* gs = reinterpret_cast<float4*>(smem + DS_GS);
    auto bar = [&](int i) { return s_bar + 8u * (uint32_t)i; };
    constexpr int B_WFULL = 0, B_WEMPTY = NST, B_AREADY = 2 * NST, B_ACCFULL = 2 * NST + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
        mbar_init(bar(B_AREADY), 128);
        mbar_init(bar(B_ACCFULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a.wf + W_BYTES);
        float4* dst = reinterpret_cast<float4*>(sp);
        for (int i = threadIdx.x; i < SP_FLOATS / 4; i += DG_THREADS) dst[i] = __ldg(src + i);
    }
    if (warp == 4) tmem_alloc<false>(s_base + DS_TMEM, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + DS_TMEM);
    const ParamOff PO = param_offsets();

    if (warp == 5) {
        // ================================================================ weight loader
        if (lane == 0) {
            uint32_t ws = 0, wph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t* srcf = a.wf;
                const uint8_t* srcb = a.wb;
                for (int k = 0; k < 20; ++k) {
                    int npe, nh, slab;
                    step_shape(k, npe, nh, slab);
                    if (k < 10) {
                        // forward weights: the forward kernel's unit stream (nf_mlp.cuh); every unit holds the two CTA-pair
                        // halves one after the other -- each is one 16 KB stage here
                        const uint32_t rpc = (k == 9) ? 32u : 64u;
                        for (int seg = 0; seg < 2; ++seg) {
                            const int nsteps = seg == 0 ? npe : nh;
                            for (int k0 = 0; k0 < nsteps; k0 += wu_ksteps(seg)) {
                                const uint32_t piece = (uint32_t)min(wu_ksteps(seg), nsteps - k0) * 2u * rpc * 16u;
                                for (int q = 0; q < 4; ++q) {          // (N-half, CTA half) = 4 pieces per K-block
                                    mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                                    mbar_arrive_expect_tx(bar(B_WFULL + ws), piece);
                                    bulk_g2s(s_wring + ws * WSTAGE, srcf, piece, bar(B_WFULL + ws));
                                    srcf += piece;
                                    if (++ws == NST) { ws = 0; wph ^= 1; }
                                }
                            }
                        }
                        continue;
                    }
                    const int fills = nh * 2;      // backward K-steps: a "hi" and a "lo" slab
                    for (int j = 0; j < fills; ++j) {
                        mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
                        mbar_arrive_expect_tx(bar(B_WFULL + ws), (uint32_t)slab);
                        bulk_g2s(s_wring + ws * WSTAGE, srcb, (uint32_t)slab, bar(B_WFULL + ws));
                        srcb += slab;
                        if (++ws == NST) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================================================================ MMA issuer (one lane)
        if (lane == 0) {
            uint32_t ws = 0, wph = 0, gstep = 0;
            const uint32_t idb256 = umma_idesc(256, true, TILE_M), idb208 = umma_idesc(208, true, TILE_M), idb64 = umma_idesc(64, true, TILE_M);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int k = 0; k < 20; ++k, ++gstep) {
                    mbar_wait(bar(B_AREADY), gstep & 1);
                    tc_fence_after();
                    int npe, nh, slab;
                    step_shape(k, npe, nh, slab);
                    uint32_t acc = 0;
                    if (k < 10) {
                        const uint32_t rpc = (k == 9) ? 32u : 64u;
                        const uint32_t idf = umma_idesc((int)rpc, FBF16, TILE_M);
                        uint32_t accq[4] = {0u, 0u, 0u, 0u};
                        for (int seg = 0; seg < 2; ++seg) {
                            const int nsteps = seg == 0 ? npe : nh;
                            const uint32_t abase = seg == 0 ? (k == 9 ? s_pedir : s_pexyz) : s_hidden;
                            for (int k0 = 0; k0 < nsteps; k0 += wu_ksteps(seg)) {
                                const int g = min(wu_ksteps(seg), nsteps - k0);
                                for (uint32_t q = 0; q < 4; ++q) {      // q = N-half * 2 + CTA half: output columns q * rpc ...
                                    mbar_wait(bar(B_WFULL + ws), wph);
                                    tc_fence_after();
                                    const uint32_t sw = s_wring + ws * WSTAGE;
                                    for (int j = 0; j < g; ++j) {
                                        umma_f16<false>(tmem_base + TM_H + q * rpc, umma_desc(abase + (uint32_t)(k0 + j) * 4096u, 2048u, 128u),
                                                        umma_desc(sw + (uint32_t)j * 2u * rpc * 16u, rpc * 16u, 128u), idf, accq[q]);
                                        accq[q] = 1;
                                    }
                                    umma_commit<false>(bar(B_WEMPTY + ws));
                                    if (++ws == NST) { ws = 0; wph ^= 1; }
                                }
                            }
                        }
                    }
                    for (int j = 0; k >= 10 && j < nh; ++j) {
                        mbar_wait(bar(B_WFULL + ws), wph);
                        tc_fence_after();
                        const uint32_t sw = s_wring + ws * WSTAGE;
                        {
                            // gradient tile = hi + lo (two bf16 tiles), weights = hi + lo (two slabs): hi*hi + lo*hi + hi*lo
                            const uint32_t ws_hi = ws;
                            if (++ws == NST) { ws = 0; wph ^= 1; }
                            mbar_wait(bar(B_WFULL + ws), wph);
                            tc_fence_after();
                            const uint32_t swl = s_wring + ws * WSTAGE;
                            const uint64_t ad_hi = umma_desc(s_hidden + (uint32_t)j * 4096u, 2048u, 128u);
                            const uint64_t ad_lo = umma_desc(s_base + DS_LO + (uint32_t)j * 4096u, 2048u, 128u);
                            const uint32_t lbo = (uint32_t)(slab >> 1);
                            for (int t = 0; t < 3; ++t) {
                                const uint64_t ad = t == 1 ? ad_lo : ad_hi;
                                const uint32_t wbase = t == 2 ? swl : sw;
                                const uint32_t ac = (t == 0) ? acc : 1u;
                                if (k == 10) {
                                    umma_f16<false>(tmem_base + TM_H, ad, umma_desc(wbase, lbo, 128u), idb256, ac);
                                    umma_f16<false>(tmem_base + TM_X, ad, umma_desc(wbase + 256u * 16u, lbo, 128u), idb64, ac);
                                } else if (k == 15) {
                                    umma_f16<false>(tmem_base + TM_H, ad, umma_desc(wbase, lbo, 128u), idb256, ac);
                                    umma_f16<false>(tmem_base + TM_X, ad, umma_desc(wbase + 256u * 16u, lbo, 128u), idb208, ac);
                                } else if (k == 19) {
                                    umma_f16<false>(tmem_base + TM_X, ad, umma_desc(wbase, lbo, 128u), idb208, 1u);   // += the skip layer's part
                                } else {
                                    umma_f16<false>(tmem_base + TM_H, ad, umma_desc(wbase, lbo, 128u), idb256, ac);
                                }
                            }
                            acc = 1;
                            umma_commit<false>(bar(B_WEMPTY + ws_hi));
                            umma_commit<false>(bar(B_WEMPTY + ws));
                            if (++ws == NST) { ws = 0; wph ^= 1; }
                        }
                    }
                    umma_commit<false>(bar(B_ACCFULL));
                }
            }
        }
    } else {
        // ================================================================ workers: thread = row of the tile
        const int tr = threadIdx.x;       // 0..127
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t gstep = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row = a.row0 + tile * TILE_M + tr;
            const bool valid = row < a.row0 + a.n_rows;
            uint8_t* img = a.scratch + (size_t)tile * TILE_SCRATCH + (size_t)tr * 16;     // + IMG_* + chunk * 2048
            float r[16];
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(a.records + (size_t)row * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t = __ldg(src + i);
                    r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
                }
                const int src_row = a.rowid ? a.rowid[row] : row;
                if (src_row >= 0) g = a.dout4[src_row];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0.f;
            }
            gs[tr] = g;
            {   // encoded features, exactly as the forward kernel writes them
                RowWriter<FBF16> w;
                w.base = s_pexyz + (uint32_t)tr * 16;
                emit_encoding<FBF16, 0, 3, 10>(w, r + 0);
                emit_encoding<FBF16, 63, 1, 4>(w, r + 3);
                emit_encoding<FBF16, 72, 3, 10>(w, r + 4);
                emit_encoding<FBF16, 135, 3, 10>(w, r + 7);
                static_for<198, 208>([&](auto ci) { w.template put<decltype(ci)::value>(0.f); });
                RowWriter<FBF16> wd;
                wd.base = s_pedir + (uint32_t)tr * 16;
                emit_encoding<FBF16, 0, 3, 4>(wd, r + 10);
                emit_encoding<FBF16, 27, 3, 4>(wd, r + 13);
                static_for<54, 64>([&](auto ci) { wd.template put<decltype(ci)::value>(0.f); });
            }
            // ... and their bf16 images for the weight gradients (each thread copies the chunks of its own row)
            for (int q = 0; q < 26; ++q) {
                const uint4 v = ld_shared_v4(s_pexyz + (uint32_t)q * CH + (uint32_t)tr * 16);
                *reinterpret_cast<uint4*>(img + IMG_XPE + q * CH) =
                    make_uint4(to_bf16x2<FBF16>(v.x), to_bf16x2<FBF16>(v.y), to_bf16x2<FBF16>(v.z), to_bf16x2<FBF16>(v.w));
            }
            for (int q = 0; q < 8; ++q) {
                const uint4 v = ld_shared_v4(s_pedir + (uint32_t)q * CH + (uint32_t)tr * 16);
                *reinterpret_cast<uint4*>(img + IMG_PD + q * CH) =
                    make_uint4(to_bf16x2<FBF16>(v.x), to_bf16x2<FBF16>(v.y), to_bf16x2<FBF16>(v.z), to_bf16x2<FBF16>(v.w));
            }
            fence_proxy_async();
            mbar_arrive(bar(B_AREADY));

            for (int k = 0; k < 20; ++k, ++gstep) {
                mbar_wait(bar(B_ACCFULL), gstep & 1);
                tc_fence_after();
                if (k < 9) {
                    // ---- forward layer k (8 = xyz_encoding_final: no ReLU): activations -> shared memory (next layer's A
                    //      operand, forward dtype) and -> HBM image (bf16)
                    const float* bias = sp + SP_BIAS + k * 256;
                    uint8_t* dst_img = img + (k < 8 ? img_h(k + 1) : IMG_F);
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        uint32_t v[32];
                        tmem_ld32(tlane + TM_H + c * 32, v);
                        tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            f[i] = __uint_as_float(v[i]) + bias[c * 32 + i];
                            if (k != 8) f[i] = fmaxf(f[i], 0.f);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            st_shared_v4(s_hidden + (uint32_t)(c * 4 + q) * CH + (uint32_t)tr * 16, pack2<FBF16>(f[8 * q], f[8 * q + 1]),
                                         pack2<FBF16>(f[8 * q + 2], f[8 * q + 3]), pack2<FBF16>(f[8 * q + 4], f[8 * q + 5]),
                                         pack2<FBF16>(f[8 * q + 6], f[8 * q + 7]));
                            *reinterpret_cast<uint4*>(dst_img + (c * 4 + q) * CH) =
                                make_uint4(pack2<true>(f[8 * q], f[8 * q + 1]), pack2<true>(f[8 * q + 2], f[8 * q + 3]),
                                           pack2<true>(f[8 * q + 4], f[8 * q + 5]), pack2<true>(f[8 * q + 6], f[8 * q + 7]));
                        }
                    }
                    if (k == 7) {
                        // sigma head (256 -> 1): d w_sigma[c] = sum_r dsigma_r h8[r][c], d b_sigma = sum_r dsigma_r
                        named_bar_sync(1, 128);
                        float s0 = 0.f, s1 = 0.f, sb = 0.f;
                        for (int rr = 0; rr < TILE_M; ++rr) {
                            const float ds = gs[rr].w;
                            const unsigned short h0 = *reinterpret_cast<const unsigned short*>(smem + DS_HIDDEN + (tr >> 3) * CH + rr * 16 + (tr & 7) * 2);
                            const unsigned short h1 = *reinterpret_cast<const unsigned short*>(smem + DS_HIDDEN + ((tr + 128) >> 3) * CH + rr * 16 + (tr & 7) * 2);
                            s0 += ds * half_bits_to_float<FBF16>(h0);
                            s1 += ds * half_bits_to_float<FBF16>(h1);
                            sb += ds;
                        }
                        atomicAdd(a.dparams + PO.w[10] + tr, s0);
                        atomicAdd(a.dparams + PO.w[10] + tr + 128, s1);
                        if (tr == 0) atomicAdd(a.dparams + PO.b[10], sb);
                    }
                } else if (k == 9) {
                    // ---- dir_encoding output d (128 columns, ReLU) -> bf16 image in shared memory + HBM; rgb head gradients
                    const float* bias = sp + SP_BIAS + 9 * 256;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld32(tlane + TM_H + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                pk[e] = pack2<true>(fmaxf(__uint_as_float(v[8 * q + 2 * e]) + bias[c * 32 + 8 * q + 2 * e], 0.f),
                                                    fmaxf(__uint_as_float(v[8 * q + 2 * e + 1]) + bias[c * 32 + 8 * q + 2 * e + 1], 0.f));
                            st_shared_v4(s_hidden + (uint32_t)(c * 4 + q) * CH + (uint32_t)tr * 16, pk[0], pk[1], pk[2], pk[3]);
                            *reinterpret_cast<uint4*>(img + IMG_D + (c * 4 + q) * CH) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    named_bar_sync(1, 128);
                    {   // d W_rgb[ch][c] = sum_r g_ch(r) d[r][c]   (thread = column c), d b_rgb = sum_r g(r)
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
                        for (int rr = 0; rr < TILE_M; ++rr) {
                            const float4 gg = gs[rr];
                            const float dv = half_bits_to_float<true>(
                                *reinterpret_cast<const unsigned short*>(smem + DS_HIDDEN + (tr >> 3) * CH + rr * 16 + (tr & 7) * 2));
                            s0 += gg.x * dv; s1 += gg.y * dv; s2 += gg.z * dv;
                            b0 += gg.x; b1 += gg.y; b2 += gg.z;
                        }
                        atomicAdd(a.dparams + PO.w[11] + tr, s0);
                        atomicAdd(a.dparams + PO.w[11] + 128 + tr, s1);
                        atomicAdd(a.dparams + PO.w[11] + 256 + tr, s2);
                        if (tr == 0) { atomicAdd(a.dparams + PO.b[11], b0); atomicAdd(a.dparams + PO.b[11] + 1, b1); atomicAdd(a.dparams + PO.b[11] + 2, b2); }
                    }
                    named_bar_sync(1, 128);
                    // dd[r][c] = (d > 0) * sum_ch g_ch W_rgb[ch][c]  -> A operand of the first backward GEMM
                    const float* wr = sp + SP_WRGB;
#pragma unroll 1
                    for (int q = 0; q < 16; ++q) {
                        const uint32_t addr = s_hidden + (uint32_t)q * CH + (uint32_t)tr * 16;
                        const uint4 dv = ld_shared_v4(addr);
                        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
                        uint32_t pk[4], pl[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c0 = q * 8 + 2 * e;
                            const float v0 = (dw[e] & 0x7fffu) ? g.x * wr[c0] + g.y * wr[128 + c0] + g.z * wr[256 + c0] : 0.f;
                            const float v1 = ((dw[e] >> 16) & 0x7fffu) ? g.x * wr[c0 + 1] + g.y * wr[128 + c0 + 1] + g.z * wr[256 + c0 + 1] : 0.f;
                            pk[e] = pack2<true>(v0, v1);
                            pl[e] = pack2<true>(v0 - __uint_as_float(pk[e] << 16), v1 - __uint_as_float(pk[e] & 0xffff0000u));
                        }
                        st_shared_v4(addr, pk[0], pk[1], pk[2], pk[3]);
                        st_shared_v4(s_base + DS_LO + (uint32_t)q * CH + (uint32_t)tr * 16, pl[0], pl[1], pl[2], pl[3]);
                        *reinterpret_cast<uint4*>(img + IMG_DD + q * CH) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                } else if (k == 19) {
                    // ---- gradient w.r.t. the xyz-like encoded features (layer 0's part + the skip layer's part)
#pragma unroll 1
                    for (int c = 0; c < 7; ++c) {           // 208 = 6 x 32 + 16
                        uint32_t v[32];
                        tmem_ld32(tlane + TM_X + c * 32, v);
                        tmem_ld_wait();
                        if (valid) {
                            float4* dst = reinterpret_cast<float4*>(a.dfeat + (size_t)row * DFEAT_W + c * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (c * 32 + i * 4 < 208)
                                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                         __uint_as_float(v[4 * i + 3]));
                        }
                    }
                } else {
                    // ---- k = 10 .. 18: a hidden-width gradient in TMEM -> (+ sigma term) -> ReLU mask -> bf16 image
                    if (k == 10) {       // the dir-like encoded features' gradient sits in the X region: 64 columns
#pragma unroll 1
                        for (int c = 0; c < 2; ++c) {
                            uint32_t v[32];
                            tmem_ld32(tlane + TM_X + c * 32, v);
                            tmem_ld_wait();
                            if (valid) {
                                float4* dst = reinterpret_cast<float4*>(a.dfeat + (size_t)row * DFEAT_W + 208 + c * 32);
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                         __uint_as_float(v[4 * i + 3]));
                            }
                        }
                    }
                    const int hi = 19 - k;                                   // k = 11..18: mask with H(hi), save as DP(hi)
                    const uint8_t* mask_img = k == 10 ? nullptr : img + img_h(hi);
                    uint8_t* dst_img = img + (k == 10 ? IMG_DF : img_dp(hi));
                    const float* wsig = sp + SP_WSIG;
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        uint32_t v[32];
                        tmem_ld32(tlane + TM_H + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * q + e]);
                            if (k == 11) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = fmaf(wsig[c * 32 + 8 * q + e], g.w, f[e]);
                            }
                            if (mask_img) {
                                const uint4 m = *reinterpret_cast<const uint4*>(mask_img + (c * 4 + q) * CH);
                                const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    if (!(mw[e] & 0x7fffu)) f[2 * e] = 0.f;
                                    if (!((mw[e] >> 16) & 0x7fffu)) f[2 * e + 1] = 0.f;
                                }
                            }
                            const uint32_t p0 = pack2<true>(f[0], f[1]), p1 = pack2<true>(f[2], f[3]), p2 = pack2<true>(f[4], f[5]),
                                           p3 = pack2<true>(f[6], f[7]);
                            auto lo2 = [](float a0, float a1, uint32_t hi) {
                                return pack2<true>(a0 - __uint_as_float(hi << 16), a1 - __uint_as_float(hi & 0xffff0000u));
                            };
                            st_shared_v4(s_hidden + (uint32_t)(c * 4 + q) * CH + (uint32_t)tr * 16, p0, p1, p2, p3);
                            st_shared_v4(s_base + DS_LO + (uint32_t)(c * 4 + q) * CH + (uint32_t)tr * 16, lo2(f[0], f[1], p0), lo2(f[2], f[3], p1),
                                         lo2(f[4], f[5], p2), lo2(f[6], f[7], p3));
                            *reinterpret_cast<uint4*>(dst_img + (c * 4 + q) * CH) = make_uint4(p0, p1, p2, p3);
                        }
                    }
                }
                if (k < 19) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(bar(B_AREADY));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<false>(tmem_base, 512);
}

// ---------------------------------------------------------------- wgrad kernel
struct WItem {
    int dy_off;      // byte offset of the dY chunks inside a tile's scratch (16 chunks = 128 output features)
    int x_off;       // byte offset of the X image
    int x_chunks;    // N / 8
    int n_valid;     // columns of X that are real input features
    int w_off;       // float offset of W[row0][col_off] in the flat gradient buffer
    int ld;          // leading dimension of W
    int b_off;       // float offset of bias[row0], or -1 (only one item per (layer, row block) sums the bias)
};
constexpr int N_WITEMS = 22;
struct WgradArgs {
    WItem items[N_WITEMS];
    const uint8_t* scratch;
    int ntiles, nsplit;
    float* dparams;
};
constexpr int WG_STAGE = 32768 + 65536;
constexpr int WG_THREADS = 6 * 32;
constexpr int WG_SMEM = 2 * WG_STAGE + 64;

__global__ void __launch_bounds__(WG_THREADS, 1) k_mlp_bwd_wgrad(const WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WItem it = a.items[blockIdx.x / a.nsplit];
    const int split = blockIdx.x % a.nsplit;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + 2 * WG_STAGE;
    auto bar = [&](int i) { return s_bar + 8u * (uint32_t)i; };      // 0,1 full; 2,3 empty; 4 done
    if (threadIdx.x == 0) {
        mbar_init(bar(0), 1); mbar_init(bar(1), 1);
        mbar_init(bar(2), 1 + 128); mbar_init(bar(3), 1 + 128);
        mbar_init(bar(4), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) tmem_alloc<false>(smem_u32(&tmem_slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    const int my_tiles = a.ntiles > split ? (a.ntiles - split + a.nsplit - 1) / a.nsplit : 0;
    const uint32_t xbytes = (uint32_t)it.x_chunks * CH;

    if (warp == 5) {
        if (lane == 0) {
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i & 1;
                const uint8_t* tb = a.scratch + (size_t)(split + i * a.nsplit) * TILE_SCRATCH;
                mbar_wait(bar(2 + s), ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(bar(s), 32768u + xbytes);
                bulk_g2s(s_base + s * WG_STAGE, tb + it.dy_off, 32768u, bar(s));
                bulk_g2s(s_base + s * WG_STAGE + 32768, tb + it.x_off, xbytes, bar(s));
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            // both operands MN-major (bits 15, 16): A = dY^T (M = 128 output features), B = X^T (N input features), K = rows
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)it.x_chunks << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            uint32_t acc = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i & 1;
                mbar_wait(bar(s), (i >> 1) & 1);
                tc_fence_after();
                const uint32_t sa = s_base + s * WG_STAGE, sb = sa + 32768;
#pragma unroll
                for (int j = 0; j < 8; ++j) {          // K-step j = rows 16 j .. 16 j + 15 of the tile
                    umma_f16<false>(tmem_base, umma_desc(sa + j * 256, 128u, 2048u), umma_desc(sb + j * 256, 128u, 2048u), idesc, acc);
                    acc = 1;
                }
                umma_commit<false>(bar(2 + s));
            }
            umma_commit<false>(bar(4));
        }
    } else {
        const int m = threadIdx.x;          // output feature of this 128-row block
        float bsum = 0.f;
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i & 1;
            mbar_wait(bar(s), (i >> 1) & 1);
            if (it.b_off >= 0) {
                const uint8_t* col = smem + s * WG_STAGE + (m >> 3) * CH + (m & 7) * 2;
                for (int rr = 0; rr < TILE_M; ++rr) bsum += half_bits_to_float<true>(*reinterpret_cast<const unsigned short*>(col + rr * 16));
            }
            mbar_arrive(bar(2 + s));
        }
        if (my_tiles > 0) {
            mbar_wait(bar(4), 0);
            tc_fence_after();
            float* wrow = a.dparams + it.w_off + (size_t)m * it.ld;
            for (int c0 = 0; c0 < it.x_chunks * 8; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < it.n_valid) atomicAdd(wrow + c0 + i, __uint_as_float(v[i]));
            }
            if (it.b_off >= 0) atomicAdd(a.dparams + it.b_off + m, bsum);
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<false>(tmem_base, 256);
}

static void build_items(WItem* it) {
    const ParamOff P = param_offsets();
    int n = 0;
    auto add = [&](int dy_img, int mb, int x_off, int x_chunks, int n_valid, int layer, int col_off, int ld, bool bias) {
        WItem w;
        w.dy_off = dy_img + mb * 16 * CH; w.x_off = x_off; w.x_chunks = x_chunks; w.n_valid = n_valid;
        w.w_off = P.w[layer] + mb * 128 * ld + col_off; w.ld = ld; w.b_off = bias ? P.b[layer] + mb * 128 : -1;
        it[n++] = w;
    };
    for (int mb = 0; mb < 2; ++mb) {
        add(img_dp(1), mb, IMG_XPE, 26, 198, 0, 0, 198, true);
        for (int l = 1; l <= 3; ++l) add(img_dp(l + 1), mb, img_h(l), 32, 256, l, 0, 256, true);
        add(img_dp(5), mb, IMG_XPE, 26, 198, 4, 0, 454, true);
        add(img_dp(5), mb, img_h(4), 32, 256, 4, 198, 454, false);
        for (int l = 5; l <= 7; ++l) add(img_dp(l + 1), mb, img_h(l), 32, 256, l, 0, 256, true);
        add(IMG_DF, mb, img_h(8), 32, 256, 8, 0, 256, true);
    }
    add(IMG_DD, 0, IMG_F, 32, 256, 9, 0, 310, true);
    add(IMG_DD, 0, IMG_PD, 8, 54, 9, 256, 310, false);
}

constexpr int CHUNK_TILES = 296;        // tiles per dgrad / wgrad round (scratch: 389 MB)

}  // namespace bwd
}  // namespace mlp
}  // namespace nf

using namespace nf;
using namespace nf::mlp;

extern "C" size_t nf_render_param_count(void) { return (size_t)bwd::param_offsets().total; }
extern "C" size_t nf_render_packed_weights_bwd_bytes(void) { return bwd::bwd_pack_bytes(); }

extern "C" int nf_render_pack_weights_bwd(const float* const* params, void* packed_out, void* stream_) {
    return nf_render_pack_weights_bwd_ex(params, NF_ENC_ALL, packed_out, stream_);
}

extern "C" int nf_render_pack_weights_bwd_ex(const float* const* params, int enc_flags, void* packed_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(enc_flags >= 0 && enc_flags <= NF_ENC_ALL, NF_E_INVALID, "nf_render_pack_weights_bwd: enc_flags %d", enc_flags);
    NF_REQUIRE(params && packed_out, NF_E_INVALID, "nf_render_pack_weights_bwd: null argument");
    bwd::PackArgs p;
    for (int i = 0; i < 12; ++i) {
        p.w[i] = params[2 * i];
        NF_REQUIRE(p.w[i], NF_E_INVALID, "nf_render_pack_weights_bwd: null parameter %d", i);
    }
    int total = 0;
    for (int g = 0; g < 10; ++g) {
        int ns, rows, l;
        bwd::bwd_gemm(g, ns, rows, l);
        total += ns * 2 * rows;      // one thread per 16-byte row piece; it writes the hi and the lo copy
    }
    bwd::k_pack_weights_bwd<<<(total + 255) / 256, 256, 0, st>>>(p, (uint8_t*)packed_out, enc_flags);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" size_t nf_nerf_mlp_backward_workspace_bytes(int n_rows) {
    if (n_rows <= 0) return 0;
    const int tiles = (n_rows + TILE_M - 1) / TILE_M;
    return (size_t)(tiles < bwd::CHUNK_TILES ? tiles : bwd::CHUNK_TILES) * bwd::TILE_SCRATCH;
}

extern "C" int nf_nerf_mlp_backward(const void* packed_fwd, const void* packed_bwd, int dtype, const float* records,
                                    const int32_t* rowid, const float* dout4, int n_rows, float* dfeat, float* dparams,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(n_rows >= 0, NF_E_INVALID, "nf_nerf_mlp_backward: negative row count");
    if (n_rows == 0) return NF_OK;
    NF_REQUIRE(packed_fwd && packed_bwd && records && dout4 && dfeat && dparams && workspace, NF_E_INVALID,
               "nf_nerf_mlp_backward: null pointer");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_nerf_mlp_backward: dtype %d", dtype);
    NF_REQUIRE(workspace_bytes >= nf_nerf_mlp_backward_workspace_bytes(n_rows), NF_E_WORKSPACE,
               "nf_nerf_mlp_backward: workspace %zu < %zu", workspace_bytes, nf_nerf_mlp_backward_workspace_bytes(n_rows));
    bwd::WgradArgs wa;
    bwd::build_items(wa.items);
    wa.scratch = (const uint8_t*)workspace;
    wa.dparams = dparams;
    const int chunk_rows = bwd::CHUNK_TILES * TILE_M;
    for (int r0 = 0; r0 < n_rows; r0 += chunk_rows) {
        const int nr = n_rows - r0 < chunk_rows ? n_rows - r0 : chunk_rows;
        const int ntiles = (nr + TILE_M - 1) / TILE_M;
        bwd::DgradArgs da;
        da.wf = (const uint8_t*)packed_fwd; da.wb = (const uint8_t*)packed_bwd;
        da.records = records; da.rowid = rowid; da.dout4 = (const float4*)dout4;
        da.row0 = r0; da.n_rows = nr;
        da.scratch = (uint8_t*)workspace; da.dfeat = dfeat; da.dparams = dparams;
        const int grid = ntiles < num_sms() ? ntiles : num_sms();
        if (dtype == NF_DTYPE_BF16) {
            NF_CUDA_OK(cudaFuncSetAttribute(bwd::k_mlp_bwd_dgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::DS_TOTAL));
            bwd::k_mlp_bwd_dgrad<true><<<grid, bwd::DG_THREADS, bwd::DS_TOTAL, st>>>(da);
        } else {
            NF_CUDA_OK(cudaFuncSetAttribute(bwd::k_mlp_bwd_dgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::DS_TOTAL));
            bwd::k_mlp_bwd_dgrad<false><<<grid, bwd::DG_THREADS, bwd::DS_TOTAL, st>>>(da);
        }
        NF_LAUNCH_OK();
        wa.ntiles = ntiles;
        wa.nsplit = ntiles < 6 ? ntiles : 6;
        NF_CUDA_OK(cudaFuncSetAttribute(bwd::k_mlp_bwd_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::WG_SMEM));
        bwd::k_mlp_bwd_wgrad<<<bwd::N_WITEMS * wa.nsplit, bwd::WG_THREADS, bwd::WG_SMEM, st>>>(wa);
        NF_LAUNCH_OK();
    }
    return NF_OK;
}
