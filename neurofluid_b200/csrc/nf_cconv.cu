// nf_cconv.cu -- Lagrangian transition model (hot path 2) on sm_100a.
//
// replaces (reference file:line):
//   ParticleNet.integrate_pos_vel / update_pos_vel      models/transmodel.py:100-104, 144-148
//   ParticleNet.compute_pose_correction                 models/transmodel.py:106-142
//     open3d.ml.torch.layers.ContinuousConv.forward x5  (:116, :118, :125)  incl. FixedRadiusSearch,
//       ball_to_cube_volume_preserving mapping, trilinear 4x4x4 filter interpolation, poly6 window
//     nn.Linear x4                                      (:117, :126)
//     ml3d.ops.reduce_subarrays_sum                     (:135-138)  -> neighbour count per particle
//
// Structure of one step (all on one stream, no host sync):
//   k_integrate        gravity half step                                   (N threads)
//   nf_grid_build      cell-sorted grid of the new positions (+ of the box points)
//   k_nbr_build        one warp per fluid particle: radius search in the 27-cell block, and for every
//                      neighbour the 8 trilinear corner (cell, weight*window) pairs of the 4^3 filter.
//                      The fluid->fluid list is reused by conv0_fluid, conv1, conv2, conv3; the
//                      fluid->box list by conv0_obstacle.  48-byte records, fixed stride per particle.
//   k_layer0           conv0_obstacle + conv0_fluid + dense0_fluid (14k MAC/particle): fp32 on CUDA cores
//   k_conv3_project/gather  conv3 + dense3 (64 -> 3): features projected through the 64 filter cells once per
//                      particle, then 8 x 3 floats gathered per neighbour pair (fp32)
//   k_cconv_tc<CIN,COUT>  conv_l + dense_l (+ residual) for l = 1, 2: per 128-particle tile, for each of the
//                      16 (z,y) filter rows the (128 x 4*CIN) slab of the patch matrix is accumulated in
//                      registers straight from the neighbour gather, written to shared memory as the fp16
//                      A operand (UMMA K-major core-matrix layout) and contracted with the matching filter
//                      slab by tcgen05.mma into one TMEM accumulator; the dense branch is a 17th slab; bias,
//                      residual and ReLU happen in the TMEM epilogue.  The (N x 64*CIN) patch matrix never
//                      exists in HBM.
//   k_update           pos/vel update
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

// ------------------------------------------------------------------------------------------------
__global__ void k_add_overflow(const int* __restrict__ flags, int* __restrict__ out) {
    if (threadIdx.x < 2 && flags[threadIdx.x]) atomicAdd(out + threadIdx.x, flags[threadIdx.x]);
}

__global__ void k_integrate(const float* __restrict__ pos, const float* __restrict__ vel, int n, float gx, float gy,
                            float gz, float dt, float* __restrict__ pos_new, float* __restrict__ vel_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 3) return;
    const int a = i % 3;
    const float g = a == 0 ? gx : (a == 1 ? gy : gz);
    const float v = vel[i];
    const float vn = v + g * dt;                  // models/transmodel.py:102
    vel_new[i] = vn;
    pos_new[i] = pos[i] + (v + vn) / 2 * dt;      // :103
}

__global__ void k_update(const float* __restrict__ pos, const float* __restrict__ pos_new, const float* __restrict__ ans3,
                         int ld3, int begin, int end, float dt, float* __restrict__ pos_out, float* __restrict__ vel_out,
                         float* __restrict__ delta_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = begin + t / 3, a = t % 3;
    if (i >= end) return;
    const float delta = ans3[(size_t)i * ld3 + a] * (1.0f / 128);   // :141
    const float pc = pos_new[3 * i + a] + delta;                     // :146
    pos_out[3 * i + a] = pc;
    vel_out[3 * i + a] = (pc - pos[3 * i + a]) / dt;                 // :147
    if (delta_out) delta_out[3 * i + a] = delta;
}

// ------------------------------------------------------------------------------------------------
// neighbour lists: one warp per output particle, row-scan of the <= 9 (y,z) rows of its cell block.
// d^2 <= r^2 (inclusive), points whose coordinates equal the query's are skipped (ignore_query_point).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_nbr_build(GridView g, const float* __restrict__ out_pos, int begin, int end,
                                                   float radius, int ignore_same, int use_window,
                                                   Pair* __restrict__ pairs, int* __restrict__ counts,
                                                   float* __restrict__ counts_f, int* __restrict__ overflow,
                                                   int* __restrict__ slab_j, float4* __restrict__ slab_w,
                                                   unsigned short* __restrict__ slab_off,
                                                   const float4* __restrict__ order /*NULL or cell-sorted copy: .w = particle index*/) {
    const int lane = threadIdx.x & 31;
    const int pos_i = begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (pos_i >= end) return;
    // neighbouring warps work on spatially neighbouring particles (cell order): their cell rows and records share L1 lines
    const int i = order ? __float_as_int(__ldg(&order[pos_i].w)) : pos_i;
    const GridHeader* h = g.hdr;
    const float qx = out_pos[3 * i], qy = out_pos[3 * i + 1], qz = out_pos[3 * i + 2];
    const float r2 = __fmul_rn(radius, radius);
    const float inv_radius = 1.0f / radius;
    const float pad = radius * 1.001f + 1e-6f;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    int n = 0;
    if (h->n > 0) {
        const int lox = cell_coord(qx - pad, ox, inv, nx), hix = cell_coord(qx + pad, ox, inv, nx);
        const int loy = cell_coord(qy - pad, oy, inv, ny), hiy = cell_coord(qy + pad, oy, inv, ny);
        const int loz = cell_coord(qz - pad, oz, inv, nz), hiz = cell_coord(qz + pad, oz, inv, nz);
        const unsigned lt = (1u << lane) - 1u;
        Pair* dst = pairs + (size_t)i * MAXNBR;
        for (int z = loz; z <= hiz; ++z)
            for (int y = loy; y <= hiy; ++y) {
                const int row = (z * ny + y) * nx;
                const int beg = __ldg(g.cell_start + row + lox), endc = __ldg(g.cell_start + row + hix + 1);
                for (int base = beg; base < endc; base += 32) {
                    const int t = base + lane;
                    bool hit = false;
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    float d2 = 0.f;
                    if (t < endc) {
                        p = __ldg(g.sorted + t);
                        const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
                        d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                        hit = d2 <= r2;
                        if (ignore_same && dx == 0.f && dy == 0.f && dz == 0.f) hit = false;
                    }
                    const unsigned m = __ballot_sync(NF_FULL, hit);
                    if (hit) {
                        const int slot = n + __popc(m & lt);
                        if (slot < MAXNBR) {
                            Pair pr;
                            pr.j = __float_as_int(p.w);
                            pr.pad = 0;
                            float a = 1.f;
                            if (use_window) {               // models/transmodel.py:73-77 on d^2 / r^2
                                const float u = 1.0f - d2 / r2;
                                a = fminf(fmaxf(u * u * u, 0.f), 1.f);
                            }
                            filter_corners(p.x - qx, p.y - qy, p.z - qz, inv_radius, a, pr);
                            dst[slot] = pr;
                        }
                    }
                    n += __popc(m);
                }
            }
    }
    if (lane == 0) {
        if (n > MAXNBR) atomicAdd(overflow, 1);
        counts[i] = min(n, MAXNBR);
        if (counts_f) counts_f[i] = (float)n;
    }
    if (slab_j) {
        // regroup this particle's pairs by filter row (stable: ballot ranks keep the pair order)
        __syncwarp();
        const int nn = min(n, MAXNBR);
        const Pair* src = pairs + (size_t)i * MAXNBR;
        int* dj = slab_j + (size_t)i * SLABCAP;
        float4* dw = slab_w + (size_t)i * SLABCAP;
        unsigned short* doff = slab_off + (size_t)i * SLABOFF;
        const unsigned lt = (1u << lane) - 1u;
        // which x cells of filter row `sidx` does a pair feed, and with what weight
        auto row_part = [](const uint4& h0, const float (&w)[8], int sidx, float (&wx)[4]) {
            bool rel = false;
            wx[0] = wx[1] = wx[2] = wx[3] = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
                if ((int)(cell >> 2) == sidx) {
                    rel = true;
#pragma unroll
                    for (int xx = 0; xx < 4; ++xx)
                        if ((int)(cell & 3) == xx) wx[xx] += w[c];
                }
            }
            return rel;
        };
        auto load_pair = [&](int t, uint4& h0, float (&w)[8]) {
            h0 = make_uint4(0u, 0xffffffffu, 0xffffffffu, 0u);      // cells 255: touches no row
#pragma unroll
            for (int c = 0; c < 8; ++c) w[c] = 0.f;
            if (t < nn) {
                h0 = *reinterpret_cast<const uint4*>(src + t);
                const float4 w0 = *(reinterpret_cast<const float4*>(src + t) + 1);
                const float4 w1 = *(reinterpret_cast<const float4*>(src + t) + 2);
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            }
        };
        int off = 0;
        if (nn <= 64) {        // the usual case: both 32-pair chunks stay in registers for all 16 rows
            // A pair's 8 corners are 4 (z,y) combinations x 2 x-neighbours: reduce them once to
            // {filter row, weight per x cell} x 4, so that the 16-row loop only compares and adds.
            auto combos = [&](int t, int& j, unsigned& rows, float (&cx)[4][4]) {
                uint4 h0;
                float w[8];
                load_pair(t, h0, w);
                j = (int)h0.x;
                rows = 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned c0 = ((q < 2 ? h0.y : h0.z) >> (16 * (q & 1))) & 0xffu;          // corner 2q
                    const unsigned c1 = ((q < 2 ? h0.y : h0.z) >> (16 * (q & 1) + 8)) & 0xffu;      // corner 2q + 1
                    rows |= (t < nn ? (c0 >> 2) : 31u) << (8 * q);                                  // 31: no row
#pragma unroll
                    for (int xx = 0; xx < 4; ++xx)
                        cx[q][xx] = ((int)(c0 & 3) == xx ? w[2 * q] : 0.f) + ((int)(c1 & 3) == xx ? w[2 * q + 1] : 0.f);
                }
            };
            int ja, jb;
            unsigned ra4, rb4;
            float ca[4][4], cb[4][4];
            combos(lane, ja, ra4, ca);
            combos(32 + lane, jb, rb4, cb);
            auto row_sum = [](unsigned rows, const float (&cx)[4][4], int sidx, float (&wx)[4]) {
                bool rel = false;
                wx[0] = wx[1] = wx[2] = wx[3] = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if ((int)((rows >> (8 * q)) & 0xffu) == sidx) {
                        rel = true;
                        wx[0] += cx[q][0]; wx[1] += cx[q][1]; wx[2] += cx[q][2]; wx[3] += cx[q][3];
                    }
                }
                return rel;
            };
#pragma unroll 1
            for (int sidx = 0; sidx < 16; ++sidx) {
                float xa[4], xb[4];
                const bool ra = row_sum(ra4, ca, sidx, xa), rb = row_sum(rb4, cb, sidx, xb);
                const unsigned ma = __ballot_sync(NF_FULL, ra), mb = __ballot_sync(NF_FULL, rb);
                if (lane == 0) doff[sidx] = (unsigned short)off;
                if (ra) {
                    const int e = off + __popc(ma & lt);
                    dj[e] = ja;
                    dw[e] = make_float4(xa[0], xa[1], xa[2], xa[3]);
                }
                off += __popc(ma);
                if (rb) {
                    const int e = off + __popc(mb & lt);
                    dj[e] = jb;
                    dw[e] = make_float4(xb[0], xb[1], xb[2], xb[3]);
                }
                off += __popc(mb);
            }
        } else {               // chunks re-read per row (L1-resident: <= 6 KB per particle)
#pragma unroll 1
            for (int sidx = 0; sidx < 16; ++sidx) {
                if (lane == 0) doff[sidx] = (unsigned short)off;
                for (int t0 = 0; t0 < nn; t0 += 32) {
                    uint4 h0;
                    float w[8], wx[4];
                    load_pair(t0 + lane, h0, w);
                    const bool rel = row_part(h0, w, sidx, wx);
                    const unsigned m = __ballot_sync(NF_FULL, rel);
                    if (rel) {
                        const int e = off + __popc(m & lt);
                        dj[e] = (int)h0.x;
                        dw[e] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                    }
                    off += __popc(m);
                }
            }
        }
        if (lane == 0) doff[16] = (unsigned short)off;
    }
}

// ------------------------------------------------------------------------------------------------
// layer 0: conv0_obstacle (3->32), conv0_fluid (4->32), dense0_fluid (4->32) -> ans0 (N,96) fp32 and
// x0 = relu(ans0) as fp16/bf16.   One warp per particle, fp32 CUDA cores.
// ------------------------------------------------------------------------------------------------
struct Layer0Args {
    const Pair* pairs_ff; const int* cnt_ff;
    const Pair* pairs_fb; const int* cnt_fb;
    const float* vel_new;        // (N,3): fluid feats = [1, vel]
    const float* box_normals;    // (M,3)
    const float* k_fluid;        // (64,4,32)
    const float* b_fluid;        // (32)
    const float* k_obst;         // (64,3,32)
    const float* b_obst;         // (32)
    const float* w_dense;        // (32,4)
    const float* b_dense;        // (32)
    float* ans0;                 // (N,96)
    void* x0;                    // (N,96) half/bf16
    int begin, end;
    int bf16;
    const float4* order;         // NULL or the fluid grid's cell-sorted copy (.w = particle index): work in cell order
};

__global__ void __launch_bounds__(256) k_layer0(const Layer0Args a) {
    __shared__ __align__(16) float sm_patch[8][NCELL * 4 + NCELL * 3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pos_i = a.begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (pos_i >= a.end) return;
    const int i = a.order ? __float_as_int(__ldg(&a.order[pos_i].w)) : pos_i;
    float* pf = sm_patch[wib];
    float* po = pf + NCELL * 4;
    // scatter: one lane per neighbour (records and features are fetched in parallel), shared-memory atomics
    for (int k = lane; k < NCELL * 7; k += 32) pf[k] = 0.f;
    __syncwarp();
    {
        const int n = a.cnt_ff[i];
        const Pair* pr = a.pairs_ff + (size_t)i * MAXNBR;
        for (int t = lane; t < n; t += 32) {
            const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));          // j, cells, pad
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
            const int j = (int)h0.x;
            const float f[4] = {1.0f, __ldg(a.vel_new + 3 * j), __ldg(a.vel_new + 3 * j + 1), __ldg(a.vel_new + 3 * j + 2)};
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) atomicAdd(pf + cell * 4 + ch, w[c] * f[ch]);
            }
        }
    }
    {
        const int n = a.cnt_fb[i];
        const Pair* pr = a.pairs_fb + (size_t)i * MAXNBR;
        for (int t = lane; t < n; t += 32) {
            const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
            const int j = (int)h0.x;
            const float f[3] = {__ldg(a.box_normals + 3 * j), __ldg(a.box_normals + 3 * j + 1), __ldg(a.box_normals + 3 * j + 2)};
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) atomicAdd(po + cell * 3 + ch, w[c] * f[ch]);
            }
        }
    }
    __syncwarp();
    // lane = output channel
    // four independent partial sums per output (a single chain of 256 dependent FMAs was latency-bound)
    float of, oo;
    {
        float s0 = __ldg(a.b_fluid + lane), s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
        for (int k = 0; k < NCELL * 4; k += 4) {
            const float4 p4 = *reinterpret_cast<const float4*>(pf + k);
            s0 += p4.x * __ldg(a.k_fluid + k * 32 + lane);
            s1 += p4.y * __ldg(a.k_fluid + (k + 1) * 32 + lane);
            s2 += p4.z * __ldg(a.k_fluid + (k + 2) * 32 + lane);
            s3 += p4.w * __ldg(a.k_fluid + (k + 3) * 32 + lane);
        }
        of = (s0 + s1) + (s2 + s3);
        s0 = __ldg(a.b_obst + lane); s1 = 0.f; s2 = 0.f; s3 = 0.f;
#pragma unroll 4
        for (int k = 0; k < NCELL * 3; k += 4) {
            const float4 p4 = *reinterpret_cast<const float4*>(po + k);
            s0 += p4.x * __ldg(a.k_obst + k * 32 + lane);
            s1 += p4.y * __ldg(a.k_obst + (k + 1) * 32 + lane);
            s2 += p4.z * __ldg(a.k_obst + (k + 2) * 32 + lane);
            s3 += p4.w * __ldg(a.k_obst + (k + 3) * 32 + lane);
        }
        oo = (s0 + s1) + (s2 + s3);
    }
    float od = __ldg(a.b_dense + lane) + __ldg(a.w_dense + lane * 4);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) od += __ldg(a.vel_new + 3 * i + ch) * __ldg(a.w_dense + lane * 4 + 1 + ch);
    float* o = a.ans0 + (size_t)i * 96;
    o[lane] = oo; o[32 + lane] = of; o[64 + lane] = od;       // cat[obstacle, fluid, dense]  (:120)
    if (a.bf16) {
        __nv_bfloat16* x = reinterpret_cast<__nv_bfloat16*>(a.x0) + (size_t)i * 96;
        x[lane] = __float2bfloat16(fmaxf(oo, 0.f)); x[32 + lane] = __float2bfloat16(fmaxf(of, 0.f));
        x[64 + lane] = __float2bfloat16(fmaxf(od, 0.f));
    } else {
        __half* x = reinterpret_cast<__half*>(a.x0) + (size_t)i * 96;
        x[lane] = __float2half(fmaxf(oo, 0.f)); x[32 + lane] = __float2half(fmaxf(of, 0.f));
        x[64 + lane] = __float2half(fmaxf(od, 0.f));
    }
    __syncwarp();
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

// ------------------------------------------------------------------------------------------------
// conv3 + dense3 (64 -> 3).  With 3 output channels the contraction is cheaper the other way round:
//   out_i = sum_j sum_c w_ijc K_c^T f_j  =  sum_j sum_c w_ijc g_j[c],   g_j[c] = K_c^T f_j  (3 values per cell)
// so every particle's features are projected through all 64 filter cells once (k_conv3_project: N x 64 x 192 MAC,
// fp32, weights in shared memory) and the neighbour pass gathers 8 x 3 floats per pair instead of building a
// 128 x 256 patch slab per filter row for 3 useful output columns (k_conv3_gather: one warp per particle, one
// lane per pair, fixed reduction tree -> deterministic).  fp32 end to end.
// ------------------------------------------------------------------------------------------------
constexpr int C3_IN = 64, C3_OUT = 3, C3_G = NCELL * C3_OUT;      // 192 projected values per particle

template <bool BF16>
__global__ void __launch_bounds__(C3_G) k_conv3_project(const void* __restrict__ x_in, int n, const float* __restrict__ kern,
                                                       float* __restrict__ g) {
    // kern: (64 cells, 64 in, 3 out) fp32 = the reference's conv3.kernel (4,4,4,64,3) flattened
    extern __shared__ float sk[];                     // [in][cell*3 + out]: thread q reads sk[ch*192 + q], conflict-free
    __shared__ float4 sx[C3_IN];                      // [in] x 4 particles of this pass
    for (int k = threadIdx.x; k < NCELL * C3_IN * C3_OUT; k += blockDim.x) {
        const int o = k % C3_OUT, ch = (k / C3_OUT) % C3_IN, cell = k / (C3_OUT * C3_IN);
        sk[ch * C3_G + cell * C3_OUT + o] = __ldg(kern + k);
    }
    const int q = threadIdx.x;
    for (int i0 = blockIdx.x * 4; i0 < n; i0 += gridDim.x * 4) {      // 4 particles per pass share every weight read
        __syncthreads();
        for (int k = threadIdx.x; k < 4 * C3_IN; k += blockDim.x) {
            const int ii = i0 + (k & 3), ch = k >> 2;
            float v = 0.f;
            if (ii < n) {
                if (BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x_in)[(size_t)ii * C3_IN + ch]);
                else v = __half2float(reinterpret_cast<const __half*>(x_in)[(size_t)ii * C3_IN + ch]);
            }
            reinterpret_cast<float*>(sx)[k] = v;
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int ch = 0; ch < C3_IN; ++ch) {
            const float w = sk[ch * C3_G + q];
            const float4 x = sx[ch];
            a0 += w * x.x; a1 += w * x.y; a2 += w * x.z; a3 += w * x.w;
        }
        if (i0 < n) g[(size_t)i0 * C3_G + q] = a0;
        if (i0 + 1 < n) g[(size_t)(i0 + 1) * C3_G + q] = a1;
        if (i0 + 2 < n) g[(size_t)(i0 + 2) * C3_G + q] = a2;
        if (i0 + 3 < n) g[(size_t)(i0 + 3) * C3_G + q] = a3;
    }
}

template <bool BF16>
__global__ void __launch_bounds__(256) k_conv3_gather(const Pair* __restrict__ pairs, const int* __restrict__ cnt,
                                                      const float* __restrict__ g, const void* __restrict__ x_in,
                                                      const float* __restrict__ b_conv, const float* __restrict__ w_dense,
                                                      const float* __restrict__ b_dense, int begin, int end,
                                                      float* __restrict__ ans3 /*(N,16)*/, const float4* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    const int pos_i = begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (pos_i >= end) return;
    const int i = order ? __float_as_int(__ldg(&order[pos_i].w)) : pos_i;
    float acc[C3_OUT] = {0.f, 0.f, 0.f};
    const int n = cnt[i];
    const Pair* pr = pairs + (size_t)i * MAXNBR;
    for (int t = lane; t < n; t += 32) {
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float* gj = g + (size_t)h0.x * C3_G;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
            const float* gc = gj + cell * C3_OUT;
            acc[0] += w[c] * __ldg(gc); acc[1] += w[c] * __ldg(gc + 1); acc[2] += w[c] * __ldg(gc + 2);
        }
    }
    // dense3 on the particle's own features: lane owns channels lane and lane + 32
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int ch = lane + 32 * m;
        float x;
        if (BF16) x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x_in)[(size_t)i * C3_IN + ch]);
        else x = __half2float(reinterpret_cast<const __half*>(x_in)[(size_t)i * C3_IN + ch]);
#pragma unroll
        for (int o = 0; o < C3_OUT; ++o) acc[o] += x * __ldg(w_dense + o * C3_IN + ch);
    }
#pragma unroll
    for (int o = 0; o < C3_OUT; ++o) acc[o] = warp_sum(acc[o]);
    if (lane < 16) {
        float v = 0.f;
        if (lane < C3_OUT) v = (lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2])) + __ldg(b_conv + lane) + __ldg(b_dense + lane);
        ans3[(size_t)i * 16 + lane] = v;
    }
}

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
__global__ void k_pack10(const float* __restrict__ pos, const float* __restrict__ vel, const float* __restrict__ nn,
                         const float* __restrict__ delta, int begin, int end, float* __restrict__ out10) {
    const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float* o = out10 + (size_t)i * 10;
    o[0] = pos[3 * i]; o[1] = pos[3 * i + 1]; o[2] = pos[3 * i + 2];
    o[3] = vel[3 * i]; o[4] = vel[3 * i + 1]; o[5] = vel[3 * i + 2];
    o[6] = nn ? nn[i] : 0.f;
    o[7] = delta ? delta[3 * i] : 0.f; o[8] = delta ? delta[3 * i + 1] : 0.f; o[9] = delta ? delta[3 * i + 2] : 0.f;
}
__global__ void k_unpack10(const float* __restrict__ out10, int n, float* __restrict__ pos, float* __restrict__ vel, float* __restrict__ nn,
                           float* __restrict__ delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* o = out10 + (size_t)i * 10;
    pos[3 * i] = o[0]; pos[3 * i + 1] = o[1]; pos[3 * i + 2] = o[2];
    vel[3 * i] = o[3]; vel[3 * i + 1] = o[4]; vel[3 * i + 2] = o[5];
    if (nn) nn[i] = o[6];
    if (delta) { delta[3 * i] = o[7]; delta[3 * i + 1] = o[8]; delta[3 * i + 2] = o[9]; }
}

template <int CIN, int COUT_PAD>
static int launch_conv(const ConvArgs& a, int dtype, cudaStream_t st) {
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


// ------------------------------------------------------------------------------------------------
// Operator-level ContinuousConv (nf_cconv_forward): one conv on arbitrary in / out point sets.
//   small shapes (cin * cout <= SMALL_MAX: the 4->32, 3->32, 64->3 layers): fp32 on CUDA cores, the whole filter in
//   shared memory, one warp per out point: patch (64 cells x cin) accumulated from the pair list, then contracted.
//   64/96 -> 64: the slab-list tensor-core kernel above with the dense slab switched off.
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_MAX = 768;

// features in one of three storage types (kind 0: fp32, 1: fp16, 2: bf16), row stride ld elements
__device__ __forceinline__ float load_feat(const void* p, size_t idx, int kind) {
    if (kind == 0) return __ldg(reinterpret_cast<const float*>(p) + idx);
    if (kind == 1) return __half2float(reinterpret_cast<const __half*>(p)[idx]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}

// patch[cell * cin + ch] = sum over the pair list of out point i of  w_c * feature_j[ch]   (one warp, patch in smem)
__device__ __forceinline__ void build_patch(const Pair* __restrict__ pr, int n, const void* __restrict__ in_feat, int ld_in, int kind,
                                            int cin, float* patch, int lane) {
    const int pn = NCELL * cin;
    for (int k = lane; k < pn; k += 32) patch[k] = 0.f;
    __syncwarp();
    for (int t0 = 0; t0 < n; t0 += 32) {
        uint4 h0 = make_uint4(0u, 0u, 0u, 0u);
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        if (t0 + lane < n) {
            h0 = __ldg(reinterpret_cast<const uint4*>(pr + t0 + lane));
            w0 = __ldg(reinterpret_cast<const float4*>(pr + t0 + lane) + 1);
            w1 = __ldg(reinterpret_cast<const float4*>(pr + t0 + lane) + 2);
        }
        const int m = min(32, n - t0);
        for (int u = 0; u < m; ++u) {            // pairs in list order: the patch sums run like a walk over the list
            const int j = __shfl_sync(NF_FULL, (int)h0.x, u);
            const unsigned c03 = __shfl_sync(NF_FULL, h0.y, u), c47 = __shfl_sync(NF_FULL, h0.z, u);
            float w[8];
            w[0] = __shfl_sync(NF_FULL, w0.x, u); w[1] = __shfl_sync(NF_FULL, w0.y, u);
            w[2] = __shfl_sync(NF_FULL, w0.z, u); w[3] = __shfl_sync(NF_FULL, w0.w, u);
            w[4] = __shfl_sync(NF_FULL, w1.x, u); w[5] = __shfl_sync(NF_FULL, w1.y, u);
            w[6] = __shfl_sync(NF_FULL, w1.z, u); w[7] = __shfl_sync(NF_FULL, w1.w, u);
            for (int ch = lane; ch < cin; ch += 32) {
                const float f = load_feat(in_feat, (size_t)j * ld_in + ch, kind);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const unsigned cell = ((c < 4 ? c03 : c47) >> (8 * (c & 3))) & 0xffu;
                    patch[cell * cin + ch] += w[c] * f;       // a lane owns its channels: no conflicts between lanes
                }
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) k_cconv_small(const Pair* __restrict__ pairs, const int* __restrict__ cnt,
                                                     const void* __restrict__ in_feat, int ld_in, int kind, int cin, int cout,
                                                     const float* __restrict__ kern /*(64, cin, cout)*/,
                                                     const float* __restrict__ bias /*(cout)*/, int n_out,
                                                     float* __restrict__ out /*(n_out, cout)*/) {
    extern __shared__ __align__(16) float sm[];
    float* sk = sm;                                   // 64 * cin * cout
    const int kn = NCELL * cin * cout, pn = NCELL * cin;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* patch = sm + kn + wib * pn;
    for (int k = threadIdx.x; k < kn; k += blockDim.x) sk[k] = __ldg(kern + k);
    __syncthreads();
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_out; i += nwarps) {
        build_patch(pairs + (size_t)i * MAXNBR, cnt[i], in_feat, ld_in, kind, cin, patch, lane);
        for (int co = 0; co < cout; ++co) {
            float acc = 0.f;
            for (int k = lane; k < pn; k += 32) acc += patch[k] * sk[k * cout + co];
            acc = warp_sum(acc);
            if (lane == 0) out[(size_t)i * cout + co] = acc + (bias ? __ldg(bias + co) : 0.f);
        }
        __syncwarp();
    }
}

template <bool BF16>
__global__ void k_to_half(const float* __restrict__ in, size_t n, void* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16(in[i]);
    else reinterpret_cast<__half*>(out)[i] = __float2half(in[i]);
}

__global__ void k_pairs_index(const Pair* __restrict__ pairs, const int* __restrict__ cnt, int n_out, int32_t* __restrict__ idx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out * MAXNBR) return;
    const int i = t / MAXNBR, k = t % MAXNBR;
    idx[t] = k < cnt[i] ? pairs[(size_t)i * MAXNBR + k].j : -1;
}

struct OpWs {
    size_t pairs, cnt, slab_j, slab_w, slab_off, x16, flags, total;
};
inline bool op_is_tc(int cin, int cout) { return (cin == 64 || cin == 96) && cout == 64; }
inline OpWs op_ws(int n_in, int n_out, int cin, int cout) {
    OpWs L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    const size_t N = (size_t)(n_out > 0 ? n_out : 1), NI = (size_t)(n_in > 0 ? n_in : 1);
    L.pairs = take(N * MAXNBR * sizeof(Pair)); L.cnt = take(N * 4);
    L.slab_j = L.slab_w = L.slab_off = L.x16 = 0;
    if (op_is_tc(cin, cout)) {
        L.slab_j = take(N * SLABCAP * 4); L.slab_w = take(N * SLABCAP * 16); L.slab_off = take(N * SLABOFF * 2);
        L.x16 = take(NI * cin * 2);
    }
    L.flags = take(256);
    L.total = o;
    return L;
}

// ================================================================================================
// Backward pass of ParticleNet.forward (training: loss.backward() at trainer/trainer_transmodel.py:197 and, through the
// renderer, trainer/trainer_e2e.py:277).  As in Open3D, a ContinuousConv has gradients w.r.t. its filter and its input
// features only -- positions enter the geometry without gradient and reach the loss through pos_new + delta
// (models/transmodel.py:146) and vel = (pos_out - pos) / dt (:147).
//
//   feature gradient of a fluid->fluid conv = the SAME conv over the same (symmetric) neighbour lists with the filter
//     flipped in all three axes and transposed: dX_j = sum_i sum_c w_ijc K_c g_i and w_ijc = w_ji,flip(c) because the
//     ball-to-cube map is odd and trilinear weights mirror.  conv1 / conv2 therefore reuse k_cconv_tc (bf16 operands:
//     gradients need the range) with re-packed weights; its epilogue applies the ReLU mask and adds the residual gradient.
//   filter gradient dK_c = sum_i P_i[c]^T g_i (P_i = the patch of layer inputs around particle i):
//     conv1 / conv2: k_cconv_wgrad rebuilds the patch slabs like the forward kernel and contracts them with the gradient
//       tile on tcgen05 with MN-major operands (K = the 128 particles of a tile), accumulating over tiles in TMEM;
//     small layers (4->32, 3->32, 64->3): fp32 patch in shared memory, outer product accumulated per block.
// ================================================================================================
struct CWgradArgs {
    const int* slab_j; const float4* slab_w; const unsigned short* slab_off;
    const void* x_in;      // (N, CIN) layer input, forward operand dtype
    const void* g;         // (N, 64) bf16: gradient w.r.t. the layer's pre-activation
    int n, ntiles, nsplit;
    const float4* order;   // NULL or the fluid grid's cell-sorted copy (.w = particle index): tiles in cell order
    float* dK;             // (64 cells, CIN, 64) accumulated
    float* dWd;            // (64, CIN) accumulated (nn.Linear layout)
};

template <int CIN>
struct WgCfg {
    static constexpr int KSLAB = 4 * CIN;
    static constexpr int MB = KSLAB / 128;
    static constexpr int SM_A = 0;                              // 128 x KSLAB bf16 (tile image)
    static constexpr int SM_G = SM_A + 128 * KSLAB * 2;          // 128 x 64 bf16 (tile image)
    static constexpr int SM_BAR = SM_G + 8 * 2048;
    static constexpr int SM_TOTAL = SM_BAR + 64;
};

template <int CIN, bool XBF16>
__global__ void __launch_bounds__(CONV_THREADS, 1) k_cconv_wgrad(const CWgradArgs a) {
    using C = WgCfg<CIN>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / a.nsplit, split = blockIdx.x % a.nsplit;      // s: filter row 0..15, 16 = dense branch
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_a = s_base + C::SM_A, s_g = s_base + C::SM_G, s_bar = s_base + C::SM_BAR;
    const uint32_t bar_a_ready = s_bar, bar_mma_done = s_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::SM_BAR + 32);
    if (threadIdx.x == 0) {
        mbar_init(bar_a_ready, WORKER_WARPS * 32);
        mbar_init(bar_mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WORKER_WARPS) tmem_alloc(smem_u32(tmem_slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const int my_tiles = a.ntiles > split ? (a.ntiles - split + a.nsplit - 1) / a.nsplit : 0;
    const int nmb = s < 16 ? C::MB : 1;

    if (warp == WORKER_WARPS) {
        if (lane == 0) {
            // A = patch slab^T (MN-major: M = slab column), B = gradient tile^T (MN-major: N = output channel), K = particle
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(bar_a_ready, it & 1);
                tc_fence_after();
                for (int mb = 0; mb < nmb; ++mb)
                    for (int j = 0; j < 8; ++j)
                        umma_f16(tmem_base + mb * 64, umma_desc(s_a + mb * 16 * 2048 + j * 256, 128, 2048),
                                 umma_desc(s_g + j * 256, 128, 2048), idesc, (it > 0 || j > 0) ? 1u : 0u);
                umma_commit(bar_mma_done);
            }
        }
    } else {
        // patch slab of filter row s for the tile's 128 particles: the forward kernel's worker loop (k_cconv_tc): NG particles
        // per warp at a time, GL lanes per particle, CPL channels per lane, EB entries per iteration with the next iteration's
        // {j, w} already in flight
        constexpr int GL = (CIN == 64) ? 8 : 16;
        constexpr int CPL = CIN / GL;
        constexpr int NG = 32 / GL;
        const int rbase = warp * ROWS_PER_WARP;
        const int gq = lane / GL, cl = lane % GL;
        const uint8_t* xin = reinterpret_cast<const uint8_t*>(a.x_in);
        auto cvt2 = [](uint32_t v, float& lo, float& hi) {
            if (XBF16) { lo = __uint_as_float(v << 16); hi = __uint_as_float(v & 0xffff0000u); }
            else { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&v)); lo = t.x; hi = t.y; }
        };
        auto packb = [](float lo, float hi) -> uint32_t {
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            return *reinterpret_cast<uint32_t*>(&h);
        };
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
        auto particle_of = [&](int tpos) -> int {      // tile position -> particle (cell order when the forward's grid is passed)
            if (tpos >= a.n) return -1;
            return a.order ? __float_as_int(__ldg(&a.order[tpos].w)) : tpos;
        };
        for (int it = 0; it < my_tiles; ++it) {
            const int row0 = (split + it * a.nsplit) * 128;
            if (it > 0) mbar_wait(bar_mma_done, (it - 1) & 1);       // the previous tile's operands have been consumed
            // gradient tile: (row, 8-column group) pieces of 16 bytes, row-major in HBM -> tile image
            for (int p = threadIdx.x; p < 128 * 8; p += WORKER_WARPS * 32) {
                const int r = p >> 3, q = p & 7;
                const int row = particle_of(row0 + r);
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (row >= 0) v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a.g) + (size_t)row * 128 + q * 16);
                *reinterpret_cast<uint4*>(smem + C::SM_G + q * 2048 + r * 16) = v;
            }
#pragma unroll 1
            for (int R = 0; R < ROWS_PER_WARP / NG; ++R) {
                const int rl = rbase + R * NG + gq;
                const int row = particle_of(row0 + rl);
                float acc[4][CPL];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int i = 0; i < CPL; ++i) acc[x][i] = 0.f;
                if (s < 16) {
                    int beg = 0, n = 0;
                    if (row >= 0) {
                        beg = __ldg(a.slab_off + (size_t)row * SLABOFF + s);
                        n = (int)__ldg(a.slab_off + (size_t)row * SLABOFF + s + 1) - beg;
                    }
                    const int nmax = __reduce_max_sync(NF_FULL, n);
                    const size_t ebase = (size_t)max(row, 0) * SLABCAP + beg;
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
                } else if (row >= 0) {
                    uint32_t f[CPL / 2];
                    load_feat(row, f);
#pragma unroll
                    for (int i = 0; i < CPL / 2; ++i) cvt2(f[i], acc[0][2 * i], acc[0][2 * i + 1]);
                }
                const int nx = (s < 16) ? 4 : 1;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (x < nx) {
                        const int k = x * CIN + cl * CPL;
                        if constexpr (CPL == 8) {
                            const uint32_t addr = s_a + (uint32_t)(k >> 3) * 2048 + (uint32_t)rl * 16;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packb(acc[x][0], acc[x][1])),
                                         "r"(packb(acc[x][2], acc[x][3])), "r"(packb(acc[x][4], acc[x][5])), "r"(packb(acc[x][6], acc[x][7]))
                                         : "memory");
                        } else {
#pragma unroll
                            for (int i = 0; i < CPL / 2; ++i) {
                                const int kk = k + 2 * i;
                                const uint32_t addr = s_a + (uint32_t)(kk >> 3) * 2048 + (uint32_t)rl * 16 + (uint32_t)(kk & 7) * 2;
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(packb(acc[x][2 * i], acc[x][2 * i + 1])) : "memory");
                            }
                        }
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_a_ready);
        }
        if (warp < 4 && my_tiles > 0) {
            mbar_wait(bar_mma_done, (my_tiles - 1) & 1);
            tc_fence_after();
            const int ml = warp * 32 + lane;
            for (int mb = 0; mb < nmb; ++mb) {
                const int m = mb * 128 + ml;
                const bool ok = s < 16 ? (m < C::KSLAB) : (m < CIN);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * 64 + c0, v);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (s < 16) atomicAdd(a.dK + ((size_t)s * C::KSLAB + m) * 64 + c0 + i, __uint_as_float(v[i]));
                            else atomicAdd(a.dWd + (size_t)(c0 + i) * CIN + m, __uint_as_float(v[i]));
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WORKER_WARPS) tmem_dealloc(tmem_base, 256);
}

// ---- conv3 (64 -> 3) backward.  With g_i = dL/d ans3_i (3 values) and G_j[c] = K_c^T f_j the forward's projection:
//   dL/dG_j[c] = sum over the pairs (i, j) of w_ijc g_i = sum over j's OWN list of w_jic' g_i with c' = 63 - c
// (fluid->fluid lists are symmetric and the filter coordinates of (j, i) mirror those of (i, j)), so one pass over the
// lists fills dG (N, 64 cells, 3) and everything else is dense:  dL/df_j = sum_c K_c dG_j[c],  dL/dK_c = sum_j f_j (x) dG_j[c].
__global__ void __launch_bounds__(256) k_conv3_bwd_scatter(const Pair* __restrict__ pairs, const int* __restrict__ cnt,
                                                           const float* __restrict__ g3 /*(N,3)*/, int n,
                                                           float* __restrict__ dG /*(N,192)*/, const float4* __restrict__ order) {
    __shared__ float acc[8][C3_G];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pos_j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pos_j >= n) return;
    const int j = order ? __float_as_int(__ldg(&order[pos_j].w)) : pos_j;
    float* a = acc[wib];
    for (int k = lane; k < C3_G; k += 32) a[k] = 0.f;
    __syncwarp();
    const int m = cnt[j];
    const Pair* pr = pairs + (size_t)j * MAXNBR;
    for (int t = lane; t < m; t += 32) {
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const int i = (int)h0.x;
        const float gx = __ldg(g3 + 3 * i), gy = __ldg(g3 + 3 * i + 1), gz = __ldg(g3 + 3 * i + 2);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned cell = (NCELL - 1) - (((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu);
            atomicAdd(a + cell * 3, w[c] * gx); atomicAdd(a + cell * 3 + 1, w[c] * gy); atomicAdd(a + cell * 3 + 2, w[c] * gz);
        }
    }
    __syncwarp();
    for (int k = lane; k < C3_G; k += 32) dG[(size_t)j * C3_G + k] = a[k];
}

// dense half of the conv3 backward, 16 particles per pass:  g_ans2 = (dG K^T + g3 Wd3) * (ans2 > 0) (fp32 + bf16) and the
// filter gradient dK3 (64,64,3) accumulated in registers over the block's particles (one atomic per element per block).
constexpr int C3B_TP = 16;
constexpr int C3B_SMEM = (C3_G * C3_IN + C3B_TP * C3_G + C3B_TP * C3_IN + C3B_TP * 4 + 3 * C3_IN) * 4;
__global__ void __launch_bounds__(256) k_conv3_bwd_dense(const float* __restrict__ dG, const void* __restrict__ x2, int xbf16,
                                                         const float* __restrict__ kern /*(64,64,3)*/,
                                                         const float* __restrict__ wd3 /*(3,64)*/, const float* __restrict__ g3,
                                                         const float* __restrict__ ans2, int n, float* __restrict__ g_ans2,
                                                         __nv_bfloat16* __restrict__ g_ans2_h, float* __restrict__ dK) {
    extern __shared__ __align__(16) float sm[];
    float* sK = sm;                                  // [q = cell*3 + o][ch]
    float* sG = sK + C3_G * C3_IN;                   // [p][q]
    float* sX = sG + C3B_TP * C3_G;                  // [p][ch]
    float* sg3 = sX + C3B_TP * C3_IN;                // [p][4]
    float* sW = sg3 + C3B_TP * 4;                    // [o][ch]
    for (int k = threadIdx.x; k < NCELL * C3_IN * C3_OUT; k += blockDim.x) {
        const int o = k % C3_OUT, ch = (k / C3_OUT) % C3_IN, cell = k / (C3_OUT * C3_IN);
        sK[(cell * C3_OUT + o) * C3_IN + ch] = __ldg(kern + k);
    }
    for (int k = threadIdx.x; k < 3 * C3_IN; k += blockDim.x) sW[k] = __ldg(wd3 + k);
    const int ch = threadIdx.x & 63, grp = threadIdx.x >> 6;       // a warp has one grp: the sG reads below are broadcasts
    float acc[48];
#pragma unroll
    for (int m = 0; m < 48; ++m) acc[m] = 0.f;
    for (int i0 = blockIdx.x * C3B_TP; i0 < n; i0 += gridDim.x * C3B_TP) {
        __syncthreads();
        for (int k = threadIdx.x; k < C3B_TP * C3_G; k += blockDim.x) {
            const int r = k / C3_G;
            sG[k] = i0 + r < n ? __ldg(dG + (size_t)i0 * C3_G + k) : 0.f;
        }
        for (int k = threadIdx.x; k < C3B_TP * C3_IN; k += blockDim.x) {
            const int r = k >> 6;
            float v = 0.f;
            if (i0 + r < n) {
                if (xbf16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x2)[(size_t)i0 * C3_IN + k]);
                else v = __half2float(reinterpret_cast<const __half*>(x2)[(size_t)i0 * C3_IN + k]);
            }
            sX[k] = v;
        }
        if (threadIdx.x < C3B_TP * 4) {
            const int r = threadIdx.x >> 2, o = threadIdx.x & 3;
            sg3[threadIdx.x] = (o < 3 && i0 + r < n) ? __ldg(g3 + (size_t)(i0 + r) * 3 + o) : 0.f;
        }
        __syncthreads();
        {   // feature gradient of particles grp*4 .. grp*4+3, channel ch
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
            const float* g0 = sG + (grp * 4) * C3_G;
#pragma unroll 8
            for (int q = 0; q < C3_G; ++q) {
                const float w = sK[q * C3_IN + ch];
                d0 += w * g0[q]; d1 += w * g0[C3_G + q]; d2 += w * g0[2 * C3_G + q]; d3 += w * g0[3 * C3_G + q];
            }
            const float d[4] = {d0, d1, d2, d3};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = grp * 4 + u, i = i0 + r;
                if (i < n) {
                    float v = d[u] + sg3[r * 4] * sW[ch] + sg3[r * 4 + 1] * sW[C3_IN + ch] + sg3[r * 4 + 2] * sW[2 * C3_IN + ch];
                    if (!(__ldg(ans2 + (size_t)i * C3_IN + ch) > 0.f)) v = 0.f;
                    g_ans2[(size_t)i * C3_IN + ch] = v;
                    g_ans2_h[(size_t)i * C3_IN + ch] = __float2bfloat16(v);
                }
            }
        }
        // filter gradient: this thread owns q = grp*48 .. +47 of channel ch
#pragma unroll 2
        for (int r = 0; r < C3B_TP; ++r) {
            const float x = sX[r * C3_IN + ch];
            const float4* gq = reinterpret_cast<const float4*>(sG + r * C3_G + grp * 48);
#pragma unroll
            for (int m = 0; m < 12; ++m) {
                const float4 gv = gq[m];
                acc[4 * m] += x * gv.x; acc[4 * m + 1] += x * gv.y; acc[4 * m + 2] += x * gv.z; acc[4 * m + 3] += x * gv.w;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 48; ++m) {
        const int q = grp * 48 + m;
        if (acc[m] != 0.f) atomicAdd(dK + ((size_t)(q / 3) * C3_IN + ch) * 3 + q % 3, acc[m]);
    }
}

// ---- layer 0 backward, filter side: dK0_fluid (64,4,32) = sum_i patch_f(i) (x) g_i[32:64], dK0_obstacle (64,3,32) likewise
// with g_i[0:32]; the two conv biases, dense0's weight (32,4) and bias ride along.  A warp rebuilds one particle's two
// patches as k_layer0 does (8 particles per pass), then all 256 threads add the pass into register accumulators.
struct L0WgradArgs {
    const Pair* pairs_ff; const int* cnt_ff;
    const Pair* pairs_fb; const int* cnt_fb;      // NULL without a container
    const float* vel_new; const float* box_normals;
    const float* g;                               // (N,96): d ans0 = [obstacle, fluid, dense]
    int n;
    float *dKf, *dbf, *dKo, *dbo, *dWd, *dbd;
    const float4* order;
};
__global__ void __launch_bounds__(256) k_layer0_wgrad(const L0WgradArgs a) {
    __shared__ __align__(16) float sm_patch[8][NCELL * 4 + NCELL * 3];
    __shared__ float sgr[8][96];
    __shared__ float sff[8][4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float accf[32], acco[24];
#pragma unroll
    for (int m = 0; m < 32; ++m) accf[m] = 0.f;
#pragma unroll
    for (int m = 0; m < 24; ++m) acco[m] = 0.f;
    float accb = 0.f, accd = 0.f;
    for (int p0 = blockIdx.x * 8; p0 < a.n; p0 += gridDim.x * 8) {
        __syncthreads();
        const int pos_i = p0 + wib;
        float* pf = sm_patch[wib];
        float* po = pf + NCELL * 4;
        for (int k = lane; k < NCELL * 7; k += 32) pf[k] = 0.f;
        __syncwarp();
        if (pos_i < a.n) {
            const int i = a.order ? __float_as_int(__ldg(&a.order[pos_i].w)) : pos_i;
            {
                const int m = a.cnt_ff[i];
                const Pair* pr = a.pairs_ff + (size_t)i * MAXNBR;
                for (int t = lane; t < m; t += 32) {
                    const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
                    const int j = (int)h0.x;
                    const float f[4] = {1.0f, __ldg(a.vel_new + 3 * j), __ldg(a.vel_new + 3 * j + 1), __ldg(a.vel_new + 3 * j + 2)};
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) atomicAdd(pf + cell * 4 + ch, w[c] * f[ch]);
                    }
                }
            }
            if (a.cnt_fb) {
                const int m = a.cnt_fb[i];
                const Pair* pr = a.pairs_fb + (size_t)i * MAXNBR;
                for (int t = lane; t < m; t += 32) {
                    const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
                    const int j = (int)h0.x;
                    const float f[3] = {__ldg(a.box_normals + 3 * j), __ldg(a.box_normals + 3 * j + 1), __ldg(a.box_normals + 3 * j + 2)};
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) atomicAdd(po + cell * 3 + ch, w[c] * f[ch]);
                    }
                }
            }
            for (int k = lane; k < 96; k += 32) sgr[wib][k] = __ldg(a.g + (size_t)i * 96 + k);
            if (lane < 4) sff[wib][lane] = lane == 0 ? 1.0f : __ldg(a.vel_new + 3 * i + lane - 1);
        } else {
            for (int k = lane; k < 96; k += 32) sgr[wib][k] = 0.f;
            if (lane < 4) sff[wib][lane] = 0.f;
        }
        __syncthreads();
        // thread (o = lane, grp = wib) owns patch rows grp*32.. of the fluid filter and grp*24.. of the obstacle filter
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const float go = sgr[r][lane], gf = sgr[r][32 + lane];
            const float4* qf = reinterpret_cast<const float4*>(sm_patch[r] + wib * 32);
            const float4* qo = reinterpret_cast<const float4*>(sm_patch[r] + NCELL * 4 + wib * 24);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float4 v = qf[m];
                accf[4 * m] += v.x * gf; accf[4 * m + 1] += v.y * gf; accf[4 * m + 2] += v.z * gf; accf[4 * m + 3] += v.w * gf;
            }
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const float4 v = qo[m];
                acco[4 * m] += v.x * go; acco[4 * m + 1] += v.y * go; acco[4 * m + 2] += v.z * go; acco[4 * m + 3] += v.w * go;
            }
            if (wib < 3) accb += sgr[r][wib * 32 + lane];                  // bias sums: obstacle, fluid, dense
            else if (wib < 7) accd += sgr[r][64 + lane] * sff[r][wib - 3];  // dense0 weight column wib - 3
        }
    }
#pragma unroll
    for (int m = 0; m < 32; ++m)
        if (accf[m] != 0.f) atomicAdd(a.dKf + (size_t)(wib * 32 + m) * 32 + lane, accf[m]);
    if (a.cnt_fb) {
#pragma unroll
        for (int m = 0; m < 24; ++m)
            if (acco[m] != 0.f) atomicAdd(a.dKo + (size_t)(wib * 24 + m) * 32 + lane, acco[m]);
    }
    if (wib == 0) { if (accb != 0.f) atomicAdd(a.dbo + lane, accb); }
    else if (wib == 1) { if (accb != 0.f) atomicAdd(a.dbf + lane, accb); }
    else if (wib == 2) { if (accb != 0.f) atomicAdd(a.dbd + lane, accb); }
    else if (wib < 7) { if (accd != 0.f) atomicAdd(a.dWd + lane * 4 + (wib - 3), accd); }
}

// ---- layer 0 backward, feature side (32 -> 4 through the mirrored, transposed fluid filter): the same project + gather
// split as conv3's forward.  H_i[c][ch] = sum_o K'[c][o][ch] g_i[o];  d ff_j = sum over j's list of w_jic H_i[c].
constexpr int L0B_IN = 32, L0B_G = NCELL * 4;
__global__ void __launch_bounds__(L0B_G) k_layer0_bwd_project(const float* __restrict__ g /*(N,96): columns 32..63*/, int n,
                                                              const float* __restrict__ kt /*(64,32,4)*/, float* __restrict__ H /*(N,256)*/) {
    __shared__ float sk[L0B_IN * L0B_G];              // [o][cell*4 + ch]
    __shared__ float4 sx[L0B_IN];
    for (int k = threadIdx.x; k < L0B_IN * L0B_G; k += blockDim.x) {
        const int ch = k & 3, o = (k >> 2) % L0B_IN, cell = k / (4 * L0B_IN);
        sk[o * L0B_G + cell * 4 + ch] = __ldg(kt + k);
    }
    const int q = threadIdx.x;
    for (int i0 = blockIdx.x * 4; i0 < n; i0 += gridDim.x * 4) {
        __syncthreads();
        if (threadIdx.x < 4 * L0B_IN) {
            const int ii = i0 + (threadIdx.x & 3), o = threadIdx.x >> 2;
            reinterpret_cast<float*>(sx)[threadIdx.x] = ii < n ? __ldg(g + (size_t)ii * 96 + 32 + o) : 0.f;
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int o = 0; o < L0B_IN; ++o) {
            const float w = sk[o * L0B_G + q];
            const float4 x = sx[o];
            a0 += w * x.x; a1 += w * x.y; a2 += w * x.z; a3 += w * x.w;
        }
        if (i0 < n) H[(size_t)i0 * L0B_G + q] = a0;
        if (i0 + 1 < n) H[(size_t)(i0 + 1) * L0B_G + q] = a1;
        if (i0 + 2 < n) H[(size_t)(i0 + 2) * L0B_G + q] = a2;
        if (i0 + 3 < n) H[(size_t)(i0 + 3) * L0B_G + q] = a3;
    }
}

__global__ void __launch_bounds__(256) k_layer0_bwd_gather(const Pair* __restrict__ pairs, const int* __restrict__ cnt,
                                                           const float* __restrict__ H, int n, float* __restrict__ out /*(N,4)*/,
                                                           const float4* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    const int pos_j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pos_j >= n) return;
    const int j = order ? __float_as_int(__ldg(&order[pos_j].w)) : pos_j;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int m = cnt[j];
    const Pair* pr = pairs + (size_t)j * MAXNBR;
    for (int t = lane; t < m; t += 32) {
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(pr + t));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(pr + t) + 1);
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(pr + t) + 2);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float4* hi = reinterpret_cast<const float4*>(H + (size_t)h0.x * L0B_G);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned cell = ((c < 4 ? h0.y : h0.z) >> (8 * (c & 3))) & 0xffu;
            const float4 v = __ldg(hi + cell);
            acc.x += w[c] * v.x; acc.y += w[c] * v.y; acc.z += w[c] * v.z; acc.w += w[c] * v.w;
        }
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0) *reinterpret_cast<float4*>(out + (size_t)j * 4) = acc;
}

// dW (cout, cin) += g^T x over all particles; db (cout) += column sums of g (optional, may alias a second target)
__global__ void __launch_bounds__(256) k_dense_wgrad(const float* __restrict__ g, int ld_g, int cout, const void* __restrict__ x, int ld_x,
                                                     int kind, int cin, int n, float* __restrict__ dW, float* __restrict__ db0,
                                                     float* __restrict__ db1) {
    extern __shared__ float sm[];
    constexpr int TP = 32;                     // particles per tile
    float* sg = sm;                            // TP x cout
    float* sx = sm + TP * cout;                // TP x cin
    const int nout = cout * cin;
    float accw[24];
#pragma unroll
    for (int t = 0; t < 24; ++t) accw[t] = 0.f;
    float accb = 0.f;
    for (int i0 = blockIdx.x * TP; i0 < n; i0 += gridDim.x * TP) {
        __syncthreads();
        for (int k = threadIdx.x; k < TP * cout; k += blockDim.x) {
            const int r = k / cout, c = k % cout;
            sg[k] = i0 + r < n ? g[(size_t)(i0 + r) * ld_g + c] : 0.f;
        }
        for (int k = threadIdx.x; k < TP * cin; k += blockDim.x) {
            const int r = k / cin, c = k % cin;
            sx[k] = i0 + r < n ? load_feat(x, (size_t)(i0 + r) * ld_x + c, kind) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 24; ++t) {
            const int o = threadIdx.x + t * 256;
            if (o < nout) {
                const int co = o / cin, ci = o % cin;
                float a = 0.f;
                for (int r = 0; r < TP; ++r) a += sg[r * cout + co] * sx[r * cin + ci];
                accw[t] += a;
            }
        }
        if ((int)threadIdx.x < cout)
            for (int r = 0; r < TP; ++r) accb += sg[r * cout + threadIdx.x];
    }
#pragma unroll
    for (int t = 0; t < 24; ++t) {
        const int o = threadIdx.x + t * 256;
        if (o < nout && accw[t] != 0.f) atomicAdd(dW + o, accw[t]);
    }
    if ((int)threadIdx.x < cout && accb != 0.f) {
        if (db0) atomicAdd(db0 + threadIdx.x, accb);
        if (db1) atomicAdd(db1 + threadIdx.x, accb);
    }
}

// K (64, cin, cout) -> K' (64, cout, cin) with the cell index mirrored: K'[c][co][ci] = K[63 - c][ci][co];  Wd (cout, cin) -> Wd^T
__global__ void k_flip_transpose(const float* __restrict__ K, int cin, int cout, float* __restrict__ Kt, const float* __restrict__ Wd,
                                 float* __restrict__ Wdt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < NCELL * cin * cout) {
        const int ci = t % cin, co = (t / cin) % cout, c = t / (cin * cout);
        Kt[t] = K[((size_t)(NCELL - 1 - c) * cin + ci) * cout + co];
    }
    if (Wd && t < cin * cout) {
        const int co = t % cout, ci = t / cout;
        Wdt[t] = Wd[(size_t)co * cin + ci];
    }
}

// start of the backward pass: gpt = g_pos_out + g_vel_out / dt;  g_ans3 = gpt / 128;  d_pos = g_pos_out;  d_vel = dt * gpt
// (+ the feature path, added by k_bwd_tail)
__global__ void k_bwd_head(const float* __restrict__ g_pos_out, const float* __restrict__ g_vel_out, const float* __restrict__ vel_new,
                           int n, float dt, float* __restrict__ g_ans3, float* __restrict__ d_pos, float* __restrict__ d_vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float gp = g_pos_out ? g_pos_out[3 * i + a] : 0.f, gv = g_vel_out ? g_vel_out[3 * i + a] : 0.f;
        const float gpt = gp + gv / dt;
        g_ans3[3 * i + a] = gpt * (1.0f / 128);
        d_pos[3 * i + a] = gp;
        d_vel[3 * i + a] = dt * gpt;
    }
}

// d_vel += conv0_fluid^T(g)[1:4] + (g_ans0[:, 64:96] Wd0)[1:4]     (fluid features are [1, vel_new]; vel_new = vel + g dt)
__global__ void k_bwd_tail(const float* __restrict__ g_ffc /*(N,4)*/, const float* __restrict__ g_ans0 /*(N,96)*/,
                           const float* __restrict__ wd0 /*(32,4)*/, int n, float* __restrict__ d_vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = g_ffc[4 * i + 1 + a];
        for (int o = 0; o < 32; ++o) v += g_ans0[(size_t)i * 96 + 64 + o] * wd0[o * 4 + 1 + a];
        d_vel[3 * i + a] += v;
    }
}

struct BwdPackLayout {
    size_t scratch_k, scratch_w, l1, l2, k0ft, total;
};
inline BwdPackLayout bwd_pack_layout() {
    BwdPackLayout L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    L.scratch_k = take((size_t)NCELL * 96 * 64 * 4);
    L.scratch_w = take((size_t)96 * 64 * 4);
    L.l1 = take(ConvCfg<64, 96>::PACKED_BYTES);
    L.l2 = take(ConvCfg<64, 64>::PACKED_BYTES);
    L.k0ft = take((size_t)NCELL * 32 * 4 * 4);
    L.total = o;
    return L;
}

struct BwdWsLayout {
    size_t g_ans3, h0, g_ans2, g_ans2_h, g_ans1, g_ans1_h, g_ans0, g_ffc, total;
};
inline BwdWsLayout bwd_ws_layout(int n) {
    BwdWsLayout L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.g_ans3 = take(N * 3 * 4); L.h0 = take(N * L0B_G * 4);
    L.g_ans2 = take(N * 64 * 4); L.g_ans2_h = take(N * 64 * 2);
    L.g_ans1 = take(N * 64 * 4); L.g_ans1_h = take(N * 64 * 2);
    L.g_ans0 = take(N * 96 * 4);
    L.g_ffc = take(N * 4 * 4);
    L.total = o;
    return L;
}

// flat parameter(-gradient) layout of ParticleNet: nf_transition_pack_weights order, tensors concatenated
struct TParamOff {
    int off[18], total;
};
inline TParamOff tparam_offsets() {
    const int sz[18] = {NCELL * 4 * 32, 32, NCELL * 3 * 32, 32, 32 * 4, 32, NCELL * 96 * 64, 64, 64 * 96, 64,
                        NCELL * 64 * 64, 64, 64 * 64, 64, NCELL * 64 * 3, 3, 3 * 64, 3};
    TParamOff P;
    int o = 0;
    for (int i = 0; i < 18; ++i) { P.off[i] = o; o += sz[i]; }
    P.total = o;
    return P;
}

template <int CIN>
static int launch_wgrad(const CWgradArgs& a, bool xbf16, cudaStream_t st) {
    using C = WgCfg<CIN>;
    if (a.ntiles <= 0) return NF_OK;
    if (xbf16) {
        NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_wgrad<CIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SM_TOTAL));
        k_cconv_wgrad<CIN, true><<<17 * a.nsplit, CONV_THREADS, C::SM_TOTAL, st>>>(a);
    } else {
        NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_wgrad<CIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SM_TOTAL));
        k_cconv_wgrad<CIN, false><<<17 * a.nsplit, CONV_THREADS, C::SM_TOTAL, st>>>(a);
    }
    NF_LAUNCH_OK();
    return NF_OK;
}

static int launch_small(const Pair* pairs, const int* cnt, const void* in_feat, int ld_in, int kind, int cin, int cout, const float* kern,
                        int n_out, float* out, cudaStream_t st) {
    const size_t smem = (size_t)NCELL * cin * cout * 4 + (size_t)8 * NCELL * cin * 4;
    NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cconv_small<<<min((n_out + 7) / 8, 2 * num_sms()), 256, smem, st>>>(pairs, cnt, in_feat, ld_in, kind, cin, cout, kern, nullptr, n_out, out);
    NF_LAUNCH_OK();
    return NF_OK;
}

static int launch_dense_wgrad(const float* g, int ld_g, int cout, const void* x, int ld_x, int kind, int cin, int n, float* dW, float* db0,
                              float* db1, cudaStream_t st) {
    const size_t smem = (size_t)32 * (cout + cin) * 4;
    k_dense_wgrad<<<min((n + 31) / 32, num_sms()), 256, smem, st>>>(g, ld_g, cout, x, ld_x, kind, cin, n, dW, db0, db1);
    NF_LAUNCH_OK();
    return NF_OK;
}

}  // namespace cconv
}  // namespace nf

using namespace nf;
using namespace nf::cconv;

extern "C" size_t nf_transition_packed_weights_bytes(void) { return packed_layout().total; }

extern "C" int nf_transition_pack_weights(const float* const* p, int dtype, void* packed_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(p && packed_out, NF_E_INVALID, "nf_transition_pack_weights: null argument");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_transition_pack_weights: dtype %d", dtype);
    for (int i = 0; i < 18; ++i) NF_REQUIRE(p[i], NF_E_INVALID, "nf_transition_pack_weights: null parameter %d", i);
    const PackedLayout L = packed_layout();
    uint8_t* b = (uint8_t*)packed_out;
    // order: conv0_fluid.{kernel,bias}, conv0_obstacle.{kernel,bias}, dense0_fluid.{weight,bias},
    //        conv1.{k,b}, dense1.{w,b}, conv2.{k,b}, dense2.{w,b}, conv3.{k,b}, dense3.{w,b}
    NF_CUDA_OK(cudaMemcpyAsync(b + L.k_fluid, p[0], 64 * 4 * 32 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.b_fluid, p[1], 32 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.k_obst, p[2], 64 * 3 * 32 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.b_obst, p[3], 32 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.w_dense0, p[4], 32 * 4 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.b_dense0, p[5], 32 * 4, cudaMemcpyDeviceToDevice, st));
    const bool bf = dtype == NF_DTYPE_BF16;
    {
        using C = ConvCfg<96, 64>;
        const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 64;
        if (bf) k_pack_conv<96, 64, true><<<(tot + 255) / 256, 256, 0, st>>>(p[6], p[7], p[8], p[9], 64, b + L.l1);
        else k_pack_conv<96, 64, false><<<(tot + 255) / 256, 256, 0, st>>>(p[6], p[7], p[8], p[9], 64, b + L.l1);
        NF_LAUNCH_OK();
    }
    {
        using C = ConvCfg<64, 64>;
        const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 64;
        if (bf) k_pack_conv<64, 64, true><<<(tot + 255) / 256, 256, 0, st>>>(p[10], p[11], p[12], p[13], 64, b + L.l2);
        else k_pack_conv<64, 64, false><<<(tot + 255) / 256, 256, 0, st>>>(p[10], p[11], p[12], p[13], 64, b + L.l2);
        NF_LAUNCH_OK();
    }
    NF_CUDA_OK(cudaMemcpyAsync(b + L.k3, p[14], NCELL * 64 * 3 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.b3, p[15], 3 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.w_dense3, p[16], 3 * 64 * 4, cudaMemcpyDeviceToDevice, st));
    NF_CUDA_OK(cudaMemcpyAsync(b + L.b_dense3, p[17], 3 * 4, cudaMemcpyDeviceToDevice, st));
    return NF_OK;
}

extern "C" size_t nf_transition_workspace_bytes(int n_fluid, int n_box) {
    if (n_fluid < 0 || n_box < 0) return 0;
    return ws_layout(n_fluid, n_box).total;
}


// ------------------------------------------------------------------------------------------------
// operator-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" size_t nf_cconv_packed_weights_bytes(int cin, int cout) {
    if (cin <= 0 || cout <= 0) return 0;
    if (op_is_tc(cin, cout)) return cin == 96 ? ConvCfg<96, 64>::PACKED_BYTES : ConvCfg<64, 64>::PACKED_BYTES;
    if (cin * cout > SMALL_MAX) return 0;
    return align_up((size_t)NCELL * cin * cout * 4, 256) + align_up((size_t)cout * 4, 256);
}

extern "C" int nf_cconv_pack_weights(const float* kernel, const float* bias, int cin, int cout, int dtype, void* packed_out,
                                     void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(kernel && packed_out, NF_E_INVALID, "nf_cconv_pack_weights: null argument");
    NF_REQUIRE(dtype == NF_DTYPE_F16 || dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_cconv_pack_weights: dtype %d", dtype);
    NF_REQUIRE(nf_cconv_packed_weights_bytes(cin, cout) > 0, NF_E_UNSUPPORTED,
               "nf_cconv_pack_weights: %d -> %d channels (supported: cin*cout <= %d in fp32, 64/96 -> 64 on tensor cores)", cin, cout, SMALL_MAX);
    uint8_t* b = (uint8_t*)packed_out;
    const bool bf = dtype == NF_DTYPE_BF16;
    if (op_is_tc(cin, cout)) {
        if (cin == 96) {
            using C = ConvCfg<96, 64>;
            const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 64;
            if (bf) k_pack_conv<96, 64, true><<<(tot + 255) / 256, 256, 0, st>>>(kernel, bias, nullptr, nullptr, 64, b);
            else k_pack_conv<96, 64, false><<<(tot + 255) / 256, 256, 0, st>>>(kernel, bias, nullptr, nullptr, 64, b);
        } else {
            using C = ConvCfg<64, 64>;
            const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 64;
            if (bf) k_pack_conv<64, 64, true><<<(tot + 255) / 256, 256, 0, st>>>(kernel, bias, nullptr, nullptr, 64, b);
            else k_pack_conv<64, 64, false><<<(tot + 255) / 256, 256, 0, st>>>(kernel, bias, nullptr, nullptr, 64, b);
        }
        NF_LAUNCH_OK();
        return NF_OK;
    }
    const size_t kb = (size_t)NCELL * cin * cout * 4;
    NF_CUDA_OK(cudaMemcpyAsync(b, kernel, kb, cudaMemcpyDeviceToDevice, st));
    if (bias) NF_CUDA_OK(cudaMemcpyAsync(b + align_up(kb, 256), bias, (size_t)cout * 4, cudaMemcpyDeviceToDevice, st));
    else NF_CUDA_OK(cudaMemsetAsync(b + align_up(kb, 256), 0, (size_t)cout * 4, st));
    return NF_OK;
}

extern "C" size_t nf_cconv_workspace_bytes(int n_in, int n_out, int cin, int cout) {
    if (n_in < 0 || n_out < 0 || cin <= 0 || cout <= 0) return 0;
    return op_ws(n_in, n_out, cin, cout).total;
}

extern "C" int nf_cconv_forward(const nf_cconv_args* a, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(a != nullptr, NF_E_INVALID, "nf_cconv_forward: null args");
    NF_REQUIRE(a->n_in >= 0 && a->n_out >= 0 && a->cin > 0 && a->cout > 0, NF_E_INVALID, "nf_cconv_forward: bad sizes");
    NF_REQUIRE(nf_cconv_packed_weights_bytes(a->cin, a->cout) > 0, NF_E_UNSUPPORTED, "nf_cconv_forward: %d -> %d channels unsupported",
               a->cin, a->cout);
    if (a->n_out == 0) return NF_OK;
    NF_REQUIRE(a->grid_in && a->out_pos && a->weights && a->out && a->workspace, NF_E_INVALID, "nf_cconv_forward: null pointer");
    NF_REQUIRE(a->n_in == 0 || a->in_feat, NF_E_INVALID, "nf_cconv_forward: null features");
    NF_REQUIRE(a->extent > 2e-3f, NF_E_INVALID, "nf_cconv_forward: bad extent");
    const OpWs L = op_ws(a->n_in, a->n_out, a->cin, a->cout);
    NF_REQUIRE(a->workspace_bytes >= L.total, NF_E_WORKSPACE, "nf_cconv_forward: workspace %zu < %zu", a->workspace_bytes, L.total);
    char* b = (char*)a->workspace;
    Pair* pairs = (Pair*)(b + L.pairs);
    int* cnt = (int*)(b + L.cnt);
    int* flags = (int*)(b + L.flags);
    const bool tc = op_is_tc(a->cin, a->cout);
    const float radius = 0.5f * a->extent;
    NF_CUDA_OK(cudaMemsetAsync(flags, 0, 256, st));
    const int blocks = (a->n_out + 7) / 8;
    k_nbr_build<<<blocks, 256, 0, st>>>(grid_view(a->grid_in, a->n_in), a->out_pos, 0, a->n_out, radius, a->ignore_same != 0,
                                       a->use_window != 0, pairs, cnt, a->count_out, flags,
                                       tc ? (int*)(b + L.slab_j) : nullptr, tc ? (float4*)(b + L.slab_w) : nullptr,
                                       tc ? (unsigned short*)(b + L.slab_off) : nullptr, nullptr);
    NF_LAUNCH_OK();
    if (a->overflow_out) {
        k_add_overflow<<<1, 32, 0, st>>>(flags, a->overflow_out);
        NF_LAUNCH_OK();
    }
    if (a->nbr_index_out) {
        k_pairs_index<<<(a->n_out * MAXNBR + 255) / 256, 256, 0, st>>>(pairs, cnt, a->n_out, a->nbr_index_out);
        NF_LAUNCH_OK();
    }
    if (tc) {
        const size_t nx = (size_t)a->n_in * a->cin;
        if (nx) {
            if (a->dtype == NF_DTYPE_BF16) k_to_half<true><<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(a->in_feat, nx, b + L.x16);
            else k_to_half<false><<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(a->in_feat, nx, b + L.x16);
            NF_LAUNCH_OK();
        }
        ConvArgs c;
        c.slab_j = (const int*)(b + L.slab_j); c.slab_w = (const float4*)(b + L.slab_w);
        c.slab_off = (const unsigned short*)(b + L.slab_off);
        c.x_in = b + L.x16; c.w_packed = (const uint8_t*)a->weights; c.residual = nullptr; c.ld_res = 0;
        c.ans = a->out; c.x_out = nullptr; c.n = a->n_in; c.begin = 0; c.end = a->n_out; c.cout = 64; c.dense = 0;
        c.mask_src = nullptr; c.ld_mask = 0; c.relu_out = 1; c.order = nullptr; c.tile_rows = 0;
        return a->cin == 96 ? launch_conv<96, 64>(c, a->dtype, st) : launch_conv<64, 64>(c, a->dtype, st);
    }
    const size_t kb = (size_t)NCELL * a->cin * a->cout * 4;
    const size_t smem = kb + (size_t)8 * NCELL * a->cin * 4;
    NF_CUDA_OK(cudaFuncSetAttribute(k_cconv_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = min(blocks, 2 * num_sms());
    k_cconv_small<<<grid, 256, smem, st>>>(pairs, cnt, a->in_feat, a->cin, 0, a->cin, a->cout, (const float*)a->weights,
                                          (const float*)((const char*)a->weights + align_up(kb, 256)), a->n_out, a->out);
    NF_LAUNCH_OK();
    return NF_OK;
}

// ------------------------------------------------------------------------------------------------
// backward entry points
// ------------------------------------------------------------------------------------------------
extern "C" size_t nf_transition_param_count(void) { return (size_t)tparam_offsets().total; }
extern "C" size_t nf_transition_packed_weights_bwd_bytes(void) { return bwd_pack_layout().total; }
extern "C" size_t nf_transition_backward_workspace_bytes(int n_fluid) { return n_fluid < 0 ? 0 : bwd_ws_layout(n_fluid).total; }

extern "C" int nf_transition_pack_weights_bwd(const float* const* p, void* packed_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(p && packed_out, NF_E_INVALID, "nf_transition_pack_weights_bwd: null argument");
    for (int i = 0; i < 18; ++i) NF_REQUIRE(p[i], NF_E_INVALID, "nf_transition_pack_weights_bwd: null parameter %d", i);
    const BwdPackLayout L = bwd_pack_layout();
    uint8_t* b = (uint8_t*)packed_out;
    float* sk = (float*)(b + L.scratch_k);
    float* sw = (float*)(b + L.scratch_w);
    // conv1 (96 -> 64) backward: a 64 -> 96 conv with the flipped, transposed filter and dense1^T
    k_flip_transpose<<<(NCELL * 96 * 64 + 255) / 256, 256, 0, st>>>(p[6], 96, 64, sk, p[8], sw);
    NF_LAUNCH_OK();
    {
        using C = ConvCfg<64, 96>;
        const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 96;
        k_pack_conv<64, 96, true><<<(tot + 255) / 256, 256, 0, st>>>(sk, nullptr, sw, nullptr, 96, b + L.l1);
        NF_LAUNCH_OK();
    }
    k_flip_transpose<<<(NCELL * 64 * 64 + 255) / 256, 256, 0, st>>>(p[10], 64, 64, sk, p[12], sw);
    NF_LAUNCH_OK();
    {
        using C = ConvCfg<64, 64>;
        const int tot = (16 * C::KSTEPS + C::KSTEPS_DENSE) * 2 * 64;
        k_pack_conv<64, 64, true><<<(tot + 255) / 256, 256, 0, st>>>(sk, nullptr, sw, nullptr, 64, b + L.l2);
        NF_LAUNCH_OK();
    }
    k_flip_transpose<<<(NCELL * 4 * 32 + 255) / 256, 256, 0, st>>>(p[0], 4, 32, (float*)(b + L.k0ft), nullptr, nullptr);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" int nf_transition_backward(const nf_transition_bwd_args* b, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(b && b->fwd, NF_E_INVALID, "nf_transition_backward: null args");
    const nf_transition_args* a = b->fwd;
    NF_REQUIRE(a->phase == -1, NF_E_UNSUPPORTED, "nf_transition_backward: only the whole-step forward (phase -1) has a backward");
    const int N = a->n_fluid, M = a->n_box;
    if (N == 0) return NF_OK;
    NF_REQUIRE(a->workspace && a->weights && b->weights_bwd && b->workspace && b->d_pos && b->d_vel && b->d_params, NF_E_INVALID,
               "nf_transition_backward: null pointer");
    const WsLayout L = ws_layout(N, M);
    const BwdWsLayout B = bwd_ws_layout(N);
    NF_REQUIRE(b->workspace_bytes >= B.total, NF_E_WORKSPACE, "nf_transition_backward: workspace %zu < %zu", b->workspace_bytes, B.total);
    const PackedLayout PL = packed_layout();
    const BwdPackLayout BL = bwd_pack_layout();
    const TParamOff PO = tparam_offsets();
    char* ws = (char*)a->workspace;
    char* bw = (char*)b->workspace;
    const uint8_t* w = (const uint8_t*)a->weights;
    const uint8_t* wb = (const uint8_t*)b->weights_bwd;
    float* dP = b->d_params;
    const Pair* pairs_ff = (const Pair*)(ws + L.pairs_ff); const int* cnt_ff = (const int*)(ws + L.cnt_ff);
    const Pair* pairs_fb = (const Pair*)(ws + L.pairs_fb); const int* cnt_fb = (const int*)(ws + L.cnt_fb);
    const float* vel_new = (const float*)(ws + L.vel_new);
    const float* ans0 = (const float*)(ws + L.ans0); const void* x0 = ws + L.x0;
    const float* ans1 = (const float*)(ws + L.ans1); const void* x1 = ws + L.x1;
    const float* ans2 = (const float*)(ws + L.ans2); const void* x2 = ws + L.x2;
    float* g_ans3 = (float*)(bw + B.g_ans3);
    float* g_ans2 = (float*)(bw + B.g_ans2); __nv_bfloat16* g_ans2_h = (__nv_bfloat16*)(bw + B.g_ans2_h);
    float* g_ans1 = (float*)(bw + B.g_ans1); void* g_ans1_h = bw + B.g_ans1_h;
    float* g_ans0 = (float*)(bw + B.g_ans0);
    float* g_ffc = (float*)(bw + B.g_ffc);
    const bool xbf = a->dtype == NF_DTYPE_BF16;
    const int xkind = xbf ? 2 : 1;
    const int ntiles = (N + 127) / 128;
    const int nsplit = ntiles < 8 ? ntiles : 8;
    int rc;

    k_bwd_head<<<(N + 255) / 256, 256, 0, st>>>(b->g_pos_out, b->g_vel_out, vel_new, N, a->dt, g_ans3, b->d_pos, b->d_vel);
    NF_LAUNCH_OK();
    // ---- layer 3: ans3 = conv3(x2) + dense3(x2)
    const float4* order = grid_view(ws + L.grid_f, N).sorted;     // the forward's fluid grid is still in its workspace
    float* dG = (float*)(ws + L.g3);                              // the forward's projection buffer is free again
    k_conv3_bwd_scatter<<<(N + 7) / 8, 256, 0, st>>>(pairs_ff, cnt_ff, g_ans3, N, dG, order);
    NF_LAUNCH_OK();
    NF_CUDA_OK(cudaFuncSetAttribute(k_conv3_bwd_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, C3B_SMEM));
    k_conv3_bwd_dense<<<min((N + C3B_TP - 1) / C3B_TP, num_sms()), 256, C3B_SMEM, st>>>(
        dG, x2, xbf ? 1 : 0, (const float*)(w + PL.k3), (const float*)(w + PL.w_dense3), g_ans3, ans2, N, g_ans2, g_ans2_h, dP + PO.off[14]);
    NF_LAUNCH_OK();
    if ((rc = launch_dense_wgrad(g_ans3, 3, 3, x2, 64, xkind, 64, N, dP + PO.off[16], dP + PO.off[15], dP + PO.off[17], st)) != NF_OK) return rc;
    // ---- layer 2: ans2 = conv2(x1) + dense2(x1) + ans1
    CWgradArgs wg;
    wg.slab_j = (const int*)(ws + L.slab_j); wg.slab_w = (const float4*)(ws + L.slab_w); wg.slab_off = (const unsigned short*)(ws + L.slab_off);
    wg.n = N; wg.ntiles = ntiles; wg.nsplit = nsplit; wg.order = order;
    wg.x_in = x1; wg.g = g_ans2_h; wg.dK = dP + PO.off[10]; wg.dWd = dP + PO.off[12];
    if ((rc = launch_wgrad<64>(wg, xbf, st)) != NF_OK) return rc;
    if ((rc = launch_dense_wgrad(g_ans2, 64, 64, x1, 64, xkind, 0, N, dP /*unused: cin = 0*/, dP + PO.off[11], dP + PO.off[13], st)) != NF_OK) return rc;
    ConvArgs c;
    c.slab_j = wg.slab_j; c.slab_w = wg.slab_w; c.slab_off = wg.slab_off; c.n = N; c.begin = 0; c.end = N; c.dense = 1; c.relu_out = 0;
    c.order = order; c.tile_rows = 0;
    c.x_in = g_ans2_h; c.w_packed = wb + BL.l2; c.residual = g_ans2; c.ld_res = 64; c.ans = g_ans1; c.x_out = g_ans1_h; c.cout = 64;
    c.mask_src = ans1; c.ld_mask = 64;
    if ((rc = launch_conv<64, 64>(c, NF_DTYPE_BF16, st)) != NF_OK) return rc;
    // ---- layer 1: ans1 = conv1(x0) + dense1(x0)
    wg.x_in = x0; wg.g = g_ans1_h; wg.dK = dP + PO.off[6]; wg.dWd = dP + PO.off[8];
    if ((rc = launch_wgrad<96>(wg, xbf, st)) != NF_OK) return rc;
    if ((rc = launch_dense_wgrad(g_ans1, 64, 64, x0, 96, xkind, 0, N, dP, dP + PO.off[7], dP + PO.off[9], st)) != NF_OK) return rc;
    c.x_in = g_ans1_h; c.w_packed = wb + BL.l1; c.residual = nullptr; c.ld_res = 0; c.ans = g_ans0; c.x_out = nullptr; c.cout = 96;
    c.mask_src = ans0; c.ld_mask = 96;
    if ((rc = launch_conv<64, 96>(c, NF_DTYPE_BF16, st)) != NF_OK) return rc;
    // ---- layer 0: ans0 = [conv0_obstacle(box normals), conv0_fluid([1, vel']), dense0([1, vel'])]
    L0WgradArgs l0;
    l0.pairs_ff = pairs_ff; l0.cnt_ff = cnt_ff; l0.pairs_fb = M > 0 ? pairs_fb : nullptr; l0.cnt_fb = M > 0 ? cnt_fb : nullptr;
    l0.vel_new = vel_new; l0.box_normals = a->box_normals; l0.g = g_ans0; l0.n = N; l0.order = order;
    l0.dKf = dP + PO.off[0]; l0.dbf = dP + PO.off[1]; l0.dKo = dP + PO.off[2]; l0.dbo = dP + PO.off[3];
    l0.dWd = dP + PO.off[4]; l0.dbd = dP + PO.off[5];
    k_layer0_wgrad<<<min((N + 7) / 8, 2 * num_sms()), 256, 0, st>>>(l0);
    NF_LAUNCH_OK();
    float* h0 = (float*)(bw + B.h0);
    k_layer0_bwd_project<<<min((N + 3) / 4, 2 * num_sms()), L0B_G, 0, st>>>(g_ans0, N, (const float*)(wb + BL.k0ft), h0);
    NF_LAUNCH_OK();
    k_layer0_bwd_gather<<<(N + 7) / 8, 256, 0, st>>>(pairs_ff, cnt_ff, h0, N, g_ffc, order);
    NF_LAUNCH_OK();
    k_bwd_tail<<<(N + 255) / 256, 256, 0, st>>>(g_ffc, g_ans0, (const float*)(w + PL.w_dense0), N, b->d_vel);
    NF_LAUNCH_OK();
    return NF_OK;
}

// phases: 0 integrate + grids + neighbour lists + layer 0;  1,2,3 conv layers;  4 position/velocity update
extern "C" int nf_transition_num_phases(void) { return 5; }

extern "C" int nf_transition_layer_buffer(int n_fluid, int n_box, int layer, size_t* off, size_t* row_bytes) {
    NF_REQUIRE(off && row_bytes && layer >= 0 && layer <= 6, NF_E_INVALID, "nf_transition_layer_buffer: bad arguments");
    const WsLayout L = ws_layout(n_fluid, n_box);
    if (layer <= 2) {
        *off = layer == 0 ? L.x0 : (layer == 1 ? L.x1 : L.x2);
        *row_bytes = (layer == 0 ? 96 : 64) * 2;
    } else {        // fp32 pre-activation outputs (the reference's ans_convs[layer - 3]); the last one is padded to 16 floats per row
        *off = layer == 3 ? L.ans0 : (layer == 4 ? L.ans1 : (layer == 5 ? L.ans2 : L.ans3));
        *row_bytes = (layer == 3 ? 96 : (layer == 6 ? 16 : 64)) * 4;
    }
    return NF_OK;
}

extern "C" int nf_transition_step(const nf_transition_args* a, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(a != nullptr, NF_E_INVALID, "nf_transition_step: null args");
    NF_REQUIRE(a->n_fluid >= 0 && a->n_box >= 0, NF_E_INVALID, "nf_transition_step: negative sizes");
    if (a->n_fluid == 0) return NF_OK;
    NF_REQUIRE(a->pos && a->vel && a->weights && a->workspace && a->pos_out && a->vel_out, NF_E_INVALID,
               "nf_transition_step: null pointer");
    NF_REQUIRE(a->n_box == 0 || (a->box && a->box_normals), NF_E_INVALID, "nf_transition_step: null box");
    NF_REQUIRE(a->dtype == NF_DTYPE_F16 || a->dtype == NF_DTYPE_BF16, NF_E_UNSUPPORTED, "nf_transition_step: dtype");
    NF_REQUIRE(a->filter_extent > 2e-3f && a->dt > 0.f, NF_E_INVALID, "nf_transition_step: bad extent/dt");
    NF_REQUIRE(a->phase >= NF_PHASE_SHARDED && a->phase < 5, NF_E_INVALID, "nf_transition_step: phase %d", a->phase);
    const int N = a->n_fluid, M = a->n_box;
    const bool sharded = a->phase == NF_PHASE_SHARDED;
    const int world = sharded ? comm::world() : 1;
    NF_REQUIRE(world <= ROW_PAD, NF_E_UNSUPPORTED, "nf_transition_step: at most %d ranks", ROW_PAD);
    const int per = (N + world - 1) / world;
    const int begin = a->phase == -1 ? 0 : (sharded ? min(comm::rank() * per, N) : a->shard_begin);
    const int end = a->phase == -1 ? N : (sharded ? min(begin + per, N) : a->shard_end);
    NF_REQUIRE(begin >= 0 && end <= N && begin <= end, NF_E_INVALID, "nf_transition_step: bad shard [%d,%d)", begin, end);
    const WsLayout L = ws_layout(N, M);
    NF_REQUIRE(a->workspace_bytes >= L.total, NF_E_WORKSPACE, "nf_transition_step: workspace %zu < %zu",
               a->workspace_bytes, L.total);
    const PackedLayout PL = packed_layout();
    char* b = (char*)a->workspace;
    const uint8_t* w = (const uint8_t*)a->weights;
    float* pos_new = (float*)(b + L.pos_new);
    float* vel_new = (float*)(b + L.vel_new);
    Pair* pairs_ff = (Pair*)(b + L.pairs_ff); int* cnt_ff = (int*)(b + L.cnt_ff);
    Pair* pairs_fb = (Pair*)(b + L.pairs_fb); int* cnt_fb = (int*)(b + L.cnt_fb);
    int* slab_j = (int*)(b + L.slab_j); float4* slab_w = (float4*)(b + L.slab_w);
    unsigned short* slab_off = (unsigned short*)(b + L.slab_off);
    float* ans0 = (float*)(b + L.ans0); void* x0 = b + L.x0;
    float* ans1 = (float*)(b + L.ans1); void* x1 = b + L.x1;
    float* ans2 = (float*)(b + L.ans2); void* x2 = b + L.x2;
    float* ans3 = (float*)(b + L.ans3);
    int* flags = (int*)(b + L.flags);
    const float radius = 0.5f * a->filter_extent;
    const int nshard = end - begin;
    const int ph_lo = a->phase < 0 ? 0 : a->phase, ph_hi = a->phase < 0 ? 4 : a->phase;
    NF_REQUIRE(!sharded || a->nnbr_out, NF_E_INVALID, "nf_transition_step: the sharded step needs nnbr_out");
    ConvArgs c;
    // Whole-step, single-GPU: every per-particle kernel walks the particles in the fluid grid's CELL ORDER (the grid is
    // built in phase 0), so that a warp's / a tile's particles are spatial neighbours and their gathers share cache lines.
    // Sharded / per-phase calls keep array order: a rank's rows must be a contiguous block for the in-place all-gather.
    const float4* order = a->phase == -1 ? grid_view(b + L.grid_f, N).sorted : nullptr;
    c.slab_j = slab_j; c.slab_w = slab_w; c.slab_off = slab_off; c.n = N; c.begin = begin; c.end = end; c.dense = 1;
    c.mask_src = nullptr; c.ld_mask = 0; c.relu_out = 1; c.order = order;
    c.tile_rows = 128;
    if (a->phase != -1)       // a shard: the smallest tile (>= 16 rows) that still fits one wave of CTAs
        while (c.tile_rows > 16 && (nshard + c.tile_rows / 2 - 1) / (c.tile_rows / 2) <= num_sms()) c.tile_rows /= 2;
    if (sharded && world > 1) {      // peers may store into this workspace from here on (peer-memory exchange, nf_comm.cu)
        const int rc = comm::enter(st);
        if (rc != NF_OK) return rc;
    }
    for (int ph = ph_lo; ph <= ph_hi; ++ph) {
    if (ph == 0) {
        NF_CUDA_OK(cudaMemsetAsync(flags, 0, 256, st));
        k_integrate<<<(3 * N + 255) / 256, 256, 0, st>>>(a->pos, a->vel, N, a->gravity[0], a->gravity[1], a->gravity[2],
                                                        a->dt, pos_new, vel_new);
        NF_LAUNCH_OK();
        int rc = nf_grid_build(pos_new, N, 1.002f * radius, b + L.grid_f, grid_layout(N).total, stream_);
        if (rc != NF_OK) return rc;
        if (!a->box_grid_ws) {
            rc = nf_grid_build(a->box, M, 1.002f * radius, b + L.grid_b, grid_layout(M).total, stream_);
            if (rc != NF_OK) return rc;
        }
        if (nshard > 0) {
            const int blocks = (nshard + 7) / 8;
            k_nbr_build<<<blocks, 256, 0, st>>>(grid_view(b + L.grid_f, N), pos_new, begin, end, radius, 1, 1, pairs_ff,
                                               cnt_ff, a->nnbr_out, flags, slab_j, slab_w, slab_off, order);
            NF_LAUNCH_OK();
            k_nbr_build<<<blocks, 256, 0, st>>>(grid_view(a->box_grid_ws ? a->box_grid_ws : b + L.grid_b, M), pos_new, begin, end, radius, 1, 1, pairs_fb,
                                               cnt_fb, nullptr, flags + 1, nullptr, nullptr, nullptr, order);
            NF_LAUNCH_OK();
            if (a->overflow_out) {
                k_add_overflow<<<1, 32, 0, st>>>(flags, a->overflow_out);
                NF_LAUNCH_OK();
            }
            Layer0Args l0;
            l0.pairs_ff = pairs_ff; l0.cnt_ff = cnt_ff; l0.pairs_fb = pairs_fb; l0.cnt_fb = cnt_fb;
            l0.vel_new = vel_new; l0.box_normals = a->box_normals;
            l0.k_fluid = (const float*)(w + PL.k_fluid); l0.b_fluid = (const float*)(w + PL.b_fluid);
            l0.k_obst = (const float*)(w + PL.k_obst); l0.b_obst = (const float*)(w + PL.b_obst);
            l0.w_dense = (const float*)(w + PL.w_dense0); l0.b_dense = (const float*)(w + PL.b_dense0);
            l0.ans0 = ans0; l0.x0 = x0; l0.begin = begin; l0.end = end; l0.bf16 = a->dtype == NF_DTYPE_BF16;
            l0.order = order;
            k_layer0<<<blocks, 256, 0, st>>>(l0);
            NF_LAUNCH_OK();
            if (a->feats0_out)
                NF_CUDA_OK(cudaMemcpyAsync(a->feats0_out + (size_t)begin * 96, ans0 + (size_t)begin * 96,
                                           (size_t)nshard * 96 * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (ph == 1) {     // conv1 + dense1 : 96 -> 64 (no residual: widths differ, :127-130)
        c.x_in = x0; c.w_packed = w + PL.l1; c.residual = nullptr; c.ld_res = 0; c.ans = ans1; c.x_out = x1; c.cout = 64;
        int rc = launch_conv<96, 64>(c, a->dtype, st);
        if (rc != NF_OK) return rc;
    }
    if (ph == 2) {     // conv2 + dense2 + residual : 64 -> 64
        c.x_in = x1; c.w_packed = w + PL.l2; c.residual = ans1; c.ld_res = 64; c.ans = ans2; c.x_out = x2; c.cout = 64;
        int rc = launch_conv<64, 64>(c, a->dtype, st);
        if (rc != NF_OK) return rc;
    }
    if (ph == 3) {     // conv3 + dense3 : 64 -> 3  (project every particle, then gather per neighbour)
        float* g3 = (float*)(b + L.g3);
        const bool bf = a->dtype == NF_DTYPE_BF16;
        const size_t smem = (size_t)NCELL * 64 * 3 * sizeof(float);
        const int pgrid = min((N + 3) / 4, 2 * num_sms());
        if (bf) {
            NF_CUDA_OK(cudaFuncSetAttribute(k_conv3_project<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_conv3_project<true><<<pgrid, C3_G, smem, st>>>(x2, N, (const float*)(w + PL.k3), g3);
        } else {
            NF_CUDA_OK(cudaFuncSetAttribute(k_conv3_project<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_conv3_project<false><<<pgrid, C3_G, smem, st>>>(x2, N, (const float*)(w + PL.k3), g3);
        }
        NF_LAUNCH_OK();
        if (nshard > 0) {
            const int blocks = (nshard + 7) / 8;
            if (bf) k_conv3_gather<true><<<blocks, 256, 0, st>>>(pairs_ff, cnt_ff, g3, x2, (const float*)(w + PL.b3), (const float*)(w + PL.w_dense3),
                                                                (const float*)(w + PL.b_dense3), begin, end, ans3, order);
            else k_conv3_gather<false><<<blocks, 256, 0, st>>>(pairs_ff, cnt_ff, g3, x2, (const float*)(w + PL.b3), (const float*)(w + PL.w_dense3),
                                                               (const float*)(w + PL.b_dense3), begin, end, ans3, order);
            NF_LAUNCH_OK();
        }
    }
    if (ph == 4 && nshard > 0) {
        k_update<<<(3 * nshard + 255) / 256, 256, 0, st>>>(a->pos, pos_new, ans3, 16, begin, end, a->dt, a->pos_out,
                                                          a->vel_out, a->delta_out);
        NF_LAUNCH_OK();
    }
    if (sharded && world > 1) {
        // the exchange step: rows produced by this rank -> every rank, in place, on the same stream
        int rc = NF_OK;
        if (ph == 0) rc = comm::allgather_inplace(x0, (size_t)per * 96 * 2, st);
        else if (ph == 1) rc = comm::allgather_inplace(x1, (size_t)per * 64 * 2, st);
        else if (ph == 2) rc = comm::allgather_inplace(x2, (size_t)per * 64 * 2, st);
        else if (ph == 4) {
            float* out10 = (float*)(b + L.out10);
            if (nshard > 0) {
                k_pack10<<<(nshard + 255) / 256, 256, 0, st>>>(a->pos_out, a->vel_out, a->nnbr_out, a->delta_out, begin, end, out10);
                NF_LAUNCH_OK();
            }
            rc = comm::allgather_inplace(out10, (size_t)per * 10 * 4, st);
            if (rc != NF_OK) return rc;
            k_unpack10<<<(N + 255) / 256, 256, 0, st>>>(out10, N, a->pos_out, a->vel_out, a->nnbr_out, a->delta_out);
            NF_LAUNCH_OK();
        }
        if (rc != NF_OK) return rc;
    }
    }   // phase loop
    return NF_OK;
}
