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
//
// The backward pass lives in nf_cconv_bwd.cu; formats, k_cconv_tc and the layouts shared by both in nf_cconv.cuh.
#include "nf_cconv.cuh"

namespace nf {
namespace cconv {

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
// conv3 + dense3 (64 -> 3).  With 3 output channels the contraction is cheaper the other way round:
//   out_i = sum_j sum_c w_ijc K_c^T f_j  =  sum_j sum_c w_ijc g_j[c],   g_j[c] = K_c^T f_j  (3 values per cell)
// so every particle's features are projected through all 64 filter cells once (k_conv3_project: N x 64 x 192 MAC,
// fp32, weights in shared memory) and the neighbour pass gathers 8 x 3 floats per pair instead of building a
// 128 x 256 patch slab per filter row for 3 useful output columns (k_conv3_gather: one warp per particle, one
// lane per pair, fixed reduction tree -> deterministic).  fp32 end to end.
// ------------------------------------------------------------------------------------------------

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


// ------------------------------------------------------------------------------------------------
// Operator-level ContinuousConv (nf_cconv_forward): one conv on arbitrary in / out point sets.
//   small shapes (cin * cout <= SMALL_MAX: the 4->32, 3->32, 64->3 layers): fp32 on CUDA cores, the whole filter in
//   shared memory, one warp per out point: patch (64 cells x cin) accumulated from the pair list, then contracted.
//   64/96 -> 64: the slab-list tensor-core kernel above with the dense slab switched off.
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_MAX = 768;


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

