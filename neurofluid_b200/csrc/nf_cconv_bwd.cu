// nf_cconv_bwd.cu -- backward pass of the Lagrangian transition model (hot path 2) on sm_100a.
//
// replaces: autograd through ParticleNet.forward (models/transmodel.py:100-163): loss.backward() at
//           trainer/trainer_transmodel.py:197 and, through the renderer, trainer/trainer_e2e.py:277.
#include "nf_cconv.cuh"

namespace nf {
namespace cconv {

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
