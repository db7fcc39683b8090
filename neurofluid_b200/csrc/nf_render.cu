// nf_render.cu -- ray-stage kernels + orchestration of the whole renderer forward.
//
// replaces (reference file:line):
//   coarse_sample_ray                utils/ray_utils.py:232-256
//   RenderNet.search                 models/renderer.py:112-122   (pytorch3d ball_query)
//   smoothing_position               models/renderer.py:96-109
//   embedding_local_geometry         models/renderer.py:125-179   (geometry part; the encodings live in nf_mlp.cu)
//   mask / use_mask                  models/renderer.py:233-237, 258-262
//   render_image                     models/renderer.py:182-208
//   sample_pdf / ImportanceSampling  utils/ray_utils.py:178-229
//   RenderNet.forward / coarse_rendering / fine_rendering   models/renderer.py:211-369
//
// One warp owns one ray.  Sample s of a ray lives in lane s%32, register slot s/32, so all per-ray
// state (depths, weights, counts) stays in registers and the along-ray products are warp scans.
//
// stage Q0 : coarse depths -> sample positions -> first-K ball query per non-empty sample ->
//            num_nn, "all K slots valid" bitmask, and one 64-byte geometry record per evaluated sample
//            appended to a compact list (rows are handed out to warps in blocks of 32).
// [MLP]    : nf_mlp.cu over the compact list -> (r,g,b,sigma) scattered to a dense per-sample array.
// stage MID: alpha-composite the coarse samples (warp scan), emit rgb0/depth0/opacity0/mask_0, build
//            the piecewise-constant pdf, draw the importance samples by inverse CDF, merge with the
//            coarse depths (rank merge), then run the ball query again for the merged samples.
// [MLP]    : fine network.
// stage FIN: alpha-composite the fine samples -> rgb1/depth1/opacity1/mask_1.
#include <vector>

#include "nf_common.cuh"
#include "nf_mlp.cuh"

namespace nf {
namespace render {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int ROW_BLOCK = 32;

struct StageArgs {
    GridView g;
    const float* particles;
    const float* rays;
    int n_rays;
    float ro[3];
    float radius;
    int K;
    int use_mask, white_bg, mode;
    const float* z_coarse;
    const float* u_imp;
    int S0, n_imp, S1;
    // outputs
    float *rgb0, *depth0, *opac0, *mask0;
    long long* num_nn0;
    float *rgb1, *depth1, *opac1, *mask1;
    long long* num_nn1;
    // workspace
    int* counters;       // [0] rows coarse, [1] rows fine, [2] active coarse, [3] active fine
    unsigned* act0;      // (R, NS0)
    unsigned* act1;      // (R, NS1)
    float* z1;           // (R, S1)
    float* rec0; int* rowid0; float4* out0; int cap0;
    float* rec1; int* rowid1; float4* out1; int cap1;
};

struct RowAlloc {
    int cur = 0, left = 0;
    __device__ __forceinline__ int take(int* counter, int lane) {
        if (left == 0) {
            int b = 0;
            if (lane == 0) b = atomicAdd(counter, ROW_BLOCK);
            cur = __shfl_sync(NF_FULL, b, 0);
            left = ROW_BLOCK;
        }
        --left;
        return cur++;
    }
    // unused rows of the last block become holes the MLP skips
    __device__ __forceinline__ void flush(int* rowid, int cap, int lane) {
        if (lane < left && cur + lane < cap) rowid[cur + lane] = -1;
        left = 0;
    }
};

// ------------------------------------------------------------------------------------------------
// neighbour search + local geometry for all samples of one ray
// ------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void ray_query(const StageArgs& p, int lane, const float (&o)[3], const float (&d)[3],
                                          const float (&z)[NS], int S, RowAlloc& ra, float* rec, int* rowid,
                                          int* row_counter, int* active_counter, int cap, int sample_base,
                                          unsigned (&fullbits)[NS], int (&cnt)[NS]) {
    float px[NS], py[NS], pz[NS];
    unsigned nonempty[NS], todo[NS];
    const int K = p.K;
    const float radius = p.radius;
#pragma unroll
    for (int slot = 0; slot < NS; ++slot) {
        const int s = slot * 32 + lane;
        // xyz = o + d * z, rounded like the eager torch expression (mul, then add)
        px[slot] = __fadd_rn(o[0], __fmul_rn(d[0], z[slot]));
        py[slot] = __fadd_rn(o[1], __fmul_rn(d[1], z[slot]));
        pz[slot] = __fadd_rn(o[2], __fmul_rn(d[2], z[slot]));
        const bool in = s < S;
        const bool ne = in && grid_maybe_nonempty(p.g, px[slot], py[slot], pz[slot], radius);
        nonempty[slot] = __ballot_sync(NF_FULL, ne);
        todo[slot] = p.use_mask ? nonempty[slot] : __ballot_sync(NF_FULL, in);
        cnt[slot] = 0;
        fullbits[slot] = 0u;
    }
    int n_active = 0;
#pragma unroll
    for (int slot = 0; slot < NS; ++slot) {
        unsigned m = todo[slot];
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float qx = __shfl_sync(NF_FULL, px[slot], src);
            const float qy = __shfl_sync(NF_FULL, py[slot], src);
            const float qz = __shfl_sync(NF_FULL, pz[slot], src);
            int best = 0x7fffffff, nsel = 0;
            if ((nonempty[slot] >> src) & 1u) nsel = warp_first_k(p.g, qx, qy, qz, radius, K, lane, best);
            const bool sel = lane < nsel;
            float nx = 0.f, ny = 0.f, nz = 0.f, d2 = 0.f;
            if (sel) {
                nx = __ldg(p.particles + 3 * (size_t)best);
                ny = __ldg(p.particles + 3 * (size_t)best + 1);
                nz = __ldg(p.particles + 3 * (size_t)best + 2);
                d2 = dist2_exact(qx, qy, qz, nx, ny, nz);
            }
            // nn_mask = dists.ne(0): a real neighbour at exactly zero distance counts as padding
            const bool valid = sel && (d2 != 0.f);
            const int nvalid = __popc(__ballot_sync(NF_FULL, valid));
            const bool full = (nvalid == K);
            if (lane == src) cnt[slot] = nvalid;
            if (full) { fullbits[slot] |= 1u << src; ++n_active; }
            if (p.use_mask && !full) continue;

            // ---- smoothing_position: every one of the K slots takes part; padded slots are zeros
            float w = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
            if (lane < K) {
                const float ex = nx - qx, ey = ny - qy, ez = nz - qz;   // nx..nz are 0 on padded slots
                const float dist = sqrtf(ex * ex + ey * ey + ez * ez);
                const float t = dist / radius;
                w = fmaxf(1.0f - t * t * t, 0.f);
                wx = w * nx; wy = w * ny; wz = w * nz;
            }
            const float density = warp_sum(w);
            const float den = density + 1e-12f;
            const float sx = warp_sum(wx) / den, sy = warp_sum(wy) / den, sz = warp_sum(wz) / den;
            // ---- variance of the valid neighbour offsets (two pass)
            const float nvf = (float)nvalid + 1e-12f;
            const float vx = valid ? nx - qx : 0.f, vy = valid ? ny - qy : 0.f, vz = valid ? nz - qz : 0.f;
            const float mx = warp_sum(vx) / nvf, my = warp_sum(vy) / nvf, mz = warp_sum(vz) / nvf;
            const float ax = valid ? (vx - mx) * (vx - mx) : 0.f;
            const float ay = valid ? (vy - my) * (vy - my) : 0.f;
            const float az = valid ? (vz - mz) * (vz - mz) : 0.f;
            const float varx = warp_sum(ax) / nvf, vary = warp_sum(ay) / nvf, varz = warp_sum(az) / nvf;
            // ---- direction from the camera to the smoothed position
            const float tx = sx - p.ro[0], ty = sy - p.ro[1], tz = sz - p.ro[2];
            const float tn = sqrtf(tx * tx + ty * ty + tz * tz);
            const int row = ra.take(row_counter, lane);
            if (row < cap) {
                if (lane < 4) {
                    float4 v;
                    if (lane == 0) v = make_float4(qx, qy, qz, density);
                    else if (lane == 1) v = make_float4(sx, sy, sz, varx);
                    else if (lane == 2) v = make_float4(vary, varz, d[0], d[1]);
                    else v = make_float4(d[2], tx / tn, ty / tn, tz / tn);
                    reinterpret_cast<float4*>(rec + (size_t)row * 16)[lane] = v;
                }
                if (lane == 0) rowid[row] = sample_base + slot * 32 + src;
            }
        }
    }
    if (lane == 0 && n_active) atomicAdd(active_counter, n_active);
}

// ------------------------------------------------------------------------------------------------
// alpha compositing along one ray (models/renderer.py:182-208)
// ------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void ray_composite(const float (&z)[NS], const float4 (&c)[NS], int S, float dnorm,
                                              int lane, bool white, float (&w)[NS], float (&rgb)[3], float& depth,
                                              float& acc) {
    float carry = 1.f;
    float r = 0.f, g = 0.f, b = 0.f, dep = 0.f, a = 0.f;
#pragma unroll
    for (int slot = 0; slot < NS; ++slot) {
        const int s = slot * 32 + lane;
        const bool in = s < S;
        float zn = __shfl_down_sync(NF_FULL, z[slot], 1);
        float z_next0 = 0.f;
        if (slot + 1 < NS) z_next0 = __shfl_sync(NF_FULL, z[slot + 1], 0);
        if (lane == 31) zn = z_next0;
        float delta = (s == S - 1) ? 1e10f : zn - z[slot];
        delta *= dnorm;
        const float alpha = in ? 1.0f - expf(-delta * fmaxf(c[slot].w, 0.f)) : 0.f;
        const float a1 = in ? (1.0f - alpha + 1e-10f) : 1.f;
        float P = a1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float t = __shfl_up_sync(NF_FULL, P, off);
            if (lane >= off) P *= t;
        }
        float excl = __shfl_up_sync(NF_FULL, P, 1);
        if (lane == 0) excl = 1.f;
        const float T = carry * excl;
        carry *= __shfl_sync(NF_FULL, P, 31);
        const float wt = alpha * T;
        w[slot] = wt;
        r += wt * c[slot].x; g += wt * c[slot].y; b += wt * c[slot].z;
        dep += wt * z[slot];
        a += wt;
    }
    r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); dep = warp_sum(dep); a = warp_sum(a);
    if (white) { r += 1.0f - a; g += 1.0f - a; b += 1.0f - a; }
    rgb[0] = r; rgb[1] = g; rgb[2] = b; depth = dep; acc = a;
}

__device__ __forceinline__ void load_ray(const float* rays, int ray, float (&o)[3], float (&d)[3]) {
    const float* r = rays + (size_t)ray * 6;
    o[0] = __ldg(r); o[1] = __ldg(r + 1); o[2] = __ldg(r + 2);
    d[0] = __ldg(r + 3); d[1] = __ldg(r + 4); d[2] = __ldg(r + 5);
}

template <int NS>
__device__ __forceinline__ void store_counts(long long* num_nn, unsigned* act, int ray, int S, int lane,
                                             const int (&cnt)[NS], const unsigned (&fullbits)[NS]) {
#pragma unroll
    for (int slot = 0; slot < NS; ++slot) {
        const int s = slot * 32 + lane;
        if (num_nn && s < S) num_nn[(size_t)ray * S + s] = cnt[slot];
        if (lane == 0) act[(size_t)ray * NS + slot] = fullbits[slot];
    }
}

// ------------------------------------------------------------------------------------------------
// stage Q0
// ------------------------------------------------------------------------------------------------
template <int NS0>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_stage_q0(const StageArgs p) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    RowAlloc ra;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        float o[3], d[3], z[NS0];
        load_ray(p.rays, ray, o, d);
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) z[slot] = __ldg(p.z_coarse + min(slot * 32 + lane, p.S0 - 1));
        unsigned fullbits[NS0];
        int cnt[NS0];
        ray_query<NS0>(p, lane, o, d, z, p.S0, ra, p.rec0, p.rowid0, p.counters + 0, p.counters + 2, p.cap0,
                       ray * p.S0, fullbits, cnt);
        store_counts<NS0>(p.num_nn0, p.act0, ray, p.S0, lane, cnt, fullbits);
    }
    ra.flush(p.rowid0, p.cap0, lane);
}

// ------------------------------------------------------------------------------------------------
// stage MID
// ------------------------------------------------------------------------------------------------
template <int NS0, int NS1>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_stage_mid(const StageArgs p) {
    __shared__ float sm_z0[WARPS_PER_BLOCK][NS0 * 32];
    __shared__ float sm_w[WARPS_PER_BLOCK][NS0 * 32];      // coarse weights, later the importance samples
    __shared__ float sm_bins[WARPS_PER_BLOCK][NS0 * 32];
    __shared__ float sm_cdf[WARPS_PER_BLOCK][NS0 * 32];
    __shared__ float sm_smp[WARPS_PER_BLOCK][NS1 * 32];
    __shared__ float sm_z1[WARPS_PER_BLOCK][NS1 * 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S0 = p.S0, S1 = p.S1, NI = p.n_imp;
    float* z0s = sm_z0[wib]; float* ws = sm_w[wib]; float* bins = sm_bins[wib]; float* cdf = sm_cdf[wib];
    float* smp = sm_smp[wib]; float* z1s = sm_z1[wib];
    RowAlloc ra;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        float o[3], d[3], z0[NS0], w0[NS0];
        float4 c0[NS0];
        load_ray(p.rays, ray, o, d);
        const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        int nfull = 0;
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int s = slot * 32 + lane;
            z0[slot] = __ldg(p.z_coarse + min(s, S0 - 1));
            const unsigned bits = p.act0[(size_t)ray * NS0 + slot];
            nfull += __popc(bits);
            const bool ev = s < S0 && (p.use_mask ? ((bits >> lane) & 1u) : true);
            c0[slot] = ev ? p.out0[(size_t)ray * S0 + s] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float rgb[3], depth, acc;
        ray_composite<NS0>(z0, c0, S0, dnorm, lane, p.white_bg != 0, w0, rgb, depth, acc);
        if (lane == 0) {
            if (p.rgb0) { p.rgb0[3 * (size_t)ray] = rgb[0]; p.rgb0[3 * (size_t)ray + 1] = rgb[1]; p.rgb0[3 * (size_t)ray + 2] = rgb[2]; }
            if (p.depth0) p.depth0[ray] = depth;
            if (p.opac0) p.opac0[ray] = acc;
            if (p.mask0) p.mask0[ray] = (float)nfull;
        }
        // ---------------- sample_pdf (utils/ray_utils.py:178-220), det=True
        __syncwarp();
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int s = slot * 32 + lane;
            if (s < S0) { z0s[s] = z0[slot]; ws[s] = w0[slot]; }
        }
        __syncwarp();
        const int nb = S0 - 1;     // bins = cdf entries
        const int npdf = S0 - 2;   // pdf entries (weights[1:-1])
        float pw[NS0], tot = 0.f;
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int i = slot * 32 + lane;
            if (i < nb) bins[i] = 0.5f * (z0s[i + 1] + z0s[i]);
            pw[slot] = (i < npdf) ? ws[i + 1] + 1e-5f : 0.f;
            tot += pw[slot];
        }
        tot = warp_sum(tot);
        float carry = 0.f;
        if (lane == 0) cdf[0] = 0.f;
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int i = slot * 32 + lane;
            float P = pw[slot] / tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float t = __shfl_up_sync(NF_FULL, P, off);
                if (lane >= off) P += t;
            }
            if (i < npdf) cdf[i + 1] = carry + P;
            carry += __shfl_sync(NF_FULL, P, 31);
        }
        __syncwarp();
        // inverse CDF; running max keeps the draws non-decreasing (they are, up to 1 ulp of rounding)
        float runmax = -3.0e38f;
        for (int j0 = 0; j0 < NI; j0 += 32) {
            const int j = j0 + lane;
            float sv = -3.0e38f;
            if (j < NI) {
                const float u = __ldg(p.u_imp + j);
                int lo = 0, hi = nb;            // count of cdf[k] <= u   (searchsorted right=True)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
                }
                const int below = max(lo - 1, 0), above = min(lo, nb - 1);
                const float c_lo = cdf[below], c_hi = cdf[above];
                float den = c_hi - c_lo;
                if (den < 1e-5f) den = 1.f;
                const float t = (u - c_lo) / den;
                sv = bins[below] + t * (bins[above] - bins[below]);
            }
            float M = sv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float t = __shfl_up_sync(NF_FULL, M, off);
                if (lane >= off) M = fmaxf(M, t);
            }
            M = fmaxf(M, runmax);
            if (j < NI) smp[j] = M;
            runmax = __shfl_sync(NF_FULL, M, 31);
        }
        __syncwarp();
        // rank merge of the two sorted lists (ties: coarse depths first)
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int i = slot * 32 + lane;
            if (i < S0) {
                const float v = z0s[i];
                int lo = 0, hi = NI;            // # samples < v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (smp[mid] < v) lo = mid + 1; else hi = mid; }
                z1s[i + lo] = v;
            }
        }
        for (int j0 = 0; j0 < NI; j0 += 32) {
            const int j = j0 + lane;
            if (j < NI) {
                const float v = smp[j];
                int lo = 0, hi = S0;            // # coarse depths <= v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (z0s[mid] <= v) lo = mid + 1; else hi = mid; }
                z1s[j + lo] = v;
            }
        }
        __syncwarp();
        float z1[NS1];
#pragma unroll
        for (int slot = 0; slot < NS1; ++slot) {
            const int s = slot * 32 + lane;
            z1[slot] = z1s[min(s, S1 - 1)];
            if (s < S1) p.z1[(size_t)ray * S1 + s] = z1[slot];
        }
        unsigned fullbits[NS1];
        int cnt[NS1];
        ray_query<NS1>(p, lane, o, d, z1, S1, ra, p.rec1, p.rowid1, p.counters + 1, p.counters + 3, p.cap1,
                       ray * S1, fullbits, cnt);
        store_counts<NS1>(p.num_nn1, p.act1, ray, S1, lane, cnt, fullbits);
    }
    ra.flush(p.rowid1, p.cap1, lane);
}

// ------------------------------------------------------------------------------------------------
// stage FIN (also the only compositing stage of coarse_rendering, with FIRST = true)
// ------------------------------------------------------------------------------------------------
template <int NS, bool FIRST>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_stage_fin(const StageArgs p) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S = FIRST ? p.S0 : p.S1;
    const unsigned* act = FIRST ? p.act0 : p.act1;
    const float4* out = FIRST ? p.out0 : p.out1;
    float* o_rgb = FIRST ? p.rgb0 : p.rgb1;
    float* o_depth = FIRST ? p.depth0 : p.depth1;
    float* o_opac = FIRST ? p.opac0 : p.opac1;
    float* o_mask = FIRST ? p.mask0 : p.mask1;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        float o[3], d[3], z[NS], w[NS];
        float4 c[NS];
        load_ray(p.rays, ray, o, d);
        const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        int nfull = 0;
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const int s = slot * 32 + lane;
            z[slot] = FIRST ? __ldg(p.z_coarse + min(s, S - 1)) : p.z1[(size_t)ray * S + min(s, S - 1)];
            const unsigned bits = act[(size_t)ray * NS + slot];
            nfull += __popc(bits);
            const bool ev = s < S && (p.use_mask ? ((bits >> lane) & 1u) : true);
            c[slot] = ev ? out[(size_t)ray * S + s] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float rgb[3], depth, acc;
        ray_composite<NS>(z, c, S, dnorm, lane, p.white_bg != 0, w, rgb, depth, acc);
        if (lane == 0) {
            if (o_rgb) { o_rgb[3 * (size_t)ray] = rgb[0]; o_rgb[3 * (size_t)ray + 1] = rgb[1]; o_rgb[3 * (size_t)ray + 2] = rgb[2]; }
            if (o_depth) o_depth[ray] = depth;
            if (o_opac) o_opac[ray] = acc;
            if (o_mask) o_mask[ray] = (float)nfull;
        }
    }
}

struct WsLayout {
    size_t counters, act0, act1, z1, rec0, rowid0, out0, rec1, rowid1, out1, total;
    int cap0, cap1, ns0, ns1;
};

static int max_stage_warps() { return num_sms() * 8 * WARPS_PER_BLOCK; }

static int pick_ns(int s, const int* opts, int n) {
    for (int i = 0; i < n; ++i)
        if (s <= opts[i] * 32) return opts[i];
    return -1;
}

static WsLayout ws_layout(int R, int S0, int NI) {
    WsLayout L;
    const int S1 = S0 + NI;
    static const int o0[] = {2, 4};
    static const int o1[] = {4, 6, 8};
    L.ns0 = pick_ns(S0, o0, 2);
    L.ns1 = NI > 0 ? pick_ns(S1, o1, 3) : 4;
    const size_t slack = (size_t)ROW_BLOCK * max_stage_warps();
    L.cap0 = (int)((size_t)R * S0 + slack);
    L.cap1 = (int)((size_t)R * S1 + slack);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    L.counters = take(64);
    L.act0 = take(sizeof(unsigned) * (size_t)R * 4);
    L.act1 = take(sizeof(unsigned) * (size_t)R * 8);
    L.z1 = take(sizeof(float) * (size_t)R * S1);
    L.rec0 = take(sizeof(float) * 16 * (size_t)L.cap0);
    L.rowid0 = take(sizeof(int) * (size_t)L.cap0);
    L.out0 = take(sizeof(float4) * (size_t)R * S0);
    L.rec1 = take(sizeof(float) * 16 * (size_t)L.cap1);
    L.rowid1 = take(sizeof(int) * (size_t)L.cap1);
    L.out1 = take(sizeof(float4) * (size_t)R * S1);
    L.total = o;
    return L;
}

template <int NS0>
static int launch_mid(int ns1, int grid, const StageArgs& p, cudaStream_t st) {
    switch (ns1) {
        case 4: k_stage_mid<NS0, 4><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(p); break;
        case 6: k_stage_mid<NS0, 6><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(p); break;
        case 8: k_stage_mid<NS0, 8><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(p); break;
        default: set_error("unsupported fine sample count"); return NF_E_UNSUPPORTED;
    }
    NF_LAUNCH_OK();
    return NF_OK;
}

// ------------------------------------------------------------------------------------------------
// optional per-stage timing with CUDA events recorded on the launching stream (bench.py roofline)
// ------------------------------------------------------------------------------------------------
constexpr int NSTAGE_T = 5;   // q0, mlp coarse, mid, mlp fine, fin
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_events;   // (NSTAGE_T + 1) events per profiled call

struct StageTimer {
    cudaStream_t st;
    bool on;
    explicit StageTimer(cudaStream_t s) : st(s), on(g_prof_on) { mark(); }
    void mark() {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
        cudaEventRecord(e, st);
        g_prof_events.push_back(e);
    }
};

}  // namespace render
}  // namespace nf

using namespace nf;
using namespace nf::render;

extern "C" int nf_profile_enable(int on) {
    for (cudaEvent_t e : g_prof_events) cudaEventDestroy(e);
    g_prof_events.clear();
    g_prof_on = on != 0;
    return NF_OK;
}

extern "C" int nf_profile_read(double* stage_ms /*[5]*/, int* n_calls) {
    NF_REQUIRE(stage_ms && n_calls, NF_E_INVALID, "nf_profile_read: null argument");
    for (int i = 0; i < NSTAGE_T; ++i) stage_ms[i] = 0.0;
    const size_t per = NSTAGE_T + 1;
    const size_t calls = g_prof_events.size() / per;
    for (size_t c = 0; c < calls; ++c) {
        NF_CUDA_OK(cudaEventSynchronize(g_prof_events[c * per + NSTAGE_T]));
        for (int i = 0; i < NSTAGE_T; ++i) {
            float ms = 0.f;
            NF_CUDA_OK(cudaEventElapsedTime(&ms, g_prof_events[c * per + i], g_prof_events[c * per + i + 1]));
            stage_ms[i] += ms;
        }
    }
    *n_calls = (int)calls;
    for (cudaEvent_t e : g_prof_events) cudaEventDestroy(e);
    g_prof_events.clear();
    return NF_OK;
}

extern "C" size_t nf_render_workspace_bytes(int n_rays, int n_coarse, int n_importance) {
    if (n_rays <= 0 || n_coarse <= 0 || n_importance < 0) return 0;
    return ws_layout(n_rays, n_coarse, n_importance).total;
}

extern "C" int nf_render_forward(const nf_render_args* a, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(a != nullptr, NF_E_INVALID, "nf_render_forward: null args");
    NF_REQUIRE(a->grid_ws && a->rays && a->z_coarse && a->workspace, NF_E_INVALID, "nf_render_forward: null pointer");
    NF_REQUIRE(a->n_particles == 0 || a->particles, NF_E_INVALID, "nf_render_forward: null particles");
    NF_REQUIRE(a->mode >= 0 && a->mode <= 2, NF_E_INVALID, "nf_render_forward: mode %d", a->mode);
    NF_REQUIRE(a->K >= 1 && a->K <= 32, NF_E_UNSUPPORTED, "nf_render_forward: K=%d not in [1,32]", a->K);
    NF_REQUIRE(a->n_coarse >= 3 && a->n_coarse <= 128, NF_E_UNSUPPORTED, "nf_render_forward: n_coarse=%d not in [3,128]", a->n_coarse);
    const bool fine = a->mode != NF_RENDER_COARSE;
    const int NI = fine ? a->n_importance : 0;
    NF_REQUIRE(!fine || (NI >= 1 && a->n_coarse + NI <= 256), NF_E_UNSUPPORTED,
               "nf_render_forward: n_coarse+n_importance=%d not in [.,256]", a->n_coarse + NI);
    NF_REQUIRE(!fine || a->u_importance, NF_E_INVALID, "nf_render_forward: null u_importance");
    NF_REQUIRE(a->weights_coarse && (!fine || a->weights_fine), NF_E_INVALID, "nf_render_forward: null weights");
    NF_REQUIRE(a->radius > 1e-3f, NF_E_INVALID, "nf_render_forward: radius too small");
    if (a->n_rays == 0) return NF_OK;
    NF_REQUIRE(a->n_rays > 0 && (size_t)a->n_rays * (a->n_coarse + NI) < (size_t)1 << 30, NF_E_UNSUPPORTED,
               "nf_render_forward: too many samples in one call (chunk the rays)");
    const WsLayout L = ws_layout(a->n_rays, a->n_coarse, NI);
    NF_REQUIRE(a->workspace_bytes >= L.total, NF_E_WORKSPACE, "nf_render_forward: workspace %zu < %zu",
               a->workspace_bytes, L.total);
    char* b = (char*)a->workspace;
    StageArgs p;
    p.g = grid_view(a->grid_ws);
    p.particles = a->particles;
    p.rays = a->rays;
    p.n_rays = a->n_rays;
    p.ro[0] = a->ro[0]; p.ro[1] = a->ro[1]; p.ro[2] = a->ro[2];
    p.radius = a->radius;
    p.K = a->K;
    p.use_mask = a->use_mask; p.white_bg = a->white_background; p.mode = a->mode;
    p.z_coarse = a->z_coarse; p.u_imp = a->u_importance;
    p.S0 = a->n_coarse; p.n_imp = NI; p.S1 = a->n_coarse + NI;
    const bool want0 = a->mode != NF_RENDER_FINE;
    p.rgb0 = want0 ? a->rgb0 : nullptr; p.depth0 = want0 ? a->depth0 : nullptr;
    p.opac0 = want0 ? a->opacity0 : nullptr; p.mask0 = want0 ? a->mask0 : nullptr;
    p.num_nn0 = want0 ? (long long*)a->num_nn0 : nullptr;
    p.rgb1 = a->rgb1; p.depth1 = a->depth1; p.opac1 = a->opacity1; p.mask1 = a->mask1;
    p.num_nn1 = (long long*)a->num_nn1;
    p.counters = (int*)(b + L.counters);
    p.act0 = (unsigned*)(b + L.act0); p.act1 = (unsigned*)(b + L.act1);
    p.z1 = (float*)(b + L.z1);
    p.rec0 = (float*)(b + L.rec0); p.rowid0 = (int*)(b + L.rowid0); p.out0 = (float4*)(b + L.out0); p.cap0 = L.cap0;
    p.rec1 = (float*)(b + L.rec1); p.rowid1 = (int*)(b + L.rowid1); p.out1 = (float4*)(b + L.out1); p.cap1 = L.cap1;

    NF_CUDA_OK(cudaMemsetAsync(p.counters, 0, 64, st));
    const int threads = WARPS_PER_BLOCK * 32;
    const int grid = min((a->n_rays + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, num_sms() * 8);

    StageTimer tm(st);
    // ---- stage Q0
    if (L.ns0 == 2) k_stage_q0<2><<<grid, threads, 0, st>>>(p);
    else k_stage_q0<4><<<grid, threads, 0, st>>>(p);
    NF_LAUNCH_OK();
    tm.mark();
    // ---- coarse network
    mlp::KernelArgs m;
    m.packed = (const uint8_t*)a->weights_coarse;
    m.records = p.rec0; m.rowid = p.rowid0; m.n_rows_dev = p.counters + 0; m.n_rows_host = 0; m.n_rows_cap = L.cap0;
    m.n_layers = (a->mode == NF_RENDER_FINE) ? 8 : 10;
    m.desc_swap = 0;
    m.out4 = p.out0;
    int rc = mlp::launch(m, a->dtype, st);
    if (rc != NF_OK) return rc;
    tm.mark();
    if (!fine) {
        if (L.ns0 == 2) k_stage_fin<2, true><<<grid, threads, 0, st>>>(p);
        else k_stage_fin<4, true><<<grid, threads, 0, st>>>(p);
        NF_LAUNCH_OK();
        tm.mark(); tm.mark(); tm.mark();
    } else {
        rc = (L.ns0 == 2) ? launch_mid<2>(L.ns1, grid, p, st) : launch_mid<4>(L.ns1, grid, p, st);
        if (rc != NF_OK) return rc;
        tm.mark();
        m.packed = (const uint8_t*)a->weights_fine;
        m.records = p.rec1; m.rowid = p.rowid1; m.n_rows_dev = p.counters + 1; m.n_rows_cap = L.cap1;
        m.n_layers = 10;
        m.out4 = p.out1;
        rc = mlp::launch(m, a->dtype, st);
        if (rc != NF_OK) return rc;
        tm.mark();
        switch (L.ns1) {
            case 4: k_stage_fin<4, false><<<grid, threads, 0, st>>>(p); break;
            case 6: k_stage_fin<6, false><<<grid, threads, 0, st>>>(p); break;
            default: k_stage_fin<8, false><<<grid, threads, 0, st>>>(p); break;
        }
        NF_LAUNCH_OK();
        tm.mark();
    }
    if (a->stats) NF_CUDA_OK(cudaMemcpyAsync(a->stats, p.counters, 16, cudaMemcpyDeviceToDevice, st));
    return NF_OK;
}
