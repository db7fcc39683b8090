// nf_render_bwd.cu -- backward pass of the renderer forward (nf_render.cu), for training through RenderNet.
//
// replaces: autograd through RenderNet.forward (models/renderer.py:211-270) as run by loss.backward() at
//           trainer/trainer_e2e.py:277 and trainer/trainer_renderer.py:96 -- gradients reach the two NeRF MLPs' parameters
//           and, through the neighbour positions (pytorch3d masked_gather is differentiable, ball_query's dists / idx are
//           not), the particle positions, i.e. the transition model.  The importance samples are detached
//           (utils/ray_utils.py:224), so sample positions carry no gradient and the two passes are independent.
//
// Needs the forward's workspace with the neighbour lists kept (NF_RENDER_SAVE_NEIGHBORS): per pass
//   k_composite_bwd   one warp per ray: recomputes alpha / transmittance from the stored per-sample (r,g,b,sigma),
//                     turns d rgb / d depth / d opacity into d(pre-sigmoid rgb, sigma) of every evaluated sample
//                     (models/renderer.py:182-208 differentiated: suffix sums by a warp scan).
//   nf_nerf_mlp_backward (nf_mlp_bwd.cu)  tcgen05 dgrad + wgrad -> d(encoded features) per row, d(parameters).
//   k_geom_bwd        one warp per record row, lane = neighbour: Embedding backward (models/nerf.py:21-38) on the ten
//                     particle-dependent record values, then smoothing_position / variance / smoothed direction
//                     (models/renderer.py:96-109, 137-175) differentiated w.r.t. the K neighbour positions; atomicAdd
//                     into d particles.
#include "nf_common.cuh"
#include "nf_mlp.cuh"

extern "C" int nf_nerf_mlp_backward(const void*, const void*, int, const float*, const int32_t*, const float*, int, float*, float*,
                                    void*, size_t, void*);
extern "C" size_t nf_nerf_mlp_backward_workspace_bytes(int);

namespace nf {
namespace render_bwd {

constexpr int WARPS = 8;
constexpr int DFEAT_W = 272;

struct CompArgs {
    const float* rays; int n_rays;
    const float* z_shared;      // (S) depths shared by all rays (coarse pass) or NULL
    const float* z_per_ray;     // (n_rays, z_stride) (fine pass; perturbed coarse pass) or NULL
    int z_stride;
    const float* noise;         // optional (n_rays, S): sigma noise of the forward
    int S;
    const unsigned* act; int act_stride;
    const unsigned char* miss;
    const float4* out4;         // (n_rays, S) forward (r,g,b,sigma); valid where evaluated
    int use_mask, white_bg;
    const float* d_rgb;         // (n_rays,3) or NULL
    const float* d_depth;       // (n_rays) or NULL
    const float* d_opac;        // (n_rays) or NULL
    float4* dout4;              // (n_rays, S): written for evaluated samples
};

template <int NS>
__global__ void __launch_bounds__(WARPS * 32) k_composite_bwd(const CompArgs p) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S = p.S;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        if (p.miss[ray]) continue;           // no evaluated sample on this ray
        const float* r = p.rays + (size_t)ray * 6;
        const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
        const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
        const float gr = p.d_rgb ? p.d_rgb[3 * (size_t)ray] : 0.f, gg = p.d_rgb ? p.d_rgb[3 * (size_t)ray + 1] : 0.f,
                    gb = p.d_rgb ? p.d_rgb[3 * (size_t)ray + 2] : 0.f;
        const float gdep = p.d_depth ? p.d_depth[ray] : 0.f, gacc = p.d_opac ? p.d_opac[ray] : 0.f;
        const float white = p.white_bg ? 1.f : 0.f;
        float z[NS], alpha[NS], T[NS], delta[NS], gw[NS], x[NS];
        float4 c[NS];
        bool ev[NS];
        // ---- forward recompute (ray_composite of nf_render.cu)
        float carry = 1.f;
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const int s = slot * 32 + lane;
            const bool in = s < S;
            z[slot] = p.z_shared ? __ldg(p.z_shared + min(s, S - 1)) : p.z_per_ray[(size_t)ray * p.z_stride + min(s, S - 1)];
        }
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const int s = slot * 32 + lane;
            const bool in = s < S;
            const unsigned bits = p.act[(size_t)ray * p.act_stride + slot];
            ev[slot] = in && (p.use_mask ? ((bits >> lane) & 1u) : true);
            c[slot] = ev[slot] ? p.out4[(size_t)ray * S + s] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.noise && in) c[slot].w += p.noise[(size_t)ray * S + s];
            float zn = __shfl_down_sync(NF_FULL, z[slot], 1);
            float z_next0 = 0.f;
            if (slot + 1 < NS) z_next0 = __shfl_sync(NF_FULL, z[slot + 1], 0);
            if (lane == 31) zn = z_next0;
            float dl = (s == S - 1) ? 1e10f : zn - z[slot];
            dl *= dnorm;
            delta[slot] = dl;
            alpha[slot] = in ? 1.0f - expf(-dl * fmaxf(c[slot].w, 0.f)) : 0.f;
            const float a1 = in ? (1.0f - alpha[slot] + 1e-10f) : 1.f;
            float P = a1;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float t = __shfl_up_sync(NF_FULL, P, off);
                if (lane >= off) P *= t;
            }
            float excl = __shfl_up_sync(NF_FULL, P, 1);
            if (lane == 0) excl = 1.f;
            T[slot] = carry * excl;
            carry *= __shfl_sync(NF_FULL, P, 31);
        }
        // ---- d L / d w_i and the suffix sums  S_i = sum_{k > i} (dL/dw_k) w_k
        float run = 0.f;
        float incl[NS];
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const float w = alpha[slot] * T[slot];
            gw[slot] = gr * (c[slot].x - white) + gg * (c[slot].y - white) + gb * (c[slot].z - white) + gdep * z[slot] + gacc;
            x[slot] = (slot * 32 + lane < S) ? gw[slot] * w : 0.f;
            float P = x[slot];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float t = __shfl_up_sync(NF_FULL, P, off);
                if (lane >= off) P += t;
            }
            incl[slot] = run + P;
            run += __shfl_sync(NF_FULL, P, 31);
        }
        const float total = run;
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const int s = slot * 32 + lane;
            if (!ev[slot]) continue;
            const float a1 = 1.0f - alpha[slot] + 1e-10f;
            const float dalpha = gw[slot] * T[slot] - (total - incl[slot]) / a1;
            const float dsigma = c[slot].w > 0.f ? dalpha * delta[slot] * (1.0f - alpha[slot]) : 0.f;
            const float w = alpha[slot] * T[slot];
            p.dout4[(size_t)ray * S + s] = make_float4(w * gr * c[slot].x * (1.0f - c[slot].x), w * gg * c[slot].y * (1.0f - c[slot].y),
                                                       w * gb * c[slot].z * (1.0f - c[slot].z), dsigma);
        }
    }
}

struct GeomArgs {
    const float* dfeat;      // (rows, 272)
    const float* rec;        // (rows, 16)
    const int* nbr;          // (rows, K)
    const float* particles;  // (P, 3)
    int n_rows, K;
    float radius;
    float ro[3];
    const float* ro_dev;
    int include_ray, same_smooth;
    float* dparticles;       // (P, 3), accumulated
};

__global__ void __launch_bounds__(WARPS * 32) k_geom_bwd(const GeomArgs p) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float rox = p.ro_dev ? __ldg(p.ro_dev) : p.ro[0], roy = p.ro_dev ? __ldg(p.ro_dev + 1) : p.ro[1],
                roz = p.ro_dev ? __ldg(p.ro_dev + 2) : p.ro[2];
    const float radius = p.radius;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.n_rows; row += nwarps) {
        const float* df = p.dfeat + (size_t)row * DFEAT_W;
        const float* rc = p.rec + (size_t)row * 16;
        // ---- Embedding backward for the ten particle-dependent values: lane 0 density, 1-3 smoothed, 4-6 variance, 7-9 smoothed dir
        float gv = 0.f;
        if (lane < 10) {
            int B, C, L, comp, ridx;
            if (lane == 0) { B = 63; C = 1; L = 4; comp = 0; ridx = 3; }
            else if (lane < 4) { B = 72; C = 3; L = 10; comp = lane - 1; ridx = 4 + comp; }
            else if (lane < 7) { B = 135; C = 3; L = 10; comp = lane - 4; ridx = 7 + comp; }
            else { B = 208 + 27; C = 3; L = 4; comp = lane - 7; ridx = 13 + comp; }
            const float v = rc[ridx];
            gv = df[B + comp];
            float fr = 1.f;
            for (int f = 0; f < L; ++f) {
                float s, c;
                sincosf(fr * v, &s, &c);
                gv += fr * (df[B + C + 2 * C * f + comp] * c - df[B + C + 2 * C * f + C + comp] * s);
                fr *= 2.f;
            }
        }
        const float g_den = __shfl_sync(NF_FULL, gv, 0);
        float g_sm[3], g_var[3], g_sd[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            g_sm[a] = __shfl_sync(NF_FULL, gv, 1 + a);
            g_var[a] = __shfl_sync(NF_FULL, gv, 4 + a);
            g_sd[a] = __shfl_sync(NF_FULL, gv, 7 + a);
        }
        // ---- local geometry, forward values (lane = neighbour slot)
        const float qx = rc[0], qy = rc[1], qz = rc[2];
        const int j = lane < p.K ? p.nbr[(size_t)row * p.K + lane] : -1;
        const bool valid = j >= 0;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (valid) { nx = __ldg(p.particles + 3 * (size_t)j); ny = __ldg(p.particles + 3 * (size_t)j + 1); nz = __ldg(p.particles + 3 * (size_t)j + 2); }
        const float ex = nx - qx, ey = ny - qy, ez = nz - qz;
        const float dist = sqrtf(ex * ex + ey * ey + ez * ez);
        const float t = dist / radius;
        const float w = valid ? fmaxf(1.0f - t * t * t, 0.f) : 0.f;
        const int cnt = __popc(__ballot_sync(NF_FULL, valid));
        float W = warp_sum(w);
        if (cnt < p.K) {      // padded slots = a phantom particle at the origin (no gradient: it is a constant)
            const float tq = sqrtf(qx * qx + qy * qy + qz * qz) / radius;
            W += (float)(p.K - cnt) * fmaxf(1.0f - tq * tq * tq, 0.f);
        }
        const float Nx = warp_sum(w * nx), Ny = warp_sum(w * ny), Nz = warp_sum(w * nz);
        const float den = W + 1e-12f;
        const float sx = Nx / den, sy = Ny / den, sz = Nz / den;
        // exclude_ray=False: the encoded position is x (1 - alpha) + smoothed * alpha -> its gradient reaches the neighbours scaled by alpha
        const int nv_cnt = __popc(__ballot_sync(NF_FULL, valid && dist2_exact(qx, qy, qz, nx, ny, nz) != 0.f));
        const float al = p.include_ray ? ((p.same_smooth || nv_cnt > 20) ? 0.9f : 0.1f) : 1.0f;
        const float bx = qx * (1.0f - al) + sx * al, by = qy * (1.0f - al) + sy * al, bz = qz * (1.0f - al) + sz * al;
        const float ux = bx - rox, uy = by - roy, uz = bz - roz;
        const float un = sqrtf(ux * ux + uy * uy + uz * uz);
        const float dxs = ux / un, dys = uy / un, dzs = uz / un;
        const float dot = dxs * g_sd[0] + dys * g_sd[1] + dzs * g_sd[2];
        // d smoothed (direct + through the smoothed direction)
        const float gsx = al * (g_sm[0] + (g_sd[0] - dxs * dot) / un), gsy = al * (g_sm[1] + (g_sd[1] - dys * dot) / un),
                    gsz = al * (g_sm[2] + (g_sd[2] - dzs * dot) / un);
        const float dNx = gsx / den, dNy = gsy / den, dNz = gsz / den;
        const float dW = g_den - (gsx * sx + gsy * sy + gsz * sz) / den;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (valid) {
            gx = w * dNx; gy = w * dNy; gz = w * dNz;
            if (w > 0.f && dist > 0.f) {
                const float g_w = dNx * nx + dNy * ny + dNz * nz + dW;
                const float g_dist = g_w * (-3.0f * t * t / radius);
                gx += g_dist * ex / dist; gy += g_dist * ey / dist; gz += g_dist * ez / dist;
            }
        }
        // ---- variance (models/renderer.py:160-166): mask = ball_query's squared distance != 0
        const bool vv = valid && dist2_exact(qx, qy, qz, nx, ny, nz) != 0.f;
        const float nvf = (float)__popc(__ballot_sync(NF_FULL, vv)) + 1e-12f;
        const float mx = warp_sum(vv ? ex : 0.f) / nvf, my = warp_sum(vv ? ey : 0.f) / nvf, mz = warp_sum(vv ? ez : 0.f) / nvf;
        const float ddx = vv ? ex - mx : 0.f, ddy = vv ? ey - my : 0.f, ddz = vv ? ez - mz : 0.f;
        const float sdx = warp_sum(ddx), sdy = warp_sum(ddy), sdz = warp_sum(ddz);
        if (vv) {
            gx += g_var[0] * (2.0f * ddx / nvf - 2.0f * sdx / (nvf * nvf));
            gy += g_var[1] * (2.0f * ddy / nvf - 2.0f * sdy / (nvf * nvf));
            gz += g_var[2] * (2.0f * ddz / nvf - 2.0f * sdz / (nvf * nvf));
        }
        if (valid) {
            atomicAdd(p.dparticles + 3 * (size_t)j, gx);
            atomicAdd(p.dparticles + 3 * (size_t)j + 1, gy);
            atomicAdd(p.dparticles + 3 * (size_t)j + 2, gz);
        }
    }
}

struct BwdWs {
    size_t dout, dfeat, mlp, total;
};
inline BwdWs bwd_ws(int R, int S0, int NI, int rows0, int rows1) {
    BwdWs L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
    const int rmax = rows0 > rows1 ? rows0 : rows1;
    L.dout = take(sizeof(float4) * (size_t)R * (S0 + NI));
    L.dfeat = take(sizeof(float) * DFEAT_W * (size_t)(rmax > 0 ? rmax : 1));
    L.mlp = take(nf_nerf_mlp_backward_workspace_bytes(rmax > 0 ? rmax : 1));
    L.total = o;
    return L;
}

static int launch_comp(int ns, int grid, const CompArgs& c, cudaStream_t st) {
    switch (ns) {
        case 2: k_composite_bwd<2><<<grid, WARPS * 32, 0, st>>>(c); break;
        case 4: k_composite_bwd<4><<<grid, WARPS * 32, 0, st>>>(c); break;
        case 6: k_composite_bwd<6><<<grid, WARPS * 32, 0, st>>>(c); break;
        case 8: k_composite_bwd<8><<<grid, WARPS * 32, 0, st>>>(c); break;
        default: set_error("nf_render_backward: unsupported sample count"); return NF_E_UNSUPPORTED;
    }
    NF_LAUNCH_OK();
    return NF_OK;
}

}  // namespace render_bwd
}  // namespace nf

using namespace nf;
using namespace nf::render_bwd;

extern "C" size_t nf_render_backward_workspace_bytes(int n_rays, int n_coarse, int n_importance, int rows_coarse, int rows_fine) {
    if (n_rays <= 0 || n_coarse <= 0 || n_importance < 0 || rows_coarse < 0 || rows_fine < 0) return 0;
    return bwd_ws(n_rays, n_coarse, n_importance, rows_coarse, rows_fine).total;
}

extern "C" int nf_render_backward(const nf_render_bwd_args* b, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(b && b->fwd, NF_E_INVALID, "nf_render_backward: null args");
    const nf_render_args* a = b->fwd;
    NF_REQUIRE(a->flags & NF_RENDER_SAVE_NEIGHBORS, NF_E_INVALID,
               "nf_render_backward: the forward call must keep the neighbour lists (NF_RENDER_SAVE_NEIGHBORS)");
    NF_REQUIRE(a->workspace && b->workspace && b->d_particles, NF_E_INVALID, "nf_render_backward: null pointer");
    if (a->n_rays == 0) return NF_OK;
    const bool fine = a->mode != NF_RENDER_COARSE;
    const int NI = fine ? a->n_importance : 0;
    const int S0 = a->n_coarse, S1 = S0 + NI;
    nf_render_ws_view v;
    int rc = nf_render_workspace_view(a->n_rays, S0, NI, a->K, a->flags, &v);
    if (rc != NF_OK) return rc;
    const char* ws = (const char*)a->workspace;
    int counters[16];
    NF_CUDA_OK(cudaMemcpyAsync(counters, ws + v.counters, 64, cudaMemcpyDeviceToHost, st));
    NF_CUDA_OK(cudaStreamSynchronize(st));       // the one host sync of the backward pass: row counts size the launches
    const int rows0 = counters[0] < v.cap0 ? counters[0] : v.cap0, rows1 = fine ? (counters[1] < v.cap1 ? counters[1] : v.cap1) : 0;
    const BwdWs L = bwd_ws(a->n_rays, S0, NI, rows0, rows1);
    NF_REQUIRE(b->workspace_bytes >= L.total, NF_E_WORKSPACE, "nf_render_backward: workspace %zu < %zu", b->workspace_bytes, L.total);
    char* bw = (char*)b->workspace;
    const int grid = min((a->n_rays + WARPS - 1) / WARPS, num_sms() * 8);
    const unsigned char* miss = (const unsigned char*)(ws + v.miss);

    for (int pass = 0; pass < 2; ++pass) {
        const bool coarse = pass == 0;
        if (coarse && (a->mode == NF_RENDER_FINE || !(b->d_rgb0 || b->d_depth0 || b->d_opacity0))) continue;
        if (!coarse && (!fine || !(b->d_rgb1 || b->d_depth1 || b->d_opacity1))) continue;
        const int rows = coarse ? rows0 : rows1;
        if (rows == 0) continue;
        const void* wfw = coarse ? a->weights_coarse : a->weights_fine;
        const void* wbw = coarse ? b->weights_coarse_bwd : b->weights_fine_bwd;
        float* dpar = coarse ? b->d_params_coarse : b->d_params_fine;
        NF_REQUIRE(wfw && wbw && dpar, NF_E_INVALID, "nf_render_backward: null weights / parameter-gradient buffer");
        CompArgs c;
        c.rays = a->rays; c.n_rays = a->n_rays;
        c.z_shared = (coarse && a->z_stride == 0) ? a->z_coarse : nullptr;
        c.z_per_ray = coarse ? (a->z_stride ? a->z_coarse : nullptr) : (const float*)(ws + v.z1);
        c.z_stride = coarse ? a->z_stride : S1;
        c.noise = coarse ? a->noise0 : a->noise1;
        c.S = coarse ? S0 : S1;
        c.act = (const unsigned*)(ws + (coarse ? v.act0 : v.act1));
        c.act_stride = coarse ? v.act_stride0 : v.act_stride1;
        c.miss = miss;
        c.out4 = (const float4*)(ws + (coarse ? v.out0 : v.out1));
        c.use_mask = a->use_mask; c.white_bg = a->white_background;
        c.d_rgb = coarse ? b->d_rgb0 : b->d_rgb1;
        c.d_depth = coarse ? b->d_depth0 : b->d_depth1;
        c.d_opac = coarse ? b->d_opacity0 : b->d_opacity1;
        c.dout4 = (float4*)(bw + L.dout);
        rc = launch_comp(c.act_stride, grid, c, st);
        if (rc != NF_OK) return rc;
        const float* rec = (const float*)(ws + (coarse ? v.rec0 : v.rec1));
        const int* rowid = (const int*)(ws + (coarse ? v.rowid0 : v.rowid1));
        rc = nf_nerf_mlp_backward(wfw, wbw, a->dtype, rec, rowid, (const float*)(bw + L.dout), rows, (float*)(bw + L.dfeat), dpar,
                                  bw + L.mlp, L.total - L.mlp, stream_);
        if (rc != NF_OK) return rc;
        GeomArgs g;
        g.dfeat = (const float*)(bw + L.dfeat); g.rec = rec;
        g.nbr = (const int*)(ws + (coarse ? v.nbr0 : v.nbr1));
        g.particles = a->particles; g.n_rows = rows; g.K = a->K; g.radius = a->radius;
        g.ro[0] = a->ro[0]; g.ro[1] = a->ro[1]; g.ro[2] = a->ro[2]; g.ro_dev = a->ro_dev;
        g.include_ray = a->include_ray; g.same_smooth = a->same_smooth_factor;
        g.dparticles = b->d_particles;
        k_geom_bwd<<<min((rows + WARPS - 1) / WARPS, num_sms() * 8), WARPS * 32, 0, st>>>(g);
        NF_LAUNCH_OK();
    }
    return NF_OK;
}
