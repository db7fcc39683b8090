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
// One warp owns one ray; inside a ray, sample s lives in lane s%32 of 32-sample step s/32.
//
// stage Q0 : coarse depths -> sample positions -> first-K ball query (search_scs / search_stream, see
//            ray_query_group) -> num_nn, "all K slots valid" bitmask, and one 64-byte geometry record per evaluated
//            sample appended to a compact list (rows are handed out per 32-sample step: one atomic, no holes).
//            Rays that never come within reach of the particle set are flagged here and only written later on.
// [MLP]    : nf_mlp.cu over the compact list -> (r,g,b,sigma) scattered to a dense per-sample array.
// stage MID: alpha-composite the coarse samples (warp scans), emit rgb0/depth0/opacity0/mask_0, build
//            the piecewise-constant pdf, draw the importance samples by inverse CDF, merge with the
//            coarse depths (rank merge in shared memory), then run the ball query for the merged samples (the
//            coarse ones among them reuse stage Q0's records and counts).
// [MLP]    : fine network.
// stage FIN: alpha-composite the fine samples -> rgb1/depth1/opacity1/mask_1.
#include <stdlib.h>
#include <vector>

#include "nf_common.cuh"
#include "nf_mlp.cuh"

namespace nf {
namespace render {

constexpr int WARPS_PER_BLOCK = 8;

struct StageArgs {
    GridView g;
    const float* particles;
    int n_points;
    const float* rays;
    int n_rays;
    float ro[3];
    const float* ro_dev;                       // optional: camera position read from device memory (overrides ro)
    float radius;
    int K;
    int use_mask, white_bg, mode;
    int include_ray, same_smooth;              // exclude_ray=False: blend the sample position into the smoothed position
    int search_mode;                           // 0 = index-order stream, 1 = sorted-candidate sweep
    float sub_span;                            // sweep: max depth span of one gather
    int sub_look;                              // sweep: ... and max samples ahead it reaches
    int solo_max_occ, peel_lanes, peel_from;   // stream: tuning, see search_stream
    const float* z_coarse;
    const float* u_imp;
    int z_stride, u_stride;                    // 0: one table shared by all rays;  else per-ray rows (perturb > 0)
    const float* noise0;                       // optional (R, S0): added to sigma before the ReLU (noise_std > 0)
    const float* noise1;                       // optional (R, S1)
    int S0, n_imp, S1;
    // outputs
    float *rgb0, *depth0, *opac0, *mask0;
    long long* num_nn0;
    float *rgb1, *depth1, *opac1, *mask1;
    long long* num_nn1;
    // workspace
    int* counters;       // [0] rows coarse, [1] rows fine, [2] active coarse, [3] active fine,
                         // [4..7] fine-pass search statistics: group scans, solo (row-scan) queries, scan steps / 64,
                         // candidates tested + row-scan iterations / 64; [8..11] the same for the coarse pass
    unsigned* act0;      // (R, NS0)
    int* base0;          // (R, NS0) first row of each 32-sample coarse step in rec0
    unsigned char* cnt0; // (R, S0) neighbour count of every coarse sample
    unsigned char* miss; // (R) 1: the ray never comes within reach of the particle set (use_mask only)
    unsigned* act1;      // (R, NS1)
    float* z1;           // (R, S1)
    float* rec0; int* rowid0; float4* out0; int cap0;
    float* rec1; int* rowid1; float4* out1; int cap1;
    int* nbr0; int* nbr1;   // optional (NF_RENDER_SAVE_NEIGHBORS): (rows, K) neighbour indices of every record row, -1 padded
};

// ------------------------------------------------------------------------------------------------
// Search flavour 1, "index-order stream" (any P).  One LANE per sample: the warp streams the particle set in
// original index order, 128 per step, through a per-warp bitmap over half-resolution grid cells (the union of
// the cells within reach of a still-unfinished lane), broadcasts the survivors and every lane appends its hits
// until it has K.  Lanes with a sparse cell neighbourhood, and stragglers, are answered one at a time by the
// warp-cooperative row scan (warp_first_k_rows).  smem: bm[BM_WORDS] | hitbuf[HITBUF].
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int sel_stride(int K) { return K | 1; }
__host__ __device__ inline size_t sel_bytes(int K) { return (size_t)32 * sel_stride(K) * sizeof(int); }

constexpr int BM_BITS = 16384;
constexpr int BM_WORDS = BM_BITS / 32;

__device__ __forceinline__ int search_stream(const StageArgs& p, int lane, float qx, float qy, float qz, bool search,
                                             int occ, QueryStats& qs, unsigned* bm, int* hitbuf, int* sel) {
    const GridHeader* h = p.g.hdr;
    const int K = p.K;
    const float radius = p.radius;
    const float r2 = __fmul_rn(radius, radius);
    const float pad = radius * 1.001f + 1e-6f;
    const int P = h->n;
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    int cnt = 0;
    // Sparse neighbourhoods (few points in the 3x3x3 cell block) cannot fill K quickly in an index-order
    // stream -- a lane with fewer than K neighbours would drag the whole group through all P particles --
    // so those lanes skip the group scan and are answered one at a time by the warp-cooperative row scan
    // below, as are stragglers still unfinished when the scan has gone PEEL_STEPS steps.
    bool solo = search && occ < p.solo_max_occ;
    if (__any_sync(NF_FULL, search && !solo)) {
        bool done = !search || solo;
        int built_for = 0;
        ++qs.n_lock;
        // half-resolution cell range of this lane's ball (<= 5 fine cells per axis when cell > reach)
        const float inv2 = __fmul_rn(inv, 2.0f);
        const float fcell = 0.5f * h->cell;
        const int fnx = 2 * nx, fny = 2 * ny, fnz = 2 * nz;
        const int flox = cell_coord(qx - pad, ox, inv2, fnx), fhix = cell_coord(qx + pad, ox, inv2, fnx);
        const int floy = cell_coord(qy - pad, oy, inv2, fny), fhiy = cell_coord(qy + pad, oy, inv2, fny);
        const int floz = cell_coord(qz - pad, oz, inv2, fnz), fhiz = cell_coord(qz + pad, oz, inv2, fnz);
        const float rm = pad + 1e-3f * fcell;          // cull margin covers the rounding of the binning
        const float rm2 = rm * rm;
        int cnext[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) cnext[t] = __ldg(p.g.fine_of + 32 * t + lane);   // padded by 128 entries of -1
        for (int j0 = 0; j0 < P; j0 += 128) {
            const unsigned pending = __ballot_sync(NF_FULL, !done);
            if (!pending) break;
            const int npend = __popc(pending);
            if (npend <= p.peel_lanes && j0 >= p.peel_from) {
                solo = solo || !done;
                break;
            }
            if (npend * 3 <= built_for || built_for == 0) {
                // (re)build the bitmap: fine cells whose box comes within reach of an unfinished lane
                __syncwarp();
                for (int w = lane; w < BM_WORDS; w += 32) bm[w] = 0u;
                __syncwarp();
                if (!done) {
                    for (int cz = floz; cz <= fhiz; ++cz) {
                        const float bz = oz + (float)cz * fcell;
                        const float dz = fmaxf(fmaxf(bz - qz, qz - (bz + fcell)), 0.f);
                        // clamped boundary cells also hold everything beyond them: never cull those
                        const bool ez = (cz == 0) || (cz == fnz - 1);
                        for (int cy = floy; cy <= fhiy; ++cy) {
                            const float by = oy + (float)cy * fcell;
                            const float dy = fmaxf(fmaxf(by - qy, qy - (by + fcell)), 0.f);
                            const bool ey = ez || (cy == 0) || (cy == fny - 1);
                            const float dzy = dz * dz + dy * dy;
                            for (int cx = flox; cx <= fhix; ++cx) {
                                const float bx = ox + (float)cx * fcell;
                                const float dx = fmaxf(fmaxf(bx - qx, qx - (bx + fcell)), 0.f);
                                const bool e = ey || (cx == 0) || (cx == fnx - 1);
                                if (e || dzy + dx * dx < rm2) {
                                    const int c = (cz * fny + cy) * fnx + cx;
                                    atomicOr(&bm[(c & (BM_BITS - 1)) >> 5], 1u << (c & 31));
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                built_for = npend;
            }
            ++qs.it_lock;
            // this step's cell ids were prefetched; fetch the next step's while we work
            int c[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) c[t] = cnext[t];
            if (j0 + 128 < P) {
#pragma unroll
                for (int t = 0; t < 4; ++t) cnext[t] = __ldg(p.g.fine_of + j0 + 128 + 32 * t + lane);
            }
            // lanes whose particle passes the bitmap fetch it themselves: coalesced, 4 loads in flight
            float4 pp[4];
            bool inb[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                inb[t] = (c[t] >= 0) && ((bm[(c[t] & (BM_BITS - 1)) >> 5] >> (c[t] & 31)) & 1u);
                pp[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb[t]) pp[t] = __ldg(p.g.orig4 + j0 + 32 * t + lane);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                unsigned m = __ballot_sync(NF_FULL, inb[t]);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const float cx = __shfl_sync(NF_FULL, pp[t].x, b);
                    const float cy = __shfl_sync(NF_FULL, pp[t].y, b);
                    const float cz = __shfl_sync(NF_FULL, pp[t].z, b);
                    ++qs.it_rows;     // statistics: candidates tested
                    if (!done && dist2_exact(qx, qy, qz, cx, cy, cz) < r2) {
                        sel[lane * sel_stride(K) + cnt] = j0 + 32 * t + b;
                        ++cnt;
                        done = cnt >= K;
                    }
                }
            }
        }
    }
    // ---- solo lanes: exact first-K by the whole warp over the <= 9 cell rows around that one sample
    {
        unsigned ms = __ballot_sync(NF_FULL, solo);
        while (ms) {
            const int b = __ffs(ms) - 1;
            ms &= ms - 1;
            const float sx = __shfl_sync(NF_FULL, qx, b), sy = __shfl_sync(NF_FULL, qy, b), sz = __shfl_sync(NF_FULL, qz, b);
            int best = 0x7fffffff;
            ++qs.n_rows;
            const int n = warp_first_k_rows(p.g, sx, sy, sz, radius, K, lane, best, qs.it_rows, hitbuf);
            if (lane < n) sel[b * sel_stride(K) + lane] = best;
            if (lane == b) cnt = n;
        }
        __syncwarp();
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------------
// Search flavour 2, "sorted-candidate sweep" (P <= SCS_MAX_POINTS; the default).  The first-K-by-index rule
// wants each sample's in-ball particles in ascending original index.  Consecutive samples of a ray whose depths
// span <= sub_span form a sub-group; the warp
//   1. gathers every particle within reach of the sub-group's segment (capsule test) from the <= 4x4 cell rows
//      around it (coalesced 16-byte records of the cell-sorted copy) and sets bit[index] in a per-warp
//      shared-memory bitmap over the index space -- a counting sort by index, for free;
//   2. walks the bitmap in ascending order, 1024 indices per batch, compacting set bits into a small ring
//      (ballot-free: popc + warp scan), and
//   3. takes 128 candidates at a time (four per LANE, positions in registers) and sweeps the still-unfinished
//      samples over them: broadcast the sample, exact distance tests, ballots; hit lanes write their index to
//      the sample's next free slots.  A sample retires at K hits, the walk stops when all have retired.
// Work is proportional to the candidates near the segment, not to P, and a sample with fewer than K neighbours
// costs one pass over its own neighbourhood only.  smem: ibm[ceil(P/32)] | ring[SCS_RING] (u16).
// ------------------------------------------------------------------------------------------------
constexpr int SCS_MAX_POINTS = 65536;
constexpr int SCS_BLOCK = 128;                  // candidates per sweep (4 per lane)
constexpr int SCS_RING = 1024 + 2 * SCS_BLOCK;  // sorted candidate list: un-swept rest (< block) + one 1024-index batch + read slack

__host__ __device__ inline size_t scs_smem_bytes(int P) {
    return (size_t)((P + 1023) / 1024) * 128 + SCS_RING * sizeof(unsigned short);
}

__device__ __forceinline__ int search_scs(const StageArgs& p, int lane, const float (&o)[3], const float (&d)[3],
                                          float zv, float qx, float qy, float qz, bool search, float& gz0, float& gz1,
                                          int& lst_len, int& lst_word, const float* zs, int S, int s0 /*sample index of lane 0*/,
                                          QueryStats& qs, unsigned* ibm, int* sel) {
    const GridHeader* h = p.g.hdr;
    const int K = p.K;
    const float radius = p.radius;
    const float r2 = __fmul_rn(radius, radius);
    const float pad = radius * 1.001f + 1e-4f;      // also absorbs the rounding of o + d*z along the segment
    const float pad2 = pad * pad;
    const int P = h->n;
    const int nwords = ((P + 1023) >> 10) << 5;          // multiple of 32 words
    unsigned short* ring = reinterpret_cast<unsigned short*>(ibm + nwords);
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const unsigned lt = (1u << lane) - 1u;
    const int KP = sel_stride(K);
    int cnt = 0;
    unsigned todo = __ballot_sync(NF_FULL, search);
    while (todo) {
        // ---- next sub-group: the searching lanes inside the depth interval [gz0, gz1] whose candidates the
        //      bitmap holds; when the first pending lane lies outside it, gather a new interval of sub_span
        //      starting there (it also serves the following 32-sample steps of this ray: importance samples
        //      cluster, so several steps usually share one gather)
        const int a = __ffs(todo) - 1;
        const float za = __shfl_sync(NF_FULL, zv, a);
        const bool fresh = !(za >= gz0 && za <= gz1);
        if (fresh) {      // at most sub_span deep and at most sub_look samples ahead (importance samples cluster)
            gz0 = za;
            gz1 = fminf(za + p.sub_span, zs[min(s0 + a + p.sub_look, S - 1)]);
        }
        const unsigned sub = __ballot_sync(NF_FULL, search && lane >= a && zv <= gz1) & todo;
        todo &= ~sub;
        if (fresh) {
        const float ax = __shfl_sync(NF_FULL, qx, a), ay = __shfl_sync(NF_FULL, qy, a), az = __shfl_sync(NF_FULL, qz, a);
        const float span = gz1 - gz0;
        const float ex = d[0] * span, ey = d[1] * span, ez = d[2] * span;
        const float len2 = ex * ex + ey * ey + ez * ez;
        const float inv_len2 = len2 > 0.f ? 1.0f / len2 : 0.f;
        const int lox = cell_coord(fminf(ax, ax + ex) - pad, ox, inv, nx), hix = cell_coord(fmaxf(ax, ax + ex) + pad, ox, inv, nx);
        const int loy = cell_coord(fminf(ay, ay + ey) - pad, oy, inv, ny), hiy = cell_coord(fmaxf(ay, ay + ey) + pad, oy, inv, ny);
        const int loz = cell_coord(fminf(az, az + ez) - pad, oz, inv, nz), hiz = cell_coord(fmaxf(az, az + ez) + pad, oz, inv, nz);
        const int wy = hiy - loy + 1;
        const int nrows = (hiz - loz + 1) * wy;
        ++qs.n_lock;
        __syncwarp();
        for (int w = lane; w < nwords; w += 32) ibm[w] = 0u;
        __syncwarp();
        // ---- 1. gather: bit[index] for every particle within reach of the segment
        for (int row0 = 0; row0 < nrows; row0 += 32) {
            const int r = row0 + lane;
            int beg = 0, end = 0;
            if (r < nrows) {
                // Row of cells (cy, cz): only the part of the segment whose y and z come within reach of the row's
                // box can have neighbours there; that parameter interval also bounds the x cells worth reading.
                // (A particle within `pad` of the segment at parameter t has y(t), z(t) within pad of its own y, z.)
                // Grid boundary rows / cells also hold whatever was clamped into them: unbounded on that side.
                const int cy = loy + r % wy, cz = loz + r / wy;
                const float big = 3.0e30f;
                const float y0 = cy == 0 ? -big : oy + (float)cy * h->cell - pad, y1 = cy == ny - 1 ? big : oy + (float)(cy + 1) * h->cell + pad;
                const float z0 = cz == 0 ? -big : oz + (float)cz * h->cell - pad, z1 = cz == nz - 1 ? big : oz + (float)(cz + 1) * h->cell + pad;
                float t0 = 0.f, t1 = 1.f;
                bool any = true;
                if (fabsf(ey) > 1e-12f) {
                    const float ta = (y0 - ay) / ey, tb = (y1 - ay) / ey;
                    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                } else any = any && ay >= y0 && ay <= y1;
                if (fabsf(ez) > 1e-12f) {
                    const float ta = (z0 - az) / ez, tb = (z1 - az) / ez;
                    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                } else any = any && az >= z0 && az <= z1;
                // the intervals were computed with rounding: widen a little, still inside [0, 1]
                t0 = fmaxf(t0 - 1e-3f, 0.f); t1 = fminf(t1 + 1e-3f, 1.f);
                if (any && t0 <= t1) {
                    const float xa = ax + t0 * ex, xb = ax + t1 * ex;
                    int rlo = cell_coord(fminf(xa, xb) - pad, ox, inv, nx), rhi = cell_coord(fmaxf(xa, xb) + pad, ox, inv, nx);
                    rlo = max(rlo, lox); rhi = min(rhi, hix);
                    if (rlo <= rhi) {
                        const int rowbase = (cz * ny + cy) * nx;
                        beg = __ldg(p.g.cell_start + rowbase + rlo);
                        end = __ldg(p.g.cell_start + rowbase + rhi + 1);
                    }
                }
            }
            const int nr = min(32, nrows - row0);
            for (int t = 0; t < nr; ++t) {
                const int rb = __shfl_sync(NF_FULL, beg, t), re = __shfl_sync(NF_FULL, end, t);
                for (int i = rb + lane; i < re; i += 128) {          // 4 independent 16-byte loads in flight
                    float4 c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = __ldg(p.g.sorted + min(i + 32 * u, re - 1));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float vx = c[u].x - ax, vy = c[u].y - ay, vz = c[u].z - az;
                        const float tt = fminf(fmaxf((vx * ex + vy * ey + vz * ez) * inv_len2, 0.f), 1.f);
                        const float wx = vx - tt * ex, wy_ = vy - tt * ey, wz = vz - tt * ez;
                        if (i + 32 * u < re && wx * wx + wy_ * wy_ + wz * wz < pad2) {
                            const int idx = __float_as_int(c[u].w);
                            atomicOr(&ibm[idx >> 5], 1u << (idx & 31));
                        }
                    }
                }
                ++qs.it_lock;
            }
        }
        }
        __syncwarp();
        // ---- 2 + 3. walk the bitmap in index order; sweep the unfinished samples over blocks of <= 128
        //      candidates (4 per lane, positions in registers): per sample one broadcast, four distance tests
        unsigned pend = sub;
        auto sweep = [&](int head, int navail) {
            float4 c[SCS_BLOCK / 32];
            int id[SCS_BLOCK / 32];
#pragma unroll
            for (int u = 0; u < SCS_BLOCK / 32; ++u) {
                const int i = 32 * u + lane;
                id[u] = (int)ring[head + i];                 // head + i < SCS_RING always
                c[u] = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.f);
                if (i < navail) c[u] = __ldg(p.g.orig4 + id[u]);
            }
            for (unsigned m = pend; m;) {
                const int s = __ffs(m) - 1;
                m &= m - 1;
                const float sx = __shfl_sync(NF_FULL, qx, s), sy = __shfl_sync(NF_FULL, qy, s), sz = __shfl_sync(NF_FULL, qz, s);
                int n = __shfl_sync(NF_FULL, cnt, s);
                bool hit[SCS_BLOCK / 32];
#pragma unroll
                for (int u = 0; u < SCS_BLOCK / 32; ++u) hit[u] = dist2_exact(sx, sy, sz, c[u].x, c[u].y, c[u].z) < r2;
#pragma unroll
                for (int u = 0; u < SCS_BLOCK / 32; ++u) {
                    if (u == 2 && n >= K) break;
                    const unsigned hm = __ballot_sync(NF_FULL, hit[u]);
                    const int slot = n + __popc(hm & lt);
                    if (hit[u] && slot < K) sel[s * KP + slot] = id[u];
                    n += __popc(hm);
                }
                ++qs.it_rows;
                n = min(n, K);
                if (lane == s) cnt = n;
                if (n >= K) pend &= ~(1u << s);
            }
        };
        // The sorted candidate list is kept in `ring` from candidate 0 on and survives to the next sub-group that
        // uses the same gather (lst_len candidates, bitmap walked up to word lst_word): a cluster of importance
        // samples spread over several 32-sample steps walks the bitmap once.  When the list would outgrow the ring
        // its already-swept prefix is dropped (lst_word < 0 from then on: the next sub-group starts over).
        if (fresh || lst_word < 0) { lst_len = 0; lst_word = 0; }
        int pos = 0;                      // sweep cursor of this sub-group
        int wnext = lst_word;
        bool dropped = false;
        while (pend) {
            const int avail = lst_len - pos;
            if (avail >= SCS_BLOCK || (avail > 0 && wnext >= nwords)) {
                const int nb = min(avail, SCS_BLOCK);
                sweep(pos, nb);
                pos += nb;
                continue;
            }
            if (wnext >= nwords) break;
            unsigned w = ibm[wnext + lane];
            const int c = __popc(w);
            int incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(NF_FULL, incl, off);
                if (lane >= off) incl += t;
            }
            const int tot = __shfl_sync(NF_FULL, incl, 31);
            const int base = (wnext + lane) << 5;
            wnext += 32;
            if (!tot) continue;
            if (lst_len + tot > SCS_RING - SCS_BLOCK) {
                // no room: keep only the < SCS_BLOCK candidates not swept yet (read all, then write: ranges may overlap)
                unsigned short v[SCS_BLOCK / 32];
#pragma unroll
                for (int u = 0; u < SCS_BLOCK / 32; ++u) v[u] = ring[pos + 32 * u + lane];
                __syncwarp();
#pragma unroll
                for (int u = 0; u < SCS_BLOCK / 32; ++u)
                    if (32 * u + lane < avail) ring[32 * u + lane] = v[u];
                lst_len = avail;
                pos = 0;
                dropped = true;
                __syncwarp();
            }
            int wpos = lst_len + incl - c;
            while (w) {
                ring[wpos++] = (unsigned short)(base + __ffs(w) - 1);
                w &= w - 1;
            }
            lst_len += tot;
            __syncwarp();
        }
        lst_word = dropped ? -1 : wnext;
        __syncwarp();
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------------
// Neighbour search + local geometry for all samples of one ray: sample s lives in lane s%32 of 32-sample
// step s/32.  The search (one of the two flavours above) leaves each lane's <= K neighbour indices, ascending,
// in sel[lane*sel_stride(K) + k] (odd stride: the k-th hits of 32 samples and the consecutive hits of one sample
// both fall in distinct banks); local geometry is then one per-lane pass over them in that order -- the
// reference's summation order.  smem per warp: sel[32*sel_stride(K)] | scratch[search_smem_bytes(FL, P)].
// FL: 0 = index-order stream, 1 = sorted-candidate sweep (separate kernel instances: the sweep alone fits 80
// registers, i.e. three 8-warp blocks per SM).
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t search_smem_bytes(int flavour, int P) {
    return flavour == 1 ? scs_smem_bytes(P) : BM_WORDS * sizeof(unsigned) + HITBUF * sizeof(int);
}

// The slot loop is deliberately NOT unrolled and keeps no per-slot register arrays: an unrolled copy per slot
// made the kernel ~190 KB of SASS and 70 % of its stall samples instruction-cache misses.
template <int FL>
__device__ __forceinline__ void ray_query_group(const StageArgs& p, int lane, const float (&o)[3],
                                             const float (&d)[3], const float* zs /*smem: S sorted depths*/, int S,
                                             float* rec, int* rowid, int* row_counter, int* active_counter, int cap,
                                             int ray, long long* num_nn, unsigned* act, int act_stride,
                                             QueryStats& qs, int* sel, unsigned* scratch,
                                             const short* src /*smem or null: coarse index of each merged sample*/,
                                             int* nbr /*null or (rows, K): neighbour list of every record row*/) {
    // Fine pass: a merged sample that IS one of the coarse samples (same depth, same position, bit for bit) was
    // already searched by stage Q0 -- its neighbour count comes from cnt0 and its geometry record is copied from
    // rec0 instead of being searched and built again (the coarse samples are the ones spread through the whole
    // fluid; the importance samples cluster).
    const bool coarse_pass = src == nullptr;
    const int act_stride0 = (p.S0 + 31) >> 5 <= 2 ? 2 : 4;      // NS0 of the coarse pass (pick_ns)
    const int K = p.K;
    const int KP = sel_stride(K);
    const float radius = p.radius;
    float gz0 = 1.f, gz1 = 0.f;      // sweep: depth interval whose candidates the index bitmap currently holds
    int lst_len = 0, lst_word = -1;  // sweep: sorted candidate list kept across sub-groups of one gather
    const unsigned lt = (1u << lane) - 1u;
    int n_active = 0;
    const int sample_base = ray * S;
    const int nslots = (S + 31) >> 5;
    const float rox = p.ro_dev ? __ldg(p.ro_dev) : p.ro[0], roy = p.ro_dev ? __ldg(p.ro_dev + 1) : p.ro[1],
                roz = p.ro_dev ? __ldg(p.ro_dev + 2) : p.ro[2];
#pragma unroll 1
    for (int slot = 0; slot < nslots; ++slot) {
        const int s = slot * 32 + lane;
        const bool in = s < S;
        const float zv = zs[min(s, S - 1)];
        const float qx = __fadd_rn(o[0], __fmul_rn(d[0], zv));
        const float qy = __fadd_rn(o[1], __fmul_rn(d[1], zv));
        const float qz = __fadd_rn(o[2], __fmul_rn(d[2], zv));
        const int from = (!coarse_pass && in) ? (int)src[s] : -1;
        const bool reused = from >= 0;
        const int pre = reused ? (int)p.cnt0[(size_t)ray * p.S0 + from] : 0;
        const int occ = (in && !reused) ? grid_occupancy(p.g, qx, qy, qz, radius) : 0;
        const bool search = occ > 0;
        // use_mask=False: every sample is evaluated, even empty ones
        const bool want = p.use_mask ? (reused ? pre == p.K : search) : in;
        if (!__any_sync(NF_FULL, want || search)) {
            if (num_nn && in) num_nn[(size_t)sample_base + s] = pre;
            if (coarse_pass && in) p.cnt0[(size_t)ray * p.S0 + s] = 0;
            if (lane == 0) act[(size_t)ray * act_stride + slot] = 0u;
            continue;
        }
        int cnt;
        if constexpr (FL == 1) cnt = search_scs(p, lane, o, d, zv, qx, qy, qz, search, gz0, gz1, lst_len, lst_word, zs, S, slot * 32, qs, scratch, sel);
        else cnt = search_stream(p, lane, qx, qy, qz, search, occ, qs, scratch, reinterpret_cast<int*>(scratch + BM_WORDS), sel);
        // ---- per-lane local geometry over the selected neighbours (ascending index, like the reference);
        //      one pass: var = (sum v^2 - 2 mean sum v + n mean^2) / n  ==  sum (v - mean)^2 / n
        float wsum = 0.f, wx = 0.f, wy = 0.f, wz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
        int nvalid = 0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            if (k < cnt) {
                const float4 pp = __ldg(p.g.orig4 + sel[lane * KP + k]);
                const float ex = pp.x - qx, ey = pp.y - qy, ez = pp.z - qz;
                const float t = sqrtf(ex * ex + ey * ey + ez * ez) / radius;
                const float w = fmaxf(1.0f - t * t * t, 0.f);
                wsum += w; wx += w * pp.x; wy += w * pp.y; wz += w * pp.z;
                if (dist2_exact(qx, qy, qz, pp.x, pp.y, pp.z) != 0.f) {
                    vx += ex; vy += ey; vz += ez;
                    ux += ex * ex; uy += ey * ey; uz += ez * ez;
                    ++nvalid;
                }
            }
        }
        if (cnt < K) {   // padded slots are zeros = a phantom particle at the origin (models/renderer.py:97-98)
            const float t = sqrtf(qx * qx + qy * qy + qz * qz) / radius;
            wsum += (float)(K - cnt) * fmaxf(1.0f - t * t * t, 0.f);
        }
        const float nvf = (float)nvalid + 1e-12f;
        const float mx = vx / nvf, my = vy / nvf, mz = vz / nvf;
        const float ax = fmaxf(ux - 2.0f * mx * vx + (float)nvalid * mx * mx, 0.f);
        const float ay = fmaxf(uy - 2.0f * my * vy + (float)nvalid * my * my, 0.f);
        const float az = fmaxf(uz - 2.0f * mz * vz + (float)nvalid * mz * mz, 0.f);
        if (reused) nvalid = pre;
        const bool full = in && (nvalid == K);
        const unsigned fb = __ballot_sync(NF_FULL, full);
        if (num_nn && in) num_nn[(size_t)sample_base + s] = nvalid;
        if (coarse_pass && in) p.cnt0[(size_t)ray * p.S0 + s] = (unsigned char)nvalid;
        if (lane == 0) act[(size_t)ray * act_stride + slot] = fb;
        n_active += __popc(fb);
        const bool eval = want && (p.use_mask ? full : true);
        const unsigned me = __ballot_sync(NF_FULL, eval);
        int base = 0;
        if (me) {
            if (lane == 0) base = atomicAdd(row_counter, __popc(me));
            base = __shfl_sync(NF_FULL, base, 0);
            const int row = base + __popc(me & lt);
            if (eval && row < cap) {
                float4* dst = reinterpret_cast<float4*>(rec + (size_t)row * 16);
                if (reused) {
                    // row of this sample in the coarse list: base of its 32-sample step + rank among the evaluated
                    const int slot0 = from >> 5, nin = min(32, p.S0 - slot0 * 32);
                    const unsigned ev0 = p.use_mask ? p.act0[(size_t)ray * act_stride0 + slot0]
                                                    : (nin == 32 ? NF_FULL : ((1u << nin) - 1u));
                    const int row0 = p.base0[(size_t)ray * act_stride0 + slot0] + __popc(ev0 & ((1u << (from & 31)) - 1u));
                    const float4* sp = reinterpret_cast<const float4*>(p.rec0 + (size_t)row0 * 16);
                    dst[0] = sp[0]; dst[1] = sp[1]; dst[2] = sp[2]; dst[3] = sp[3];
                    if (nbr)
                        for (int k = 0; k < K; ++k) nbr[(size_t)row * K + k] = p.nbr0[(size_t)row0 * K + k];
                } else {
                    const float den = wsum + 1e-12f;
                    float sx = wx / den, sy = wy / den, sz = wz / den;
                    if (p.include_ray) {          // models/renderer.py:100-109
                        const float al = (p.same_smooth || nvalid > 20) ? 0.9f : 0.1f;
                        sx = qx * (1.0f - al) + sx * al; sy = qy * (1.0f - al) + sy * al; sz = qz * (1.0f - al) + sz * al;
                    }
                    const float tx = sx - rox, ty = sy - roy, tz = sz - roz;
                    const float tn = sqrtf(tx * tx + ty * ty + tz * tz);
                    dst[0] = make_float4(qx, qy, qz, wsum);
                    dst[1] = make_float4(sx, sy, sz, ax / nvf);
                    dst[2] = make_float4(ay / nvf, az / nvf, d[0], d[1]);
                    dst[3] = make_float4(d[2], tx / tn, ty / tn, tz / tn);
                    if (nbr)
                        for (int k = 0; k < K; ++k) nbr[(size_t)row * K + k] = k < cnt ? sel[lane * KP + k] : -1;
                }
                rowid[row] = sample_base + s;
            }
        }
        if (coarse_pass && lane == 0) p.base0[(size_t)ray * act_stride + slot] = base;
    }
    // mask words of unused 32-sample steps (S not a multiple that fills act_stride) must read as "none valid"
    if (lane >= nslots && lane < act_stride) act[(size_t)ray * act_stride + lane] = 0u;
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

// ------------------------------------------------------------------------------------------------
// Rays that never come within reach of the particle set's bounding box (most of an image: the fluid covers a
// fraction of it) need no search, no pdf and no merge: every sample has zero neighbours, so with use_mask every
// output is known -- white / zero, counts zero -- and identical to what the general path computes (all weights
// are exactly 0).  Stage Q0 decides it once per ray (slab test of the sampled depth range against the box grown
// by the search reach, with margin) and the later stages follow the flag.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ray_misses_points(const StageArgs& p, const float (&o)[3], const float (&d)[3], float z_lo,
                                                  float z_hi) {
    const GridHeader* h = p.g.hdr;
    if (h->n == 0) return true;
    const float pad = p.radius * 1.002f + 1e-3f;
    float t0 = z_lo - 1e-3f, t1 = z_hi + 1e-3f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = h->bmin[a] - pad, hi = h->bmax[a] + pad;
        if (fabsf(d[a]) > 1e-9f) {
            const float ta = (lo - o[a]) / d[a], tb = (hi - o[a]) / d[a];
            t0 = fmaxf(t0, fminf(ta, tb) - 1e-3f);
            t1 = fminf(t1, fmaxf(ta, tb) + 1e-3f);
        } else if (o[a] < lo || o[a] > hi) {
            return true;
        }
    }
    return t0 > t1;
}

// ------------------------------------------------------------------------------------------------
// stage Q0
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t q0_smem_per_warp(int fl, int K, int P) {
    return sel_bytes(K) + search_smem_bytes(fl, P);
}

template <int NS0, int FL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, FL == 1 ? 3 : 2) k_stage_q0(const StageArgs p) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ float sm_zw[WARPS_PER_BLOCK][NS0 * 32];     // this warp's ray's coarse depths (a shared table, or its perturbed row)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float* sm_z = sm_zw[wib];
    for (int s = lane; s < p.S0; s += 32) sm_z[s] = __ldg(p.z_coarse + s);
    __syncwarp();
    int* sel = reinterpret_cast<int*>(dyn_smem + wib * q0_smem_per_warp(FL, p.K, p.n_points));
    unsigned* scratch = reinterpret_cast<unsigned*>(sel + 32 * sel_stride(p.K));
    QueryStats qs;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        float o[3], d[3];
        load_ray(p.rays, ray, o, d);
        if (p.z_stride) {
            __syncwarp();
            for (int s = lane; s < p.S0; s += 32) sm_z[s] = __ldg(p.z_coarse + (size_t)ray * p.z_stride + s);
            __syncwarp();
        }
        // with sigma noise an empty ray still composites relu(noise): no shortcut for rays that miss the particles
        const bool miss = p.use_mask && !p.noise0 && !p.noise1 && ray_misses_points(p, o, d, sm_z[0], sm_z[p.S0 - 1]);
        if (lane == 0) p.miss[ray] = miss ? 1 : 0;
        if (miss) {
            for (int s = lane; s < p.S0; s += 32) {
                if (p.num_nn0) p.num_nn0[(size_t)ray * p.S0 + s] = 0;
                p.cnt0[(size_t)ray * p.S0 + s] = 0;
            }
            if (lane < NS0) { p.act0[(size_t)ray * NS0 + lane] = 0u; p.base0[(size_t)ray * NS0 + lane] = 0; }
            continue;
        }
        ray_query_group<FL>(p, lane, o, d, sm_z, p.S0, p.rec0, p.rowid0, p.counters + 0, p.counters + 2, p.cap0, ray,
                        p.num_nn0, p.act0, NS0, qs, sel, scratch, nullptr, p.nbr0);
    }
    if (lane == 0) {
        atomicAdd(p.counters + 8, qs.n_lock);
        atomicAdd(p.counters + 9, qs.n_rows);
        atomicAdd(p.counters + 10, qs.it_lock >> 6);
        atomicAdd(p.counters + 11, qs.it_rows >> 6);
    }
}

// ------------------------------------------------------------------------------------------------
// stage MID
// ------------------------------------------------------------------------------------------------
// per warp: z1s[NS1*32] f32 | src1[NS1*32] i16 | sel | union { pdf scratch: z0s, ws, bins, cdf [NS0*32] f32, smp[NS1*32] f32 ;
//                                                             search scratch }   (the pdf arrays are dead once the merged
// depths exist, the search scratch is dead between rays)
__host__ __device__ inline size_t mid_smem_per_warp(int fl, int ns0, int ns1, int K, int P) {
    const size_t pdf = (size_t)(4 * ns0 * 32 + ns1 * 32) * sizeof(float);
    const size_t srch = search_smem_bytes(fl, P);
    return (size_t)ns1 * 32 * (sizeof(float) + sizeof(short)) + sel_bytes(K) + (pdf > srch ? pdf : srch);
}

template <int NS0, int NS1, int FL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, FL == 1 ? 3 : 2) k_stage_mid(const StageArgs p) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S0 = p.S0, S1 = p.S1, NI = p.n_imp;
    float* z1s = reinterpret_cast<float*>(dyn_smem + wib * mid_smem_per_warp(FL, NS0, NS1, p.K, p.n_points));   // merged depths
    short* src1 = reinterpret_cast<short*>(z1s + NS1 * 32);      // merged sample -> coarse index or -1
    int* sel = reinterpret_cast<int*>(src1 + NS1 * 32);
    unsigned* scratch = reinterpret_cast<unsigned*>(sel + 32 * sel_stride(p.K));
    float* z0s = reinterpret_cast<float*>(scratch);   // coarse depths            (pdf scratch overlays search scratch)
    float* ws = z0s + NS0 * 32;          // coarse weights
    float* bins = ws + NS0 * 32;
    float* cdf = bins + NS0 * 32;
    float* smp = cdf + NS0 * 32;         // importance samples
    QueryStats qs;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < p.n_rays; ray += nwarps) {
        if (p.miss[ray]) {          // nothing within reach anywhere along the ray: every output is known
            if (lane == 0) {
                const float bg = p.white_bg ? 1.f : 0.f;
                if (p.rgb0) { p.rgb0[3 * (size_t)ray] = bg; p.rgb0[3 * (size_t)ray + 1] = bg; p.rgb0[3 * (size_t)ray + 2] = bg; }
                if (p.depth0) p.depth0[ray] = 0.f;
                if (p.opac0) p.opac0[ray] = 0.f;
                if (p.mask0) p.mask0[ray] = 0.f;
            }
            if (p.num_nn1)
                for (int s = lane; s < S1; s += 32) p.num_nn1[(size_t)ray * S1 + s] = 0;
            if (lane < NS1) p.act1[(size_t)ray * NS1 + lane] = 0u;
            continue;
        }
        float o[3], d[3], z0[NS0], w0[NS0];
        float4 c0[NS0];
        load_ray(p.rays, ray, o, d);
        const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        int nfull = 0;
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int s = slot * 32 + lane;
            z0[slot] = __ldg(p.z_coarse + (size_t)ray * p.z_stride + min(s, S0 - 1));
            const unsigned bits = p.act0[(size_t)ray * NS0 + slot];
            nfull += __popc(bits);
            const bool ev = s < S0 && (p.use_mask ? ((bits >> lane) & 1u) : true);
            c0[slot] = ev ? p.out0[(size_t)ray * S0 + s] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.noise0 && s < S0) c0[slot].w += p.noise0[(size_t)ray * S0 + s];      // models/renderer.py:192-196
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
                const float u = __ldg(p.u_imp + (size_t)ray * p.u_stride + j);
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
            if (p.u_stride) {          // random inverse-CDF arguments: the draws come in no order; sorted below
                if (j < NI) smp[j] = sv;
                continue;
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
        if (p.u_stride) {
            // torch.sort(cat[z, z_samples]) (utils/ray_utils.py:225): bitonic sort of the draws in shared memory, then the
            // same rank merge as the deterministic path
            int n2 = 1;
            while (n2 < NI) n2 <<= 1;
            for (int i = NI + lane; i < n2; i += 32) smp[i] = 3.0e38f;
            __syncwarp();
            for (int k = 2; k <= n2; k <<= 1)
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    for (int i = lane; i < n2; i += 32) {
                        const int ixj = i ^ jj;
                        if (ixj > i) {
                            const float va = smp[i], vb = smp[ixj];
                            if ((va > vb) == ((i & k) == 0)) { smp[i] = vb; smp[ixj] = va; }
                        }
                    }
                    __syncwarp();
                }
        }
        // rank merge of the two sorted lists (ties: coarse depths first)
#pragma unroll
        for (int slot = 0; slot < NS0; ++slot) {
            const int i = slot * 32 + lane;
            if (i < S0) {
                const float v = z0s[i];
                int lo = 0, hi = NI;            // # samples < v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (smp[mid] < v) lo = mid + 1; else hi = mid; }
                z1s[i + lo] = v;
                src1[i + lo] = (short)i;
            }
        }
        for (int j0 = 0; j0 < NI; j0 += 32) {
            const int j = j0 + lane;
            if (j < NI) {
                const float v = smp[j];
                int lo = 0, hi = S0;            // # coarse depths <= v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (z0s[mid] <= v) lo = mid + 1; else hi = mid; }
                z1s[j + lo] = v;
                src1[j + lo] = (short)-1;
            }
        }
        __syncwarp();
        for (int s = lane; s < S1; s += 32) p.z1[(size_t)ray * S1 + s] = z1s[s];
        ray_query_group<FL>(p, lane, o, d, z1s, S1, p.rec1, p.rowid1, p.counters + 1, p.counters + 3, p.cap1, ray,
                        p.num_nn1, p.act1, NS1, qs, sel, scratch, src1, p.nbr1);
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(p.counters + 4, qs.n_lock);
        atomicAdd(p.counters + 5, qs.n_rows);
        atomicAdd(p.counters + 6, qs.it_lock >> 6);
        atomicAdd(p.counters + 7, qs.it_rows >> 6);
    }
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
        if (p.miss[ray]) {          // see ray_misses_points: merged depths were never written for this ray
            if (lane == 0) {
                const float bg = p.white_bg ? 1.f : 0.f;
                if (o_rgb) { o_rgb[3 * (size_t)ray] = bg; o_rgb[3 * (size_t)ray + 1] = bg; o_rgb[3 * (size_t)ray + 2] = bg; }
                if (o_depth) o_depth[ray] = 0.f;
                if (o_opac) o_opac[ray] = 0.f;
                if (o_mask) o_mask[ray] = 0.f;
            }
            continue;
        }
        float o[3], d[3], z[NS], w[NS];
        float4 c[NS];
        load_ray(p.rays, ray, o, d);
        const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        int nfull = 0;
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {
            const int s = slot * 32 + lane;
            z[slot] = FIRST ? __ldg(p.z_coarse + (size_t)ray * p.z_stride + min(s, S - 1)) : p.z1[(size_t)ray * S + min(s, S - 1)];
            const unsigned bits = act[(size_t)ray * NS + slot];
            nfull += __popc(bits);
            const bool ev = s < S && (p.use_mask ? ((bits >> lane) & 1u) : true);
            c[slot] = ev ? out[(size_t)ray * S + s] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float* noise = FIRST ? p.noise0 : p.noise1;
            if (noise && s < S) c[slot].w += noise[(size_t)ray * S + s];
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

#ifdef NF_TUNING
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif

struct WsLayout {
    size_t counters, pe_scratch, act0, act1, base0, cnt0, miss, z1, rec0, rowid0, out0, rec1, rowid1, out1, nbr0, nbr1, total;
    int cap0, cap1, ns0, ns1;
};

static int pick_ns(int s, const int* opts, int n) {
    for (int i = 0; i < n; ++i)
        if (s <= opts[i] * 32) return opts[i];
    return -1;
}

static WsLayout ws_layout(int R, int S0, int NI, int K = 0 /* > 0: room for the saved neighbour lists */) {
    WsLayout L;
    const int S1 = S0 + NI;
    static const int o0[] = {2, 4};
    static const int o1[] = {4, 6, 8};
    L.ns0 = pick_ns(S0, o0, 2);
    L.ns1 = NI > 0 ? pick_ns(S1, o1, 3) : 4;
    L.cap0 = (int)((size_t)R * S0);
    L.cap1 = (int)((size_t)R * S1);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    L.counters = take(64);
    L.pe_scratch = take(mlp::PE_SCRATCH_BYTES);      // encoded-feature scratch of the MLP kernel (nf_mlp.cuh), shared by both networks
    L.act0 = take(sizeof(unsigned) * (size_t)R * 4);
    L.act1 = take(sizeof(unsigned) * (size_t)R * 8);
    L.base0 = take(sizeof(int) * (size_t)R * 4);
    L.cnt0 = take((size_t)R * S0);
    L.miss = take((size_t)R);
    L.z1 = take(sizeof(float) * (size_t)R * S1);
    L.rec0 = take(sizeof(float) * 16 * (size_t)L.cap0);
    L.rowid0 = take(sizeof(int) * (size_t)L.cap0);
    L.out0 = take(sizeof(float4) * (size_t)R * S0);
    L.rec1 = take(sizeof(float) * 16 * (size_t)L.cap1);
    L.rowid1 = take(sizeof(int) * (size_t)L.cap1);
    L.out1 = take(sizeof(float4) * (size_t)R * S1);
    L.nbr0 = L.nbr1 = 0;
    if (K > 0) {
        L.nbr0 = take(sizeof(int) * (size_t)L.cap0 * K);
        L.nbr1 = take(sizeof(int) * (size_t)L.cap1 * K);
    }
    L.total = o;
    return L;
}

template <int NS0, int NS1, int FL>
static int launch_mid3(int grid, const StageArgs& p, cudaStream_t st) {
    const size_t smem = WARPS_PER_BLOCK * mid_smem_per_warp(FL, NS0, NS1, p.K, p.n_points);
    NF_CUDA_OK(cudaFuncSetAttribute(k_stage_mid<NS0, NS1, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_stage_mid<NS0, NS1, FL><<<grid, WARPS_PER_BLOCK * 32, smem, st>>>(p);
    NF_LAUNCH_OK();
    return NF_OK;
}

template <int NS0, int NS1>
static int launch_mid2(int grid, const StageArgs& p, cudaStream_t st) {
    return p.search_mode == 1 ? launch_mid3<NS0, NS1, 1>(grid, p, st) : launch_mid3<NS0, NS1, 0>(grid, p, st);
}

template <int NS0>
static int launch_mid(int ns1, int grid, const StageArgs& p, cudaStream_t st) {
    switch (ns1) {
        case 4: return launch_mid2<NS0, 4>(grid, p, st);
        case 6: return launch_mid2<NS0, 6>(grid, p, st);
        case 8: return launch_mid2<NS0, 8>(grid, p, st);
        default: set_error("unsupported fine sample count"); return NF_E_UNSUPPORTED;
    }
}

template <int NS0, int FL>
static int launch_q0(int grid, const StageArgs& p, cudaStream_t st) {
    const size_t smem = WARPS_PER_BLOCK * q0_smem_per_warp(FL, p.K, p.n_points);
    NF_CUDA_OK(cudaFuncSetAttribute(k_stage_q0<NS0, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_stage_q0<NS0, FL><<<grid, WARPS_PER_BLOCK * 32, smem, st>>>(p);
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

extern "C" size_t nf_render_workspace_bytes_ex(int n_rays, int n_coarse, int n_importance, int K, int flags) {
    if (n_rays <= 0 || n_coarse <= 0 || n_importance < 0) return 0;
    return ws_layout(n_rays, n_coarse, n_importance, (flags & NF_RENDER_SAVE_NEIGHBORS) ? K : 0).total;
}

extern "C" int nf_render_workspace_view(int n_rays, int n_coarse, int n_importance, int K, int flags, nf_render_ws_view* v) {
    NF_REQUIRE(v && n_rays > 0 && n_coarse > 0 && n_importance >= 0, NF_E_INVALID, "nf_render_workspace_view: bad arguments");
    const WsLayout L = ws_layout(n_rays, n_coarse, n_importance, (flags & NF_RENDER_SAVE_NEIGHBORS) ? K : 0);
    v->counters = L.counters; v->act0 = L.act0; v->act1 = L.act1; v->z1 = L.z1;
    v->rec0 = L.rec0; v->rowid0 = L.rowid0; v->out0 = L.out0;
    v->rec1 = L.rec1; v->rowid1 = L.rowid1; v->out1 = L.out1;
    v->nbr0 = L.nbr0; v->nbr1 = L.nbr1; v->miss = L.miss;
    v->act_stride0 = L.ns0; v->act_stride1 = L.ns1; v->cap0 = L.cap0; v->cap1 = L.cap1;
    v->total = L.total;
    return NF_OK;
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
    const bool save_nbr = (a->flags & NF_RENDER_SAVE_NEIGHBORS) != 0;
    const WsLayout L = ws_layout(a->n_rays, a->n_coarse, NI, save_nbr ? a->K : 0);
    NF_REQUIRE(a->workspace_bytes >= L.total, NF_E_WORKSPACE, "nf_render_forward: workspace %zu < %zu",
               a->workspace_bytes, L.total);
    char* b = (char*)a->workspace;
    StageArgs p;
    p.g = grid_view(a->grid_ws, a->n_particles);
    p.particles = a->particles;
    p.n_points = a->n_particles;
    p.rays = a->rays;
    p.n_rays = a->n_rays;
    p.ro[0] = a->ro[0]; p.ro[1] = a->ro[1]; p.ro[2] = a->ro[2];
    p.ro_dev = a->ro_dev;
    p.radius = a->radius;
    p.K = a->K;
    p.use_mask = a->use_mask; p.white_bg = a->white_background; p.mode = a->mode;
    p.include_ray = a->include_ray; p.same_smooth = a->same_smooth_factor;
    {
        // search tuning: compile-time constants.  A build with -DNF_TUNING (tests/gpu_tune.py) reads NF_* environment
        // overrides instead; the release library never looks at the environment.
        struct Tune { int solo_max_occ, peel_lanes, peel_from, sub_look; float sub_span_r; };
#ifdef NF_TUNING
        auto read = [] {
            Tune t;
            t.solo_max_occ = env_int("NF_SOLO_MAX_OCC", 600);
            t.peel_lanes = env_int("NF_PEEL_LANES", 4);
            t.peel_from = env_int("NF_PEEL_FROM", 1536);
            t.sub_look = env_int("NF_SUB_LOOK", 96);
            const char* span = getenv("NF_SUB_SPAN");
            t.sub_span_r = span ? -(float)atof(span) : 3.5f;      // negative: absolute length, positive: multiples of r
            return t;
        };
        const Tune t = read();
#else
        const Tune t = {600, 4, 1536, 96, 3.5f};
#endif
        p.solo_max_occ = t.solo_max_occ; p.peel_lanes = t.peel_lanes; p.peel_from = t.peel_from; p.sub_look = t.sub_look;
        p.sub_span = t.sub_span_r < 0.f ? -t.sub_span_r : t.sub_span_r * a->radius;
        NF_REQUIRE(a->search >= NF_SEARCH_AUTO && a->search <= NF_SEARCH_SWEEP, NF_E_INVALID, "nf_render_forward: search %d", a->search);
        NF_REQUIRE(a->search != NF_SEARCH_SWEEP || a->n_particles <= SCS_MAX_POINTS, NF_E_UNSUPPORTED,
                   "nf_render_forward: NF_SEARCH_SWEEP needs n_particles <= %d", SCS_MAX_POINTS);
        p.search_mode = (a->search == NF_SEARCH_STREAM || a->n_particles > SCS_MAX_POINTS) ? 0 : 1;
    }
    p.z_coarse = a->z_coarse; p.u_imp = a->u_importance;
    p.z_stride = a->z_stride; p.u_stride = a->u_stride;
    NF_REQUIRE((a->z_stride == 0 || a->z_stride >= a->n_coarse) && (a->u_stride == 0 || a->u_stride >= NI), NF_E_INVALID,
               "nf_render_forward: z_stride / u_stride must be 0 or at least a row");
    p.noise0 = a->noise0; p.noise1 = fine ? a->noise1 : nullptr;
    p.S0 = a->n_coarse; p.n_imp = NI; p.S1 = a->n_coarse + NI;
    const bool want0 = a->mode != NF_RENDER_FINE;
    p.rgb0 = want0 ? a->rgb0 : nullptr; p.depth0 = want0 ? a->depth0 : nullptr;
    p.opac0 = want0 ? a->opacity0 : nullptr; p.mask0 = want0 ? a->mask0 : nullptr;
    p.num_nn0 = want0 ? (long long*)a->num_nn0 : nullptr;
    p.rgb1 = a->rgb1; p.depth1 = a->depth1; p.opac1 = a->opacity1; p.mask1 = a->mask1;
    p.num_nn1 = (long long*)a->num_nn1;
    p.counters = (int*)(b + L.counters);
    p.act0 = (unsigned*)(b + L.act0); p.act1 = (unsigned*)(b + L.act1);
    p.base0 = (int*)(b + L.base0); p.cnt0 = (unsigned char*)(b + L.cnt0); p.miss = (unsigned char*)(b + L.miss);
    p.z1 = (float*)(b + L.z1);
    p.rec0 = (float*)(b + L.rec0); p.rowid0 = (int*)(b + L.rowid0); p.out0 = (float4*)(b + L.out0); p.cap0 = L.cap0;
    p.rec1 = (float*)(b + L.rec1); p.rowid1 = (int*)(b + L.rowid1); p.out1 = (float4*)(b + L.out1); p.cap1 = L.cap1;
    p.nbr0 = save_nbr ? (int*)(b + L.nbr0) : nullptr;
    p.nbr1 = save_nbr ? (int*)(b + L.nbr1) : nullptr;

    NF_CUDA_OK(cudaMemsetAsync(p.counters, 0, 64, st));
    const int threads = WARPS_PER_BLOCK * 32;
    const int grid = min((a->n_rays + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, num_sms() * 8);

    StageTimer tm(st);
    // ---- stage Q0
    {
        const int rc0 = (L.ns0 == 2) ? (p.search_mode == 1 ? launch_q0<2, 1>(grid, p, st) : launch_q0<2, 0>(grid, p, st))
                                     : (p.search_mode == 1 ? launch_q0<4, 1>(grid, p, st) : launch_q0<4, 0>(grid, p, st));
        if (rc0 != NF_OK) return rc0;
    }
    tm.mark();
    // ---- coarse network
    mlp::KernelArgs m;
    m.pe_scratch = (uint8_t*)(b + L.pe_scratch);
    m.packed = (const uint8_t*)a->weights_coarse;
    m.records = p.rec0; m.rowid = p.rowid0; m.n_rows_dev = p.counters + 0; m.n_rows_host = 0; m.n_rows_cap = L.cap0;
    m.n_layers = (a->mode == NF_RENDER_FINE) ? 8 : 10;
#ifdef NF_TUNING
    m.desc_swap = env_int("NF_MLP_DESC_SWAP", 0);
#else
    m.desc_swap = 0;
#endif
    m.trace = nullptr;
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
    if (a->stats) NF_CUDA_OK(cudaMemcpyAsync(a->stats, p.counters, 64, cudaMemcpyDeviceToDevice, st));
    return NF_OK;
}
