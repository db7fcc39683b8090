/*
 * nf_oracle.c -- CPU oracle for the two NeuroFluid hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (neurofluid_b200/) never links, imports or calls it.
 *
 * It restates, in plain C, the arithmetic of the three third-party operators the reference
 * delegates its hot paths to (none of them is vendored under /root/reference):
 *
 *   nfo_ball_query        PyTorch3D v0.6.1  pytorch3d.ops.ball_query   (reference call site:
 *                         models/renderer.py:116-118)   first-K-by-index, squared distances,
 *                         strict "<", idx padded with -1, dists padded with 0.
 *   nfo_radius_search     open3d 0.15.2  FixedRadiusSearch(ignore_query_point=True)
 *                         (inside ContinuousConv.forward, models/transmodel.py:116,118,125)
 *   nfo_cconv_forward     open3d 0.15.2  ml.torch.layers.ContinuousConv forward with
 *                         kernel_size 4x4x4, interpolation='linear', align_corners=True,
 *                         coordinate_mapping='ball_to_cube_volume_preserving', normalize=False,
 *                         window = clamp((1 - d^2/r^2)^3, 0, 1)  (models/transmodel.py:73-98)
 *
 * PARITY STATUS: "parity unpinned" w.r.t. the upstream binaries -- the reference ships no tests,
 * golden vectors or fixtures and PyTorch3D / Open3D are not installable in
 * this image, so the semantics above are restated from the published algorithms.  What *is*
 * pinned: the reference's own Python (models/renderer.py, models/nerf.py, utils/ray_utils.py,
 * models/transmodel.py) is executed unmodified on top of these operators by
 * oracle/make_golden.py and its outputs are committed under tests/golden/.
 *
 * Build: see oracle/build_oracle.py (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off matters: the in-radius test must round exactly like the non-FMA x86 build
 * of the upstream CPU kernel, ((dx*dx + dy*dy) + dz*dz) < r*r, so that the CUDA path can be
 * compared bit-exactly on neighbour sets.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NFO_API __attribute__((visibility("default")))

NFO_API int nfo_version(void) { return 1; }

NFO_API int nfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * pytorch3d.ops.ball_query, one shared point cloud p2 for every query (the reference repeats
 * the particle tensor per ray, models/renderer.py:113, which is the same thing).
 * q: (nq,3)  p: (np,3)  idx: (nq,K) int64, -1 padded   dists: (nq,K) float, 0 padded (squared)
 * ------------------------------------------------------------------------------------------ */
NFO_API void nfo_ball_query(const float* q, int64_t nq, const float* p, int64_t np_, float radius,
                            int K, int64_t* idx, float* dists, int nthreads) {
    const float r2 = radius * radius;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
        const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        int64_t* oi = idx + i * K;
        float* od = dists + i * K;
        for (int k = 0; k < K; ++k) { oi[k] = -1; od[k] = 0.0f; }
        int count = 0;
        for (int64_t j = 0; j < np_ && count < K; ++j) {
            const float dx = qx - p[3 * j], dy = qy - p[3 * j + 1], dz = qz - p[3 * j + 2];
            const float d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 < r2) { oi[count] = j; od[count] = d2; ++count; }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Uniform-grid helper for the radius search (Open3D uses a spatial hash; any exact search gives
 * the same neighbour *set*, we emit neighbours in ascending index order for determinism).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    float ox, oy, oz, inv;
    int nx, ny, nz;
    int64_t* start; /* ncell+1 */
    int32_t* items; /* n */
} nfo_grid;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void grid_build(nfo_grid* g, const float* p, int64_t n, float cell) {
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = p[3 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    if (n == 0) { mn[0] = mn[1] = mn[2] = 0; mx[0] = mx[1] = mx[2] = 0; }
    g->ox = mn[0]; g->oy = mn[1]; g->oz = mn[2]; g->inv = 1.0f / cell;
    g->nx = clampi((int)((mx[0] - mn[0]) * g->inv) + 1, 1, 256);
    g->ny = clampi((int)((mx[1] - mn[1]) * g->inv) + 1, 1, 256);
    g->nz = clampi((int)((mx[2] - mn[2]) * g->inv) + 1, 1, 256);
    int64_t nc = (int64_t)g->nx * g->ny * g->nz;
    g->start = (int64_t*)calloc(nc + 1, sizeof(int64_t));
    g->items = (int32_t*)malloc((n > 0 ? n : 1) * sizeof(int32_t));
    int32_t* cellof = (int32_t*)malloc((n > 0 ? n : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) {
        int cx = clampi((int)floorf((p[3 * i] - g->ox) * g->inv), 0, g->nx - 1);
        int cy = clampi((int)floorf((p[3 * i + 1] - g->oy) * g->inv), 0, g->ny - 1);
        int cz = clampi((int)floorf((p[3 * i + 2] - g->oz) * g->inv), 0, g->nz - 1);
        int32_t c = (cz * g->ny + cy) * g->nx + cx;
        cellof[i] = c;
        g->start[c + 1]++;
    }
    for (int64_t c = 0; c < nc; ++c) g->start[c + 1] += g->start[c];
    int64_t* fill = (int64_t*)malloc(nc * sizeof(int64_t));
    memcpy(fill, g->start, nc * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) g->items[fill[cellof[i]]++] = (int32_t)i; /* ascending idx per cell */
    free(fill);
    free(cellof);
}

static void grid_free(nfo_grid* g) { free(g->start); free(g->items); }

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

/* neighbours of one query, ascending index; returns count (writes up to cap entries) */
static int grid_query(const nfo_grid* g, const float* p, const float* qp, float radius,
                      int ignore_same_pos, int32_t* out, float* out_d2, int cap) {
    const float r2 = radius * radius;
    const float pad = radius * 1.001f + 1e-6f;
    int lo[3], hi[3];
    const float o[3] = {g->ox, g->oy, g->oz};
    const int dim[3] = {g->nx, g->ny, g->nz};
    for (int a = 0; a < 3; ++a) {
        lo[a] = clampi((int)floorf((qp[a] - pad - o[a]) * g->inv), 0, dim[a] - 1);
        hi[a] = clampi((int)floorf((qp[a] + pad - o[a]) * g->inv), 0, dim[a] - 1);
    }
    int cnt = 0;
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y) {
            int64_t c0 = ((int64_t)z * g->ny + y) * g->nx + lo[0];
            int64_t c1 = ((int64_t)z * g->ny + y) * g->nx + hi[0];
            for (int64_t s = g->start[c0]; s < g->start[c1 + 1]; ++s) {
                int32_t j = g->items[s];
                const float dx = p[3 * j] - qp[0], dy = p[3 * j + 1] - qp[1], dz = p[3 * j + 2] - qp[2];
                if (ignore_same_pos && dx == 0.0f && dy == 0.0f && dz == 0.0f) continue;
                const float d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 <= r2) {
                    if (cnt < cap) out[cnt] = j;
                    ++cnt;
                }
            }
        }
    int m = cnt < cap ? cnt : cap;
    qsort(out, m, sizeof(int32_t), cmp_i32);
    if (out_d2)
        for (int k = 0; k < m; ++k) {
            int32_t j = out[k];
            const float dx = p[3 * j] - qp[0], dy = p[3 * j + 1] - qp[1], dz = p[3 * j + 2] - qp[2];
            out_d2[k] = (dx * dx + dy * dy) + dz * dz;
        }
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * FixedRadiusSearch: counts only (row_splits) or full lists.
 * Pass 1 (nbr == NULL): fills counts[n_out].  Pass 2: fills nbr[row_splits[i]..] ascending.
 * inclusive d^2 <= r^2; points whose coordinates equal the query's are skipped when
 * ignore_same_pos != 0 (radius_search_ignore_query_points=True, models/transmodel.py:93).
 * ------------------------------------------------------------------------------------------ */
#define NFO_MAX_NBR 4096

NFO_API void nfo_radius_search(const float* in_pos, int64_t n_in, const float* out_pos, int64_t n_out,
                               float radius, int ignore_same_pos, int64_t* counts,
                               const int64_t* row_splits, int32_t* nbr, float* nbr_d2, int nthreads) {
    nfo_grid g;
    grid_build(&g, in_pos, n_in, radius * 1.002f);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int32_t* tmp = (int32_t*)malloc(NFO_MAX_NBR * sizeof(int32_t));
        float* tmpd = (float*)malloc(NFO_MAX_NBR * sizeof(float));
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n_out; ++i) {
            int c = grid_query(&g, in_pos, out_pos + 3 * i, radius, ignore_same_pos, tmp, tmpd, NFO_MAX_NBR);
            if (c > NFO_MAX_NBR) c = NFO_MAX_NBR;
            if (counts) counts[i] = c;
            if (nbr) {
                memcpy(nbr + row_splits[i], tmp, c * sizeof(int32_t));
                if (nbr_d2) memcpy(nbr_d2 + row_splits[i], tmpd, c * sizeof(float));
            }
        }
        free(tmp);
        free(tmpd);
    }
    grid_free(&g);
}

/* ------------------------------------------------------------------------------------------
 * ContinuousConv filter coordinates: relative position (in units where the search ball is the
 * unit ball) -> volume-preserving ball->cylinder->cube map -> trilinear corner weights on the
 * size^3 filter grid (align_corners=True).  Shared verbatim (by restatement, in Python) with
 * oracle/third_party_ops.py so both can be cross-checked.
 * ------------------------------------------------------------------------------------------ */
static float sgnf(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

NFO_API void nfo_ball_to_cube(float* x, float* y, float* z) {
    float X = *x, Y = *y, Z = *z;
    const float sq = X * X + Y * Y + Z * Z;
    const float n = sqrtf(sq);
    if (sq < 1e-12f) { *x = *y = *z = 0.0f; return; }
    /* sphere -> cylinder */
    const float xy2 = X * X + Y * Y;
    if (1.25f * Z * Z > xy2) {
        const float s = sqrtf(3.0f * n / (n + fabsf(Z)));
        X *= s; Y *= s; Z = sgnf(Z) * n;
    } else {
        const float s = n / sqrtf(xy2);
        X *= s; Y *= s; Z *= 1.5f;
    }
    /* cylinder -> cube */
    const float nxy2 = X * X + Y * Y;
    if (nxy2 < 1e-12f) {
        X = 0.0f; Y = 0.0f;
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
    *x = X; *y = Y; *z = Z;
}

/* corner cell indices (8) and weights (8) for one relative position; size = filter size per axis */
NFO_API void nfo_filter_corners(float rx, float ry, float rz, float inv_radius, int size,
                                const float* offset, int32_t* cell, float* w) {
    float x = rx * inv_radius, y = ry * inv_radius, z = rz * inv_radius;
    nfo_ball_to_cube(&x, &y, &z);
    float c[3] = {x * 0.5f, y * 0.5f, z * 0.5f};
    int i0[3], i1[3];
    float f[3];
    for (int a = 0; a < 3; ++a) {
        float t = (c[a] + 0.5f + (offset ? offset[a] : 0.0f)) * (float)(size - 1);
        float fl = floorf(t);
        f[a] = t - fl;
        i0[a] = clampi((int)fl, 0, size - 1);
        i1[a] = clampi((int)fl + 1, 0, size - 1);
    }
    for (int k = 0; k < 8; ++k) {
        const int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
        const int ix = bx ? i1[0] : i0[0], iy = by ? i1[1] : i0[1], iz = bz ? i1[2] : i0[2];
        cell[k] = (iz * size + iy) * size + ix;
        w[k] = (bx ? f[0] : 1.0f - f[0]) * (by ? f[1] : 1.0f - f[1]) * (bz ? f[2] : 1.0f - f[2]);
    }
}

/* ------------------------------------------------------------------------------------------
 * ContinuousConv forward.
 *  kernel: (size,size,size,cin,cout) float, [z][y][x][cin][cout]; bias: (cout) or NULL
 *  out[i] = sum_j window(d2_ij / r^2) * sum_c w_ijc * kernel[c]^T feat[j]  + bias
 *  extent: filter diameter (radius = extent/2).  counts_out (optional): neighbours per out point.
 * ------------------------------------------------------------------------------------------ */
NFO_API void nfo_cconv_forward(const float* in_pos, const float* in_feat, int64_t n_in, int cin,
                               const float* out_pos, int64_t n_out, float extent, int size,
                               const float* kernel, const float* bias, const float* offset,
                               int cout, int ignore_same_pos, int use_window, float* out,
                               int64_t* counts_out, int nthreads) {
    const float radius = 0.5f * extent;
    const float r2 = radius * radius;
    const float inv_radius = 2.0f / extent;
    const int ncell = size * size * size;
    nfo_grid g;
    grid_build(&g, in_pos, n_in, radius * 1.002f);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int32_t* nb = (int32_t*)malloc(NFO_MAX_NBR * sizeof(int32_t));
        float* nd = (float*)malloc(NFO_MAX_NBR * sizeof(float));
        float* patch = (float*)malloc((size_t)ncell * cin * sizeof(float));
#pragma omp for schedule(dynamic, 32)
        for (int64_t i = 0; i < n_out; ++i) {
            const float* qp = out_pos + 3 * i;
            int c = grid_query(&g, in_pos, qp, radius, ignore_same_pos, nb, nd, NFO_MAX_NBR);
            if (c > NFO_MAX_NBR) c = NFO_MAX_NBR;
            if (counts_out) counts_out[i] = c;
            memset(patch, 0, (size_t)ncell * cin * sizeof(float));
            for (int k = 0; k < c; ++k) {
                const int32_t j = nb[k];
                float a = 1.0f;
                if (use_window) {
                    float t = 1.0f - nd[k] / r2;
                    a = t * t * t;
                    a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
                }
                int32_t cell[8];
                float w[8];
                nfo_filter_corners(in_pos[3 * j] - qp[0], in_pos[3 * j + 1] - qp[1],
                                   in_pos[3 * j + 2] - qp[2], inv_radius, size, offset, cell, w);
                const float* f = in_feat + (size_t)j * cin;
                for (int cc = 0; cc < 8; ++cc) {
                    const float ww = a * w[cc];
                    float* pr = patch + (size_t)cell[cc] * cin;
                    for (int ch = 0; ch < cin; ++ch) pr[ch] += ww * f[ch];
                }
            }
            float* o = out + (size_t)i * cout;
            for (int oc = 0; oc < cout; ++oc) o[oc] = bias ? bias[oc] : 0.0f;
            for (int cc = 0; cc < ncell; ++cc)
                for (int ch = 0; ch < cin; ++ch) {
                    const float pv = patch[(size_t)cc * cin + ch];
                    if (pv == 0.0f) continue;
                    const float* kr = kernel + ((size_t)cc * cin + ch) * cout;
                    for (int oc = 0; oc < cout; ++oc) o[oc] += pv * kr[oc];
                }
        }
        free(nb);
        free(nd);
        free(patch);
    }
    grid_free(&g);
}
