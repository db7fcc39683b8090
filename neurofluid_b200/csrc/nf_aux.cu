// nf_aux.cu -- the callers' side of the two hot paths, on the device so that
// an evaluation rollout never leaves it.
//
// replaces (reference file:line):
//   get_ray_directions / get_rays            utils/ray_utils.py:85-104, 107-130   (camera rays of one view)
//   _ground_truth_to_prediction_distance     utils/point_eval.py:11-14            (scipy cKDTree nearest neighbour)
//   _distance                                 utils/point_eval.py:7-8
//   img2mse                                   trainer/trainer_e2e.py:24            (mean squared error; PSNR on the host
//                                                                                   side is -10 log10 of it, :25)
#include "nf_common.cuh"

namespace nf {
namespace aux {

// One thread per pixel.  i = column (x), j = row (y) of kornia.create_meshgrid(H, W, normalized_coordinates=False);
// direction = ((i - W/2)/f, -(j - H/2)/f, -1) rotated by c2w[:, :3] and normalised; origin = c2w[:, 3].
__global__ void k_generate_rays(int H, int W, float focal, const float* __restrict__ c2w /*3x4 row-major, device*/,
                                float* __restrict__ rays /*(H*W,6)*/) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int i = p % W, j = p / W;
    const float dx = __fdiv_rn(__fsub_rn((float)i, 0.5f * (float)W), focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)j, 0.5f * (float)H), focal);
    const float dz = -1.0f;
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)      // directions @ c2w[:, :3].T, left-to-right sum like a plain dot product
        d[a] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[4 * a]), __fmul_rn(dy, c2w[4 * a + 1])), __fmul_rn(dz, c2w[4 * a + 2]));
    const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    float* o = rays + (size_t)p * 6;
    o[0] = c2w[3]; o[1] = c2w[7]; o[2] = c2w[11];
    o[3] = __fdiv_rn(d[0], n); o[4] = __fdiv_rn(d[1], n); o[5] = __fdiv_rn(d[2], n);
}

// Nearest point of a grid-sorted set for every query: one warp per query scans the cubes of cells of growing
// Chebyshev radius R around the query's cell until the best distance found is <= R * cell (nothing outside the
// scanned cube can be closer).  Only the shell added by each step is scanned.  Exact (not approximate).
__global__ void __launch_bounds__(256) k_nearest(GridView g, const float* __restrict__ q, int nq,
                                                 float* __restrict__ dist_out, int* __restrict__ idx_out) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq) return;
    const GridHeader* h = g.hdr;
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    float best = 3.0e38f;
    int bi = -1;
    if (h->n > 0) {
        const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
        const float cell = h->cell;
        // the query may lie outside the grid: distances are measured from its clamped cell, and the stopping rule
        // uses the distance from the query to that cell's cube, which only grows with the shell radius
        const int cx = cell_coord(qx, h->origin[0], h->inv_cell, nx);
        const int cy = cell_coord(qy, h->origin[1], h->inv_cell, ny);
        const int cz = cell_coord(qz, h->origin[2], h->inv_cell, nz);
        const int rmax = max(max(max(cx, nx - 1 - cx), max(cy, ny - 1 - cy)), max(cz, nz - 1 - cz));
        for (int R = 0; R <= rmax; ++R) {
            const int z0 = max(cz - R, 0), z1 = min(cz + R, nz - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, ny - 1);
            const int x0 = max(cx - R, 0), x1 = min(cx + R, nx - 1);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const bool face = (z == cz - R) || (z == cz + R) || (y == cy - R) || (y == cy + R);
                    const int row = (z * ny + y) * nx;
                    // on a face row the whole x range is new; elsewhere only the two end cells of the shell
                    for (int part = 0; part < (face ? 1 : 2); ++part) {
                        int xa, xb;
                        if (face) { xa = x0; xb = x1; }
                        else if (part == 0) { xa = cx - R; xb = cx - R; if (xa < 0) continue; }
                        else { xa = cx + R; xb = cx + R; if (xb > nx - 1 || R == 0) continue; }
                        const int beg = __ldg(g.cell_start + row + xa), end = __ldg(g.cell_start + row + xb + 1);
                        for (int t = beg + lane; t < end; t += 32) {
                            const float4 p = __ldg(g.sorted + t);
                            const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                            if (d2 < best) { best = d2; bi = __float_as_int(p.w); }
                        }
                    }
                }
            float wb = best;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wb = fminf(wb, __shfl_xor_sync(NF_FULL, wb, o));
            // everything not yet scanned lies outside the cube of half-width (R + frac) cells around the query
            const float fx = fminf(qx - (h->origin[0] + (float)cx * cell), (h->origin[0] + (float)(cx + 1) * cell) - qx);
            const float fy = fminf(qy - (h->origin[1] + (float)cy * cell), (h->origin[1] + (float)(cy + 1) * cell) - qy);
            const float fz = fminf(qz - (h->origin[2] + (float)cz * cell), (h->origin[2] + (float)(cz + 1) * cell) - qz);
            const float margin = fmaxf(fminf(fminf(fx, fy), fz), 0.f) + (float)R * cell * 0.999f;
            if (wb <= margin * margin) break;
        }
        // warp arg-min (ties: smallest index, like a left-to-right scan)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(NF_FULL, best, o);
            const int oi = __shfl_xor_sync(NF_FULL, bi, o);
            if (ob < best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
        }
    }
    if (lane == 0) {
        dist_out[i] = bi >= 0 ? sqrtf(best) : 3.0e38f;
        if (idx_out) idx_out[i] = bi;
    }
}

// sum of squared differences (double accumulator) -> out[0]; the caller divides by n.
__global__ void __launch_bounds__(256) k_sqdiff_sum(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                    double* __restrict__ out) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        s += (double)d * (double)d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NF_FULL, s, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(out, t);
    }
}

// per-point Euclidean distance between two equally ordered sets (utils/point_eval.py:7-8)
__global__ void k_pair_distance(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float dx = a[3 * i] - b[3 * i], dy = a[3 * i + 1] - b[3 * i + 1], dz = a[3 * i + 2] - b[3 * i + 2];
    out[i] = sqrtf(dx * dx + dy * dy + dz * dz);
}

}  // namespace aux
}  // namespace nf

using namespace nf;
using namespace nf::aux;

extern "C" int nf_generate_rays(int H, int W, float focal, const float* c2w_dev, float* rays_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(H >= 0 && W >= 0 && (long long)H * W < (1ll << 31), NF_E_INVALID, "nf_generate_rays: bad image size %dx%d", H, W);
    NF_REQUIRE(focal > 0.f, NF_E_INVALID, "nf_generate_rays: focal must be positive");
    if (H == 0 || W == 0) return NF_OK;
    NF_REQUIRE(c2w_dev && rays_out, NF_E_INVALID, "nf_generate_rays: null pointer");
    k_generate_rays<<<(H * W + 255) / 256, 256, 0, st>>>(H, W, focal, c2w_dev, rays_out);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" int nf_nearest_distance(const void* grid_ws, int n_points, const float* queries, int nq, float* dist_out,
                                   int32_t* idx_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(nq >= 0 && n_points >= 0, NF_E_INVALID, "nf_nearest_distance: negative size");
    if (nq == 0) return NF_OK;
    NF_REQUIRE(grid_ws && queries && dist_out, NF_E_INVALID, "nf_nearest_distance: null pointer");
    k_nearest<<<(nq + 7) / 8, 256, 0, st>>>(grid_view(grid_ws, n_points), queries, nq, dist_out, idx_out);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" int nf_pair_distance(const float* a, const float* b, int n, float* dist_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(n >= 0, NF_E_INVALID, "nf_pair_distance: negative size");
    if (n == 0) return NF_OK;
    NF_REQUIRE(a && b && dist_out, NF_E_INVALID, "nf_pair_distance: null pointer");
    k_pair_distance<<<(n + 255) / 256, 256, 0, st>>>(a, b, n, dist_out);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" int nf_sqdiff_sum(const float* a, const float* b, long long n, double* sum_out_dev, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(n >= 0, NF_E_INVALID, "nf_sqdiff_sum: negative size");
    NF_REQUIRE(sum_out_dev, NF_E_INVALID, "nf_sqdiff_sum: null output");
    NF_CUDA_OK(cudaMemsetAsync(sum_out_dev, 0, sizeof(double), st));
    if (n == 0) return NF_OK;
    NF_REQUIRE(a && b, NF_E_INVALID, "nf_sqdiff_sum: null pointer");
    const int grid = (int)min((long long)num_sms() * 8, (n + 255) / 256);
    k_sqdiff_sum<<<grid, 256, 0, st>>>(a, b, n, sum_out_dev);
    NF_LAUNCH_OK();
    return NF_OK;
}
