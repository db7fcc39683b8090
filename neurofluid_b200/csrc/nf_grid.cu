// nf_grid.cu -- dense clamped spatial grid (counting sort by cell) + first-K ball query entry point.
//
// replaces: the brute-force scan of pytorch3d.ops.ball_query behind models/renderer.py:112-122 and the
// FixedRadiusSearch table build behind open3d ContinuousConv (models/transmodel.py:116,118,125).
//
// Layout in HBM (caller workspace, see grid_layout()):
//   GridHeader | cell_start[ncells+1] | fill[ncells] | occ27[ncells] | sorted float4[n] | cell_of[n] | unordered[n]
// `sorted` holds (x,y,z,bit-cast original index) grouped by cell, so a query streams 16-byte records
// from at most 9 contiguous ranges (the x-neighbours of a (y,z) row are adjacent in memory).
// Inside a cell the points are ascending in original index (atomic scatter + rank pass = a stable
// counting sort), which is what lets the first-K query stop early and makes every neighbour list
// deterministic.
#include "nf_common.cuh"

namespace nf {

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_grid_init(GridHeader* h) {
    if (threadIdx.x < 3) {
        h->bbox_bits[threadIdx.x] = 0xffffffffu;
        h->bbox_bits[3 + threadIdx.x] = 0u;
    }
}

__global__ void k_grid_bbox(const float* __restrict__ pos, int n, GridHeader* h) {
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pos[3 * i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(NF_FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(NF_FULL, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&h->bbox_bits[a], f2ord(mn[a]));
            atomicMax(&h->bbox_bits[3 + a], f2ord(mx[a]));
        }
    }
}

__global__ void k_grid_setup(GridHeader* h, int n, float cell) {
    if (threadIdx.x != 0) return;
    float mn[3], mx[3], ext = 0.f;
    for (int a = 0; a < 3; ++a) {
        mn[a] = n > 0 ? ord2f(h->bbox_bits[a]) : 0.f;
        mx[a] = n > 0 ? ord2f(h->bbox_bits[3 + a]) : 0.f;
        if (!(mn[a] > -1e30f && mn[a] < 1e30f)) mn[a] = 0.f;   // NaN / inf guards
        if (!(mx[a] > -1e30f && mx[a] < 1e30f)) mx[a] = mn[a];
        ext = fmaxf(ext, mx[a] - mn[a]);
    }
    const float c = fmaxf(cell, ext / (float)(GRID_MAX_DIM - 1));
    h->cell = c;
    h->inv_cell = 1.0f / c;
    int nc = 1;
    for (int a = 0; a < 3; ++a) {
        h->origin[a] = mn[a];
        h->bmin[a] = mn[a];
        h->bmax[a] = mx[a];
        int d = (int)floorf((mx[a] - mn[a]) * h->inv_cell) + 1;
        d = min(max(d, 1), GRID_MAX_DIM);
        h->dim[a] = d;
        nc *= d;
    }
    h->ncells = nc;
    h->n = n;
}

__global__ void k_grid_count(const float* __restrict__ pos, int n, const GridHeader* __restrict__ h,
                             int* __restrict__ counts /* = cell_start + 1 */, int* __restrict__ cell_of,
                             float4* __restrict__ orig4, int* __restrict__ fine_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coord(pos[3 * i], h->origin[0], h->inv_cell, h->dim[0]);
    const int cy = cell_coord(pos[3 * i + 1], h->origin[1], h->inv_cell, h->dim[1]);
    const int cz = cell_coord(pos[3 * i + 2], h->origin[2], h->inv_cell, h->dim[2]);
    const int c = (cz * h->dim[1] + cy) * h->dim[0] + cx;
    cell_of[i] = c;
    {
        const float inv2 = __fmul_rn(h->inv_cell, 2.0f);
        const int fx = cell_coord(pos[3 * i], h->origin[0], inv2, 2 * h->dim[0]);
        const int fy = cell_coord(pos[3 * i + 1], h->origin[1], inv2, 2 * h->dim[1]);
        const int fz = cell_coord(pos[3 * i + 2], h->origin[2], inv2, 2 * h->dim[2]);
        fine_of[i] = (fz * 2 * h->dim[1] + fy) * 2 * h->dim[0] + fx;
    }
    orig4[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __int_as_float(c));
    atomicAdd(counts + c, 1);
}

// single-CTA inclusive scan of counts[0..ncells) in place (cell_start[0] stays 0)
__global__ void __launch_bounds__(1024) k_grid_scan(const GridHeader* __restrict__ h, int* __restrict__ counts) {
    __shared__ int warp_tot[32];
    const int nc = h->ncells;
    const int per = (nc + 1023) / 1024;
    const int beg = min(threadIdx.x * per, nc), end = min(beg + per, nc);
    int s = 0;
    for (int i = beg; i < end; ++i) s += counts[i];
    // block exclusive scan of s
    int incl = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(NF_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(NF_FULL, v, o);
            if (lane >= o) v += t;
        }
        warp_tot[lane] = v;
    }
    __syncthreads();
    int run = incl - s + (w > 0 ? warp_tot[w - 1] : 0);
    for (int i = beg; i < end; ++i) {
        run += counts[i];
        counts[i] = run;
    }
}

__global__ void k_grid_scatter(int n, const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                               int* __restrict__ fill, int* __restrict__ unordered) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of[i];
    unordered[cell_start[c] + atomicAdd(fill + c, 1)] = i;
}

// Make every cell ascending in original index (what a stable counting sort would give): element e of a
// cell goes to position #{e' in the same cell : e' < e}.  O(cell population) per point; adjacent threads
// read the same segment, so the loads are L1 broadcasts.
__global__ void k_grid_rank(const float* __restrict__ pos, int n, const int* __restrict__ cell_of,
                            const int* __restrict__ cell_start, const int* __restrict__ unordered,
                            float4* __restrict__ sorted) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = unordered[s];
    const int c = cell_of[i];
    const int beg = cell_start[c], end = cell_start[c + 1];
    int rank = 0;
    for (int t = beg; t < end; ++t) rank += (unordered[t] < i);
    sorted[beg + rank] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __int_as_float(i));
}

__global__ void k_grid_occ(const GridHeader* __restrict__ h, const int* __restrict__ cell_start,
                           int* __restrict__ occ27) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= h->ncells) return;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const int x = c % nx, y = (c / nx) % ny, z = c / (nx * ny);
    const int x0 = max(x - 1, 0), x1 = min(x + 1, nx - 1);
    int tot = 0;
    for (int zz = max(z - 1, 0); zz <= min(z + 1, nz - 1); ++zz)
        for (int yy = max(y - 1, 0); yy <= min(y + 1, ny - 1); ++yy) {
            const int row = (zz * ny + yy) * nx;
            tot += cell_start[row + x1 + 1] - cell_start[row + x0];
        }
    occ27[c] = tot;
}

__global__ void __launch_bounds__(256) k_ballquery(GridView g, const float* __restrict__ q, int nq, float radius,
                                                   int K, int* __restrict__ idx_out, int* __restrict__ cnt_out) {
    __shared__ int sm_hits[8][HITBUF];
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int* buf = sm_hits[threadIdx.x >> 5];
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += nwarps) {
        const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        int best = 0x7fffffff, cnt = 0;
        const int occ = grid_occupancy(g, qx, qy, qz, radius);
        QueryStats qs;
        if (occ > 0) cnt = warp_first_k(g, qx, qy, qz, radius, K, lane, best, occ, LOCKSTEP_MIN_OCC, qs, buf);
        if (lane < K) idx_out[(size_t)i * K + lane] = lane < cnt ? best : -1;
        if (lane == 0) cnt_out[i] = cnt;
    }
}

}  // namespace nf

using namespace nf;

extern "C" size_t nf_grid_workspace_bytes(int n_points) { return grid_layout(n_points).total; }

extern "C" int nf_grid_build(const float* pos, int n, float cell, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(ws != nullptr && n >= 0 && cell > 0.f, NF_E_INVALID, "nf_grid_build: bad arguments");
    NF_REQUIRE(n == 0 || pos != nullptr, NF_E_INVALID, "nf_grid_build: null positions");
    const GridLayout L = grid_layout(n);
    NF_REQUIRE(ws_bytes >= L.total, NF_E_WORKSPACE, "nf_grid_build: workspace %zu < %zu", ws_bytes, L.total);
    char* b = (char*)ws;
    GridHeader* h = (GridHeader*)(b + L.off_hdr);
    int* cell_start = (int*)(b + L.off_start);
    int* fill = (int*)(b + L.off_fill);
    int* occ = (int*)(b + L.off_occ);
    float4* sorted = (float4*)(b + L.off_sorted);
    int* cell_of = (int*)(b + L.off_cellof);
    int* unordered = (int*)(b + L.off_unordered);
    float4* orig4 = (float4*)(b + L.off_orig4);
    int* fine_of = (int*)(b + L.off_fineof);
    // the tail padding of fine_of (read by 128-wide filter steps) must never match a marked cell
    NF_CUDA_OK(cudaMemsetAsync(fine_of + n, 0xff, 128 * sizeof(int), st));
    // zero cell_start and fill in one memset (they are adjacent)
    NF_CUDA_OK(cudaMemsetAsync(cell_start, 0, L.off_occ - L.off_start, st));
    k_grid_init<<<1, 32, 0, st>>>(h);
    NF_LAUNCH_OK();
    if (n > 0) {
        const int nb = min((n + 255) / 256, 4 * num_sms());
        k_grid_bbox<<<nb, 256, 0, st>>>(pos, n, h);
        NF_LAUNCH_OK();
    }
    k_grid_setup<<<1, 32, 0, st>>>(h, n, cell);
    NF_LAUNCH_OK();
    if (n > 0) {
        k_grid_count<<<(n + 255) / 256, 256, 0, st>>>(pos, n, h, cell_start + 1, cell_of, orig4, fine_of);
        NF_LAUNCH_OK();
    }
    k_grid_scan<<<1, 1024, 0, st>>>(h, cell_start + 1);
    NF_LAUNCH_OK();
    if (n > 0) {
        k_grid_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, cell_of, cell_start, fill, unordered);
        NF_LAUNCH_OK();
        k_grid_rank<<<(n + 255) / 256, 256, 0, st>>>(pos, n, cell_of, cell_start, unordered, sorted);
        NF_LAUNCH_OK();
    }
    k_grid_occ<<<(GRID_MAX_CELLS + 255) / 256, 256, 0, st>>>(h, cell_start, occ);
    NF_LAUNCH_OK();
    return NF_OK;
}

extern "C" int nf_ballquery_firstk(const void* grid_ws, int n_points, const float* queries, int nq, float radius, int K,
                                   int32_t* idx_out, int32_t* count_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NF_REQUIRE(K >= 1 && K <= 32, NF_E_UNSUPPORTED, "nf_ballquery_firstk: K=%d not in [1,32]", K);
    NF_REQUIRE(nq >= 0, NF_E_INVALID, "nf_ballquery_firstk: negative query count");
    if (nq == 0) return NF_OK;
    NF_REQUIRE(grid_ws && idx_out && count_out, NF_E_INVALID, "nf_ballquery_firstk: null pointer");
    if (nq == 0) return NF_OK;
    NF_REQUIRE(queries != nullptr, NF_E_INVALID, "nf_ballquery_firstk: null queries");
    const GridView g = grid_view(grid_ws, n_points);
    const int nb = min((nq + 7) / 8, 16 * num_sms());
    k_ballquery<<<nb, 256, 0, st>>>(g, queries, nq, radius, K, idx_out, count_out);
    NF_LAUNCH_OK();
    return NF_OK;
}
