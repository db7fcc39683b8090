// nf_common.cuh -- shared helpers for libnf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/nf_b200.h"

#define NF_FULL 0xffffffffu

namespace nf {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define NF_CUDA_OK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            nf::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NF_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

#define NF_LAUNCH_OK()                                                                            \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            nf::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NF_E_CUDA;                                                                     \
        }                                                                                         \
        nf::count_launch();                                                                       \
    } while (0)

#define NF_REQUIRE(cond, code, ...)                                                               \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            nf::set_error(__VA_ARGS__);                                                           \
            return (code);                                                                        \
        }                                                                                         \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int num_sms();

// ------------------------------------------------------------------------------------------------
// Spatial grid: dense, clamped, cell-sorted copy of the points.  Lives in a caller workspace.
// ------------------------------------------------------------------------------------------------
constexpr int GRID_MAX_DIM = 64;
constexpr int GRID_MAX_CELLS = GRID_MAX_DIM * GRID_MAX_DIM * GRID_MAX_DIM;

struct GridHeader {
    float origin[3];
    float cell;
    float inv_cell;
    float bmin[3];
    float bmax[3];
    int dim[3];
    int ncells;
    int n;
    unsigned bbox_bits[6];  // scratch: ordered-uint encodings of min/max
    int pad[10];
};
static_assert(sizeof(GridHeader) == 128, "GridHeader layout");

struct GridView {
    const GridHeader* hdr;
    const int* cell_start;   // ncells + 1
    const int* occ27;        // ncells: points in the 3x3x3 block around each cell
    const float4* sorted;    // n: xyz + original index (bit-cast) in cell order
};

struct GridLayout {
    size_t off_hdr, off_start, off_fill, off_occ, off_sorted, off_cellof, total;
};
inline GridLayout grid_layout(int n) {
    GridLayout L;
    size_t o = 0;
    L.off_hdr = o; o += align_up(sizeof(GridHeader), 256);
    L.off_start = o; o += align_up(sizeof(int) * (GRID_MAX_CELLS + 1), 256);
    L.off_fill = o; o += align_up(sizeof(int) * GRID_MAX_CELLS, 256);
    L.off_occ = o; o += align_up(sizeof(int) * GRID_MAX_CELLS, 256);
    L.off_sorted = o; o += align_up(sizeof(float4) * (size_t)(n > 0 ? n : 1), 256);
    L.off_cellof = o; o += align_up(sizeof(int) * (size_t)(n > 0 ? n : 1), 256);
    L.total = o;
    return L;
}
inline GridView grid_view(const void* ws, int n_unused = 0) {
    // offsets before `sorted` do not depend on n
    GridLayout L = grid_layout(1);
    const char* b = (const char*)ws;
    GridView g;
    g.hdr = (const GridHeader*)(b + L.off_hdr);
    g.cell_start = (const int*)(b + L.off_start);
    g.occ27 = (const int*)(b + L.off_occ);
    g.sorted = (const float4*)(b + L.off_sorted);
    return g;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NF_FULL, v, o);
    return v;
}

// exact (non-contracted) squared distance, rounding like ((dx*dx + dy*dy) + dz*dz) on a non-FMA CPU
__device__ __forceinline__ float dist2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// monotone cell coordinate used for both binning and query ranges
__device__ __forceinline__ int cell_coord(float v, float origin, float inv_cell, int dim) {
    const int c = __float2int_rd(__fmul_rn(__fsub_rn(v, origin), inv_cell));
    return min(max(c, 0), dim - 1);
}

// First-K-by-index ball query by one warp.  On return lane k holds the k-th smallest in-radius
// particle index (INT_MAX if fewer than k+1 found); returns min(#in-radius, K).  K <= 32.
__device__ __forceinline__ int warp_first_k(const GridView& g, float qx, float qy, float qz, float radius,
                                            int K, int lane, int& best) {
    const GridHeader* h = g.hdr;
    const float r2 = __fmul_rn(radius, radius);
    const float pad = radius * 1.001f + 1e-6f;
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const int lox = cell_coord(qx - pad, ox, inv, nx), hix = cell_coord(qx + pad, ox, inv, nx);
    const int loy = cell_coord(qy - pad, oy, inv, ny), hiy = cell_coord(qy + pad, oy, inv, ny);
    const int loz = cell_coord(qz - pad, oz, inv, nz), hiz = cell_coord(qz + pad, oz, inv, nz);
    best = 0x7fffffff;
    int cnt = 0, kth = 0x7fffffff;
    for (int z = loz; z <= hiz; ++z)
        for (int y = loy; y <= hiy; ++y) {
            const int row = (z * ny + y) * nx;
            const int beg = __ldg(g.cell_start + row + lox), end = __ldg(g.cell_start + row + hix + 1);
            for (int base = beg; base < end; base += 32) {
                const int i = base + lane;
                int idx = 0x7fffffff;
                bool hit = false;
                if (i < end) {
                    const float4 p = __ldg(g.sorted + i);
                    idx = __float_as_int(p.w);
                    hit = (dist2_exact(qx, qy, qz, p.x, p.y, p.z) < r2) && (idx < kth);
                }
                unsigned m = __ballot_sync(NF_FULL, hit);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int v = __shfl_sync(NF_FULL, idx, src);
                    if (v < kth) {  // warp-uniform
                        const int pos = __popc(__ballot_sync(NF_FULL, best < v));
                        const int up = __shfl_up_sync(NF_FULL, best, 1);
                        if (lane == pos) best = v;
                        else if (lane > pos) best = up;
                        if (cnt < K) ++cnt;
                        kth = __shfl_sync(NF_FULL, best, K - 1);
                    }
                }
            }
        }
    return cnt;
}

// conservative "anything within reach?" test for a query point (per lane, no warp cooperation)
__device__ __forceinline__ bool grid_maybe_nonempty(const GridView& g, float qx, float qy, float qz, float radius) {
    const GridHeader* h = g.hdr;
    const float pad = radius * 1.001f + 1e-6f;
    if (h->n == 0) return false;
    if (h->cell <= pad) return true;  // occ27 is only conservative when one cell covers the reach
    if (qx < h->bmin[0] - pad || qx > h->bmax[0] + pad || qy < h->bmin[1] - pad || qy > h->bmax[1] + pad ||
        qz < h->bmin[2] - pad || qz > h->bmax[2] + pad)
        return false;
    const int cx = cell_coord(qx, h->origin[0], h->inv_cell, h->dim[0]);
    const int cy = cell_coord(qy, h->origin[1], h->inv_cell, h->dim[1]);
    const int cz = cell_coord(qz, h->origin[2], h->inv_cell, h->dim[2]);
    return __ldg(g.occ27 + (cz * h->dim[1] + cy) * h->dim[0] + cx) > 0;
}
#endif  // __CUDACC__

}  // namespace nf
