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

// nf_comm.cu: the process's NCCL communicator (world 1 until nf_comm_init)
namespace comm {
int world();
int rank();
int allgather_inplace(void* buf, size_t bytes_per_rank, cudaStream_t st);
int enter(cudaStream_t st);      // start of a sequence of exchanges on the registered buffer (peer-memory path; no-op otherwise)
}

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
    const float4* sorted;    // n: xyz + original index (bit-cast) in cell order, index-ascending per cell
    const float4* orig4;     // n: xyz + cell id (bit-cast) in ORIGINAL order
    const int* cell_of;      // n: cell id in ORIGINAL order
    const int* fine_of;      // n: half-resolution cell id ((fz*2ny + fy)*2nx + fx) in ORIGINAL order
};

struct GridLayout {
    size_t off_hdr, off_start, off_fill, off_occ, off_sorted, off_cellof, off_unordered, off_orig4, off_fineof, total;
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
    L.off_unordered = o; o += align_up(sizeof(int) * (size_t)(n > 0 ? n : 1), 256);
    L.off_orig4 = o; o += align_up(sizeof(float4) * (size_t)(n > 0 ? n : 1), 256);
    L.off_fineof = o; o += align_up(sizeof(int) * ((size_t)(n > 0 ? n : 1) + 128), 256);
    L.total = o;
    return L;
}
inline GridView grid_view(const void* ws, int n) {
    GridLayout L = grid_layout(n);
    const char* b = (const char*)ws;
    GridView g;
    g.hdr = (const GridHeader*)(b + L.off_hdr);
    g.cell_start = (const int*)(b + L.off_start);
    g.occ27 = (const int*)(b + L.off_occ);
    g.sorted = (const float4*)(b + L.off_sorted);
    g.orig4 = (const float4*)(b + L.off_orig4);
    g.cell_of = (const int*)(b + L.off_cellof);
    g.fine_of = (const int*)(b + L.off_fineof);
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

// Sorted-insert of index v into the warp-distributed ascending list (lane k holds the k-th smallest).
__device__ __forceinline__ void warp_list_insert(int& best, int v, int lane) {
    const int pos = __popc(__ballot_sync(NF_FULL, best < v));
    const int up = __shfl_up_sync(NF_FULL, best, 1);
    if (lane == pos) best = v;
    else if (lane > pos) best = up;
}

constexpr int HITBUF = 256;  // per-warp shared-memory buffer of hit indices

// Keep the K smallest of buf[0..n) (n > K, values distinct): compacts them into buf[0..K) and returns the
// K-th smallest.  Selection by bisection on the index value with warp-wide counting (REDUX).
__device__ __forceinline__ int warp_select_k(int* buf, int n, int K, int lane, int vmax) {
    int e[HITBUF / 32];
#pragma unroll
    for (int t = 0; t < HITBUF / 32; ++t) {
        const int i = t * 32 + lane;
        e[t] = i < n ? buf[i] : 0x7fffffff;
    }
    int lo = 0, hi = vmax;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
#pragma unroll
        for (int t = 0; t < HITBUF / 32; ++t) c += (e[t] <= mid);
        c = __reduce_add_sync(NF_FULL, c);
        if (c >= K) hi = mid; else lo = mid + 1;
    }
    __syncwarp();
    int base = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < HITBUF / 32; ++t) {
        const bool keep = e[t] <= lo;
        const unsigned m = __ballot_sync(NF_FULL, keep);
        if (keep) buf[base + __popc(m & lt)] = e[t];
        base += __popc(m);
    }
    __syncwarp();
    return lo;
}

// First-K-by-index ball query by one warp, row-scan flavour: streams the <= 9 contiguous (y,z) rows of the
// query's cell neighbourhood 32 candidates at a time, appends in-radius indices to `buf` (shared memory,
// HITBUF ints owned by this warp) and selects the K smallest at the end (or whenever the buffer fills, after
// which later candidates are pre-filtered against the K-th best so far).
// On return lane k holds the k-th smallest in-radius particle index (INT_MAX if fewer than k+1 found);
// returns min(#in-radius, K).  K <= 32.
__device__ __forceinline__ int warp_first_k_rows(const GridView& g, float qx, float qy, float qz, float radius,
                                                 int K, int lane, int& best, int& iters, int* buf) {
    const GridHeader* h = g.hdr;
    const float r2 = __fmul_rn(radius, radius);
    const float pad = radius * 1.001f + 1e-6f;
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const int lox = cell_coord(qx - pad, ox, inv, nx), hix = cell_coord(qx + pad, ox, inv, nx);
    const int loy = cell_coord(qy - pad, oy, inv, ny), hiy = cell_coord(qy + pad, oy, inv, ny);
    const int loz = cell_coord(qz - pad, oz, inv, nz), hiz = cell_coord(qz + pad, oz, inv, nz);
    const unsigned lt = (1u << lane) - 1u;
    const int vmax = h->n;
    int n = 0, kth = 0x7fffffff;
    __syncwarp();
    for (int z = loz; z <= hiz; ++z)
        for (int y = loy; y <= hiy; ++y) {
            const int row = (z * ny + y) * nx;
            const int beg = __ldg(g.cell_start + row + lox), end = __ldg(g.cell_start + row + hix + 1);
            for (int base = beg; base < end; base += 32) {
                const int i = base + lane;
                int idx = 0x7fffffff;
                bool hit = false;
                ++iters;
                if (i < end) {
                    const float4 p = __ldg(g.sorted + i);
                    idx = __float_as_int(p.w);
                    hit = (idx < kth) && (dist2_exact(qx, qy, qz, p.x, p.y, p.z) < r2);
                }
                unsigned m = __ballot_sync(NF_FULL, hit);
                if (m) {
                    if (n > HITBUF - 32) {      // uniform: make room, tighten the pre-filter
                        kth = warp_select_k(buf, n, K, lane, vmax);
                        n = K;
                        hit = hit && (idx < kth);
                        m = __ballot_sync(NF_FULL, hit);
                    }
                    if (hit) buf[n + __popc(m & lt)] = idx;
                    n += __popc(m);
                }
            }
        }
    __syncwarp();
    if (n > K) {
        warp_select_k(buf, n, K, lane, vmax);
        n = K;
    }
    // ascending order across lanes (rank sort of <= K values)
    const int v = lane < n ? buf[lane] : 0x7fffffff;
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (__shfl_sync(NF_FULL, v, j) < v);
    __syncwarp();
    if (lane < n) buf[rank] = v;
    __syncwarp();
    best = lane < n ? buf[lane] : 0x7fffffff;
    return n;
}

// Lockstep flavour: lane c owns one of the <= 27 neighbourhood cells and walks it in ascending original
// index (cells are index-sorted), one element per iteration.  A lane retires as soon as its next index
// is not below the current K-th best, so a query deep inside the fluid touches only the ~K/(hit rate)
// lowest-index candidates instead of all of them.  Requires cell > reach (<= 3 cells per axis).
__device__ __forceinline__ int warp_first_k_lockstep(const GridView& g, float qx, float qy, float qz, float radius,
                                                     int K, int lane, int& best, int& iters) {
    const GridHeader* h = g.hdr;
    const float r2 = __fmul_rn(radius, radius);
    const float pad = radius * 1.001f + 1e-6f;
    const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2], inv = h->inv_cell;
    const int nx = h->dim[0], ny = h->dim[1], nz = h->dim[2];
    const int lox = cell_coord(qx - pad, ox, inv, nx), hix = cell_coord(qx + pad, ox, inv, nx);
    const int loy = cell_coord(qy - pad, oy, inv, ny), hiy = cell_coord(qy + pad, oy, inv, ny);
    const int loz = cell_coord(qz - pad, oz, inv, nz), hiz = cell_coord(qz + pad, oz, inv, nz);
    // lane -> (dx,dy,dz) in a 3x3x3 block anchored at (lox,loy,loz)
    const int dx = lane % 3, dy = (lane / 3) % 3, dz = lane / 9;
    int cur = 0, end = 0;
    if (lane < 27 && lox + dx <= hix && loy + dy <= hiy && loz + dz <= hiz) {
        const int c = ((loz + dz) * ny + (loy + dy)) * nx + (lox + dx);
        cur = __ldg(g.cell_start + c);
        end = __ldg(g.cell_start + c + 1);
    }
    best = 0x7fffffff;
    int cnt = 0, kth = 0x7fffffff;
    while (true) {
        int idx = 0x7fffffff;
        bool hit = false;
        ++iters;
        if (cur < end) {
            const float4 p = __ldg(g.sorted + cur);
            idx = __float_as_int(p.w);
            if (idx < kth) {
                hit = dist2_exact(qx, qy, qz, p.x, p.y, p.z) < r2;
                ++cur;
            } else {
                cur = end;  // ascending list: nothing further in this cell can enter the best K
            }
        }
        if (!__any_sync(NF_FULL, cur < end || hit)) {
            // nobody has anything left and no pending hit
            break;
        }
        unsigned m = __ballot_sync(NF_FULL, hit);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int v = __shfl_sync(NF_FULL, idx, src);
            if (v < kth) {
                warp_list_insert(best, v, lane);
                if (cnt < K) ++cnt;
                kth = __shfl_sync(NF_FULL, best, K - 1);
            }
        }
    }
    return cnt;
}

// occ: number of points in the 3x3x3 cell block around the query (from grid_occupancy); picks the flavour.
struct QueryStats {
    int n_lock = 0, n_rows = 0, it_lock = 0, it_rows = 0;
};
constexpr int LOCKSTEP_MIN_OCC = 2000;
__device__ __forceinline__ int warp_first_k(const GridView& g, float qx, float qy, float qz, float radius, int K,
                                            int lane, int& best, int occ, int min_occ, QueryStats& qs, int* buf) {
    const float pad = radius * 1.001f + 1e-6f;
    if (occ >= min_occ && g.hdr->cell > pad) {
        ++qs.n_lock;
        return warp_first_k_lockstep(g, qx, qy, qz, radius, K, lane, best, qs.it_lock);
    }
    ++qs.n_rows;
    return warp_first_k_rows(g, qx, qy, qz, radius, K, lane, best, qs.it_rows, buf);
}

// Conservative "anything within reach?" test for a query point (per lane, no warp cooperation):
// returns an upper bound on the number of candidate points (0 = certainly no neighbour).
__device__ __forceinline__ int grid_occupancy(const GridView& g, float qx, float qy, float qz, float radius) {
    const GridHeader* h = g.hdr;
    const float pad = radius * 1.001f + 1e-6f;
    if (h->n == 0) return 0;
    if (qx < h->bmin[0] - pad || qx > h->bmax[0] + pad || qy < h->bmin[1] - pad || qy > h->bmax[1] + pad ||
        qz < h->bmin[2] - pad || qz > h->bmax[2] + pad)
        return 0;
    if (h->cell <= pad) return 1;  // occ27 is only an upper bound when one cell covers the reach
    const int cx = cell_coord(qx, h->origin[0], h->inv_cell, h->dim[0]);
    const int cy = cell_coord(qy, h->origin[1], h->inv_cell, h->dim[1]);
    const int cz = cell_coord(qz, h->origin[2], h->inv_cell, h->dim[2]);
    return __ldg(g.occ27 + (cz * h->dim[1] + cy) * h->dim[0] + cx);
}
__device__ __forceinline__ bool grid_maybe_nonempty(const GridView& g, float qx, float qy, float qz, float radius) {
    return grid_occupancy(g, qx, qy, qz, radius) > 0;
}
#endif  // __CUDACC__

}  // namespace nf
