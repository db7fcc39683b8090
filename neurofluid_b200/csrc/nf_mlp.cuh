// nf_mlp.cuh -- launch interface, constants and packed-weight layout of the fused PE + NeRF MLP kernels (nf_mlp.cu, nf_mlp2.cu)
#pragma once
#include "nf_common.cuh"

namespace nf {
namespace mlp {

constexpr int PACKED_BYTES = 154 * 8192 + 20 * 4096 + 3208 * 4;
constexpr int RECORD_FLOATS = 16;
// algorithmic multiply-accumulates per evaluated row (models/nerf.py layer shapes)
constexpr long long MAC_PER_ROW = 665984;

constexpr int TILE_M = 128;
constexpr int KX_STEPS = 13;   // xyz-like features 198 -> 208 = 13 K-steps of 16
constexpr int KD_STEPS = 4;    // dir-like features 54 -> 64
constexpr int KH_STEPS = 16;   // hidden width 256
constexpr int STAGE_BYTES = 8192;  // one K-step of a 256-row weight slab: 2 k-chunks x 256 rows x 16 B
constexpr int N256_STEPS = 154;
constexpr int N128_STEPS = 20;
constexpr int W_BYTES = N256_STEPS * 8192 + N128_STEPS * 4096;
// small fp32 params appended after the weight slabs
constexpr int SP_BIAS = 0;       // [10][256]
constexpr int SP_WSIG = 2560;    // [256]
constexpr int SP_BSIG = 2816;    // [1] (+3 pad)
constexpr int SP_WRGB = 2820;    // [3][128]
constexpr int SP_BRGB = 3204;    // [3] (+1 pad)
constexpr int SP_FLOATS = 3208;
static_assert(PACKED_BYTES == W_BYTES + SP_FLOATS * 4, "packed size");

// Packed weight stream = a sequence of UNITS in consumption order:
//   for layer l (0..9; 8 = xyz_encoding_final, 9 = dir_encoding), for segment (0: encoded-feature columns, 1: hidden columns),
//   for K-block k0 = 0, kb, 2 kb ... of the segment's K-steps (16 input columns each; kb = wu_ksteps(segment): 4 for the
//   encoded-feature segments, 8 for the hidden ones), for N-half nh = 0, 1 of the layer's outputs:
//   unit = [CTA r of the pair (2)][K-step j < g][8-column chunk kc (2)][row (rpc)][8 halves]
//   with rpc = 64 (32 for the dir layer) rows per CTA: output feature nh * 2 rpc + r * rpc + row.
constexpr int WU_KSTEPS = 8;      // the most K-steps a unit holds
__host__ __device__ inline int wu_ksteps(int seg) { return seg == 0 ? 4 : 8; }
__host__ __device__ inline int layer_pe_steps(int l) { return (l == 0 || l == 4) ? KX_STEPS : (l == 9 ? KD_STEPS : 0); }


// Encoding ablations: fixed-layout column k of the xyz-like block (0..207) / dir-like block (0..63) -> column of the
// network's own (narrower) input, or -1 when the block is disabled or k is padding.  flags = NF_ENC_*.
__host__ __device__ inline int enc_in_xyz(int flags) { return 63 + ((flags & 1) ? 9 : 0) + ((flags & 2) ? 63 : 0) + ((flags & 4) ? 63 : 0); }
__host__ __device__ inline int enc_in_dir(int flags) { return 27 + ((flags & 8) ? 27 : 0); }
__host__ __device__ inline int enc_col_xyz(int k, int flags) {
    if (k < 63) return k;
    const int d = (flags & 1) ? 9 : 0, s = (flags & 2) ? 63 : 0;
    if (k < 72) return (flags & 1) ? 63 + (k - 63) : -1;
    if (k < 135) return (flags & 2) ? 63 + d + (k - 72) : -1;
    if (k < 198) return (flags & 4) ? 63 + d + s + (k - 135) : -1;
    return -1;
}
__host__ __device__ inline int enc_col_dir(int k, int flags) {
    if (k < 27) return k;
    if (k < 54) return (flags & 8) ? 27 + (k - 27) : -1;
    return -1;
}

// Encoded-feature scratch of the two-tile kernel (nf_mlp2.cu): per CTA, 2 passes x 2 tiles x (xyz-like 26 + dir-like 8 chunks
// of 2 KB).  Part of the caller's workspace; sized for the largest grid the kernel launches (one CTA per SM, <= 160).
constexpr int PE_SCRATCH_CTAS = 160;
constexpr size_t PE_SCRATCH_BYTES = (size_t)PE_SCRATCH_CTAS * 2 * 2 * (26 + 8) * 2048;

struct KernelArgs {
    uint8_t* pe_scratch;     // PE_SCRATCH_BYTES of device memory, 16-byte aligned (L2-resident while a launch runs)
    const uint8_t* packed;   // weight slabs + small params
    const float* records;    // (n_rows,16)
    const int* rowid;        // (n_rows) destination index in out4, <0 = skip; NULL = identity
    const int* n_rows_dev;   // device-resident row count (NULL -> n_rows_host)
    int n_rows_host;
    int n_rows_cap;          // upper bound used to clamp the device count
    int n_layers;            // 10 = full net, 8 = sigma only
    int desc_swap;           // debug: swap LBO/SBO
    float4* out4;
    long long* trace;        // debug: per-role clock64 timeline of one tile of CTA 0 (NULL = off)
};


int launch(const KernelArgs& a, int dtype, cudaStream_t st);        // nf_mlp.cu: dispatch (production: launch2)
int launch2(const KernelArgs& a, int dtype, cudaStream_t st);       // nf_mlp2.cu: CTA pairs, two 128-row tiles per CTA sharing the weight stream

}  // namespace mlp
}  // namespace nf
