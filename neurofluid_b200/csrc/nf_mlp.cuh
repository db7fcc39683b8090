// nf_mlp.cuh -- launch interface of the fused PE + NeRF MLP kernel (nf_mlp.cu)
#pragma once
#include "nf_common.cuh"

namespace nf {
namespace mlp {

constexpr int PACKED_BYTES = 154 * 8192 + 20 * 4096 + 3208 * 4;
constexpr int RECORD_FLOATS = 16;
// algorithmic multiply-accumulates per evaluated row (models/nerf.py layer shapes)
constexpr long long MAC_PER_ROW = 665984;

struct KernelArgs {
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


int launch(const KernelArgs& a, int dtype, cudaStream_t st);

}  // namespace mlp
}  // namespace nf
