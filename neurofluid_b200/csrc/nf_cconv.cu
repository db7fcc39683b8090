// nf_cconv.cu -- transition model kernels (placeholder until the ContinuousConv path lands)
#include "nf_common.cuh"
using namespace nf;
extern "C" size_t nf_transition_packed_weights_bytes(void) { return 0; }
extern "C" int nf_transition_pack_weights(const float* const*, int, void*, void*) {
    set_error("transition model not built yet"); return NF_E_UNSUPPORTED; }
extern "C" size_t nf_transition_workspace_bytes(int, int) { return 0; }
extern "C" int nf_transition_num_phases(void) { return 0; }
extern "C" int nf_transition_step(const nf_transition_args*, void*) {
    set_error("transition model not built yet"); return NF_E_UNSUPPORTED; }
extern "C" size_t nf_cconv_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int nf_cconv_forward(const float*, const float*, int, int, const float*, int, float, const float*,
                                const float*, int, int, int, float*, int32_t*, void*, size_t, void*) {
    set_error("transition model not built yet"); return NF_E_UNSUPPORTED; }
