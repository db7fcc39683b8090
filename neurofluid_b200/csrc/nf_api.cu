// nf_api.cu -- version / error plumbing of libnf_b200.so
#include <stdarg.h>
#include <string.h>

#include "nf_common.cuh"

namespace nf {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}
}  // namespace nf

extern "C" int nf_version(void) { return NF_B200_VERSION; }
extern "C" const char* nf_last_error(void) { return nf::g_err; }
extern "C" int64_t nf_launch_count(void) { return (int64_t)nf::g_launches.load(); }
