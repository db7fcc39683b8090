// nf_comm.cu -- NCCL over NVLink / NVSwitch behind the C ABI: one communicator per process (= per GPU), in-place
// all-gathers of equal row blocks on the caller's stream, launched right behind the kernel that produced the rows.
//
// The reference has no distributed code: this is the "position all-gather per step" of
// BASELINE.json's north_star, used by the particle-block sharded transition step (nf_transition_step, phase
// NF_PHASE_SHARDED).  NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the library
// torch already loaded), so libnf_b200.so itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include "nf_common.cuh"

namespace nf {
namespace comm {

struct Api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
};
static Api g_api;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

static bool load_api() {
    if (g_api.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("NCCL not found: %s", dlerror()); return false; }
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_api.AllGather = (decltype(g_api.AllGather))dlsym(h, "ncclAllGather");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllGather && g_api.GetErrorString;
    if (!g_api.ok) set_error("NCCL symbols missing in libnccl");
    return g_api.ok;
}

int world() { return g_world; }
int rank() { return g_rank; }
bool ready() { return g_comm != nullptr || g_world == 1; }

// in place: rank g's block of bytes_per_rank bytes sits at buf + g * bytes_per_rank on every rank
int allgather_inplace(void* buf, size_t bytes_per_rank, cudaStream_t st) {
    if (g_world == 1) return NF_OK;
    NF_REQUIRE(g_comm != nullptr, NF_E_INVALID, "nf_allgather_rows: nf_comm_init has not been called");
    const ncclResult_t r = g_api.AllGather((const char*)buf + (size_t)g_rank * bytes_per_rank, buf, bytes_per_rank, ncclChar, g_comm, st);
    NF_REQUIRE(r == ncclSuccess, NF_E_CUDA, "ncclAllGather failed: %s", g_api.GetErrorString(r));
    return NF_OK;
}

}  // namespace comm
}  // namespace nf

using namespace nf;

extern "C" int nf_comm_unique_id(void* id_out_host /*NF_COMM_ID_BYTES*/) {
    NF_REQUIRE(id_out_host != nullptr, NF_E_INVALID, "nf_comm_unique_id: null argument");
    if (!comm::load_api()) return NF_E_UNSUPPORTED;
    static_assert(NF_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    const ncclResult_t r = comm::g_api.GetUniqueId((ncclUniqueId*)id_out_host);
    NF_REQUIRE(r == ncclSuccess, NF_E_CUDA, "ncclGetUniqueId failed: %s", comm::g_api.GetErrorString(r));
    return NF_OK;
}

extern "C" int nf_comm_init(const void* id_host, int rank, int world) {
    NF_REQUIRE(world >= 1 && rank >= 0 && rank < world, NF_E_INVALID, "nf_comm_init: rank %d of %d", rank, world);
    if (comm::g_comm) { comm::g_api.CommDestroy(comm::g_comm); comm::g_comm = nullptr; }
    comm::g_rank = rank; comm::g_world = world;
    if (world == 1) return NF_OK;
    NF_REQUIRE(id_host != nullptr, NF_E_INVALID, "nf_comm_init: null id");
    if (!comm::load_api()) return NF_E_UNSUPPORTED;
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    const ncclResult_t r = comm::g_api.CommInitRank(&comm::g_comm, world, id, rank);
    if (r != ncclSuccess) {
        comm::g_comm = nullptr; comm::g_world = 1; comm::g_rank = 0;
        set_error("ncclCommInitRank failed: %s", comm::g_api.GetErrorString(r));
        return NF_E_CUDA;
    }
    return NF_OK;
}

extern "C" int nf_comm_finalize(void) {
    if (comm::g_comm) { comm::g_api.CommDestroy(comm::g_comm); comm::g_comm = nullptr; }
    comm::g_rank = 0; comm::g_world = 1;
    return NF_OK;
}

extern "C" int nf_comm_info(int* rank_host, int* world_host) {
    if (rank_host) *rank_host = comm::g_rank;
    if (world_host) *world_host = comm::g_world;
    return NF_OK;
}

extern "C" int nf_allgather_rows(void* buf, size_t bytes_per_rank, void* stream_) {
    NF_REQUIRE(buf != nullptr || bytes_per_rank == 0, NF_E_INVALID, "nf_allgather_rows: null buffer");
    if (bytes_per_rank == 0) return NF_OK;
    return comm::allgather_inplace(buf, bytes_per_rank, (cudaStream_t)stream_);
}
