// nf_comm.cu -- NCCL over NVLink / NVSwitch behind the C ABI: one communicator per process (= per GPU), in-place
// all-gathers of equal row blocks on the caller's stream, launched right behind the kernel that produced the rows.
//
// The reference has no distributed code: this is the "position all-gather per step" of
// BASELINE.json's north_star, used by the particle-block sharded transition step (nf_transition_step, phase
// NF_PHASE_SHARDED).  NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the library
// torch already loaded), so libnf_b200.so itself has no link-time dependency on it.
//
// Peer-memory exchange (nf_comm_register_buffer): for a buffer every rank has registered, the all-gather is not an NCCL
// call but ONE small kernel that stores this rank's rows straight into every peer's copy of the buffer through NVLink
// (CUDA IPC mappings), raises a per-rank epoch flag in every peer's memory, and a one-warp kernel that waits for the
// peers' flags.  The sharded transition step does four such exchanges per step on blocks of 0.15-0.7 MB: it is bound by
// per-collective latency, and this path costs two ~3 us launches where ncclAllGather costs ~35 us.
#include <cuda.h>
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

// ---- peer-memory exchange state
constexpr int MAX_RANKS = 8;
constexpr size_t FLAG_BYTES = 4096;        // [0,64): data epochs written by the peers; [64,68): ticket; [72,80) / [80,88): this rank's
                                           // own data / "entered" epoch counters (device-resident, so that a captured CUDA graph
                                           // of a step replays correctly); [128,132): time-outs; [192,256): "entered" epochs
                                           // written by the peers; [256,..): staging
struct Peers {
    char* buf[MAX_RANKS];                  // the registered buffer of every rank, mapped into this process (own entry: own pointer)
    unsigned long long* flag[MAX_RANKS];   // every rank's flag block
};
static Peers g_peers = {};
static char* g_flags = nullptr;            // this rank's flag block (cudaMalloc in nf_comm_init, shared through CUDA IPC)
static void* g_opened_flags[MAX_RANKS] = {};
static void* g_opened_buf[MAX_RANKS] = {};
static char* g_reg_base = nullptr;         // registered buffer of this rank
static size_t g_reg_bytes = 0;
static bool g_enter_pending = false;       // enter() has signalled; the next exchange waits for the peers' signals first

struct IpcRecord {                         // what a rank tells the others about one allocation
    cudaIpcMemHandle_t handle;             // of the cudaMalloc allocation that contains the pointer
    unsigned long long offset;             // of the pointer inside that allocation
    unsigned long long bytes;
};
static_assert(sizeof(IpcRecord) == 80, "IPC record");

static int ipc_record(void* ptr, size_t bytes, IpcRecord* rec) {
    typedef CUresult (*GetRange)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRange get_range = nullptr;
    if (!get_range) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        NF_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
        NF_REQUIRE(fn && q == cudaDriverEntryPointSuccess, NF_E_UNSUPPORTED, "cuMemGetAddressRange not available");
        get_range = (GetRange)fn;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    NF_REQUIRE(get_range(&base, &size, (CUdeviceptr)ptr) == CUDA_SUCCESS, NF_E_CUDA, "cuMemGetAddressRange failed");
    memset(rec, 0, sizeof(*rec));
    const cudaError_t e = cudaIpcGetMemHandle(&rec->handle, (void*)base);
    if (e != cudaSuccess) {        // e.g. memory from a stream-ordered / expandable-segment allocator
        cudaGetLastError();
        set_error("cudaIpcGetMemHandle: %s (the buffer must come from a plain cudaMalloc allocation)", cudaGetErrorString(e));
        return NF_E_UNSUPPORTED;
    }
    rec->offset = (unsigned long long)((CUdeviceptr)ptr - base);
    rec->bytes = bytes;
    return NF_OK;
}

// every rank's record -> every rank (NCCL all-gather through the staging area of the flag block, then to the host)
static int exchange_records(const IpcRecord* mine, IpcRecord* all) {
    char* stage = g_flags + 256;
    NF_CUDA_OK(cudaMemcpy(stage + (size_t)g_rank * sizeof(IpcRecord), mine, sizeof(IpcRecord), cudaMemcpyHostToDevice));
    const ncclResult_t r = g_api.AllGather(stage + (size_t)g_rank * sizeof(IpcRecord), stage, sizeof(IpcRecord), ncclChar, g_comm, (cudaStream_t)0);
    NF_REQUIRE(r == ncclSuccess, NF_E_CUDA, "ncclAllGather (IPC handles) failed: %s", g_api.GetErrorString(r));
    NF_CUDA_OK(cudaStreamSynchronize((cudaStream_t)0));
    NF_CUDA_OK(cudaMemcpy(all, stage, sizeof(IpcRecord) * (size_t)g_world, cudaMemcpyDeviceToHost));
    return NF_OK;
}

static void close_buffer_mappings() {
    for (int p = 0; p < MAX_RANKS; ++p) {
        if (g_opened_buf[p]) { cudaIpcCloseMemHandle(g_opened_buf[p]); g_opened_buf[p] = nullptr; }
        g_peers.buf[p] = nullptr;
    }
    g_reg_base = nullptr; g_reg_bytes = 0;
}

static void close_all_mappings() {
    close_buffer_mappings();
    for (int p = 0; p < MAX_RANKS; ++p) {
        if (g_opened_flags[p]) { cudaIpcCloseMemHandle(g_opened_flags[p]); g_opened_flags[p] = nullptr; }
        g_peers.flag[p] = nullptr;
    }
    if (g_flags) { cudaFree(g_flags); g_flags = nullptr; }
    g_enter_pending = false;
}

// this rank's block -> the same offset of every peer's buffer; the last block to finish raises this rank's flag everywhere
__global__ void __launch_bounds__(256) k_push_block(Peers P, size_t off, size_t n8, int rank, int world, unsigned long long* my_epoch,
                                                    unsigned int* ticket) {
    const uint2* src = reinterpret_cast<const uint2*>(P.buf[rank] + off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 v = src[i];
        for (int p = 0; p < world; ++p)
            if (p != rank) reinterpret_cast<uint2*>(P.buf[p] + off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {            // every other block's stores are fenced before its ticket
            *ticket = 0u;
            const unsigned long long epoch = *my_epoch + 1;     // every rank counts its exchanges the same way
            *my_epoch = epoch;
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                if (p != rank) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.flag[p] + rank), "l"(epoch) : "memory");
        }
    }
}

// "this rank has entered the exchange sequence: its buffer may be written": one lane per peer
__global__ void k_signal_enter(Peers P, int rank, int world, unsigned long long* my_epoch) {
    const int p = threadIdx.x;
    const unsigned long long epoch = *my_epoch + 1;
    __syncwarp();
    if (p == 0) *my_epoch = epoch;
    if (p < world && p != rank) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.flag[p] + 24 + rank), "l"(epoch) : "memory");
}

// one lane per peer: wait until its flag in THIS rank's memory has reached the epoch (bounded: a lost peer must not hang the GPU
// for ever; nf_comm_exchange_timeouts reports a wait that gave up)
__global__ void k_wait_peers(const unsigned long long* flags, int rank, int world, const unsigned long long* my_epoch, unsigned int* timeouts) {
    const int p = threadIdx.x;
    if (p >= world || p == rank) return;
    const unsigned long long epoch = *my_epoch;          // written by the signal / push kernel just before on this stream
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + p) : "memory");
        if (v < epoch && clock64() - t0 > 200000000000ll) { atomicAdd(timeouts, 1u); break; }     // ~100 s: a peer that is busy
                                                                                                  // elsewhere (host work between steps) is waited for
    } while (v < epoch);
}

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

// Start of a sequence of exchanges on the registered buffer (the sharded transition step calls it before its first kernel):
// tells every peer that this rank's buffer may be written from now on.  Peers that are ahead would otherwise store their
// rows into memory this rank is still using for something else (a replicated step on the same workspace, say).  Within a
// sequence the last exchange is the barrier: nobody starts the next sequence before everybody has pushed its last block.
int enter(cudaStream_t st) {
    if (g_world == 1 || !g_reg_base) return NF_OK;
    k_signal_enter<<<1, 32, 0, st>>>(g_peers, g_rank, g_world, (unsigned long long*)(g_flags + 80));
    NF_LAUNCH_OK();
    g_enter_pending = true;
    return NF_OK;
}

// in place: rank g's block of bytes_per_rank bytes sits at buf + g * bytes_per_rank on every rank
int allgather_inplace(void* buf, size_t bytes_per_rank, cudaStream_t st) {
    if (g_world == 1) return NF_OK;
    NF_REQUIRE(g_comm != nullptr, NF_E_INVALID, "nf_allgather_rows: nf_comm_init has not been called");
    const size_t span = bytes_per_rank * (size_t)g_world;
    if (g_reg_base && (char*)buf >= g_reg_base && (char*)buf + span <= g_reg_base + g_reg_bytes && bytes_per_rank % 8 == 0 &&
        ((uintptr_t)buf & 7) == 0) {
        // peer-memory path: push my block into every peer's buffer, raise my flag there, wait for theirs
        const size_t off = (size_t)((char*)buf - g_reg_base) + (size_t)g_rank * bytes_per_rank;
        const size_t n8 = bytes_per_rank / 8;
        if (g_enter_pending) {
            // a peer's buffer may be written once that peer has entered the sequence (it may have been using the memory otherwise)
            k_wait_peers<<<1, 32, 0, st>>>((const unsigned long long*)(g_flags + 192), g_rank, g_world, (const unsigned long long*)(g_flags + 80),
                                           (unsigned int*)(g_flags + 128));
            NF_LAUNCH_OK();
            g_enter_pending = false;
        }
        const int blocks = (int)((n8 + 255) / 256 < 64 ? (n8 + 255) / 256 : 64);
        k_push_block<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(g_peers, off, n8, g_rank, g_world, (unsigned long long*)(g_flags + 72),
                                                              (unsigned int*)(g_flags + 64));
        NF_LAUNCH_OK();
        k_wait_peers<<<1, 32, 0, st>>>((const unsigned long long*)g_flags, g_rank, g_world, (const unsigned long long*)(g_flags + 72),
                                       (unsigned int*)(g_flags + 128));
        NF_LAUNCH_OK();
        return NF_OK;
    }
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
    comm::close_all_mappings();
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

extern "C" int nf_comm_register_buffer(void* buf, size_t bytes) {
    using namespace comm;
    if (g_world == 1) return NF_OK;
    NF_REQUIRE(g_comm != nullptr, NF_E_INVALID, "nf_comm_register_buffer: nf_comm_init has not been called");
    NF_REQUIRE(g_world <= MAX_RANKS, NF_E_UNSUPPORTED, "nf_comm_register_buffer: at most %d ranks", MAX_RANKS);
    NF_CUDA_OK(cudaDeviceSynchronize());           // nothing of a previous registration is still in flight
    close_buffer_mappings();
    IpcRecord all[MAX_RANKS];
    int rc;
    // Every step below is collective: a rank that fails locally still takes part in the exchanges (with bytes = 0 in its
    // record), so that all ranks see the failure and fall back to NCCL together instead of leaving the others inside a collective.
    auto all_have = [&](unsigned long long want) {
        for (int p = 0; p < g_world; ++p)
            if (all[p].bytes != want) return false;
        return true;
    };
    if (!g_flags) {      // first registration: the flag blocks
        NF_CUDA_OK(cudaMalloc((void**)&g_flags, FLAG_BYTES));
        NF_CUDA_OK(cudaMemset(g_flags, 0, FLAG_BYTES));
        IpcRecord mine;
        if (ipc_record(g_flags, FLAG_BYTES, &mine) != NF_OK) memset(&mine, 0, sizeof(mine));
        if ((rc = exchange_records(&mine, all)) != NF_OK) return rc;
        bool ok = all_have(FLAG_BYTES);
        for (int p = 0; ok && p < g_world; ++p) {
            if (p == g_rank) { g_peers.flag[p] = (unsigned long long*)g_flags; continue; }
            void* base = nullptr;
            if (cudaIpcOpenMemHandle(&base, all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                set_error("cudaIpcOpenMemHandle (flag block of rank %d): %s", p, cudaGetErrorString(cudaGetLastError()));
                ok = false;
                break;
            }
            g_opened_flags[p] = base;
            g_peers.flag[p] = (unsigned long long*)((char*)base + all[p].offset);
        }
        mine.bytes = ok ? FLAG_BYTES : 0;
        if ((rc = exchange_records(&mine, all)) != NF_OK) return rc;
        if (!all_have(FLAG_BYTES)) {
            const bool mine_ok = ok;
            close_all_mappings();
            if (mine_ok) set_error("nf_comm_register_buffer: a rank could not export or map the flag blocks (CUDA IPC)");
            return NF_E_UNSUPPORTED;
        }
    }
    if (buf == nullptr || bytes == 0) return NF_OK;
    IpcRecord mine;
    if (ipc_record(buf, bytes, &mine) != NF_OK) memset(&mine, 0, sizeof(mine));
    if ((rc = exchange_records(&mine, all)) != NF_OK) return rc;
    bool ok = all_have(bytes);
    if (!ok && mine.bytes == bytes)
        set_error("nf_comm_register_buffer: the ranks registered different sizes, or a rank could not export its buffer (CUDA IPC)");
    for (int p = 0; ok && p < g_world; ++p) {
        if (p == g_rank) { g_peers.buf[p] = (char*)buf; continue; }
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle (buffer of rank %d): %s", p, cudaGetErrorString(cudaGetLastError()));
            ok = false;
            break;
        }
        g_opened_buf[p] = base;
        g_peers.buf[p] = (char*)base + all[p].offset;
    }
    // nobody pushes into a peer before every rank has finished mapping (the flags of the next exchange are the barrier for the
    // data, this exchange is the one for the mappings -- and tells everybody whether everybody succeeded)
    mine.bytes = ok ? bytes : 0;
    if ((rc = exchange_records(&mine, all)) != NF_OK) return rc;
    if (!all_have(bytes)) {
        const bool mine_ok = ok;
        close_buffer_mappings();
        if (mine_ok) set_error("nf_comm_register_buffer: a rank could not map a peer's buffer (CUDA IPC)");
        return NF_E_UNSUPPORTED;
    }
    g_reg_base = (char*)buf; g_reg_bytes = bytes;
    return NF_OK;
}

extern "C" int nf_comm_exchange_timeouts(unsigned int* count_host) {
    NF_REQUIRE(count_host != nullptr, NF_E_INVALID, "nf_comm_exchange_timeouts: null argument");
    *count_host = 0;
    if (comm::g_flags) NF_CUDA_OK(cudaMemcpy(count_host, comm::g_flags + 128, 4, cudaMemcpyDeviceToHost));
    return NF_OK;
}

extern "C" int nf_comm_finalize(void) {
    comm::close_all_mappings();
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
    int rc = comm::enter((cudaStream_t)stream_);       // stand-alone use: the exchange is a sequence of its own
    if (rc != NF_OK) return rc;
    return comm::allgather_inplace(buf, bytes_per_rank, (cudaStream_t)stream_);
}
