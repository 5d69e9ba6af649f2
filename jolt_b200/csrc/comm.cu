// NCCL plumbing for the index-sharded sumcheck. One process per GPU; the
// process group / rendezvous belongs to the caller (torch.distributed broadcasts the unique id), the
// per-round collectives are issued here, on the context's stream, between the round kernel and the
// tiny publish kernel - no host synchronisation besides the Fiat-Shamir round trip itself.
// libnccl is resolved at run time (dlopen by soname finds the copy torch already loaded), so
// libjolt_b200.so has no link-time dependency on it and still loads on a CPU-only box.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "ctx.hpp"
#include "poly_kernels.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
std::mutex g_api_mu;

const NcclApi* nccl_api(const char* path, std::string* err) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (g_api.lib) return &g_api;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        if (err) *err = std::string("cannot load libnccl: ") + (dlerror() ? dlerror() : "not found");
        return nullptr;
    }
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        if (err) *err = "libnccl is missing a required symbol";
        dlclose(lib);
        return nullptr;
    }
    g_api = a;
    return &g_api;
}

__global__ void publish_lanes_kernel(const uint64_t* lanes, int n_u64, uint64_t* result, volatile uint64_t* flag,
                                     uint64_t seq) {
    int i = threadIdx.x;
    for (; i < n_u64; i += blockDim.x) result[i] = lanes[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *flag = seq;
    }
}

}  // namespace

int jb_ctx::comm_allreduce_lanes(uint64_t* d_lanes, size_t n_u64) {
    const NcclApi* a = nccl_api(nullptr, &err);
    if (!a || !nccl_comm) return fail(JB_ERR_INVALID, "no NCCL communicator on this context (jb_comm_init)");
    ncclResult_t r = a->AllReduce(d_lanes, d_lanes, n_u64, ncclUint64, ncclSum, (ncclComm_t)nccl_comm, stream);
    if (r != ncclSuccess) return fail(JB_ERR_CUDA, a->GetErrorString(r));
    return JB_OK;
}

int jb_ctx::comm_allgather(const uint64_t* d_send, uint64_t* d_recv, size_t n_u64_per_rank) {
    const NcclApi* a = nccl_api(nullptr, &err);
    if (!a || !nccl_comm) return fail(JB_ERR_INVALID, "no NCCL communicator on this context (jb_comm_init)");
    ncclResult_t r = a->AllGather(d_send, d_recv, n_u64_per_rank, ncclUint64, (ncclComm_t)nccl_comm, stream);
    if (r != ncclSuccess) return fail(JB_ERR_CUDA, a->GetErrorString(r));
    return JB_OK;
}

// lanes (device) -> host-mapped result + sequence flag; the caller then spins on the flag.
int jb_ctx::publish_lanes(const uint64_t* d_lanes, int n_u64) {
    ++result_seq;
    publish_lanes_kernel<<<1, 64, 0, stream>>>(d_lanes, n_u64, d_result_alias, d_result_alias + 64, result_seq);
    launches++;
    return check(cudaGetLastError(), "publish_lanes launch");
}

extern "C" {

int jb_comm_unique_id(uint8_t out[128], const char* libnccl_path_or_null) {
    if (!out) return JB_ERR_INVALID;
    std::string err;
    const NcclApi* a = nccl_api(libnccl_path_or_null, &err);
    if (!a) return JB_ERR_UNSUPPORTED;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (a->GetUniqueId(&id) != ncclSuccess) return JB_ERR_CUDA;
    std::memcpy(out, &id, 128);
    return JB_OK;
}

int jb_comm_init(jb_ctx* c, int nranks, int rank, const uint8_t id_bytes[128], const char* libnccl_path_or_null) {
    if (!c || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return JB_ERR_INVALID;
    if (nranks & (nranks - 1)) return c->fail(JB_ERR_UNSUPPORTED, "comm: world size must be a power of two");
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    const NcclApi* a = nccl_api(libnccl_path_or_null, &c->err);
    if (!a) return JB_ERR_UNSUPPORTED;
    if (c->nccl_comm) return c->fail(JB_ERR_INVALID, "comm: already initialised");
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return c->fail(JB_ERR_CUDA, a->GetErrorString(r));
    c->nccl_comm = comm;
    c->world = nranks;
    c->rank = rank;
    if (!c->d_lanes && cudaMalloc((void**)&c->d_lanes, 64 * 8) != cudaSuccess) return c->fail(JB_ERR_OOM, "comm: lanes buffer");
    return JB_OK;
}

// ---- peer-memory exchange buffers (CUDA IPC): the handle of this rank's buffer goes out through the
// caller's rendezvous, the peers' handles come back and are opened here.
int jb_comm_p2p_handle(jb_ctx* c, uint8_t out[64]) {
    if (!c || !out) return JB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (c->world > 16 || c->rank >= 16) return c->fail(JB_ERR_UNSUPPORTED, "p2p: at most 16 ranks (the NCCL path stays)");
    c->quiesce_resident(true);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!c->xch_peer[c->rank]) {
        void* p = nullptr;
        static_assert(jb::XCH_BYTES <= jb::XCH_ARENA_OFFSET, "exchange area too small");
        // exchange slots + flags, then the gather arena (two halves)
        if (cudaMalloc(&p, jb::XCH_TOTAL_BYTES) != cudaSuccess) return c->fail(JB_ERR_OOM, "p2p: exchange buffer");
        cudaMemset(p, 0, jb::XCH_ARENA_OFFSET);
        cudaDeviceSynchronize();
        c->xch_peer[c->rank] = (uint64_t*)p;
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->xch_peer[c->rank]);
    if (e != cudaSuccess) return c->check(e, "cudaIpcGetMemHandle");
    std::memcpy(out, &h, 64);
    return JB_OK;
}

int jb_comm_p2p_open(jb_ctx* c, const uint8_t* handles) {
    if (!c) return JB_ERR_INVALID;
    if (!handles) {  // NULL: switch the exchange off (the NCCL all-reduce path stays)
        std::lock_guard<std::mutex> lk(c->mu);
        c->xch_ready = false;
        return JB_OK;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (c->world > 16) return c->fail(JB_ERR_UNSUPPORTED, "p2p: at most 16 ranks");
    if (!c->xch_peer[c->rank]) return c->fail(JB_ERR_INVALID, "p2p: call jb_comm_p2p_handle first");
    for (int g = 0; g < c->world; ++g) {
        if (g == c->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * g, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return c->check(e, "cudaIpcOpenMemHandle (peer exchange unavailable; NCCL path stays)");
        c->xch_peer[g] = (uint64_t*)p;
    }
    c->xch_ready = true;
    c->xch_seq = 0;
    return JB_OK;
}

int jb_comm_destroy(jb_ctx* c) {
    if (!c) return JB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->nccl_comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        const NcclApi* a = nccl_api(nullptr, nullptr);
        if (a) a->CommDestroy((ncclComm_t)c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    if (c->world <= 16 && c->rank < 16 && (c->xch_ready || c->xch_peer[c->rank])) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        for (int g = 0; g < 16; ++g) {
            if (!c->xch_peer[g]) continue;
            if (g == c->rank) cudaFree(c->xch_peer[g]);
            else cudaIpcCloseMemHandle(c->xch_peer[g]);
            c->xch_peer[g] = nullptr;
        }
        c->xch_ready = false;
    }
    c->world = 1;
    c->rank = 0;
    return JB_OK;
}

}  // extern "C"
