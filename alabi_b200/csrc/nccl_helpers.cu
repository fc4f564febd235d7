// NCCL helpers behind the C ABI ("NCCL init / bcast / allgather helpers"; SURVEY 8e:
// train once on one GPU, broadcast L and alpha over NVLink, all_gather the acquisition argmin and
// the chain blocks).  The reference has no such layer (its parallelism is multiprocessing pools,
// alabi/core.py:309-314); these entry points let a binding move the factor state between the
// handles of different processes without staging copies and without torch.distributed.
//
// NCCL is resolved at RUN time (dlopen of libnccl.so.2: the copy already loaded into the process —
// e.g. PyTorch's — or the system's), so libalabi_b200.so itself has no link-time dependency on it
// and loads on machines without NCCL; the helpers then return -6.
#include <dlfcn.h>
#include <string.h>
#include "handle.h"
#include "alabi_b200.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { kNcclUint8 = 1, kNcclSuccess = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(a.lib, "ncclBroadcast"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.AllGather;
    return a;
}

int nccl_fail(const char* what, int rc) {
    NcclApi& a = api();
    ab_set_error("%s: NCCL error %d (%s)", what, rc, a.GetErrorString ? a.GetErrorString(rc) : "?");
    return -7;
}

#define AB_NCCL_READY()                                                                                \
    NcclApi& A = api();                                                                                \
    if (!A.ok) { ab_set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return -6; }

}  // namespace

struct ab_comm {
    NcclComm comm = nullptr;
    int world = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
};

extern "C" int ab_nccl_unique_id(unsigned char* h_id) {
    if (!h_id) { ab_set_error("null argument"); return -1; }
    AB_NCCL_READY();
    NcclUniqueId id;
    int rc = A.GetUniqueId(&id);
    if (rc != kNcclSuccess) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(h_id, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" int ab_nccl_init(ab_comm** out, int world, int rank, const unsigned char* h_id, int device, void* cuda_stream) {
    if (!out || !h_id || world < 1 || rank < 0 || rank >= world) { ab_set_error("ab_nccl_init: bad argument"); return -1; }
    AB_NCCL_READY();
    AB_CUDA(cudaSetDevice(device));
    NcclUniqueId id;
    memcpy(id.internal, h_id, sizeof(id.internal));
    ab_comm* c = new ab_comm();
    c->world = world; c->rank = rank; c->device = device; c->stream = (cudaStream_t)cuda_stream;
    int rc = A.CommInitRank(&c->comm, world, id, rank);
    if (rc != kNcclSuccess) { delete c; return nccl_fail("ncclCommInitRank", rc); }
    *out = c;
    return 0;
}

extern "C" int ab_nccl_destroy(ab_comm* c) {
    if (!c) return 0;
    NcclApi& A = api();
    if (A.ok && c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        A.CommDestroy(c->comm);
    }
    delete c;
    return 0;
}

// d_buf (nbytes) of rank `root` -> the same buffer on every rank; asynchronous on the communicator's stream
extern "C" int ab_nccl_broadcast(ab_comm* c, void* d_buf, int64_t nbytes, int root) {
    if (!c || !d_buf || nbytes < 0 || root < 0 || root >= c->world) { ab_set_error("ab_nccl_broadcast: bad argument"); return -1; }
    AB_NCCL_READY();
    AB_CUDA(cudaSetDevice(c->device));
    int rc = A.Broadcast(d_buf, d_buf, (size_t)nbytes, kNcclUint8, root, c->comm, c->stream);
    if (rc != kNcclSuccess) return nccl_fail("ncclBroadcast", rc);
    return 0;
}

// every rank contributes nbytes_per_rank; d_recv (world x nbytes_per_rank) holds them in rank order
extern "C" int ab_nccl_allgather(ab_comm* c, const void* d_send, void* d_recv, int64_t nbytes_per_rank) {
    if (!c || !d_send || !d_recv || nbytes_per_rank < 0) { ab_set_error("ab_nccl_allgather: bad argument"); return -1; }
    AB_NCCL_READY();
    AB_CUDA(cudaSetDevice(c->device));
    int rc = A.AllGather(d_send, d_recv, (size_t)nbytes_per_rank, kNcclUint8, c->comm, c->stream);
    if (rc != kNcclSuccess) return nccl_fail("ncclAllGather", rc);
    return 0;
}

// Factor state of the trained handle of rank `root` -> the handles of all ranks, IN PLACE (no staging
// copies): L, the diagonal-block inverses (so every replica derives L^-1 with identical bits), the
// log-determinant parts and alpha.  Every rank must have called ab_gp_set_inputs / ab_gp_set_kernel
// with the same inputs and hyper-parameters (they travel on the host side).  Synchronises the
// communicator's stream with the handle's on both ends.
extern "C" int ab_nccl_broadcast_gp(ab_comm* c, ab_gp* h, int root) {
    if (!c || !h) { ab_set_error("ab_nccl_broadcast_gp: null argument"); return -1; }
    AB_NCCL_READY();
    if (!h->have_inputs || !h->have_kernel) { ab_set_error("ab_nccl_broadcast_gp: inputs and kernel must be set on every rank"); return -2; }
    if (c->rank == root && !(h->factored && h->have_alpha)) {
        ab_set_error("ab_nccl_broadcast_gp: the root handle is not trained (ab_gp_factor + ab_gp_set_targets)");
        return -2;
    }
    AB_CUDA(cudaSetDevice(h->device));
    if (!h->scaled) {
        int rc = ab_launch_scale_inputs(h);
        if (rc) return rc;
        h->scaled = true;
    }
    AB_CUDA(cudaStreamSynchronize(h->stream));                 // the root's factor / this rank's scaling is complete
    const size_t np = (size_t)h->npad;
    struct { void* p; size_t bytes; } parts[] = {
        {h->L, np * np * sizeof(double)}, {h->Dinv, np * AB_NB * sizeof(double)},
        {h->logdet_parts, np / AB_NB * sizeof(double)}, {h->alpha, np * sizeof(double)}};
    for (auto& part : parts) {
        int rc = A.Broadcast(part.p, part.p, part.bytes, kNcclUint8, root, c->comm, c->stream);
        if (rc != kNcclSuccess) return nccl_fail("ncclBroadcast", rc);
    }
    AB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->rank != root) {
        h->factored = true;
        h->have_alpha = true;
        h->have_linv = h->have_kinv = false;
        h->info = 0;
    }
    return 0;
}

extern "C" int ab_nccl_sync(ab_comm* c) {
    if (!c) { ab_set_error("null argument"); return -1; }
    AB_CUDA(cudaSetDevice(c->device));
    AB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- peer memory (CUDA IPC) for the fused chain all_gather of the sampler ----------------------
extern "C" int ab_peer_alloc(int device, size_t bytes, void** d_ptr, unsigned char* h_handle) {
    if (!d_ptr || !h_handle || bytes == 0) { ab_set_error("ab_peer_alloc: bad argument"); return -1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == AB_PEER_HANDLE_BYTES, "handle size");
    AB_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    AB_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t hnd;
    cudaError_t e = cudaIpcGetMemHandle(&hnd, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        ab_set_error("ab_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return -100 - (int)e;
    }
    memcpy(h_handle, &hnd, sizeof(hnd));
    *d_ptr = p;
    return 0;
}

extern "C" int ab_peer_open(int device, const unsigned char* h_handle, void** d_ptr) {
    if (!d_ptr || !h_handle) { ab_set_error("ab_peer_open: bad argument"); return -1; }
    AB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, h_handle, sizeof(hnd));
    void* p = nullptr;
    AB_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return 0;
}

extern "C" int ab_peer_close(int device, void* d_ptr) {
    if (!d_ptr) return 0;
    AB_CUDA(cudaSetDevice(device));
    AB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

extern "C" int ab_peer_free(int device, void* d_ptr) {
    if (!d_ptr) return 0;
    AB_CUDA(cudaSetDevice(device));
    AB_CUDA(cudaFree(d_ptr));
    return 0;
}
