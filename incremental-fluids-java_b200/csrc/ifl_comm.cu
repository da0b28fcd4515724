/*
 * ifl_comm.cu -- slab-decomposed (multi-GPU) plumbing: NCCL (resolved at run time with dlopen, so the
 * library has no link-time NCCL dependency and shares the copy a host process such as torch already
 * loaded) and CUDA IPC mappings of the neighbours' strip-boundary inboxes.
 *
 * What crosses NVLink per PCG iteration:
 *   - the strip hand-over of the two triangular sweeps: the last strip of rank r stores its edge row with plain
 *     8-byte stores straight into the flag-in-data inbox of rank r+1 (peer memory, IPC-mapped);
 *   - one boundary row of z and of s in each direction (ncclSend/ncclRecv) for the fused s-update + mat-vec;
 *   - per-strip partial sums of the two dot products (grouped ncclBroadcast of every rank's strip range) and the
 *     per-rank max (ncclAllGather), summed by every rank in
 *     global strip order -- so the result does not depend on the number of GPUs.
 */
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <nccl.h>
#include "ifl_internal.h"

namespace ifl {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static const char* nccl_load() {
    if (g_nccl.lib) return nullptr;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 could not be loaded";
#define IFL_SYM(field, name) \
    *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return "NCCL symbol missing: " name;
    IFL_SYM(GetUniqueId, "ncclGetUniqueId")
    IFL_SYM(CommInitRank, "ncclCommInitRank")
    IFL_SYM(CommDestroy, "ncclCommDestroy")
    IFL_SYM(AllGather, "ncclAllGather")
    IFL_SYM(Broadcast, "ncclBroadcast")
    IFL_SYM(Send, "ncclSend")
    IFL_SYM(Recv, "ncclRecv")
    IFL_SYM(GroupStart, "ncclGroupStart")
    IFL_SYM(GroupEnd, "ncclGroupEnd")
    IFL_SYM(GetErrorString, "ncclGetErrorString")
#undef IFL_SYM
    g_nccl.lib = lib;
    return nullptr;
}

static thread_local char g_comm_err[256] = "";
const char* comm_last_error() { return g_comm_err; }
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    snprintf(g_comm_err, sizeof g_comm_err, "NCCL error: %s at %s", g_nccl.GetErrorString(r_), #call); return -1; } } while (0)
#define CC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_comm_err, sizeof g_comm_err, "CUDA error: %s at %s", cudaGetErrorString(e_), #call); return -1; } } while (0)

int comm_unique_id(unsigned char out[128]) {
    const char* e = nccl_load();
    if (e) { snprintf(g_comm_err, sizeof g_comm_err, "%s", e); return -1; }
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

int comm_init(Comm& c, int rank, int world, const unsigned char id_bytes[128], cudaStream_t st) {
    const char* e = nccl_load();
    if (e) { snprintf(g_comm_err, sizeof g_comm_err, "%s", e); return -1; }
    c.rank = rank; c.world = world; c.stream = st;
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    ncclComm_t comm;
    NC(g_nccl.CommInitRank(&comm, world, id, rank));
    c.nccl = comm;
    return 0;
}
void comm_destroy(Comm& c) {
    for (void* p : {c.peer_prev_base, c.peer_next_base}) if (p) cudaIpcCloseMemHandle(p);
    c.peer_prev_base = c.peer_next_base = nullptr;
    if (c.nccl) g_nccl.CommDestroy((ncclComm_t)c.nccl);
    c.nccl = nullptr;
}

/* Exchange the CUDA IPC handle of `inbox_base` (one allocation holding bnd_f and bnd_b) with every rank and map the
 * allocations of rank-1 and rank+1. */
int comm_map_neighbour_inboxes(Comm& c, void* inbox_base, void* scratch_dev /* >= world*64 bytes */) {
    cudaIpcMemHandle_t mine;
    CC(cudaIpcGetMemHandle(&mine, inbox_base));
    unsigned char* dev = (unsigned char*)scratch_dev;
    CC(cudaMemcpyAsync(dev + (size_t)c.rank * 64, &mine, 64, cudaMemcpyHostToDevice, c.stream));
    NC(g_nccl.AllGather(dev + (size_t)c.rank * 64, dev, 64, ncclUint8, (ncclComm_t)c.nccl, c.stream));
    cudaIpcMemHandle_t all[64];
    if (c.world > 64) { snprintf(g_comm_err, sizeof g_comm_err, "world_size > 64"); return -1; }
    CC(cudaMemcpyAsync(all, dev, (size_t)c.world * 64, cudaMemcpyDeviceToHost, c.stream));
    CC(cudaStreamSynchronize(c.stream));
    if (c.rank > 0) CC(cudaIpcOpenMemHandle(&c.peer_prev_base, all[c.rank - 1], cudaIpcMemLazyEnablePeerAccess));
    if (c.rank < c.world - 1) CC(cudaIpcOpenMemHandle(&c.peer_next_base, all[c.rank + 1], cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int comm_all_gather(Comm& c, const void* send, void* recv, size_t bytes_per_rank) {
    NC(g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)c.nccl, c.stream));
    return 0;
}
/* Per-strip values (partial sums of a dot product): every rank broadcasts the range it owns, in place, as one grouped
 * operation.  The ranges of the balanced partition differ by at most one strip, so an all-gather does not fit. */
int comm_all_gather_strips(Comm& c, double* strip_vals, int nstrips, size_t per_strip) {
    NC(g_nccl.GroupStart());
    for (int r = 0; r < c.world; r++) {
        int s0, s1;
        slab_range(nstrips, c.world, r, &s0, &s1);
        NC(g_nccl.Broadcast(strip_vals + (size_t)s0 * per_strip, strip_vals + (size_t)s0 * per_strip, (size_t)(s1 - s0) * per_strip, ncclFloat64, r,
                             (ncclComm_t)c.nccl, c.stream));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}
/* exchange one packed row with each neighbour: send_up -> rank-1, send_down -> rank+1 */
int comm_exchange_rows(Comm& c, const double* send_up, double* recv_up, const double* send_down, double* recv_down, size_t n) {
    NC(g_nccl.GroupStart());
    if (c.rank > 0) {
        NC(g_nccl.Send(send_up, n, ncclFloat64, c.rank - 1, (ncclComm_t)c.nccl, c.stream));
        NC(g_nccl.Recv(recv_up, n, ncclFloat64, c.rank - 1, (ncclComm_t)c.nccl, c.stream));
    }
    if (c.rank < c.world - 1) {
        NC(g_nccl.Send(send_down, n, ncclFloat64, c.rank + 1, (ncclComm_t)c.nccl, c.stream));
        NC(g_nccl.Recv(recv_down, n, ncclFloat64, c.rank + 1, (ncclComm_t)c.nccl, c.stream));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}

} // namespace ifl

namespace ifl {
int comm_transfers(Comm& c, const CommXfer* x, int n) {
    if (n <= 0) return 0;
    NC(g_nccl.GroupStart());
    for (int i = 0; i < n; i++) {
        if (x[i].count == 0) continue;
        if (x[i].send) NC(g_nccl.Send(x[i].ptr, x[i].count, ncclFloat64, x[i].peer, (ncclComm_t)c.nccl, c.stream));
        else           NC(g_nccl.Recv(x[i].ptr, x[i].count, ncclFloat64, x[i].peer, (ncclComm_t)c.nccl, c.stream));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}
int comm_all_gather_u64(Comm& c, unsigned long long* vals) {
    NC(g_nccl.AllGather(vals + c.rank, vals, 1, ncclUint64, (ncclComm_t)c.nccl, c.stream));
    return 0;
}
} // namespace ifl
