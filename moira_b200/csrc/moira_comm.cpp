// moira_comm.cpp -- the path's only collective, inside the library: the sum over GPUs of the MOIRA_N_COUNTERS
// uint64 counters (good / bad counts + floor(ee) histogram), the device-side counterpart of the reference's scalar
// accumulators and final report (moira.py:406-408, 483-485, 509-519; SURVEY.md 8b "moira_reduce_counters", 8e).
// NCCL over NVLink 5 / NVSwitch, for hosts that have no torch.distributed: one communicator rank per context, either
// one process per GPU (moira_comm_unique_id + moira_comm_init) or one process driving several GPUs
// (moira_comm_init_all, the CLI's --devices).  Read data never crosses GPUs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a host that never reduces does not need it, and a process that
// already carries a copy (torch bundles one) shares it instead of loading a second.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <new>
#include <vector>

#include "moira_internal.h"

using namespace moira;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.GroupStart &&
                 api.GroupEnd && api.GetErrorString;
    });
    return api;
}

int need_nccl()
{
    if (!nccl().ok) return fail(MOIRA_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    return MOIRA_OK;
}

#define NC(call)                                                                                             \
    do {                                                                                                     \
        ncclResult_t r_ = (call);                                                                            \
        if (r_ != ncclSuccess) return fail(MOIRA_ERR_NCCL, "%s failed: %s", #call, nccl().GetErrorString(r_)); \
    } while (0)
#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(MOIRA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

}  // namespace

// the communicator state of a context lives here, keyed by the context's comm slot (moira_api.cu owns moira_ctx)
namespace moira {
struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    int device = 0;
    cudaStream_t stream = nullptr;          // collectives run here, behind an event of the caller's stream
    cudaEvent_t ev = nullptr;
    uint64_t *d_buf = nullptr;              // MOIRA_N_COUNTERS staging on the device
    uint64_t *h_buf = nullptr;              // pinned
};

int comm_state_init(CommState *s, int device)
{
    s->device = device;
    CU(cudaSetDevice(device));
    if (!s->stream) CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    if (!s->ev) CU(cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming));
    if (!s->d_buf) CU(cudaMalloc(&s->d_buf, MOIRA_N_COUNTERS * sizeof(uint64_t)));
    if (!s->h_buf) CU(cudaHostAlloc((void **)&s->h_buf, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaHostAllocDefault));
    return MOIRA_OK;
}

CommState *comm_state_new() { return new (std::nothrow) CommState(); }

int comm_info(const CommState *s, int *rank, int *n_ranks)
{
    if (rank) *rank = s->comm ? s->rank : 0;
    if (n_ranks) *n_ranks = s->comm ? s->n_ranks : 1;
    return MOIRA_OK;
}

static void comm_state_destroy(CommState *s);
void comm_state_free(CommState *s)
{
    if (!s) return;
    comm_state_destroy(s);
    delete s;
}

static void comm_state_destroy(CommState *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->comm && nccl().ok) nccl().CommDestroy(s->comm);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->d_buf) cudaFree(s->d_buf);
    if (s->h_buf) cudaFreeHost(s->h_buf);
    *s = CommState();
}

int comm_unique_id(uint8_t id[MOIRA_COMM_ID_BYTES])
{
    int rc = need_nccl();
    if (rc) return rc;
    static_assert(MOIRA_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    NC(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return MOIRA_OK;
}

int comm_init_rank(CommState *s, int device, const uint8_t id[MOIRA_COMM_ID_BYTES], int rank, int n_ranks)
{
    int rc = need_nccl();
    if (rc) return rc;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MOIRA_ERR_BAD_ARG, "bad rank %d of %d", rank, n_ranks);
    if (s->comm) { nccl().CommDestroy(s->comm); s->comm = nullptr; }
    if ((rc = comm_state_init(s, device))) return rc;
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    NC(nccl().CommInitRank(&s->comm, n_ranks, u, rank));
    s->rank = rank;
    s->n_ranks = n_ranks;
    return MOIRA_OK;
}

int comm_init_all(CommState **states, const int *devices, int n)
{
    int rc = need_nccl();
    if (rc) return rc;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return fail(MOIRA_ERR_BAD_ARG, "contexts %d and %d share device %d: one communicator rank per GPU", j, i, devices[i]);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; i++) {
        if (states[i]->comm) { nccl().CommDestroy(states[i]->comm); states[i]->comm = nullptr; }
        if ((rc = comm_state_init(states[i], devices[i]))) return rc;
    }
    NC(nccl().CommInitAll(comms.data(), n, devices));
    for (int i = 0; i < n; i++) { states[i]->comm = comms[i]; states[i]->rank = i; states[i]->n_ranks = n; }
    return MOIRA_OK;
}

// all-reduce (sum) of d_counters in place, ordered behind the work already enqueued on `stream`; the caller's stream
// waits for the result (so the next thing it enqueues sees the sums)
int comm_reduce_device(CommState *s, uint64_t *d_counters, cudaStream_t stream)
{
    if (!s->comm) return fail(MOIRA_ERR_BAD_ARG, "context has no communicator (moira_comm_init / moira_comm_init_all)");
    CU(cudaSetDevice(s->device));
    NC(nccl().AllReduce(d_counters, d_counters, MOIRA_N_COUNTERS, ncclUint64, ncclSum, s->comm, stream));
    return MOIRA_OK;
}

// host counters in / out, blocking
int comm_reduce_host(CommState *s, uint64_t *counters)
{
    if (!s->comm) return fail(MOIRA_ERR_BAD_ARG, "context has no communicator (moira_comm_init / moira_comm_init_all)");
    CU(cudaSetDevice(s->device));
    memcpy(s->h_buf, counters, MOIRA_N_COUNTERS * sizeof(uint64_t));
    CU(cudaMemcpyAsync(s->d_buf, s->h_buf, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    NC(nccl().AllReduce(s->d_buf, s->d_buf, MOIRA_N_COUNTERS, ncclUint64, ncclSum, s->comm, s->stream));
    CU(cudaMemcpyAsync(s->h_buf, s->d_buf, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(counters, s->h_buf, MOIRA_N_COUNTERS * sizeof(uint64_t));
    return MOIRA_OK;
}

// one process, n contexts: every context's host counters become the sum (one NCCL group call)
int comm_reduce_all(CommState **states, int n, uint64_t *const *counters)
{
    int rc = need_nccl();
    if (rc) return rc;
    for (int i = 0; i < n; i++) {
        CommState *s = states[i];
        if (!s->comm) return fail(MOIRA_ERR_BAD_ARG, "context %d has no communicator (moira_comm_init_all)", i);
        CU(cudaSetDevice(s->device));
        memcpy(s->h_buf, counters[i], MOIRA_N_COUNTERS * sizeof(uint64_t));
        CU(cudaMemcpyAsync(s->d_buf, s->h_buf, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    }
    NC(nccl().GroupStart());
    for (int i = 0; i < n; i++) {
        CommState *s = states[i];
        ncclResult_t r = nccl().AllReduce(s->d_buf, s->d_buf, MOIRA_N_COUNTERS, ncclUint64, ncclSum, s->comm, s->stream);
        if (r != ncclSuccess) { nccl().GroupEnd(); return fail(MOIRA_ERR_NCCL, "ncclAllReduce failed: %s", nccl().GetErrorString(r)); }
    }
    NC(nccl().GroupEnd());
    for (int i = 0; i < n; i++) {
        CommState *s = states[i];
        CU(cudaSetDevice(s->device));
        CU(cudaMemcpyAsync(s->h_buf, s->d_buf, MOIRA_N_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    }
    for (int i = 0; i < n; i++) {
        CommState *s = states[i];
        CU(cudaSetDevice(s->device));
        CU(cudaStreamSynchronize(s->stream));
        memcpy(counters[i], s->h_buf, MOIRA_N_COUNTERS * sizeof(uint64_t));
    }
    return MOIRA_OK;
}

}  // namespace moira
