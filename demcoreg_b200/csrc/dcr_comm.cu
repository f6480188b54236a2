// Row-band mode, the collective side: an NCCL communicator owned by (or handed to) the library, and the whole row-band
// iteration as ONE native call -- the phases of dcr_band_phase with their ncclAllReduce's enqueued back to back on the
// caller's stream, no interpreter (and no host synchronisation) between them.  (dcr_set_comm; 8e.)
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch has already
// loaded), so the library has no link-time dependency on it and pair-sharded users never touch it.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <nccl.h>
#include "dcr_common.cuh"
#include "dcr_internal.cuh"

struct NcclApi {
    void* lib;
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclAllReduce) AllReduce;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclGetErrorString) GetErrorString;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;   // 0 untried, 1 ok, -1 failed   (first call happens before any threads race: comm creation)
    if (state == 1) return &api;
    if (state == -1) return nullptr;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { dcr_set_error("libnccl.so.2 not found: %s", dlerror()); state = -1; return nullptr; }
    api.lib = h;
#define DCR_NCCL_SYM(name) \
    api.name = (decltype(api.name))dlsym(h, "nccl" #name); \
    if (!api.name) { dcr_set_error("nccl" #name " missing from libnccl"); state = -1; return nullptr; }
    DCR_NCCL_SYM(GetUniqueId) DCR_NCCL_SYM(CommInitRank) DCR_NCCL_SYM(CommDestroy) DCR_NCCL_SYM(AllReduce)
    DCR_NCCL_SYM(Send) DCR_NCCL_SYM(Recv) DCR_NCCL_SYM(GroupStart) DCR_NCCL_SYM(GroupEnd) DCR_NCCL_SYM(GetErrorString)
#undef DCR_NCCL_SYM
    state = 1;
    return &api;
}

struct dcr_comm {
    ncclComm_t comm;
    int rank, nranks;
    int owned;   // created by dcr_comm_init (destroyed with the handle) vs injected by dcr_comm_set
};

#define DCR_NCCL_OK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    dcr_set_error("%s: %s", #call, api->GetErrorString(r_)); return 100 + (int)r_; } } while (0)

extern "C" int dcr_comm_unique_id(void* id, size_t id_bytes) {
    DCR_ARG(id && id_bytes >= sizeof(ncclUniqueId), "id buffer must hold 128 bytes");
    NcclApi* api = nccl_api();
    if (!api) return -7;
    ncclUniqueId u;
    DCR_NCCL_OK(api->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return 0;
}

extern "C" int dcr_comm_init(const void* id, size_t id_bytes, int rank, int nranks, dcr_comm** out) {
    DCR_ARG(id && out && id_bytes >= sizeof(ncclUniqueId), "null pointer / short id");
    DCR_ARG(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "1 <= nranks <= 8");
    NcclApi* api = nccl_api();
    if (!api) return -7;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c;
    DCR_NCCL_OK(api->CommInitRank(&c, nranks, u, rank));
    dcr_comm* h = new dcr_comm;
    h->comm = c; h->rank = rank; h->nranks = nranks; h->owned = 1;
    *out = h;
    return 0;
}

extern "C" int dcr_comm_set(void* nccl_comm, int rank, int nranks, dcr_comm** out) {
    DCR_ARG(nccl_comm && out, "null pointer");
    DCR_ARG(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "1 <= nranks <= 8");
    if (!nccl_api()) return -7;
    dcr_comm* h = new dcr_comm;
    h->comm = (ncclComm_t)nccl_comm; h->rank = rank; h->nranks = nranks; h->owned = 0;
    *out = h;
    return 0;
}

extern "C" int dcr_comm_destroy(dcr_comm* h) {
    if (!h) return 0;
    NcclApi* api = nccl_api();
    if (api && h->owned) api->CommDestroy(h->comm);
    delete h;
    return 0;
}

extern "C" int dcr_comm_rank(const dcr_comm* h, int* rank, int* nranks) {
    DCR_ARG(h && rank && nranks, "null pointer");
    *rank = h->rank; *nranks = h->nranks;
    return 0;
}

// Halo rows of one band raster: rows [send_up, send_up + n_up) go to rank - 1, [send_dn, send_dn + n_dn) to rank + 1;
// [recv_up, ...) / [recv_dn, ...) arrive from them.  One grouped send/recv per call (bytes: rows * row_bytes).
extern "C" int dcr_band_exchange_halos(dcr_comm* h, void* band, int64_t row_bytes, int send_up_row, int n_send_up,
                                       int recv_up_row, int n_recv_up, int send_dn_row, int n_send_dn, int recv_dn_row,
                                       int n_recv_dn, void* stream) {
    DCR_ARG(h && band && row_bytes > 0, "null pointer");
    NcclApi* api = nccl_api();
    if (!api) return -7;
    char* b = (char*)band;
    cudaStream_t s = (cudaStream_t)stream;
    DCR_NCCL_OK(api->GroupStart());
    if (h->rank > 0) {
        if (n_recv_up > 0) DCR_NCCL_OK(api->Recv(b + (int64_t)recv_up_row * row_bytes, (size_t)n_recv_up * row_bytes, ncclUint8, h->rank - 1, h->comm, s));
        if (n_send_up > 0) DCR_NCCL_OK(api->Send(b + (int64_t)send_up_row * row_bytes, (size_t)n_send_up * row_bytes, ncclUint8, h->rank - 1, h->comm, s));
    }
    if (h->rank < h->nranks - 1) {
        if (n_recv_dn > 0) DCR_NCCL_OK(api->Recv(b + (int64_t)recv_dn_row * row_bytes, (size_t)n_recv_dn * row_bytes, ncclUint8, h->rank + 1, h->comm, s));
        if (n_send_dn > 0) DCR_NCCL_OK(api->Send(b + (int64_t)send_dn_row * row_bytes, (size_t)n_send_dn * row_bytes, ncclUint8, h->rank + 1, h->comm, s));
    }
    DCR_NCCL_OK(api->GroupEnd());
    return 0;
}

// One whole row-band iteration: every phase of dcr_band_phase followed by the all-reduce (sum) of the buffers it declares,
// all on `stream`; then the result block (identical on every rank) to the host.  Collective: every rank of the
// communicator calls it with its own band.
extern "C" int dcr_band_iteration(const dcr_front_args* front, int flags, dcr_comm* h, dcr_iter_result* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    DCR_ARG(front && h && out && workspace, "null pointer");
    NcclApi* api = nccl_api();
    if (!api) return -7;
    cudaStream_t s = (cudaStream_t)stream;
    static const ncclDataType_t kind_type[3] = {ncclUint32, ncclInt64, ncclFloat64};
    const int nph = dcr_band_nphases();
    int f = flags;
    for (int attempt = 0; attempt < 3; ++attempt) {
        for (int ph = 0; ph < nph; ++ph) {
            void* ptrs[8];
            int64_t counts[8];
            int kinds[8], nred = 0;
            int rc = dcr_band_phase(front, f, h->rank, h->nranks, ph, workspace, workspace_bytes, ptrs, counts, kinds, &nred, stream);
            if (rc) return rc;
            if (h->nranks > 1 && nred > 0) {
                if (nred > 1) DCR_NCCL_OK(api->GroupStart());
                for (int k = 0; k < nred; ++k)
                    DCR_NCCL_OK(api->AllReduce(ptrs[k], ptrs[k], (size_t)counts[k], kind_type[kinds[k]], ncclSum, h->comm, s));
                if (nred > 1) DCR_NCCL_OK(api->GroupEnd());
            }
        }
        int rc = dcr_band_result(front, workspace, workspace_bytes, out, stream);
        if (rc) return rc;
        // a declined one-pass selection / fast bin path (seen by every rank alike): once more with the radix engine for it
        if (out->status != DCR_STATUS_RERUN_RADIX) break;
        f |= out->select_engine ? DCR_ITER_RADIX_BINS : DCR_ITER_RADIX_ONLY;
    }
    return 0;
}
