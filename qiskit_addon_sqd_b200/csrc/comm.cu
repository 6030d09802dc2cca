// NCCL exchange for ONE large diagonalisation sharded over GPUs (BASELINE.json config 5).
// Every rank holds the full Davidson vectors (replicated, bit-identical arithmetic) and builds only its
// block of rows of sigma; the blocks are combined by an all-reduce (sum, FP64) over NVLink.  NCCL is
// resolved at run time from the copy torch has already loaded (dlopen by soname), so the library has
// no link-time dependency on it and still loads on a machine without NCCL.
#include <dlfcn.h>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSum = 0, kNcclDouble = 8 };

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a{};
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
            a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
            a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
            a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
            a.Broadcast = (decltype(a.Broadcast))dlsym(h, "ncclBroadcast");
            a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
            a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString &&
                   a.Broadcast && a.GroupStart && a.GroupEnd;
        }
        return a;
    }();
    return api;
}

int nccl_allreduce_sum_f64(void* comm, double* buf, int64_t n, cudaStream_t st) {
    NcclApi& a = nccl();
    SQD_REQUIRE(a.ok, "NCCL is not available in this process");
    const ncclResult_t r = a.AllReduce(buf, buf, (size_t)n, kNcclDouble, kNcclSum, (ncclComm_t)comm, st);
    SQD_REQUIRE(r == 0, "ncclAllReduce failed: %s", a.GetErrorString(r));
    return 0;
}

// Exchange of DISJOINT blocks: rank r has written elements [offs[r], offs[r+1]) of buf; afterwards every rank
// holds all of buf.  One grouped call of `world` broadcasts (the blocks have unequal sizes, which rules out
// ncclAllGather): each rank receives (world-1)/world of the vector -- half of what the all-reduce of
// zero-padded full vectors moved (round 1) -- and nothing is added, so the result is the unsharded sigma
// bit for bit by construction.
int nccl_allgather_blocks(void* comm, double* buf, const long long* offs, int world, cudaStream_t st) {
    NcclApi& a = nccl();
    SQD_REQUIRE(a.ok, "NCCL is not available in this process");
    ncclResult_t r = a.GroupStart();
    SQD_REQUIRE(r == 0, "ncclGroupStart failed: %s", a.GetErrorString(r));
    for (int root = 0; root < world; ++root) {
        const long long cnt = offs[root + 1] - offs[root];
        if (cnt <= 0) continue;
        r = a.Broadcast(buf + offs[root], buf + offs[root], (size_t)cnt, kNcclDouble, root, (ncclComm_t)comm, st);
        SQD_REQUIRE(r == 0, "ncclBroadcast failed: %s", a.GetErrorString(r));
    }
    r = a.GroupEnd();
    SQD_REQUIRE(r == 0, "ncclGroupEnd failed: %s", a.GetErrorString(r));
    return 0;
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_nccl_unique_id(char* h_id128) {
    NcclApi& a = nccl();
    SQD_REQUIRE(a.ok, "sqd_nccl_unique_id: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    SQD_REQUIRE(r == 0, "ncclGetUniqueId failed: %s", a.GetErrorString(r));
    memcpy(h_id128, id.internal, 128);
    return 0;
}

int sqd_nccl_init(const char* h_id128, int rank, int world, void** comm_out) {
    NcclApi& a = nccl();
    SQD_REQUIRE(a.ok, "sqd_nccl_init: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(id.internal, h_id128, 128);
    ncclComm_t c = nullptr;
    const ncclResult_t r = a.CommInitRank(&c, world, id, rank);
    SQD_REQUIRE(r == 0, "ncclCommInitRank failed: %s", a.GetErrorString(r));
    *comm_out = (void*)c;
    return 0;
}

int sqd_nccl_destroy(void* comm) {
    NcclApi& a = nccl();
    if (a.ok && comm) a.CommDestroy((ncclComm_t)comm);
    return 0;
}

int sqd_allreduce_sum_f64(void* comm, double* d_buf, int64_t n, void* stream) {
    return nccl_allreduce_sum_f64(comm, d_buf, n, (cudaStream_t)stream);
}

}  // extern "C"
