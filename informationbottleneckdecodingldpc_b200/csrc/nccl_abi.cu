// nccl_abi.cu -- the one collective of the path, exported through the C ABI:
// an in-place ncclAllReduce(sum) over a short int64 vector {bit errors, frame errors, frames, iterations}, so that
// every rank takes the same `while errors < min_errors` decision as the single-device reference loop
// (Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:98).  Codewords are independent: there is no collective
// inside a decode.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy a host process such as PyTorch has already mapped,
// or the path in IBLDPC_NCCL_LIB), so libibldpc.so itself has no link-time dependency on it and non-Python hosts
// can all-reduce their counters through the same handle.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "ibldpc_internal.h"

using ibldpc::fail_msg;

namespace {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
using NcclComm = void*;
constexpr int kNcclInt64 = 4;   // ncclInt64
constexpr int kNcclSum = 0;     // ncclSum

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

NcclApi& api()
{
    static NcclApi a = [] {
        NcclApi x;
        const char* names[] = {getenv("IBLDPC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            x.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (x.lib) break;
            x.why = dlerror();
        }
        if (!x.lib) return x;
        x.GetUniqueId = (decltype(x.GetUniqueId))dlsym(x.lib, "ncclGetUniqueId");
        x.CommInitRank = (decltype(x.CommInitRank))dlsym(x.lib, "ncclCommInitRank");
        x.AllReduce = (decltype(x.AllReduce))dlsym(x.lib, "ncclAllReduce");
        x.CommDestroy = (decltype(x.CommDestroy))dlsym(x.lib, "ncclCommDestroy");
        x.GetErrorString = (decltype(x.GetErrorString))dlsym(x.lib, "ncclGetErrorString");
        if (!x.GetUniqueId || !x.CommInitRank || !x.AllReduce || !x.CommDestroy) {
            x.why = "libnccl lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy";
            x.lib = nullptr;
        }
        return x;
    }();
    return a;
}

int nccl_fail(const char* what, int rc)
{
    NcclApi& a = api();
    return fail_msg(IBLDPC_E_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error ") +
                                       " (" + std::to_string(rc) + ")");
}

int need_api()
{
    if (!api().lib) return fail_msg(IBLDPC_E_STATE, "NCCL is not available: " + api().why + " (set IBLDPC_NCCL_LIB to libnccl.so.2)");
    return IBLDPC_OK;
}

}  // namespace

extern "C" {

int ibldpc_nccl_unique_id(uint8_t* id128)
{
    if (!id128) return fail_msg(IBLDPC_E_INVALID, "null argument");
    if (int rc = need_api()) return rc;
    NcclUniqueId id;
    const int rc = api().GetUniqueId(&id);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128, id.internal, 128);
    return IBLDPC_OK;
}

int ibldpc_nccl_init(ibldpc_handle h, const uint8_t* id128, int rank, int world)
{
    if (!h || !id128) return fail_msg(IBLDPC_E_INVALID, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail_msg(IBLDPC_E_INVALID, "need 0 <= rank < world");
    if (int rc = need_api()) return rc;
    ibldpc::DeviceGuard guard_(h->device);
    IBLDPC_CK(guard_.err);
    if (h->nccl_comm) {
        api().CommDestroy((NcclComm)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    NcclComm comm = nullptr;
    const int rc = api().CommInitRank(&comm, world, id, rank);
    if (rc) return nccl_fail("ncclCommInitRank", rc);
    h->nccl_comm = comm;
    return IBLDPC_OK;
}

int ibldpc_allreduce_counters(ibldpc_handle h, int64_t* counters_dev, int n, void* stream)
{
    if (!h || !counters_dev || n < 1) return fail_msg(IBLDPC_E_INVALID, "bad arguments");
    if (!h->nccl_comm) return fail_msg(IBLDPC_E_STATE, "ibldpc_nccl_init must be called first");
    ibldpc::DeviceGuard guard_(h->device);
    IBLDPC_CK(guard_.err);
    const int rc = api().AllReduce(counters_dev, counters_dev, (size_t)n, kNcclInt64, kNcclSum, (NcclComm)h->nccl_comm,
                                   (cudaStream_t)stream);
    if (rc) return nccl_fail("ncclAllReduce", rc);
    return IBLDPC_OK;
}

int ibldpc_nccl_finalize(ibldpc_handle h)
{
    if (!h) return IBLDPC_OK;
    if (h->nccl_comm && api().lib) {
        ibldpc::DeviceGuard guard_(h->device);
        api().CommDestroy((NcclComm)h->nccl_comm);
    }
    h->nccl_comm = nullptr;
    return IBLDPC_OK;
}

}  // extern "C"
