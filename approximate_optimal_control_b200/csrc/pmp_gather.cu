// pmp_gather.cu -- the one collective of the multi-GPU path behind the C ABI: all-gather of the endpoint records
// (x, t, v, lambda) of every rank's shard, so that a host that is not PyTorch (the reference is JAX) can shard the
// terminal samples over the GPUs of a node and still end up with the whole dataset on every rank (BASELINE.json
// north_star: "only an NCCL all-gather of endpoint (x, lambda, v) datasets").  NCCL is resolved at run time (dlopen of
// libnccl.so.2: the process's own copy if one is loaded already, e.g. the one PyTorch ships), so libb200pmp.so carries
// no link-time dependency on it and single-GPU users never touch it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/b200pmp.h"

namespace pmp {
namespace {

struct UniqueId { char internal[128]; };  // ncclUniqueId
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn error_string = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) {
            n.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        n.get_unique_id = (GetUniqueIdFn)dlsym(n.handle, "ncclGetUniqueId");
        n.comm_init_rank = (CommInitRankFn)dlsym(n.handle, "ncclCommInitRank");
        n.comm_destroy = (CommDestroyFn)dlsym(n.handle, "ncclCommDestroy");
        n.all_gather = (AllGatherFn)dlsym(n.handle, "ncclAllGather");
        n.error_string = (GetErrorStringFn)dlsym(n.handle, "ncclGetErrorString");
        if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_gather) {
            n.why = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
            n.handle = nullptr;
        }
    });
    return n;
}

}  // namespace

// each returns 0, or a negative code with *msg set
int comm_unique_id(void* id128, std::string* msg) {
    Nccl& n = nccl();
    if (!n.handle) { *msg = n.why; return -1; }
    UniqueId id;
    if (int rc = n.get_unique_id(&id)) { *msg = std::string("ncclGetUniqueId: ") + (n.error_string ? n.error_string(rc) : "error"); return -2; }
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int comm_init(void** comm, int world, int rank, const void* id128, std::string* msg) {
    Nccl& n = nccl();
    if (!n.handle) { *msg = n.why; return -1; }
    UniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    if (int rc = n.comm_init_rank(comm, world, id, rank)) { *msg = std::string("ncclCommInitRank: ") + (n.error_string ? n.error_string(rc) : "error"); return -2; }
    return 0;
}

int comm_destroy(void* comm, std::string* msg) {
    Nccl& n = nccl();
    if (!n.handle) { *msg = n.why; return -1; }
    if (int rc = n.comm_destroy(comm)) { *msg = std::string("ncclCommDestroy: ") + (n.error_string ? n.error_string(rc) : "error"); return -2; }
    return 0;
}

int allgather(void* comm, const void* local, void* global, size_t count, int dtype, cudaStream_t st, std::string* msg) {
    Nccl& n = nccl();
    if (!n.handle) { *msg = n.why; return -1; }
    const int nccl_type = dtype == PMP_F32 ? 7 /* ncclFloat32 */ : 8 /* ncclFloat64 */;
    if (int rc = n.all_gather(local, global, count, nccl_type, comm, st)) { *msg = std::string("ncclAllGather: ") + (n.error_string ? n.error_string(rc) : "error"); return -2; }
    return 0;
}

}  // namespace pmp
