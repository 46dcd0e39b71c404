#include "comm.h"

#include <dlfcn.h>

#include <cstring>

namespace dotgpu {
namespace {
struct Id128 { char b[128]; };
typedef int (*fn_get_id)(Id128*);
typedef int (*fn_init_rank)(void**, int, Id128, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

void* open_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    const char* env = getenv("DOTGPU_NCCL_LIB");
    if (env) {
        void* h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    throw Error(DOTGPU_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
}
template <class F>
F sym(void* lib, const char* name) {
    void* p = dlsym(lib, name);
    if (!p) throw Error(DOTGPU_ERR_NCCL, std::string("libnccl lacks ") + name);
    return (F)p;
}
void check(void* lib, int rc, const char* what) {
    if (rc != 0) {
        fn_errstr es = (fn_errstr)dlsym(lib, "ncclGetErrorString");
        throw Error(DOTGPU_ERR_NCCL, std::string(what) + ": " + (es ? es(rc) : "nccl error"));
    }
}
}  // namespace

Comm::~Comm() {
    peer.reset();
    if (comm && lib) {
        fn_destroy d = (fn_destroy)dlsym(lib, "ncclCommDestroy");
        if (d) d(comm);
    }
}

void Comm::unique_id(void* out128) {
    void* lib = open_nccl();
    Id128 id;
    std::memset(&id, 0, sizeof(id));
    check(lib, sym<fn_get_id>(lib, "ncclGetUniqueId")(&id), "ncclGetUniqueId");
    std::memcpy(out128, &id, 128);
}

void Comm::init(const void* uid, int rank_, int world_) {
    DG_REQUIRE(uid != nullptr, "world > 1 needs an ncclUniqueId");
    lib = open_nccl();
    rank = rank_;
    world = world_;
    Id128 id;
    std::memcpy(&id, uid, 128);
    check(lib, sym<fn_init_rank>(lib, "ncclCommInitRank")(&comm, world, id, rank), "ncclCommInitRank");
}

bool Comm::enable_peer(long long cap_doubles, cudaStream_t st) {
    std::unique_ptr<PeerReduce> p(new PeerReduce());
    if (p->init(*this, cap_doubles, st)) peer = std::move(p);
    return (bool)peer;
}

void Comm::all_reduce_sum(double* buf, long long n, cudaStream_t st) {
    if (peer && n <= peer->cap) {
        peer->push(buf, n, st);
        peer->wait_sum(buf, n, st);
        return;
    }
    nccl_all_reduce_sum(buf, n, st);
}

void Comm::all_gather_bytes(const void* send_dev, void* recv_dev, size_t bytes_per_rank, cudaStream_t st) {
    static fn_allgather ag = nullptr;
    if (!ag) ag = sym<fn_allgather>(lib, "ncclAllGather");
    check(lib, ag(send_dev, recv_dev, bytes_per_rank, 0 /* ncclInt8 */, comm, st), "ncclAllGather");
}

void Comm::nccl_all_reduce_sum(double* buf, long long n, cudaStream_t st) {
    static fn_allreduce ar = nullptr;
    if (!ar) ar = sym<fn_allreduce>(lib, "ncclAllReduce");
    // ncclFloat64 = 8, ncclSum = 0
    check(lib, ar(buf, buf, (size_t)n, 8, 0, comm, st), "ncclAllReduce");
}

}  // namespace dotgpu
