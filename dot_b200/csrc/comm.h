// The exchange steps of the path on several GPUs: per L-BFGS iteration the sum over the ranks of [g ; E] and of the subdomains'
// search-direction contributions (DOTTimeStepper.cpp:434-450 is a serial shared-memory loop in the reference).
//   Comm        NCCL communicator (libnccl is dlopen()ed, so single-GPU use has no NCCL dependency): bootstrap, fallback
//               all-reduce, exchange of the IPC handles;
//   PeerReduce  the same sums over cudaIpc-mapped NVLink peer memory, split into a push half and a wait+sum half that producer
//               and consumer kernels fuse into their own passes (peer_reduce.cu, peer.cuh).
#pragma once
#include "common.h"

#include <memory>

namespace dotgpu {
constexpr int PEER_MAX_RANKS = 16;
struct Comm;

// kernel-side views of one exchange (peer.cuh): where a producer stores this rank's vector / where a consumer finds all of them
// Two shapes of the same exchange.  ONE-SHOT (world <= 4): every rank stores its whole vector into its slot on every rank, the
// consumer adds world slots - one flag round, bytes per rank grow with world.  TWO-SHOT (world > 4): entry i belongs to the rank
// i / slice; producers store it into that rank's stage only, a small kernel on the owner adds the world contributions in rank
// order and stores the result into every rank's result vector, the consumer reads its local copy - two flag rounds, 2 x the
// vector per rank whatever the world size.
struct PeerDst {
    double* slot[PEER_MAX_RANKS];    // one-shot: MY slot in rank r's buffer; two-shot: MY stage in rank r's buffer (indices local to r's slice)
    unsigned* flag[PEER_MAX_RANKS];  // MY flag word in rank r's buffer
    unsigned* counter;               // CTA counter of the producer kernel (local)
    long long slice;                 // 0: one-shot; else entries per owner rank
    int world;
    unsigned epoch;
};
struct PeerSrc {
    const double* slots;             // one-shot: [world][cap] of this epoch's parity (local memory, written by the peers); two-shot: the result vector
    const unsigned* flags;           // [world] flags to wait for (round 1 resp. round 2)
    long long cap;
    int world;                       // 0: unused
    int nsum;                        // slots to add: world (one-shot) or 1 (two-shot)
    unsigned epoch;
};
struct PeerBcast {                   // two-shot, second round: the owner of a slice stores the sums into every rank's result vector
    double* res[PEER_MAX_RANKS];     // rank r's result vector + my slice offset
    unsigned* flag2[PEER_MAX_RANKS]; // MY round-2 flag word in rank r's buffer
    unsigned* counter;
};

// One-shot all-reduce over cudaIpc-mapped peer memory (peer_reduce.cu): push = store my vector into every rank's slot + flag,
// wait_sum = wait for every rank's flag, add the slots in rank order.
struct PeerReduce {
    int rank = 0, world = 1;
    long long cap = 0;                        // doubles per slot
    void* base = nullptr;                     // own buffer: [flags | CTA counter | 2 parities x world slots x cap]
    void* peer_base[PEER_MAX_RANKS] = {};     // the same buffer of every rank, mapped here
    unsigned epoch = 0;
    bool ok = false;
    long long slice = 0;                      // two-shot: entries per owner rank (0: one-shot)
    ~PeerReduce();
    bool init(Comm& c, long long cap_doubles, cudaStream_t st);
    void push(const double* buf, long long n, cudaStream_t st);
    void wait_sum(double* out, long long n, cudaStream_t st);
    // fused use: a producer kernel stores into begin() and ends with peer_publish(); a consumer kernel enqueued after it waits on
    // src() and adds the slots itself (peer.cuh)
    PeerDst begin();      // opens the next epoch
    void after_push(cudaStream_t st);  // two-shot: enqueues the owner-side reduce + broadcast (call right after the producer kernel)
    PeerSrc src() const;  // the current epoch
};

struct Comm {
    void* lib = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    std::unique_ptr<PeerReduce> peer;  // set by enable_peer when every rank could map every other rank's buffer
    ~Comm();
    static void unique_id(void* out128);
    void init(const void* unique_id128, int rank, int world);
    // vectors of up to cap_doubles go through peer memory from now on (DOTGPU_PEER_REDUCE=0: keep NCCL); returns whether enabled
    bool enable_peer(long long cap_doubles, cudaStream_t st);
    void all_reduce_sum(double* buf, long long n, cudaStream_t st);
    void nccl_all_reduce_sum(double* buf, long long n, cudaStream_t st);
    void all_gather_bytes(const void* send_dev, void* recv_dev, size_t bytes_per_rank, cudaStream_t st);
};
}  // namespace dotgpu
