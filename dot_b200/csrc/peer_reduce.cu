// All-reduce(sum) of the two vectors a multi-GPU L-BFGS iteration exchanges - the gradient [g ; E] and the search direction p
// (DOTTimeStepper.cpp:406-450 is a serial shared-memory loop in the reference) - over NVLink PEER MEMORY instead of an NCCL call:
//
//   push      every rank stores its vector straight into a slot of every peer's buffer (cudaIpc-mapped device memory, plain
//             coalesced stores that travel over NVLink / NVSwitch), then publishes an epoch number in the peer's flag word
//             (fence.sys + st.release.sys by the last CTA to finish);
//   wait_sum  every rank waits until the flags of all ranks show the epoch (ld.acquire.sys) and adds the slots IN RANK ORDER:
//             the result is bit-identical on every rank and from run to run (no ring / tree whose order depends on timing).
//
// One-shot, one flag round, no intermediate kernel on a third GPU: for the 4.4 MB vectors of the 1M-tet workload the exchange is
// latency-bound and this costs about half of an ncclAllReduce launch.  Slots are double-buffered by epoch parity: a rank can only
// start epoch e+2 after it has finished epoch e+1, which needs every peer's push e+1, which follows that peer's wait_sum(e) in
// stream order - so nobody overwrites a slot that is still being read.  A spin that outlives ~2 s traps (the host sees a CUDA
// error instead of a hung GPU).  NCCL stays the bootstrap (exchange of the IPC handles) and the fallback.
#include "peer.cuh"

namespace dotgpu {

namespace {

constexpr int PR_TPB = 256;
// bytes in front of the data: round-1 flags[world] at 0 (one per source rank), round-2 flags[world] at 64, CTA counters at 128 / 132
constexpr size_t PR_HEADER = 256;

__global__ void __launch_bounds__(PR_TPB) k_peer_push(long long n, const double* __restrict__ buf, PeerDst D) {
    for (long long i = (long long)blockIdx.x * PR_TPB + threadIdx.x; i < n; i += (long long)gridDim.x * PR_TPB) {
        peer_store(D, i, buf[i]);
    }
    peer_publish(D);
}

// two-shot, on the owner of a slice: wait for every rank's contribution, add them in rank order, store the sums into every rank's
// result vector, publish round 2
__global__ void __launch_bounds__(PR_TPB) k_peer_reduce_bcast(long long nslice, const double* stage, long long slice, int world,
                                                              const unsigned* flags1, unsigned epoch, PeerBcast B) {
    {
        PeerSrc W;
        W.flags = flags1;
        W.world = world;
        W.epoch = epoch;
        peer_wait_flags(W);
    }
    for (long long j = (long long)blockIdx.x * PR_TPB + threadIdx.x; j < nslice; j += (long long)gridDim.x * PR_TPB) {
        double s = __ldcg(stage + j);
        for (int r = 1; r < world; ++r) s += __ldcg(stage + (long long)r * slice + j);
        for (int r = 0; r < world; ++r) B.res[r][j] = s;
    }
    PeerDst D;  // only the fields peer_publish reads
    for (int r = 0; r < world; ++r) D.flag[r] = B.flag2[r];
    D.counter = B.counter;
    D.world = world;
    D.epoch = epoch;
    peer_publish(D);
}

__global__ void __launch_bounds__(PR_TPB) k_peer_wait_sum(long long n, double* __restrict__ out, PeerSrc S) {
    peer_wait_flags(S);
    for (long long i = (long long)blockIdx.x * PR_TPB + threadIdx.x; i < n; i += (long long)gridDim.x * PR_TPB) out[i] = peer_sum(S, i);
}

}  // namespace

PeerReduce::~PeerReduce() {
    for (int r = 0; r < world; ++r)
        if (r != rank && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
    if (base) cudaFree(base);
}

bool PeerReduce::init(Comm& c, long long cap_doubles, cudaStream_t st) {
    rank = c.rank;
    world = c.world;
    if (world < 2 || world > PEER_MAX_RANKS) return false;
    if (const char* e = std::getenv("DOTGPU_PEER_REDUCE"))
        if (e[0] == '0') return false;
    // two-shot above 4 ranks (DOTGPU_PEER_TWO_SHOT=0/1 overrides): see comm.h
    bool two = world > 4;
    if (const char* e = std::getenv("DOTGPU_PEER_TWO_SHOT")) two = e[0] == '1';
    if (two) {
        slice = ((cap_doubles + world - 1) / world + 15) / 16 * 16;
        cap = slice * world;
    } else {
        slice = 0;
        cap = (cap_doubles + 15) / 16 * 16;
    }
    // one-shot: 2 parities x world slots x cap; two-shot: 2 parities x (world stages x slice + result vector of cap) = the same size
    const size_t bytes = PR_HEADER + (two ? 4 : 2 * (size_t)world) * cap * sizeof(double);
    DG_CUDA(cudaMalloc(&base, bytes));
    DG_CUDA(cudaMemsetAsync(base, 0, bytes, st));
    // exchange the IPC handles through the communicator that already exists
    cudaIpcMemHandle_t mine;
    DG_CUDA(cudaIpcGetMemHandle(&mine, base));
    DevBuf<char> send(sizeof(mine)), recv(sizeof(mine) * (size_t)world);
    DG_CUDA(cudaMemcpyAsync(send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    c.all_gather_bytes(send.p, recv.p, sizeof(mine), st);
    std::vector<cudaIpcMemHandle_t> all(world);
    DG_CUDA(cudaMemcpyAsync(all.data(), recv.p, sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    int failed = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            peer_base[r] = base;
            continue;
        }
        if (cudaIpcOpenMemHandle(&peer_base[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            peer_base[r] = nullptr;
            failed = 1;
        }
    }
    // all ranks must agree (a rank without peer access to somebody would otherwise wait for flags nobody writes)
    DevBuf<double> f(1);
    const double fv = failed;
    DG_CUDA(cudaMemcpyAsync(f.p, &fv, sizeof(double), cudaMemcpyHostToDevice, st));
    c.nccl_all_reduce_sum(f.p, 1, st);
    double tot = 0.0;
    DG_CUDA(cudaMemcpyAsync(&tot, f.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    ok = tot == 0.0;
    return ok;
}

PeerDst PeerReduce::begin() {
    DG_REQUIRE(ok, "peer all-reduce not initialised");
    ++epoch;
    PeerDst D;
    const size_t par = epoch & 1u;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) {
        D.slot[r] = nullptr;
        D.flag[r] = nullptr;
    }
    for (int r = 0; r < world; ++r) {
        char* b = static_cast<char*>(peer_base[r]);
        double* data = reinterpret_cast<double*>(b + PR_HEADER);
        // two-shot data layout per parity: [world stages x slice | result vector of cap]  (2 * cap doubles)
        D.slot[r] = slice ? data + par * 2 * cap + (size_t)rank * slice : data + (par * world + rank) * cap;
        D.flag[r] = reinterpret_cast<unsigned*>(b) + rank;
    }
    D.counter = reinterpret_cast<unsigned*>(static_cast<char*>(base) + 128);
    D.slice = slice;
    D.world = world;
    D.epoch = epoch;
    return D;
}

void PeerReduce::after_push(cudaStream_t st) {
    if (!slice) return;
    const size_t par = epoch & 1u;
    PeerBcast B;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) {
        B.res[r] = nullptr;
        B.flag2[r] = nullptr;
    }
    for (int r = 0; r < world; ++r) {
        char* b = static_cast<char*>(peer_base[r]);
        B.res[r] = reinterpret_cast<double*>(b + PR_HEADER) + par * 2 * cap + cap + (size_t)rank * slice;
        B.flag2[r] = reinterpret_cast<unsigned*>(b + 64) + rank;
    }
    B.counter = reinterpret_cast<unsigned*>(static_cast<char*>(base) + 132);
    const double* stage = reinterpret_cast<const double*>(static_cast<const char*>(base) + PR_HEADER) + par * 2 * cap;
    const int grid = (int)std::min<long long>(296, std::max<long long>(1, (slice + PR_TPB - 1) / PR_TPB));
    k_peer_reduce_bcast<<<grid, PR_TPB, 0, st>>>(slice, stage, slice, world, reinterpret_cast<const unsigned*>(base), epoch, B);
    count_launch();
}

PeerSrc PeerReduce::src() const {
    PeerSrc S;
    const double* data = reinterpret_cast<const double*>(static_cast<const char*>(base) + PR_HEADER);
    const size_t par = epoch & 1u;
    if (slice) {
        S.slots = data + par * 2 * cap + cap;   // the result vector the slice owners wrote
        S.flags = reinterpret_cast<const unsigned*>(static_cast<const char*>(base) + 64);
        S.nsum = 1;
    } else {
        S.slots = data + par * world * cap;
        S.flags = reinterpret_cast<const unsigned*>(base);
        S.nsum = world;
    }
    S.cap = cap;
    S.world = world;
    S.epoch = epoch;
    return S;
}

void PeerReduce::push(const double* buf, long long n, cudaStream_t st) {
    DG_REQUIRE(ok && n <= cap, "peer all-reduce: vector longer than the slots");
    const PeerDst D = begin();
    const int grid = (int)std::min<long long>(592, std::max<long long>(1, (n + PR_TPB - 1) / PR_TPB));
    k_peer_push<<<grid, PR_TPB, 0, st>>>(n, buf, D);
    count_launch();
    after_push(st);
}

void PeerReduce::wait_sum(double* out, long long n, cudaStream_t st) {
    const int grid = (int)std::min<long long>(592, std::max<long long>(1, (n + PR_TPB - 1) / PR_TPB));
    k_peer_wait_sum<<<grid, PR_TPB, 0, st>>>(n, out, src());
    count_launch();
}

}  // namespace dotgpu
