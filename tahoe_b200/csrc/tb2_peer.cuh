// tb2_peer.cuh -- the interface exchange over NVLink peer memory.
//
// The reference sums ghost-node contributions with point-to-point MPI messages and dot products with MPI_Allreduce
// (CommManagerT.cpp:424-436, SolverT.cpp:854-860).  Here every rank owns an "exchange window" in its own HBM that all peers map
// (cudaIpc): a rank PUBLISHES its partial interface values into its own window, raises a flag in every peer's window, and each
// sharer PULLS the partials of the other sharers over NVLink inside the kernel that consumes them (interface node update,
// interface rows of A u) -- compute and exchange in one kernel, no collective library on the data path.  Sums run over the
// sharers in ascending rank order on every rank, so all copies of an interface node get bitwise the same value.
//
// Window layout (bytes from the base):  [0,256) arrival words of the interface exchange, one per peer; [256,512) arrival words
// of the scalar exchange; [512,1536) scalar mailboxes [2][kMaxPeers][4] doubles; [2048, ...) interface values [2][3 n_glob].
// Both exchanges are double-buffered on the parity of their epoch: a rank rewrites buffer (e & 1) at epoch e + 2, after it has
// seen every peer's flag of epoch e + 1, which a peer raises only after its own pull of epoch e.
#pragma once
#include <cstdint>

namespace tb2 {

constexpr int kMaxPeers = 16;
constexpr int64_t kPeerFlagsOff = 0, kPeerSFlagsOff = 256, kPeerMailOff = 512, kPeerDataOff = 2048;

struct PeerView {
    char* win[kMaxPeers]; // window base of every rank (own included), mapped into this process
    int rank, nranks;
    int64_t n3;           // 3 n_glob: doubles per interface buffer
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double peer_ld(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double* peer_data(const PeerView& pv, int r, unsigned long long epoch)
{
    return (double*)(pv.win[r] + kPeerDataOff) + (epoch & 1ull) * pv.n3;
}

// end of a kernel that wrote this rank's partials of `epoch` into its own window: the last CTA to finish raises the flag in
// every peer's window.  counter: a device word at 0, returned to 0.
__device__ __forceinline__ void peer_publish(const PeerView& pv, unsigned long long epoch, unsigned* counter, int64_t flags_off)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            for (int r = 0; r < pv.nranks; r++)
                if (r != pv.rank) peer_st_release((unsigned long long*)(pv.win[r] + flags_off) + pv.rank, epoch);
        }
    }
}

// start of a kernel that pulls the peers' partials of `epoch`: every CTA waits for all arrival words in its own window.
// A peer that never arrives (dead process) ends the wait after timeout_s seconds with a trap: every later CUDA call fails.
__device__ __forceinline__ void peer_wait(const PeerView& pv, unsigned long long epoch, int64_t flags_off, unsigned timeout_s = 300, bool trap = true)
{
    if ((int)threadIdx.x < pv.nranks && (int)threadIdx.x != pv.rank) {
        const unsigned long long* f = (const unsigned long long*)(pv.win[pv.rank] + flags_off) + threadIdx.x;
        if (peer_ld_acquire(f) < epoch) {
            const unsigned long long t0 = peer_timer_ns();
            while (peer_ld_acquire(f) < epoch) {
                __nanosleep(100);
                if (peer_timer_ns() - t0 > (unsigned long long)timeout_s * 1000000000ull) {
                    if (trap) __trap();
                    break;
                }
            }
        }
    }
    __syncthreads();
}

// the sum over the sharers (bit r of mask: rank r shares the node) of entry `idx` of the interface buffers of `epoch`
__device__ __forceinline__ double peer_sum(const PeerView& pv, unsigned long long epoch, unsigned mask, int64_t idx)
{
    double v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; r++) // all loads in flight before the first add: one NVLink round trip, not one per sharer
        v[r] = ((mask >> r) & 1u) ? peer_ld(peer_data(pv, r, epoch) + idx) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; r++) s += v[r]; // x + 0.0 is exact: the order over the sharers is what counts
    return s;
}
#endif

} // namespace tb2
