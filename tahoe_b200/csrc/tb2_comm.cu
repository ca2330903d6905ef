// tb2_comm.cu -- multi-GPU exchange: element-partitioned mesh, one process per GPU.
//
// The reference exchanges ghost-node field updates point-to-point and sums dot products with MPI_Allreduce
// (CommManagerT::AllGather CommManagerT.cpp:424-436, SolverT::InnerProduct SolverT.cpp:854-860).  Here every rank owns a
// set of elements; nodes on partition faces are replicated on every touching rank.  The one exchange per sweep is the sum
// of partial nodal quantities (internal force, lumped mass, A_loc p) over the sharers, done as ONE ncclAllReduce on the
// packed global interface vector (every rank contributes zeros for interface nodes it does not touch): every sharer
// receives bitwise the same sum, so the redundant node updates stay identical on all ranks.
//
// NCCL is bound at run time (dlopen libnccl.so.2) so that a single-GPU host needs no NCCL at all and a torch-hosted
// harness shares the NCCL already loaded in the process.
#include <dlfcn.h>

#include <cstring>

#include <cub/cub.cuh>

#include "tb2_internal.h"

namespace tb2 {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSum = 0, kNcclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.lib) return TB2_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot load NCCL: %s", dlerror());
        return TB2_ERR_COMM;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        set_error("NCCL library lacks required symbols");
        return TB2_ERR_COMM;
    }
    g_nccl.lib = h;
    return TB2_OK;
}
static int nccl_fail(int r, const char* what)
{
    set_error("NCCL error %d (%s) in %s", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return TB2_ERR_COMM;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int64_t n_if = 0, n_glob = 0;
    DevBuf<int> nodes, slots;
    DevBuf<double> packed;          // [n_glob][3]
    DevBuf<unsigned char> owned;    // [nn]
    DevBuf<double> scal;            // all-reduce scratch for scalars
    // overlapped explicit step: node -> slot map, the elements touching interface nodes, a stream for the collective
    DevBuf<int> node_slot;          // [nn], -1 = private node
    DevBuf<int> belems;             // [nb]
    DevBuf<unsigned char> belem_flag; // [ne]
    int64_t nb = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_reduced = nullptr, ev_done = nullptr;
    // peer-memory exchange (tb2_peer.cuh): this rank's window, the peers' windows as mapped here, the sharers of every
    // interface node, last-CTA counters, the epochs of the two exchanges
    char* win = nullptr;
    size_t win_bytes = 0;
    void* opened[kMaxPeers] = {};
    PeerView pv{};
    bool peer = false;
    DevBuf<unsigned> share_mask; // [n_if] bit r: rank r touches the node
    DevBuf<unsigned> slot_mask;  // [n_glob] the same by slot of the packed interface vector (0 for slots this rank does not touch)
    DevBuf<unsigned> counter;    // [2]
    unsigned long long xe = 0, se = 0;
    ~Comm()
    {
        for (void* q : opened)
            if (q) cudaIpcCloseMemHandle(q);
        if (win) cudaFree(win);
        if (ev_packed) cudaEventDestroy(ev_packed);
        if (ev_reduced) cudaEventDestroy(ev_reduced);
        if (ev_done) cudaEventDestroy(ev_done);
        if (stream) cudaStreamDestroy(stream);
    }
};

__global__ void k_fill_int(int64_t n, int v, int* __restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
__global__ void k_node_slots(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots, int* __restrict__ node_slot)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_if) node_slot[nodes[k]] = slots[k];
}
// flag[e] = 1 if element e has a node on the partition interface
__global__ void k_boundary_flags(int64_t ne, int64_t stride, const int* __restrict__ conn, const int* __restrict__ node_slot,
                                 unsigned char* __restrict__ flag)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int any = 0;
#pragma unroll
    for (int a = 0; a < 8; a++) any |= node_slot[conn[a * stride + e]] >= 0;
    flag[e] = (unsigned char)any;
}

__global__ void k_pack(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots, const double* __restrict__ nodal,
                       double* __restrict__ packed)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_if) return;
    const int64_t k = t / 3;
    const int i = (int)(t % 3);
    packed[3 * (int64_t)slots[k] + i] = nodal[3 * (int64_t)nodes[k] + i];
}
__global__ void k_unpack(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots, const double* __restrict__ packed,
                         double* __restrict__ nodal)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_if) return;
    const int64_t k = t / 3;
    const int i = (int)(t % 3);
    nodal[3 * (int64_t)nodes[k] + i] = packed[3 * (int64_t)slots[k] + i];
}

// equation-space variant: v[neq] holds values of the active dofs only (prescribed dofs contribute / receive nothing)
__global__ void k_pack_eq(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots, const int* __restrict__ eqnos,
                          const double* __restrict__ v, double* __restrict__ packed)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_if) return;
    const int64_t k = t / 3;
    const int i = (int)(t % 3);
    const int eq = eqnos[3 * (int64_t)nodes[k] + i];
    packed[3 * (int64_t)slots[k] + i] = eq > 0 ? v[eq - 1] : 0.0;
}
__global__ void k_unpack_eq(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots, const int* __restrict__ eqnos,
                            const double* __restrict__ packed, double* __restrict__ v)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_if) return;
    const int64_t k = t / 3;
    const int i = (int)(t % 3);
    const int eq = eqnos[3 * (int64_t)nodes[k] + i];
    if (eq > 0) v[eq - 1] = packed[3 * (int64_t)slots[k] + i];
}

// ---- peer-memory exchange -------------------------------------------------------------------------------------------------------
__global__ void k_pack_rank_bit(int64_t n_if, const int* __restrict__ slots, double bit, double* __restrict__ packed)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_if) packed[3 * (int64_t)slots[k]] = bit;
}
__global__ void k_share_mask(int64_t n_if, const int* __restrict__ slots, const double* __restrict__ packed, unsigned* __restrict__ mask,
                             unsigned* __restrict__ slot_mask)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_if) return;
    const unsigned mk = (unsigned)packed[3 * (int64_t)slots[k]];
    mask[k] = mk;
    slot_mask[slots[k]] = mk;
}
// publish the interface entries of an equation vector (prescribed dofs: 0) into this rank's window and raise the peers' flags
__global__ void __launch_bounds__(256) k_peer_pack_eq(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                     const int* __restrict__ eqnos, const double* __restrict__ v, PeerView pv,
                                                     unsigned long long epoch, unsigned* counter)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < 3 * n_if) {
        const int64_t k = t / 3;
        const int i = (int)(t % 3);
        const int eq = eqnos[3 * (int64_t)nodes[k] + i];
        peer_data(pv, pv.rank, epoch)[3 * (int64_t)slots[k] + i] = eq > 0 ? v[eq - 1] : 0.0;
    }
    peer_publish(pv, epoch, counter, kPeerFlagsOff);
}
// wait for the peers' flags, pull and sum the sharers' entries over NVLink, write them back into the equation vector
__global__ void __launch_bounds__(256) k_peer_pull_eq(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                     const unsigned* __restrict__ share_mask, const int* __restrict__ eqnos,
                                                     double* __restrict__ v, PeerView pv, unsigned long long epoch)
{
    peer_wait(pv, epoch, kPeerFlagsOff);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_if) return;
    const int64_t k = t / 3;
    const int i = (int)(t % 3);
    const int eq = eqnos[3 * (int64_t)nodes[k] + i];
    if (eq > 0) v[eq - 1] = peer_sum(pv, epoch, share_mask[k], 3 * (int64_t)slots[k] + i);
}
// all-reduce of n <= 4 scalars: every rank writes its values into every peer's mailbox, raises the flag, waits for the others
// and sums the mailbox in rank order (identical bits on every rank).  One CTA.
__global__ void k_peer_scalars(double* __restrict__ vals, int n, PeerView pv, unsigned long long epoch, unsigned timeout_s, bool trap)
{
    const int t = threadIdx.x;
    if (t < pv.nranks) {
        double* box = (double*)(pv.win[t] + kPeerMailOff) + ((epoch & 1ull) * kMaxPeers + pv.rank) * 4;
        for (int i = 0; i < n; i++) box[i] = vals[i];
        __threadfence_system();
        if (t != pv.rank) peer_st_release((unsigned long long*)(pv.win[t] + kPeerSFlagsOff) + pv.rank, epoch);
    }
    peer_wait(pv, epoch, kPeerSFlagsOff, timeout_s, trap);
    if (t < n) {
        const double* box = (const double*)(pv.win[pv.rank] + kPeerMailOff) + (epoch & 1ull) * kMaxPeers * 4;
        double s = 0.0;
        for (int r = 0; r < pv.nranks; r++) s += peer_ld(box + r * 4 + t);
        vals[t] = s;
    }
}

bool comm_active(tb2_mesh* m) { return m->comm && m->comm->nranks > 1 && m->comm->n_glob > 0; }

bool comm_plan(tb2_mesh* m, CommPlan* out)
{
    if (!comm_active(m) || !m->comm->stream) return false;
    Comm* c = m->comm;
    out->n_if = c->n_if;
    out->n_glob = c->n_glob;
    out->nb = c->nb;
    out->nodes = c->nodes.p;
    out->slots = c->slots.p;
    out->node_slot = c->node_slot.p;
    out->belems = c->belems.p;
    out->belem_flag = c->belem_flag.p;
    out->packed = c->packed.p;
    out->stream = c->stream;
    out->ev_packed = c->ev_packed;
    out->ev_reduced = c->ev_reduced;
    out->ev_done = c->ev_done;
    out->peer = c->peer;
    out->pv = c->pv;
    out->share_mask = c->share_mask.p;
    out->slot_mask = c->slot_mask.p;
    out->counter = c->counter.p;
    out->epoch = &c->xe;
    out->sepoch = &c->se;
    return true;
}

// the packed-vector all-reduce on the communicator's own stream (ordering is the caller's: events in CommPlan)
int comm_allreduce_packed(tb2_mesh* m)
{
    Comm* c = m->comm;
    if (c->peer) return TB2_OK; // the consumer pulls the peers' partials itself
    ProfScope ps(m, kProfComm, 1, c->stream);
    const int r = g_nccl.AllReduce(c->packed.p, c->packed.p, (size_t)(3 * c->n_glob), kNcclFloat64, kNcclSum, c->comm, c->stream);
    return r ? nccl_fail(r, "ncclAllReduce(interface, overlapped)") : TB2_OK;
}
const unsigned char* comm_owned_mask(tb2_mesh* m) { return m->comm ? m->comm->owned.p : nullptr; }

// the two halves of comm_sum_interface_eq for callers that overlap the all-reduce with their own work (distributed PCG):
// zero + pack the interface entries of an equation vector on stream st; copy the reduced entries back
int comm_pack_eq(tb2_mesh* m, const int* d_eqnos, const double* d_eqvec, cudaStream_t st)
{
    Comm* c = m->comm;
    const int T = 256;
    if (c->peer) {
        ProfScope ps(m, kProfComm, 1, st);
        c->xe++;
        const int64_t nt = 3 * c->n_if > 0 ? 3 * c->n_if : 1;
        k_peer_pack_eq<<<(unsigned)((nt + T - 1) / T), T, 0, st>>>(c->n_if, c->nodes.p, c->slots.p, d_eqnos, d_eqvec, c->pv, c->xe, c->counter.p);
        TB2_CUDA(cudaGetLastError());
        return TB2_OK;
    }
    ProfScope ps(m, kProfComm, 2, st);
    TB2_CUDA(cudaMemsetAsync(c->packed.p, 0, 3 * c->n_glob * sizeof(double), st));
    if (c->n_if) k_pack_eq<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, st>>>(c->n_if, c->nodes.p, c->slots.p, d_eqnos, d_eqvec, c->packed.p);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}
int comm_unpack_eq(tb2_mesh* m, const int* d_eqnos, double* d_eqvec, cudaStream_t st)
{
    Comm* c = m->comm;
    ProfScope ps(m, kProfComm, 1, st);
    const int T = 256;
    if (c->peer) {
        const int64_t nt = 3 * c->n_if > 0 ? 3 * c->n_if : 1;
        k_peer_pull_eq<<<(unsigned)((nt + T - 1) / T), T, 0, st>>>(c->n_if, c->nodes.p, c->slots.p, c->share_mask.p, d_eqnos, d_eqvec, c->pv, c->xe);
        TB2_CUDA(cudaGetLastError());
        return TB2_OK;
    }
    if (c->n_if) k_unpack_eq<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, st>>>(c->n_if, c->nodes.p, c->slots.p, d_eqnos, c->packed.p, d_eqvec);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

// in-place sum of n doubles over ranks (PCG scalars); no-op without a communicator
int comm_allreduce_scalars(tb2_mesh* m, double* d_vals, int n)
{
    if (!m->comm || m->comm->nranks == 1) return TB2_OK;
    if (m->comm->peer && n <= 4) {
        Comm* c = m->comm;
        k_peer_scalars<<<1, 32, 0, m->stream>>>(d_vals, n, c->pv, ++c->se, 300, true);
        TB2_CUDA(cudaGetLastError());
        return TB2_OK;
    }
    const int r = g_nccl.AllReduce(d_vals, d_vals, (size_t)n, kNcclFloat64, kNcclSum, m->comm->comm, m->stream);
    return r ? nccl_fail(r, "ncclAllReduce(scalars)") : TB2_OK;
}

// v[neq] += contributions of the other sharers on interface equations (A_loc p, diagonal) -- same packed all-reduce
int comm_sum_interface_eq(tb2_mesh* m, const int* d_eqnos, double* d_eqvec)
{
    Comm* c = m->comm;
    if (!c || c->nranks == 1 || c->n_glob == 0) return TB2_OK;
    ProfScope ps(m, kProfComm, 2);
    TB2_CUDA(cudaMemsetAsync(c->packed.p, 0, 3 * c->n_glob * sizeof(double), m->stream));
    const int T = 256;
    if (c->n_if) k_pack_eq<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->nodes.p, c->slots.p, d_eqnos, d_eqvec, c->packed.p);
    const int r = g_nccl.AllReduce(c->packed.p, c->packed.p, (size_t)(3 * c->n_glob), kNcclFloat64, kNcclSum, c->comm, m->stream);
    if (r) return nccl_fail(r, "ncclAllReduce(interface, eq)");
    if (c->n_if) k_unpack_eq<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->nodes.p, c->slots.p, d_eqnos, c->packed.p, d_eqvec);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

} // namespace tb2

using namespace tb2;

extern "C" {

int tb2_comm_unique_id(char h_id[128])
{
    TB2_ARG(h_id);
    TB2_CHECK(load_nccl());
    ncclUniqueId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(h_id, id.internal, 128);
    return TB2_OK;
}

int tb2_comm_init(tb2_mesh* m, int rank, int nranks, const char h_id[128], int64_t n_if, const int32_t* h_nodes, const int32_t* h_slots,
                  int64_t n_glob, const uint8_t* h_owned)
{
    TB2_ARG(m && h_id && nranks >= 1 && rank >= 0 && rank < nranks && n_if >= 0 && n_glob >= n_if);
    TB2_ARG(n_if == 0 || (h_nodes && h_slots));
    DeviceGuard dg(m->device);
    TB2_CHECK(load_nccl());
    if (m->comm) tb2_comm_destroy(m);
    Comm* c = new Comm;
    c->rank = rank;
    c->nranks = nranks;
    c->n_if = n_if;
    c->n_glob = n_glob;
    ncclUniqueId id;
    memcpy(id.internal, h_id, 128);
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r) {
        delete c;
        return nccl_fail(r, "ncclCommInitRank");
    }
    cudaError_t e = c->nodes.alloc(n_if > 0 ? n_if : 1);
    if (e == cudaSuccess) e = c->slots.alloc(n_if > 0 ? n_if : 1);
    if (e == cudaSuccess) e = c->packed.alloc(3 * (n_glob > 0 ? n_glob : 1));
    if (e == cudaSuccess) e = c->owned.alloc(m->nn);
    if (e == cudaSuccess) e = c->scal.alloc(8);
    if (e == cudaSuccess && n_if) e = cudaMemcpy(c->nodes.p, h_nodes, n_if * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_if) e = cudaMemcpy(c->slots.p, h_slots, n_if * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        if (h_owned) e = cudaMemcpy(c->owned.p, h_owned, m->nn, cudaMemcpyHostToDevice);
        else e = cudaMemset(c->owned.p, 1, m->nn);
    }
    // overlap plan: node -> slot map and the (ascending) list of elements that touch an interface node
    if (e == cudaSuccess) e = c->node_slot.alloc(m->nn);
    if (e == cudaSuccess && nranks > 1 && n_glob > 0) {
        const int T = 256;
        DevBuf<unsigned char>& flag = c->belem_flag;
        DevBuf<unsigned char> tmp;
        DevBuf<int> nsel;
        k_fill_int<<<(unsigned)((m->nn + T - 1) / T), T, 0, m->stream>>>(m->nn, -1, c->node_slot.p);
        if (n_if) k_node_slots<<<(unsigned)((n_if + T - 1) / T), T, 0, m->stream>>>(n_if, c->nodes.p, c->slots.p, c->node_slot.p);
        e = flag.alloc(m->ne);
        if (e == cudaSuccess) e = nsel.alloc(1);
        if (e == cudaSuccess) e = c->belems.alloc(m->ne);
        if (e == cudaSuccess) {
            k_boundary_flags<<<(unsigned)((m->ne + T - 1) / T), T, 0, m->stream>>>(m->ne, m->stride, m->conn.p, c->node_slot.p, flag.p);
            size_t bytes = 0;
            cub::CountingInputIterator<int> ids(0);
            e = cub::DeviceSelect::Flagged(nullptr, bytes, ids, flag.p, c->belems.p, nsel.p, (int)m->ne, m->stream);
            if (e == cudaSuccess) e = tmp.alloc(bytes);
            if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp.p, bytes, ids, flag.p, c->belems.p, nsel.p, (int)m->ne, m->stream);
            int h_n = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&h_n, nsel.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
            c->nb = h_n;
        }
        // the exchange lane outranks the bulk sweeps: its few CTAs (boundary elements, NCCL) must not queue behind a full-GPU grid
        int prio_lo = 0, prio_hi = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_reduced, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        g_nccl.CommDestroy(c->comm);
        delete c;
        return cuda_fail(e, "communicator buffers", __FILE__, __LINE__);
    }
    m->comm = c;
    return TB2_OK;
}

int tb2_comm_peer_export(tb2_mesh* m, char h_handle[64])
{
    TB2_ARG(m && h_handle);
    Comm* c = m->comm;
    if (!c) {
        set_error("tb2_comm_peer_export: no communicator on this mesh");
        return TB2_ERR_ARG;
    }
    if (c->nranks > kMaxPeers) {
        set_error("peer-memory exchange supports up to %d ranks (one NVSwitch domain), got %d", kMaxPeers, c->nranks);
        return TB2_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard dg(m->device);
    if (!c->win) {
        c->win_bytes = (size_t)kPeerDataOff + 2 * 3 * (size_t)(c->n_glob > 0 ? c->n_glob : 1) * sizeof(double);
        TB2_CUDA(cudaMalloc((void**)&c->win, c->win_bytes));
    }
    TB2_CUDA(cudaMemset(c->win, 0, c->win_bytes));
    TB2_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    TB2_CUDA(cudaIpcGetMemHandle(&h, c->win));
    memcpy(h_handle, &h, 64);
    return TB2_OK;
}

int tb2_comm_peer_import(tb2_mesh* m, const char* h_handles)
{
    TB2_ARG(m && h_handles);
    Comm* c = m->comm;
    if (!c || !c->win) {
        set_error("tb2_comm_peer_import: call tb2_comm_peer_export on every rank first");
        return TB2_ERR_ARG;
    }
    DeviceGuard dg(m->device);
    PeerView pv{};
    pv.rank = c->rank;
    pv.nranks = c->nranks;
    pv.n3 = 3 * (c->n_glob > 0 ? c->n_glob : 1);
    for (int r = 0; r < c->nranks; r++) {
        if (r == c->rank) {
            pv.win[r] = c->win;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles + 64 * (size_t)r, 64);
        void* q = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (void*& o : c->opened)
                if (o) {
                    cudaIpcCloseMemHandle(o);
                    o = nullptr;
                }
            return cuda_fail(e, "cudaIpcOpenMemHandle (peer window)", __FILE__, __LINE__);
        }
        c->opened[r] = q;
        pv.win[r] = (char*)q;
    }
    // the sharers of every interface node: each rank adds its bit 2^rank on the slots it touches (exact in a double)
    const int T = 256;
    TB2_CUDA(c->share_mask.alloc(c->n_if > 0 ? c->n_if : 1));
    TB2_CUDA(c->slot_mask.alloc(c->n_glob > 0 ? c->n_glob : 1));
    TB2_CUDA(cudaMemsetAsync(c->slot_mask.p, 0, (c->n_glob > 0 ? c->n_glob : 1) * sizeof(unsigned), m->stream));
    TB2_CUDA(c->counter.alloc(2));
    TB2_CUDA(cudaMemsetAsync(c->counter.p, 0, 2 * sizeof(unsigned), m->stream));
    TB2_CUDA(cudaMemsetAsync(c->packed.p, 0, 3 * (c->n_glob > 0 ? c->n_glob : 1) * sizeof(double), m->stream));
    if (c->n_if) k_pack_rank_bit<<<(unsigned)((c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->slots.p, (double)(1u << c->rank), c->packed.p);
    if (c->n_glob > 0) {
        const int r = g_nccl.AllReduce(c->packed.p, c->packed.p, (size_t)(3 * c->n_glob), kNcclFloat64, kNcclSum, c->comm, m->stream);
        if (r) return nccl_fail(r, "ncclAllReduce(sharer masks)");
    }
    if (c->n_if) k_share_mask<<<(unsigned)((c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->slots.p, c->packed.p, c->share_mask.p, c->slot_mask.p);
    TB2_CUDA(cudaGetLastError());
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    c->pv = pv;
    c->xe = c->se = 0;
    c->peer = true;
    return TB2_OK;
}

int tb2_comm_peer_enabled(tb2_mesh* m) { return m && m->comm && m->comm->peer ? 1 : 0; }

int tb2_comm_peer_disable(tb2_mesh* m)
{
    TB2_ARG(m);
    if (!m->comm) return TB2_OK;
    DeviceGuard dg(m->device);
    cudaStreamSynchronize(m->stream);
    if (m->comm->stream) cudaStreamSynchronize(m->comm->stream);
    m->comm->peer = false;
    return TB2_OK;
}

int tb2_comm_destroy(tb2_mesh* m)
{
    if (!m || !m->comm) return TB2_OK;
    DeviceGuard dg(m->device);
    cudaStreamSynchronize(m->stream);
    if (m->comm->stream) cudaStreamSynchronize(m->comm->stream);
    if (m->comm->comm) g_nccl.CommDestroy(m->comm->comm);
    delete m->comm;
    m->comm = nullptr;
    return TB2_OK;
}

int tb2_comm_sum_interface(tb2_mesh* m, double* d_nodal)
{
    TB2_ARG(m && d_nodal);
    Comm* c = m->comm;
    if (!c || c->nranks == 1 || c->n_glob == 0) return TB2_OK;
    DeviceGuard dg(m->device);
    ProfScope ps(m, kProfComm, 2);
    TB2_CUDA(cudaMemsetAsync(c->packed.p, 0, 3 * c->n_glob * sizeof(double), m->stream));
    const int T = 256;
    if (c->n_if) k_pack<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->nodes.p, c->slots.p, d_nodal, c->packed.p);
    const int r = g_nccl.AllReduce(c->packed.p, c->packed.p, (size_t)(3 * c->n_glob), kNcclFloat64, kNcclSum, c->comm, m->stream);
    if (r) return nccl_fail(r, "ncclAllReduce(interface)");
    if (c->n_if) k_unpack<<<(unsigned)((3 * c->n_if + T - 1) / T), T, 0, m->stream>>>(c->n_if, c->nodes.p, c->slots.p, c->packed.p, d_nodal);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

} // extern "C"
