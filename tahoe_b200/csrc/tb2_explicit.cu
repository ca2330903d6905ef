// tb2_explicit.cu -- K5: explicit central-difference update on device-resident fields.
//
// Reference order of one step:
//   FieldT::InitStep -> nExplicitCD::Predictor (nExplicitCD.cpp:72-96): d += dt v + dt^2/2 a ; v += dt/2 a ; a = 0,
//                       then ConsistentKBC on prescribed dofs (nExplicitCD.cpp:20-69)
//   LinearSolver::Solve (LinearSolver.cpp:37-101): R = fext - fint(d) ; update = M^-1 R (DiagonalMatrixT.cpp:267-323)
//   FieldT::AssembleUpdate (FieldT.cpp:531-556: prescribed dofs get 0) ; nExplicitCD::Corrector (:98-139): v += dt/2 upd ; a += upd
// On the device this is two launches per step: the element sweep (K1) and one node kernel (K5) that gathers the element
// forces, forms R, applies M^-1, the corrector and -- when another step follows -- the next step's predictor.  In that steady
// state a node moves d, v in and out and 1/m, the boundary codes (and fext, if there is one) in: 123-147 B/node/step beside
// the 192 B/element scratch read.
//
// Schedules.  One GPU: K1, K5 back to back on the mesh stream.  Round 1 cut both kernels into slabs on two streams so that
// they might overlap; ncu showed that they never did (the sweep owns the register file), and round 2's lab
// (profiles/tools/k1_lab.cu) measured every other way of hiding the node work behind the sweep -- register-lean sweeps beside
// the node kernel, forces kept in shared memory with block-local node updates, a persistent warp-specialised kernel -- slower than
// the two kernels back to back, so the slab machinery is gone.
// Several GPUs: two lanes.  comm stream: K1 over the elements touching interface nodes -> their partial forces packed ->
// ncclAllReduce -> update of the interface nodes.  Main stream: K1 over the other elements (beside the all-reduce) -> K5 over the
// private nodes.  Two events per step tie the lanes together.
#include <chrono>
#include <cstdlib>

#include "tb2_internal.h"
#include "tb2_node_update.cuh"

namespace tb2 {

int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration);
int launch_node_gather(tb2_mesh* m, double* d_out, bool per_dof);
int launch_element_forces_range(tb2_group* g, const double* d_u, const double* d_ul, int iteration, int64_t e0, int64_t e1, cudaStream_t st,
                                const int* d_elist = nullptr, const unsigned char* d_skip = nullptr);
bool comm_active(tb2_mesh* m);
bool comm_plan(tb2_mesh* m, CommPlan* out);
int comm_allreduce_packed(tb2_mesh* m);

// predictor + ConsistentKBC, one thread per dof
__global__ void __launch_bounds__(256) k_cd_predictor(int64_t ndof, double dt, double* __restrict__ d, double* __restrict__ v,
                                                     double* __restrict__ a, const unsigned char* __restrict__ code,
                                                     const double* __restrict__ bcval, double value_scale)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ndof) return;
    const unsigned char c = code[i];
    double di = d[i], vi = v[i];
    const double ai = a[i];
    cd_predict(dt, di, vi, ai);
    if (c == TB2_BC_FIX) { di = 0.0; vi = 0.0; }
    else if (c == TB2_BC_DSP) di = value_scale * bcval[i];
    d[i] = di;
    v[i] = vi;
    a[i] = 0.0;
}

// node kernel (see cd_node_update_one).  skip_slot (multi-GPU): nodes with skip_slot[n] >= 0 lie on the partition interface and
// are updated by k_cd_interface_update once the summed force has arrived.
template <bool GATHER, bool NEXT_PREDICTOR>
__global__ void __launch_bounds__(256, 6) k_cd_node_update(int64_t n_begin, int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                       const int4* __restrict__ inc8, const double* __restrict__ fe, int64_t stride,
                                                       const StepConsts sc, const NodeArrays na, const int* __restrict__ skip_slot)
{
    const int64_t n = n_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    if (skip_slot && skip_slot[n] >= 0) return;
    cd_node_update_one<GATHER, NEXT_PREDICTOR>(n, inc_ptr, inc, inc8, fe, stride, sc, na);
}

// multi-GPU, step 1: the partial internal force of this rank on its interface nodes, summed in the same ascending
// element order as everywhere else, written straight into the packed global interface vector (zeroed beforehand)
__global__ void __launch_bounds__(256) k_gather_pack(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                    const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                    const double* __restrict__ fe, int64_t stride, double* __restrict__ packed)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_if) return;
    const int64_t n = nodes[k];
    const int k0 = inc_ptr[n], k1 = inc_ptr[n + 1];
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int q = k0; q < k1; q++) {
        const int ent = __ldg(inc + q);
        const int64_t e = ent >> 3;
        const int a3 = 3 * (ent & 7);
        f0 += __ldg(fe + (int64_t)(a3)*stride + e);
        f1 += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
        f2 += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
    }
    double* out = packed + 3 * (int64_t)slots[k];
    out[0] = f0;
    out[1] = f1;
    out[2] = f2;
}

// multi-GPU, step 2: the node update of the interface nodes from the all-reduced force (same arithmetic as
// k_cd_node_update, so every sharer of a node computes bitwise the same d, v, a)
template <bool NEXT_PREDICTOR>
__global__ void __launch_bounds__(256) k_cd_interface_update(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                            const double* __restrict__ packed, const StepConsts sc, const NodeArrays na)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_if) return;
    const int64_t n = nodes[k];
    const double* f = packed + 3 * (int64_t)slots[k];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        const unsigned char c = na.code[q];
        double di = NEXT_PREDICTOR ? na.d[q] : 0.0, vi = na.v[q], ai;
        const double bcv = (NEXT_PREDICTOR && c == TB2_BC_DSP) ? na.bcval[q] : 0.0;
        cd_update_dof<NEXT_PREDICTOR>(sc, c, f[i], nodal_load(sc, na, q), na.minv[q], bcv, di, vi, ai);
        if (NEXT_PREDICTOR) na.d[q] = di;
        else {
            na.a[q] = ai;
            na.fint[q] = f[i];
        }
        na.v[q] = vi;
    }
}

// multi-GPU over peer memory (tb2_peer.cuh), step 1: the same partial forces, published into this rank's exchange window; the
// last CTA raises the arrival flag in every peer's window
__global__ void __launch_bounds__(256) k_peer_gather_publish(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                            const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                            const double* __restrict__ fe, int64_t stride, PeerView pv,
                                                            unsigned long long epoch, unsigned* counter)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_if) {
        const int64_t n = nodes[k];
        const int k0 = inc_ptr[n], k1 = inc_ptr[n + 1];
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        for (int q = k0; q < k1; q++) {
            const int ent = __ldg(inc + q);
            const int64_t e = ent >> 3;
            const int a3 = 3 * (ent & 7);
            f0 += __ldg(fe + (int64_t)(a3)*stride + e);
            f1 += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
            f2 += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
        }
        double* out = peer_data(pv, pv.rank, epoch) + 3 * (int64_t)slots[k];
        out[0] = f0;
        out[1] = f1;
        out[2] = f2;
    }
    peer_publish(pv, epoch, counter, kPeerFlagsOff);
}

// step 2, exchange and node update in one kernel: wait for the peers' flags, pull the sharers' partial forces over NVLink, sum
// them in rank order (bitwise the same on every sharer) and update the interface node
template <bool NEXT_PREDICTOR>
__global__ void __launch_bounds__(256) k_peer_interface_update(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                              const unsigned* __restrict__ share_mask, PeerView pv,
                                                              unsigned long long epoch, const StepConsts sc, const NodeArrays na)
{
    peer_wait(pv, epoch, kPeerFlagsOff);
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_if) return;
    const int64_t n = nodes[k];
    const unsigned mask = share_mask[k];
    const int64_t base = 3 * (int64_t)slots[k];
    double f[3];
#pragma unroll
    for (int i = 0; i < 3; i++) f[i] = peer_sum(pv, epoch, mask, base + i);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        const unsigned char c = na.code[q];
        double di = NEXT_PREDICTOR ? na.d[q] : 0.0, vi = na.v[q], ai;
        const double bcv = (NEXT_PREDICTOR && c == TB2_BC_DSP) ? na.bcval[q] : 0.0;
        cd_update_dof<NEXT_PREDICTOR>(sc, c, f[i], nodal_load(sc, na, q), na.minv[q], bcv, di, vi, ai);
        if (NEXT_PREDICTOR) na.d[q] = di;
        else {
            na.a[q] = ai;
            na.fint[q] = f[i];
        }
        na.v[q] = vi;
    }
}

// sparse refresh of the prescribed values (KBC controllers with a schedule): out[dof[k]] = value[k]
__global__ void k_scatter_values(int64_t n, const int64_t* __restrict__ dof, const double* __restrict__ value, double* __restrict__ out)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[dof[k]] = value[k];
}

// DiagonalMatrixT::Factorize (DiagonalMatrixT.cpp:267-310): reciprocal, |m| < 1e-12 left as is
__global__ void k_invert_diagonal(int64_t n, const double* __restrict__ m, double* __restrict__ minv)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = m[i];
    minv[i] = fabs(x) > 1.0e-12 ? 1.0 / x : x;
}

// FEManagerT::InitialCondition: a = minv (fext - fint) on free dofs, 0 elsewhere
__global__ void k_initial_acceleration(int64_t n, const double* __restrict__ fext, const double* __restrict__ fadd, const double* __restrict__ fint,
                                       const double* __restrict__ minv, const unsigned char* __restrict__ code, double* __restrict__ a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double load = fadd ? fext[i] + fadd[i] : fext[i];
    a[i] = code[i] ? 0.0 : (load - fint[i]) * minv[i];
}

} // namespace tb2

using namespace tb2;

static NodeArrays node_arrays(tb2_explicit* ex)
{
    NodeArrays na;
    na.fext = ex->has_fext ? ex->fext.p : nullptr;
    na.fadd = ex->contact ? ex->fadd.p : nullptr;
    na.minv = ex->minv.p;
    na.code = ex->bccode.p;
    na.bcval = ex->bcval.p;
    na.d = ex->d.p;
    na.v = ex->v.p;
    na.a = ex->a.p;
    na.fint = ex->fint.p;
    return na;
}

template <bool GATHER>
static void launch_node_update(tb2_explicit* ex, const StepConsts& sc, bool next, const int* skip, cudaStream_t st)
{
    tb2_mesh* m = ex->group->mesh;
    const int T = 256;
    const unsigned nb = (unsigned)((m->nn + T - 1) / T);
    const NodeArrays na = node_arrays(ex);
    ProfScope ps(m, kProfNodeUpdate, 1, st);
    if (next)
        k_cd_node_update<GATHER, true><<<nb, T, 0, st>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, sc, na, skip);
    else
        k_cd_node_update<GATHER, false><<<nb, T, 0, st>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, sc, na, skip);
}

// the exchange lane of one multi-GPU step on the communicator's stream, after the boundary-element sweep: partial interface
// forces -> sum over the sharers -> interface node update.  ev_packed marks the point where the scratch of the boundary
// elements is complete (the private nodes of the main lane gather from it).
static int interface_lane_peer(tb2_explicit* ex, const CommPlan& cp, const StepConsts& sc, bool next)
{
    tb2_mesh* m = ex->group->mesh;
    const int T = 256;
    const unsigned long long epoch = ++*cp.epoch;
    const unsigned nb = (unsigned)(((cp.n_if > 0 ? cp.n_if : 1) + T - 1) / T); // a rank without interface nodes still raises its flag
    const NodeArrays na = node_arrays(ex);
    {
        ProfScope ps(m, kProfComm, 1, cp.stream);
        k_peer_gather_publish<<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, m->inc_ptr.p, m->inc.p, m->fe.p, m->stride, cp.pv, epoch,
                                                      cp.counter);
    }
    TB2_CUDA(cudaEventRecord(cp.ev_packed, cp.stream));
    ProfScope ps(m, kProfNodeUpdate, 1, cp.stream);
    if (next) k_peer_interface_update<true><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.share_mask, cp.pv, epoch, sc, na);
    else k_peer_interface_update<false><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.share_mask, cp.pv, epoch, sc, na);
    return TB2_OK;
}
static int interface_lane_nccl(tb2_explicit* ex, const CommPlan& cp, const StepConsts& sc, bool next)
{
    tb2_mesh* m = ex->group->mesh;
    const int T = 256;
    {
        ProfScope ps(m, kProfComm, 2, cp.stream);
        TB2_CUDA(cudaMemsetAsync(cp.packed, 0, 3 * cp.n_glob * sizeof(double), cp.stream));
        if (cp.n_if)
            k_gather_pack<<<(unsigned)((cp.n_if + T - 1) / T), T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, m->inc_ptr.p, m->inc.p, m->fe.p,
                                                                                m->stride, cp.packed);
    }
    TB2_CUDA(cudaEventRecord(cp.ev_packed, cp.stream));
    TB2_CHECK(comm_allreduce_packed(m));
    if (cp.n_if) {
        ProfScope ps(m, kProfNodeUpdate, 1, cp.stream);
        const unsigned nb = (unsigned)((cp.n_if + T - 1) / T);
        const NodeArrays na = node_arrays(ex);
        if (next) k_cd_interface_update<true><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.packed, sc, na);
        else k_cd_interface_update<false><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.packed, sc, na);
    }
    return TB2_OK;
}

// the attached contact group's force on the current (predicted) state -> fadd; fadd is cleared only when the pair list changed.
// The two small kernels run on their own (high-priority) stream beside the element sweep: ev_state = the state is ready, ev_loads = fadd is.
static int contact_loads(tb2_explicit* ex)
{
    tb2_mesh* m = ex->group->mesh;
    if (!ex->stream_aux) {
        int lo = 0, hi = 0;
        TB2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        TB2_CUDA(cudaStreamCreateWithPriority(&ex->stream_aux, cudaStreamNonBlocking, hi));
        TB2_CUDA(cudaEventCreateWithFlags(&ex->ev_state, cudaEventDisableTiming));
        TB2_CUDA(cudaEventCreateWithFlags(&ex->ev_loads, cudaEventDisableTiming));
    }
    TB2_CUDA(cudaEventRecord(ex->ev_state, m->stream));
    TB2_CUDA(cudaStreamWaitEvent(ex->stream_aux, ex->ev_state, 0));
    if (ex->contact_version != contact_version(ex->contact)) {
        TB2_CUDA(cudaMemsetAsync(ex->fadd.p, 0, (size_t)m->nn * 3 * sizeof(double), ex->stream_aux));
        ex->contact_version = contact_version(ex->contact);
    }
    TB2_CHECK(contact_form_touched(ex->contact, 1.0 /* nExplicitCD: FormKd = 1 */, ex->d.p, ex->v.p, ex->fadd.p, ex->stream_aux));
    TB2_CUDA(cudaEventRecord(ex->ev_loads, ex->stream_aux));
    return TB2_OK;
}

// nsteps explicit steps on the device-resident state; fs / vs: per-step scales of fext and of the prescribed displacements
static int explicit_steps(tb2_explicit* ex, double dt, int nsteps, const double* fs, const double* vs)
{
    tb2_group* g = ex->group;
    tb2_mesh* m = g->mesh;
    const int64_t ndof = 3 * m->nn;
    const int T = 256;
    if (nsteps <= 0) return TB2_OK;
    {
        ProfScope ps(m, kProfPredictor);
        k_cd_predictor<<<(unsigned)((ndof + T - 1) / T), T, 0, m->stream>>>(ndof, dt, ex->d.p, ex->v.p, ex->a.p, ex->bccode.p, ex->bcval.p,
                                                                           vs ? vs[0] : 1.0);
    }
    CommPlan cp;
    const bool multi = comm_plan(m, &cp);
    if (multi && !m->ev_join) {
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_k5, cudaEventDisableTiming));
    }
    if (multi) { // the comm lane starts after everything queued on the mesh stream so far
        TB2_CUDA(cudaEventRecord(m->ev_k5, m->stream));
    }
    for (int s = 0; s < nsteps; s++) {
        const bool next = s + 1 < nsteps;
        const StepConsts sc{dt, fs ? fs[s] : 1.0, (next && vs) ? vs[s + 1] : 1.0};
        if (!multi) {
            if (ex->contact) TB2_CHECK(contact_loads(ex)); // on the predicted d, v -- what the group's RHSDriver sees in FEManagerT::FormRHS
            TB2_CHECK(launch_element_forces(g, ex->d.p, nullptr, 0));
            if (ex->contact) TB2_CUDA(cudaStreamWaitEvent(m->stream, ex->ev_loads, 0));
            launch_node_update<true>(ex, sc, next, nullptr, m->stream);
            continue;
        }
        // comm lane: boundary elements -> partial interface forces -> sum over the sharers (peer memory, else ncclAllReduce) -> interface nodes
        TB2_CUDA(cudaStreamWaitEvent(cp.stream, m->ev_k5, 0)); // the private nodes the boundary elements read are up to date
        TB2_CHECK(launch_element_forces_range(g, ex->d.p, nullptr, 0, 0, cp.nb, cp.stream, cp.belems));
        TB2_CHECK(cp.peer ? interface_lane_peer(ex, cp, sc, next) : interface_lane_nccl(ex, cp, sc, next));
        TB2_CUDA(cudaEventRecord(cp.ev_done, cp.stream));
        // main lane: the other elements beside the all-reduce, then the private nodes (they gather boundary-element forces too)
        TB2_CHECK(launch_element_forces_range(g, ex->d.p, nullptr, 0, 0, m->ne, m->stream, nullptr, cp.belem_flag));
        TB2_CUDA(cudaStreamWaitEvent(m->stream, cp.ev_packed, 0));
        launch_node_update<true>(ex, sc, next, cp.node_slot, m->stream);
        TB2_CUDA(cudaEventRecord(m->ev_k5, m->stream));
        // the next sweep of the main lane reads no interface node, but the one after the last step must see them all
        if (!next) TB2_CUDA(cudaStreamWaitEvent(m->stream, cp.ev_done, 0));
    }
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

extern "C" {

int tb2_explicit_create(tb2_group* g, tb2_explicit** out)
{
    TB2_ARG(g && out);
    if (g->mat.kind == TB2_J2_SIMO) {
        set_error("explicit central difference with J2Simo3D is not supported (needs per-step history commit)");
        return TB2_ERR_ARG;
    }
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    tb2_explicit* ex = new tb2_explicit;
    ex->group = g;
    ex->device = m->device;
    const size_t n = 3 * m->nn;
    cudaError_t e = cudaSuccess;
    DevBuf<double>* bufs[] = {&ex->d, &ex->v, &ex->a, &ex->mass, &ex->minv, &ex->fext, &ex->fint, &ex->bcval};
    for (auto* b : bufs) {
        if (e == cudaSuccess) e = b->alloc(n);
        if (e == cudaSuccess) e = cudaMemsetAsync(b->p, 0, n * sizeof(double), m->stream);
    }
    if (e == cudaSuccess) e = ex->bccode.alloc(n);
    if (e == cudaSuccess) e = cudaMemsetAsync(ex->bccode.p, 0, n, m->stream);
    if (e != cudaSuccess) {
        delete ex;
        return cuda_fail(e, "explicit state allocation", __FILE__, __LINE__);
    }
    int s = tb2_form_lumped_mass(g, ex->mass.p);
    if (s == TB2_OK) s = tb2_comm_sum_interface(m, ex->mass.p); // interface-node mass is summed once over the sharers
    if (s == TB2_OK) {
        k_invert_diagonal<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>((int64_t)n, ex->mass.p, ex->minv.p);
        s = tb2_group_status(g, nullptr);
    }
    if (s != TB2_OK) {
        delete ex;
        return s;
    }
    *out = ex;
    return TB2_OK;
}

int tb2_explicit_destroy(tb2_explicit* ex)
{
    if (!ex) return TB2_OK;
    // a host may destroy the mesh / group first (Tahoe deletes its element groups before its solvers): use the copies taken at creation
    DeviceGuard dg(ex->device);
    if (cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError(); // not the mesh stream: it may have been destroyed with the mesh
    if (ex->stream_copy) {
        cudaStreamSynchronize(ex->stream_copy);
        cudaStreamDestroy(ex->stream_copy);
        cudaFreeHost(ex->h_status);
        for (int b = 0; b < 2; b++) {
            cudaEventDestroy(ex->ev_snap[b]);
            cudaEventDestroy(ex->ev_copied[b]);
        }
    }
    if (ex->stream_aux) {
        cudaStreamDestroy(ex->stream_aux);
        cudaEventDestroy(ex->ev_state);
        cudaEventDestroy(ex->ev_loads);
    }
    delete ex;
    return TB2_OK;
}

int tb2_explicit_set_state(tb2_explicit* ex, const double* h_d, const double* h_v, const double* h_a)
{
    TB2_ARG(ex);
    ex->contact_searched = false; // a new configuration: an attached group with surfaces is searched again before the next step
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    if (h_d) TB2_CUDA(cudaMemcpyAsync(ex->d.p, h_d, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_v) TB2_CUDA(cudaMemcpyAsync(ex->v.p, h_v, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_a) TB2_CUDA(cudaMemcpyAsync(ex->a.p, h_a, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}
int tb2_explicit_get_state(tb2_explicit* ex, double* h_d, double* h_v, double* h_a)
{
    TB2_ARG(ex);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    if (h_d) TB2_CUDA(cudaMemcpyAsync(h_d, ex->d.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    if (h_v) TB2_CUDA(cudaMemcpyAsync(h_v, ex->v.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    if (h_a) TB2_CUDA(cudaMemcpyAsync(h_a, ex->a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}
int tb2_explicit_set_bc(tb2_explicit* ex, const uint8_t* h_code, const double* h_value, const double* h_fext)
{
    TB2_ARG(ex);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t n = 3 * m->nn;
    if (h_code) TB2_CUDA(cudaMemcpyAsync(ex->bccode.p, h_code, n, cudaMemcpyHostToDevice, m->stream));
    if (h_value) TB2_CUDA(cudaMemcpyAsync(ex->bcval.p, h_value, n * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    if (h_fext) {
        TB2_CUDA(cudaMemcpyAsync(ex->fext.p, h_fext, n * sizeof(double), cudaMemcpyHostToDevice, m->stream));
        ex->has_fext = false; // an all-zero external force is not read by the node kernel
        for (size_t i = 0; i < n && !ex->has_fext; i++) ex->has_fext = h_fext[i] != 0.0;
    }
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}

int tb2_explicit_update_bc_values(tb2_explicit* ex, int64_t count, const int64_t* h_dofs, const double* h_values)
{
    TB2_ARG(ex && count >= 0 && (count == 0 || (h_dofs && h_values)));
    if (count == 0) return TB2_OK;
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const int64_t ndof = 3 * m->nn;
    for (int64_t k = 0; k < count; k++) TB2_ARG(h_dofs[k] >= 0 && h_dofs[k] < ndof);
    DevBuf<int64_t> idx;
    DevBuf<double> val;
    TB2_CUDA(idx.alloc(count));
    TB2_CUDA(val.alloc(count));
    TB2_CUDA(cudaMemcpyAsync(idx.p, h_dofs, count * sizeof(int64_t), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(val.p, h_values, count * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    k_scatter_values<<<(unsigned)((count + 255) / 256), 256, 0, m->stream>>>(count, idx.p, val.p, ex->bcval.p);
    TB2_CUDA(cudaGetLastError());
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}

int tb2_explicit_attach_contact(tb2_explicit* ex, tb2_contact* contact)
{
    TB2_ARG(ex);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    if (contact && comm_active(m)) {
        set_error("tb2_explicit_attach_contact: contact pairs across ranks are not supported (single-GPU runs only)");
        return TB2_ERR_ARG;
    }
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    ex->contact = contact;
    ex->contact_version = ~0ull;
    ex->contact_searched = false;
    if (contact && !ex->fadd.p) TB2_CUDA(ex->fadd.alloc((size_t)m->nn * 3));
    return TB2_OK;
}

int tb2_explicit_initial_condition(tb2_explicit* ex)
{
    TB2_ARG(ex);
    tb2_group* g = ex->group;
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CHECK(tb2_form_internal_force(g, ex->d.p, nullptr, 0, ex->fint.p));
    TB2_CHECK(tb2_comm_sum_interface(m, ex->fint.p));
    const int64_t n = 3 * m->nn;
    if (ex->contact) {
        TB2_CHECK(contact_loads(ex));
        TB2_CUDA(cudaStreamWaitEvent(m->stream, ex->ev_loads, 0));
    }
    k_initial_acceleration<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(n, ex->fext.p, ex->contact ? ex->fadd.p : nullptr, ex->fint.p, ex->minv.p,
                                                                               ex->bccode.p, ex->a.p);
    TB2_CUDA(cudaGetLastError());
    return tb2_group_status(g, nullptr);
}

// An attached contact group that carries its surfaces (tb2_contact_set_surfaces) is searched on the device after every step, on the
// corrected displacements, as LinearSolver::Solve relaxes the system after its update (LinearSolver.cpp:76-90 -> ContactT::RelaxSystem):
// the steps then run one at a time (the next predictor cannot be fused into the node kernel), and once before the first step.
static int explicit_steps_searching(tb2_explicit* ex, double dt, int nsteps, const double* fs, const double* vs)
{
    if (!ex->contact_searched) {
        TB2_CHECK(tb2_contact_search(ex->contact, ex->d.p, nullptr));
        ex->contact_searched = true;
    }
    for (int s = 0; s < nsteps; s++) {
        TB2_CHECK(explicit_steps(ex, dt, 1, fs ? fs + s : nullptr, vs ? vs + s : nullptr));
        TB2_CHECK(tb2_contact_search(ex->contact, ex->d.p, nullptr));
    }
    return TB2_OK;
}

int tb2_explicit_run(tb2_explicit* ex, double dt, int nsteps, const double* h_fext_scale, const double* h_value_scale)
{
    TB2_ARG(ex && nsteps >= 0);
    DeviceGuard dg(ex->group->mesh->device);
    if (ex->contact && tb2_contact_has_surfaces(ex->contact)) TB2_CHECK(explicit_steps_searching(ex, dt, nsteps, h_fext_scale, h_value_scale));
    else TB2_CHECK(explicit_steps(ex, dt, nsteps, h_fext_scale, h_value_scale));
    return tb2_group_status(ex->group, nullptr);
}

int tb2_explicit_step_host(tb2_explicit* ex, double dt, double* h_d, double* h_v, double* h_a)
{
    TB2_ARG(ex && h_d && h_v && h_a);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    TB2_CUDA(cudaMemcpyAsync(ex->d.p, h_d, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(ex->v.p, h_v, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(ex->a.p, h_a, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(explicit_steps(ex, dt, 1, nullptr, nullptr));
    TB2_CUDA(cudaMemcpyAsync(h_d, ex->d.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_v, ex->v.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_a, ex->a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    return tb2_group_status(ex->group, nullptr);
}

int tb2_explicit_run_async(tb2_explicit* ex, double dt, int nsteps, const double* h_fext_scale, const double* h_value_scale, double* h_d,
                           int* ticket)
{
    TB2_ARG(ex && nsteps >= 0 && h_d && ticket);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    if (!ex->stream_copy) {
        TB2_CUDA(cudaStreamCreateWithFlags(&ex->stream_copy, cudaStreamNonBlocking));
        TB2_CUDA(cudaHostAlloc((void**)&ex->h_status, 4 * sizeof(unsigned long long), cudaHostAllocDefault));
        for (int b = 0; b < 2; b++) {
            TB2_CUDA(ex->dsnap[b].alloc(3 * m->nn));
            TB2_CUDA(cudaEventCreateWithFlags(&ex->ev_snap[b], cudaEventDisableTiming));
            TB2_CUDA(cudaEventCreateWithFlags(&ex->ev_copied[b], cudaEventDisableTiming));
            TB2_CUDA(cudaEventRecord(ex->ev_copied[b], ex->stream_copy));
        }
    }
    const int t = ex->tickets++;
    const int b = t & 1;
    // the snapshot buffer is free once the copy that last read it has finished
    TB2_CUDA(cudaStreamWaitEvent(m->stream, ex->ev_copied[b], 0));
    if (ex->contact && tb2_contact_has_surfaces(ex->contact)) TB2_CHECK(explicit_steps_searching(ex, dt, nsteps, h_fext_scale, h_value_scale));
    else TB2_CHECK(explicit_steps(ex, dt, nsteps, h_fext_scale, h_value_scale));
    TB2_CUDA(cudaMemcpyAsync(ex->dsnap[b].p, ex->d.p, bytes, cudaMemcpyDeviceToDevice, m->stream));
    TB2_CUDA(cudaEventRecord(ex->ev_snap[b], m->stream));
    TB2_CUDA(cudaStreamWaitEvent(ex->stream_copy, ex->ev_snap[b], 0));
    TB2_CUDA(cudaMemcpyAsync(h_d, ex->dsnap[b].p, bytes, cudaMemcpyDeviceToHost, ex->stream_copy));
    // the element status of everything up to these steps rides along (tb2_explicit_wait must not queue behind a later copy)
    TB2_CUDA(cudaMemcpyAsync(ex->h_status + 2 * b, ex->group->status.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ex->stream_copy));
    TB2_CUDA(cudaEventRecord(ex->ev_copied[b], ex->stream_copy));
    *ticket = t;
    return TB2_OK;
}

int tb2_explicit_wait(tb2_explicit* ex, int ticket)
{
    TB2_ARG(ex && ticket >= 0 && ticket < ex->tickets);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    // only the latest user of a buffer can still be in flight: waiting for it covers the ticket asked for
    TB2_CUDA(cudaEventSynchronize(ex->ev_copied[ticket & 1]));
    if (ex->h_status[2 * (ticket & 1)]) return tb2_group_status(ex->group, nullptr);
    return TB2_OK;
}

double* tb2_explicit_device_array(tb2_explicit* ex, int which)
{
    if (!ex) return nullptr;
    switch (which) {
    case 0: return ex->d.p;
    case 1: return ex->v.p;
    case 2: return ex->a.p;
    case 3: return ex->mass.p;
    case 4: return ex->fext.p;
    case 5: return ex->fint.p;
    }
    return nullptr;
}

} // extern "C"
