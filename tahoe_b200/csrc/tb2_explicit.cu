// tb2_explicit.cu -- K5: explicit central-difference update on device-resident fields.
//
// Reference order of one step:
//   FieldT::InitStep -> nExplicitCD::Predictor (nExplicitCD.cpp:72-96): d += dt v + dt^2/2 a ; v += dt/2 a ; a = 0,
//                       then ConsistentKBC on prescribed dofs (nExplicitCD.cpp:20-69)
//   LinearSolver::Solve (LinearSolver.cpp:37-101): R = fext - fint(d) ; update = M^-1 R (DiagonalMatrixT.cpp:267-323)
//   FieldT::AssembleUpdate (FieldT.cpp:531-556: prescribed dofs get 0) ; nExplicitCD::Corrector (:98-139): v += dt/2 upd ; a += upd
// On the device this is two launches per step: the element sweep (K1) and one node kernel that gathers the element
// forces, forms R, applies M^-1, the corrector and -- when another step follows -- the next step's predictor, so d, v, a
// are read and written once per step (192 B/node/step).
#include <chrono>
#include <cstdlib>

#include "tb2_internal.h"
#include "tb2_node_update.cuh"

namespace tb2 {

int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration);
int launch_node_gather(tb2_mesh* m, double* d_out, bool per_dof);
int launch_element_forces_range(tb2_group* g, const double* d_u, const double* d_ul, int iteration, int64_t e0, int64_t e1, cudaStream_t st,
                                const int* d_elist = nullptr, const unsigned char* d_skip = nullptr);
bool fused_step_supported(tb2_group* g);
int launch_fused_forces_nodes(tb2_group* g, const double* d_u, int64_t e0, int64_t e1, cudaStream_t st, const NodeArgs& q,
                              const unsigned char* d_skip);
bool comm_active(tb2_mesh* m);
bool comm_plan(tb2_mesh* m, CommPlan* out);
int comm_allreduce_packed(tb2_mesh* m);

// predictor + ConsistentKBC, one thread per dof
__global__ void __launch_bounds__(256) k_cd_predictor(int64_t ndof, double dt, double* __restrict__ d, double* __restrict__ v,
                                                     double* __restrict__ a, const unsigned char* __restrict__ code,
                                                     const double* __restrict__ bcval, double value_scale,
                                                     double* __restrict__ host_d = nullptr)
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
    if (host_d) host_d[i] = di; // host-buffer step: d is final for this step, it goes straight back over PCIe (mapped pinned memory)
}

// node kernel: gather fint, R = s*fext - fint, upd = minv*R on free dofs, corrector; optionally the next predictor.
// a_in is the acceleration left by the predictor (0 on every dof), kept as an input for generality (a += upd).
// GATHER = false: fint already holds the (interface-summed) internal force (multi-GPU path)
// skip_slot (multi-GPU overlap): nodes with skip_slot[n] >= 0 lie on the partition interface and are updated by
// k_cd_interface_update once the summed force has arrived.
// GATHER reads the node's incidence from the fixed-width table inc8 (one 32-byte load) and issues every load of the node --
// the <= 8 x 3 element forces and the nodal fields -- before the first use: the kernel runs beside the element sweep, which
// holds most of the registers of every SM, so what counts is how briefly a node-kernel CTA occupies its slot.
template <bool GATHER, bool NEXT_PREDICTOR>
__global__ void __launch_bounds__(256) k_cd_node_update(int64_t n_begin, int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                       const int4* __restrict__ inc8, const double* __restrict__ fe, int64_t stride, double dt,
                                                       double fext_scale, double next_value_scale, const double* __restrict__ fext,
                                                       const double* __restrict__ minv, const unsigned char* __restrict__ code,
                                                       const double* __restrict__ bcval, double* __restrict__ d,
                                                       double* __restrict__ v, double* __restrict__ a, double* __restrict__ fint,
                                                       const int* __restrict__ skip_slot = nullptr)
{
    const int64_t n = n_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    cd_node_update_one<GATHER, NEXT_PREDICTOR>(n, inc_ptr, inc, inc8, fe, stride, dt, fext_scale, next_value_scale, fext, minv, code, bcval, d, v, a,
                                               fint, skip_slot);
}

// Host-buffer step: k_cd_node_update<gather, no next predictor> that also writes v and a to the caller's (mapped, pinned) host
// arrays.  The block's 3 x 256 results are staged in shared memory so that every warp store to host memory is a contiguous
// 256-byte run (posted PCIe writes; the copy engines stay free for the host -> device direction).
__global__ void __launch_bounds__(256) k_cd_node_update_hostout(int64_t n_begin, int64_t nn, const int* __restrict__ inc_ptr,
                                                               const int* __restrict__ inc, const double* __restrict__ fe, int64_t stride,
                                                               double dt, double fext_scale, const double* __restrict__ fext,
                                                               const double* __restrict__ minv, const unsigned char* __restrict__ code,
                                                               double* __restrict__ v, double* __restrict__ a, double* __restrict__ fint,
                                                               double* __restrict__ host_v, double* __restrict__ host_a)
{
    __shared__ double sv[768], sa[768];
    const int64_t nb0 = n_begin + blockIdx.x * (int64_t)blockDim.x;
    const int64_t n = nb0 + threadIdx.x;
    if (n < nn) {
        double f[3] = {0.0, 0.0, 0.0};
        const int k0 = inc_ptr[n], k1 = inc_ptr[n + 1];
        for (int k = k0; k < k1; k++) {
            const int ent = __ldg(inc + k);
            const int64_t e = ent >> 3;
            const int a3 = 3 * (ent & 7);
            f[0] += __ldg(fe + (int64_t)(a3)*stride + e);
            f[1] += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
            f[2] += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int64_t q = 3 * n + i;
            const unsigned char c = code[q];
            const double R = __dsub_rn(__dmul_rn(fext_scale, fext[q]), f[i]);
            const double upd = c ? 0.0 : __dmul_rn(R, minv[q]);
            double vi = v[q], ai = a[q];
            cd_correct(dt, vi, ai, upd);
            fint[q] = f[i];
            v[q] = vi;
            a[q] = ai;
            sv[3 * threadIdx.x + i] = vi;
            sa[3 * threadIdx.x + i] = ai;
        }
    }
    __syncthreads();
    const int64_t nb1 = nb0 + blockDim.x < nn ? nb0 + blockDim.x : nn;
    const int cnt = (int)(3 * (nb1 - nb0));
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        host_v[3 * nb0 + t] = sv[t];
        host_a[3 * nb0 + t] = sa[t];
    }
}

// multi-GPU overlap, step 1: the partial internal force of this rank on its interface nodes, summed in the same ascending
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

// multi-GPU overlap, step 2: the node update of the interface nodes from the all-reduced force (same arithmetic as
// k_cd_node_update, so every sharer of a node computes bitwise the same d, v, a)
template <bool NEXT_PREDICTOR>
__global__ void __launch_bounds__(256) k_cd_interface_update(int64_t n_if, const int* __restrict__ nodes, const int* __restrict__ slots,
                                                            const double* __restrict__ packed, double dt, double fext_scale,
                                                            double next_value_scale, const double* __restrict__ fext,
                                                            const double* __restrict__ minv, const unsigned char* __restrict__ code,
                                                            const double* __restrict__ bcval, double* __restrict__ d,
                                                            double* __restrict__ v, double* __restrict__ a, double* __restrict__ fint)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_if) return;
    const int64_t n = nodes[k];
    const double* f = packed + 3 * (int64_t)slots[k];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        const unsigned char c = code[q];
        const double R = __dsub_rn(__dmul_rn(fext_scale, fext[q]), f[i]);
        const double upd = c ? 0.0 : __dmul_rn(R, minv[q]);
        double vi = v[q], ai = a[q];
        cd_correct(dt, vi, ai, upd);
        fint[q] = f[i];
        if (NEXT_PREDICTOR) {
            double di = d[q];
            cd_predict(dt, di, vi, ai);
            ai = 0.0;
            if (c == TB2_BC_FIX) { di = 0.0; vi = 0.0; }
            else if (c == TB2_BC_DSP) di = next_value_scale * bcval[q];
            d[q] = di;
        }
        v[q] = vi;
        a[q] = ai;
    }
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
__global__ void k_initial_acceleration(int64_t n, const double* __restrict__ fext, const double* __restrict__ fint,
                                       const double* __restrict__ minv, const unsigned char* __restrict__ code, double* __restrict__ a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = code[i] ? 0.0 : (fext[i] - fint[i]) * minv[i];
}

} // namespace tb2

using namespace tb2;

// Slab pipeline (single GPU).  K1 is FP64-pipe bound and K5 is HBM bound, and K5 of a node chunk only needs K1 of the element
// chunks touching it, so the two kernels of a step -- and of consecutive steps -- are overlapped: element chunks run in index
// order on the mesh stream, node chunks in index order on a second stream, ordered by events:
//   K5(s, nc) waits for K1(s, last element chunk touching nc);   K1(s+1, ec) waits for K5(s, last node chunk touched by ec).
// The same two conditions cover the write-after-read hazards on d (K5 writes what K1 reads) and on the force scratch.
// Arithmetic and summation order are those of the serial schedule: results are bitwise identical (tested).
//
// Multi-GPU (element-partitioned mesh).  The interface exchange is a third, independent lane beside the same pipeline:
//   comm stream    K1 over the elements touching interface nodes (index list) -> k_gather_pack of this rank's partial interface
//                  forces into the packed vector -> ncclAllReduce -> k_cd_interface_update of the interface nodes;
//   main / second  the slab pipeline, with K1 leaving out the boundary elements and K5 leaving out the interface nodes.
// Interface nodes are written by the comm lane only and read by the boundary elements only, so the lanes meet in two places:
// the boundary sweep of step s+1 waits for the last K5 chunk of step s (it reads private nodes too), and the K5 chunks of a
// step wait for that step's boundary sweep (they gather its element forces and overwrite the d it reads).
static int explicit_steps_pipelined(tb2_explicit* ex, double dt, int nsteps, const double* fs, const double* vs)
{
    tb2_group* g = ex->group;
    tb2_mesh* m = g->mesh;
    CommPlan cp;
    const bool multi = comm_plan(m, &cp);
    const int* skip = multi ? cp.node_slot : nullptr;
    const int C = (int)m->pipe_e0.size() - 1;
    const int64_t ndof = 3 * m->nn;
    const int T = 256;
    // node-kernel CTA size in the pipeline (experiment knob; 64 / 128 / 256 measured within 3 % of each other on B200)
    static const int T5 = (getenv("TB2_K5_THREADS") && atoi(getenv("TB2_K5_THREADS")) >= 32) ? atoi(getenv("TB2_K5_THREADS")) : 256;
    if (!m->stream2) {
        // the HBM-bound node kernels outrank the FP64-bound element sweep they run beside (their CTAs are small and short)
        int prio_lo = 0, prio_hi = 0;
        TB2_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const char* p2 = getenv("TB2_K5_PRIORITY"); // experiment knob: 1 = run the node lane at elevated priority
        TB2_CUDA(cudaStreamCreateWithPriority(&m->stream2, cudaStreamNonBlocking, (p2 && p2[0] == '1') ? (prio_hi < prio_lo ? prio_hi + 1 : prio_hi) : prio_lo));
        m->ev_k1.resize(C);
        m->ev_k5.resize(C);
        for (int c = 0; c < C; c++) {
            TB2_CUDA(cudaEventCreateWithFlags(&m->ev_k1[c], cudaEventDisableTiming));
            TB2_CUDA(cudaEventCreateWithFlags(&m->ev_k5[c], cudaEventDisableTiming));
        }
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
    {
        ProfScope ps(m, kProfPredictor);
        k_cd_predictor<<<(unsigned)((ndof + T - 1) / T), T, 0, m->stream>>>(ndof, dt, ex->d.p, ex->v.p, ex->a.p, ex->bccode.p, ex->bcval.p,
                                                                           vs ? vs[0] : 1.0);
    }
    // Experiment knob TB2_FUSED_STEP=1 (single GPU): fused element + node launches on ONE stream (see k_fused_force_update in
    // tb2_elements.cu).  Bitwise the same fields (tested), but measured slower on B200, 1M elements: 0.299 ms/step with 4 slabs
    // (0.333 / 0.376 / 0.512 with 8 / 16 / 32) against 0.266 for the two-stream form below -- a node CTA holds a third of an SM's
    // registers under the sweep's 168-register allocation while it waits on memory, and back-to-back launches on one stream lose
    // the tail filling the second stream provides.  Off by default.
    static const bool fused_on = getenv("TB2_FUSED_STEP") && getenv("TB2_FUSED_STEP")[0] == '1';
    if (!multi && fused_on && C >= 3 && fused_step_supported(g)) {
        NodeArgs base{};
        base.inc_ptr = m->inc_ptr.p;
        base.inc = m->inc.p;
        base.inc8 = (const int4*)m->inc8.p;
        base.fe = m->fe.p;
        base.stride = m->stride;
        base.dt = dt;
        base.fext = ex->fext.p;
        base.minv = ex->minv.p;
        base.code = ex->bccode.p;
        base.bcval = ex->bcval.p;
        base.d = ex->d.p;
        base.v = ex->v.p;
        base.a = ex->a.p;
        base.fint = ex->fint.p;
        base.skip_slot = nullptr;
        NodeArgs pending = base; // node slabs whose forces are complete, waiting for a launch to ride on
        int pending_first = -1;
        auto flush_alone = [&]() -> int {
            if (pending.n1 > pending.n0) {
                ProfScope ps(m, kProfNodeUpdate);
                const unsigned nb = (unsigned)((pending.n1 - pending.n0 + T - 1) / T);
                if (pending.next_predictor)
                    k_cd_node_update<true, true><<<nb, T, 0, m->stream>>>(pending.n0, pending.n1, base.inc_ptr, base.inc, base.inc8, base.fe, base.stride, dt,
                                                                         pending.fext_scale, pending.next_value_scale, base.fext, base.minv, base.code,
                                                                         base.bcval, base.d, base.v, base.a, base.fint, nullptr);
                else
                    k_cd_node_update<true, false><<<nb, T, 0, m->stream>>>(pending.n0, pending.n1, base.inc_ptr, base.inc, base.inc8, base.fe, base.stride, dt,
                                                                          pending.fext_scale, 1.0, base.fext, base.minv, base.code, base.bcval, base.d,
                                                                          base.v, base.a, base.fint, nullptr);
            }
            pending.n0 = pending.n1 = 0;
            return TB2_OK;
        };
        for (int s = 0; s < nsteps; s++) {
            const double fsc = fs ? fs[s] : 1.0;
            int nc = 0;
            for (int c = 0; c < C; c++) {
                // the nodes carried over from the previous step must not be read by this element slab (c = 0 only)
                if (c == 0 && pending.n1 > pending.n0 && pending_first <= m->pipe_nmax_of_ec[0]) TB2_CHECK(flush_alone());
                TB2_CHECK(launch_fused_forces_nodes(g, ex->d.p, m->pipe_e0[c], m->pipe_e0[c + 1], m->stream, pending, nullptr));
                pending.n0 = pending.n1 = 0;
                const int first = nc;
                for (; nc < C && m->pipe_emax_of_nc[nc] <= c; nc++) {}
                if (nc > first) {
                    pending = base;
                    pending.n0 = m->pipe_n0[first];
                    pending.n1 = m->pipe_n0[nc];
                    pending.fext_scale = fsc;
                    pending.next_predictor = s + 1 < nsteps ? 1 : 0;
                    pending.next_value_scale = (s + 1 < nsteps && vs) ? vs[s + 1] : 1.0;
                    pending_first = first;
                }
            }
            if (nc < C) { // node slabs no element slab closes (cannot happen with monotone maps; kept for safety)
                TB2_CHECK(flush_alone());
                pending = base;
                pending.n0 = m->pipe_n0[nc];
                pending.n1 = m->pipe_n0[C];
                pending.fext_scale = fsc;
                pending.next_predictor = s + 1 < nsteps ? 1 : 0;
                pending.next_value_scale = (s + 1 < nsteps && vs) ? vs[s + 1] : 1.0;
                pending_first = nc;
            }
        }
        TB2_CHECK(flush_alone());
        TB2_CUDA(cudaGetLastError());
        return TB2_OK;
    }
    // Experiment knob TB2_K1_STREAMS=2: consecutive element chunks alternate between two streams (they are independent of each
    // other; only the events order them against the node chunks), so that chunk c+1 fills the SMs the last wave of chunk c is
    // leaving.  Measured on B200 (1M elements): 0.269 ms/step against 0.263 with one stream -- off by default.
    static const bool two_k1_streams = getenv("TB2_K1_STREAMS") && getenv("TB2_K1_STREAMS")[0] == '2';
    if (two_k1_streams && !m->stream1b) {
        TB2_CUDA(cudaStreamCreateWithFlags(&m->stream1b, cudaStreamNonBlocking));
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_join1b, cudaEventDisableTiming));
    }
    TB2_CUDA(cudaEventRecord(m->ev_join, m->stream));
    TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_join, 0));
    if (two_k1_streams) TB2_CUDA(cudaStreamWaitEvent(m->stream1b, m->ev_join, 0));
    const auto t_enqueue0 = std::chrono::steady_clock::now();
    for (int s = 0; s < nsteps; s++) {
        const double fsc = fs ? fs[s] : 1.0;
        int nc = 0;
        if (multi) {
            // comm lane of step s (its previous interface update is ahead of it on the same stream)
            TB2_CUDA(cudaStreamWaitEvent(cp.stream, s > 0 ? m->ev_k5[C - 1] : m->ev_join, 0));
            TB2_CHECK(launch_element_forces_range(g, ex->d.p, nullptr, 0, 0, cp.nb, cp.stream, cp.belems));
            {
                ProfScope ps(m, kProfComm, 2, cp.stream);
                TB2_CUDA(cudaMemsetAsync(cp.packed, 0, 3 * cp.n_glob * sizeof(double), cp.stream));
                if (cp.n_if)
                    k_gather_pack<<<(unsigned)((cp.n_if + T - 1) / T), T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, m->inc_ptr.p, m->inc.p,
                                                                                        m->fe.p, m->stride, cp.packed);
            }
            TB2_CUDA(cudaEventRecord(cp.ev_packed, cp.stream));
            TB2_CHECK(comm_allreduce_packed(m));
            if (cp.n_if) {
                ProfScope ps(m, kProfNodeUpdate, 1, cp.stream);
                const unsigned nb = (unsigned)((cp.n_if + T - 1) / T);
                if (s + 1 < nsteps)
                    k_cd_interface_update<true><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.packed, dt, fsc, vs ? vs[s + 1] : 1.0,
                                                                        ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p, ex->v.p,
                                                                        ex->a.p, ex->fint.p);
                else
                    k_cd_interface_update<false><<<nb, T, 0, cp.stream>>>(cp.n_if, cp.nodes, cp.slots, cp.packed, dt, fsc, 1.0, ex->fext.p,
                                                                         ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p, ex->v.p, ex->a.p,
                                                                         ex->fint.p);
            }
            TB2_CUDA(cudaEventRecord(cp.ev_done, cp.stream));
            TB2_CUDA(cudaStreamWaitEvent(m->stream2, cp.ev_packed, 0)); // K5 chunks of this step follow the boundary sweep
        }
        for (int c = 0; c < C; c++) {
            cudaStream_t sk = (two_k1_streams && (c & 1)) ? m->stream1b : m->stream;
            if (s > 0 && m->pipe_nmax_of_ec[c] >= 0) TB2_CUDA(cudaStreamWaitEvent(sk, m->ev_k5[m->pipe_nmax_of_ec[c]], 0));
            TB2_CHECK(launch_element_forces_range(g, ex->d.p, nullptr, 0, m->pipe_e0[c], m->pipe_e0[c + 1], sk, nullptr,
                                                  multi ? cp.belem_flag : nullptr));
            TB2_CUDA(cudaEventRecord(m->ev_k1[c], sk));
            for (; nc < C && m->pipe_emax_of_nc[nc] <= c; nc++) {
                const int64_t n0 = m->pipe_n0[nc], n1 = m->pipe_n0[nc + 1];
                if (m->pipe_emax_of_nc[nc] >= 0) {
                    // "all element chunks <= emax are done": each of the two element streams runs its chunks in order
                    TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_k1[m->pipe_emax_of_nc[nc]], 0));
                    if (two_k1_streams && m->pipe_emax_of_nc[nc] >= 1)
                        TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_k1[m->pipe_emax_of_nc[nc] - 1], 0));
                }
                if (n1 > n0) {
                    ProfScope ps(m, kProfNodeUpdate, 1, m->stream2);
                    const unsigned nb = (unsigned)((n1 - n0 + T5 - 1) / T5);
                    if (s + 1 < nsteps)
                        k_cd_node_update<true, true><<<nb, T5, 0, m->stream2>>>(n0, n1, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc,
                                                                              vs ? vs[s + 1] : 1.0, ex->fext.p, ex->minv.p, ex->bccode.p,
                                                                              ex->bcval.p, ex->d.p, ex->v.p, ex->a.p, ex->fint.p, skip);
                    else
                        k_cd_node_update<true, false><<<nb, T5, 0, m->stream2>>>(n0, n1, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc, 1.0,
                                                                                ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p,
                                                                                ex->v.p, ex->a.p, ex->fint.p, skip);
                }
                TB2_CUDA(cudaEventRecord(m->ev_k5[nc], m->stream2));
            }
        }
    }
    if (multi) TB2_CUDA(cudaStreamWaitEvent(m->stream, cp.ev_done, 0));
    if (two_k1_streams) {
        TB2_CUDA(cudaEventRecord(m->ev_join1b, m->stream1b));
        TB2_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join1b, 0));
    }
    TB2_CUDA(cudaEventRecord(m->ev_join, m->stream2));
    TB2_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
    TB2_CUDA(cudaGetLastError());
    if (getenv("TB2_DEBUG_TIMING")) // host enqueue cost of the step loop (is the pipeline launch-bound?)
        fprintf(stderr, "[tb2] explicit pipeline: %d steps enqueued in %.3f ms (%.1f us/step host)\n", nsteps,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enqueue0).count(),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_enqueue0).count() / nsteps);
    return TB2_OK;
}

static bool pipeline_enabled();

// One step through HOST arrays (Tahoe's FieldT stays authoritative: d, v, a come in and go back every step), as a slab pipeline
// over four streams so that PCIe runs in both directions at once:
//   h2d stream   d, v, a of node slab 0, 1, 2, ...                                      (host -> device)
//   main stream  predictor of a slab as soon as it has landed; K1 of an element slab once the node slabs it touches are predicted
//   second       K5 (gather, M^-1 R, corrector) of a node slab once the element slabs touching it are done
//   d2h stream   d of a slab right after its predictor, v and a right after its K5         (device -> host)
// Same kernels, same arithmetic and summation order as the serial path: bitwise identical fields (tested).
// device alias of a host pointer when the memory is pinned and mapped (cudaHostAlloc / cudaHostRegister), else null
static double* mapped_host_alias(double* h)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? (double*)at.devicePointer : nullptr;
}

static int explicit_step_host_pipelined(tb2_explicit* ex, double dt, double* h_d, double* h_v, double* h_a)
{
    tb2_group* g = ex->group;
    tb2_mesh* m = g->mesh;
    const PipePlan& P = m->hplan;
    const int C = P.chunks();
    const int T = 256;
    // Opt-in (TB2_HOST_FUSED_D2H=1) for pinned + mapped host arrays: the kernels write the results to the host themselves and the
    // copy engines only upload.  Measured on the B200 box (profiles/r01c_summary.md): slower than the copy-engine download
    // (2.7 vs 2.4 ms per 1M-element step) -- 8-byte posted writes from the SMs reach 12-26 GB/s and slow the upload beside them.
    static const bool fuse_allowed = getenv("TB2_HOST_FUSED_D2H") && getenv("TB2_HOST_FUSED_D2H")[0] == '1';
    double *md = fuse_allowed ? mapped_host_alias(h_d) : nullptr, *mv = fuse_allowed ? mapped_host_alias(h_v) : nullptr,
           *ma = fuse_allowed ? mapped_host_alias(h_a) : nullptr;
    const bool fused = fuse_allowed && md && mv && ma;
    if (!m->stream2) {
        TB2_CUDA(cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking));
        const int Cc = (int)m->pipe_e0.size() - 1;
        m->ev_k1.resize(Cc);
        m->ev_k5.resize(Cc);
        for (int c = 0; c < Cc; c++) {
            TB2_CUDA(cudaEventCreateWithFlags(&m->ev_k1[c], cudaEventDisableTiming));
            TB2_CUDA(cudaEventCreateWithFlags(&m->ev_k5[c], cudaEventDisableTiming));
        }
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
    if (!m->stream_h2d) {
        TB2_CUDA(cudaStreamCreateWithFlags(&m->stream_h2d, cudaStreamNonBlocking));
        TB2_CUDA(cudaStreamCreateWithFlags(&m->stream_d2h, cudaStreamNonBlocking));
        for (auto* v : {&m->ev_h2d, &m->ev_pred, &m->ev_hk1, &m->ev_hk5}) {
            v->resize(C);
            for (int c = 0; c < C; c++) TB2_CUDA(cudaEventCreateWithFlags(&(*v)[c], cudaEventDisableTiming));
        }
        TB2_CUDA(cudaEventCreateWithFlags(&m->ev_d2h_done, cudaEventDisableTiming));
    }
    // everything queued on the mesh stream so far (state uploads, earlier steps) precedes the copies
    TB2_CUDA(cudaEventRecord(m->ev_join, m->stream));
    TB2_CUDA(cudaStreamWaitEvent(m->stream_h2d, m->ev_join, 0));
    TB2_CUDA(cudaStreamWaitEvent(m->stream_d2h, m->ev_join, 0));
    TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_join, 0));
    for (int nc = 0; nc < C; nc++) {
        const int64_t o = 3 * P.n0[nc];
        const size_t bytes = (size_t)(3 * (P.n0[nc + 1] - P.n0[nc])) * sizeof(double);
        if (bytes) {
            ProfScope ps(m, kProfOther, 0, m->stream_h2d);
            TB2_CUDA(cudaMemcpyAsync(ex->d.p + o, h_d + o, bytes, cudaMemcpyHostToDevice, m->stream_h2d));
            TB2_CUDA(cudaMemcpyAsync(ex->v.p + o, h_v + o, bytes, cudaMemcpyHostToDevice, m->stream_h2d));
            TB2_CUDA(cudaMemcpyAsync(ex->a.p + o, h_a + o, bytes, cudaMemcpyHostToDevice, m->stream_h2d));
        }
        TB2_CUDA(cudaEventRecord(m->ev_h2d[nc], m->stream_h2d));
    }
    int np = 0, nc = 0;
    // predictor + ConsistentKBC of node slabs [np, upto], each as soon as its copy has landed; d of the slab goes straight back
    auto predict_upto = [&](int upto) -> int {
        for (; np < C && np <= upto; np++) {
            const int64_t o = 3 * P.n0[np], cnt = 3 * (P.n0[np + 1] - P.n0[np]);
            TB2_CUDA(cudaStreamWaitEvent(m->stream, m->ev_h2d[np], 0));
            if (cnt) {
                ProfScope ps(m, kProfPredictor);
                k_cd_predictor<<<(unsigned)((cnt + T - 1) / T), T, 0, m->stream>>>(cnt, dt, ex->d.p + o, ex->v.p + o, ex->a.p + o, ex->bccode.p + o,
                                                                                  ex->bcval.p + o, 1.0, fused ? md + o : nullptr);
            }
            TB2_CUDA(cudaEventRecord(m->ev_pred[np], m->stream));
            if (!fused) {
                TB2_CUDA(cudaStreamWaitEvent(m->stream_d2h, m->ev_pred[np], 0));
                if (cnt) {
                    ProfScope ps(m, kProfOther, 0, m->stream_d2h);
                    TB2_CUDA(cudaMemcpyAsync(h_d + o, ex->d.p + o, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, m->stream_d2h));
                }
            }
        }
        return TB2_OK;
    };
    for (int c = 0; c < C; c++) {
        TB2_CHECK(predict_upto(P.nmax_of_ec[c])); // the node slabs this element slab reads
        TB2_CHECK(launch_element_forces_range(g, ex->d.p, nullptr, 0, P.e0[c], P.e0[c + 1], m->stream));
        TB2_CUDA(cudaEventRecord(m->ev_hk1[c], m->stream));
        for (; nc < C && P.emax_of_nc[nc] <= c; nc++) {
            const int64_t n0 = P.n0[nc], n1 = P.n0[nc + 1];
            const int64_t o = 3 * n0;
            if (np <= nc) TB2_CHECK(predict_upto(nc)); // a node slab no element of slabs <= c touches
            TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_pred[nc], 0));
            if (P.emax_of_nc[nc] >= 0) TB2_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_hk1[P.emax_of_nc[nc]], 0));
            if (n1 > n0) {
                ProfScope ps(m, kProfNodeUpdate, 1, m->stream2);
                const unsigned nb = (unsigned)((n1 - n0 + T - 1) / T);
                if (fused)
                    k_cd_node_update_hostout<<<nb, T, 0, m->stream2>>>(n0, n1, m->inc_ptr.p, m->inc.p, m->fe.p, m->stride, dt, 1.0, ex->fext.p,
                                                                      ex->minv.p, ex->bccode.p, ex->v.p, ex->a.p, ex->fint.p, mv, ma);
                else
                    k_cd_node_update<true, false><<<nb, T, 0, m->stream2>>>(n0, n1, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, 1.0, 1.0,
                                                                           ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p, ex->v.p,
                                                                           ex->a.p, ex->fint.p);
            }
            TB2_CUDA(cudaEventRecord(m->ev_hk5[nc], m->stream2));
            if (!fused) {
                TB2_CUDA(cudaStreamWaitEvent(m->stream_d2h, m->ev_hk5[nc], 0));
                if (n1 > n0) {
                    const size_t bytes = (size_t)(3 * (n1 - n0)) * sizeof(double);
                    ProfScope ps(m, kProfOther, 0, m->stream_d2h);
                    TB2_CUDA(cudaMemcpyAsync(h_v + o, ex->v.p + o, bytes, cudaMemcpyDeviceToHost, m->stream_d2h));
                    TB2_CUDA(cudaMemcpyAsync(h_a + o, ex->a.p + o, bytes, cudaMemcpyDeviceToHost, m->stream_d2h));
                }
            }
        }
    }
    TB2_CUDA(cudaEventRecord(m->ev_d2h_done, m->stream_d2h));
    TB2_CUDA(cudaStreamWaitEvent(m->stream, m->ev_d2h_done, 0));
    TB2_CUDA(cudaEventRecord(m->ev_join, m->stream2));
    TB2_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

static bool pipeline_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char* s = getenv("TB2_PIPELINE");
        on = (s && s[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

static int explicit_steps(tb2_explicit* ex, double dt, int nsteps, const double* fs, const double* vs)
{
    tb2_group* g = ex->group;
    tb2_mesh* m = g->mesh;
    const int64_t ndof = 3 * m->nn;
    const int T = 256;
    const unsigned nbn = (unsigned)((m->nn + T - 1) / T), nbd = (unsigned)((ndof + T - 1) / T);
    if (nsteps <= 0) return TB2_OK;
    if (m->pipe_e0.size() > 2 && nsteps > 1 && pipeline_enabled()) return explicit_steps_pipelined(ex, dt, nsteps, fs, vs);
    {
        ProfScope ps(m, kProfPredictor);
        k_cd_predictor<<<nbd, T, 0, m->stream>>>(ndof, dt, ex->d.p, ex->v.p, ex->a.p, ex->bccode.p, ex->bcval.p, vs ? vs[0] : 1.0);
    }
    const bool multi = comm_active(m);
    for (int s = 0; s < nsteps; s++) {
        TB2_CHECK(launch_element_forces(g, ex->d.p, nullptr, 0));
        const double fsc = fs ? fs[s] : 1.0;
        if (multi) {
            // partial nodal forces -> sum over the ranks sharing interface nodes -> update from the summed force
            TB2_CHECK(launch_node_gather(m, ex->fint.p, true));
            TB2_CHECK(tb2_comm_sum_interface(m, ex->fint.p));
            ProfScope ps(m, kProfNodeUpdate);
            if (s + 1 < nsteps)
                k_cd_node_update<false, true><<<nbn, T, 0, m->stream>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc,
                                                                       vs ? vs[s + 1] : 1.0, ex->fext.p, ex->minv.p, ex->bccode.p,
                                                                       ex->bcval.p, ex->d.p, ex->v.p, ex->a.p, ex->fint.p);
            else
                k_cd_node_update<false, false><<<nbn, T, 0, m->stream>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc, 1.0,
                                                                        ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p,
                                                                        ex->v.p, ex->a.p, ex->fint.p);
        } else if (s + 1 < nsteps) {
            ProfScope ps(m, kProfNodeUpdate);
            k_cd_node_update<true, true><<<nbn, T, 0, m->stream>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc,
                                                            vs ? vs[s + 1] : 1.0, ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p,
                                                            ex->d.p, ex->v.p, ex->a.p, ex->fint.p);
        } else {
            ProfScope ps(m, kProfNodeUpdate);
            k_cd_node_update<true, false><<<nbn, T, 0, m->stream>>>(0, m->nn, m->inc_ptr.p, m->inc.p, (const int4*)m->inc8.p, m->fe.p, m->stride, dt, fsc, 1.0,
                                                             ex->fext.p, ex->minv.p, ex->bccode.p, ex->bcval.p, ex->d.p, ex->v.p,
                                                             ex->a.p, ex->fint.p);
        }
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
    DeviceGuard dg(ex->group->mesh->device);
    cudaStreamSynchronize(ex->group->mesh->stream);
    delete ex;
    return TB2_OK;
}

int tb2_explicit_set_state(tb2_explicit* ex, const double* h_d, const double* h_v, const double* h_a)
{
    TB2_ARG(ex);
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
    if (h_fext) TB2_CUDA(cudaMemcpyAsync(ex->fext.p, h_fext, n * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
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
    k_initial_acceleration<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(n, ex->fext.p, ex->fint.p, ex->minv.p, ex->bccode.p, ex->a.p);
    TB2_CUDA(cudaGetLastError());
    return tb2_group_status(g, nullptr);
}

int tb2_explicit_run(tb2_explicit* ex, double dt, int nsteps, const double* h_fext_scale, const double* h_value_scale)
{
    TB2_ARG(ex && nsteps >= 0);
    DeviceGuard dg(ex->group->mesh->device);
    TB2_CHECK(explicit_steps(ex, dt, nsteps, h_fext_scale, h_value_scale));
    return tb2_group_status(ex->group, nullptr);
}

int tb2_explicit_step_host(tb2_explicit* ex, double dt, double* h_d, double* h_v, double* h_a)
{
    TB2_ARG(ex && h_d && h_v && h_a);
    tb2_mesh* m = ex->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    if (!comm_active(m) && m->hplan.chunks() > 1 && pipeline_enabled()) {
        TB2_CHECK(explicit_step_host_pipelined(ex, dt, h_d, h_v, h_a));
        return tb2_group_status(ex->group, nullptr);
    }
    TB2_CUDA(cudaMemcpyAsync(ex->d.p, h_d, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(ex->v.p, h_v, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(ex->a.p, h_a, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(explicit_steps(ex, dt, 1, nullptr, nullptr));
    TB2_CUDA(cudaMemcpyAsync(h_d, ex->d.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_v, ex->v.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_a, ex->a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    return tb2_group_status(ex->group, nullptr);
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
