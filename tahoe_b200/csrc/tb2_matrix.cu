// tb2_matrix.cu -- K9 (equation numbers, sparsity, element->slot map) and K6-K8 (CSR SpMV, Jacobi-PCG).
//
// Equation numbering follows NodeManagerT::SetEquationNumbers / FieldT::InitEquations (NodeManagerT.cpp:712-767,
// FieldT.cpp:635-659): node-major, dof-minor, 1-based, prescribed dofs = -1.  Because the numbering is monotone in the
// node id, the sorted column list of a row is the concatenation, over the row node's neighbour nodes in ascending node
// order, of the neighbours' active equations -- which is what GraphT::MakeGraph + MSRBuilderT's per-row sort produce
// (GraphT.cpp:376-488, MSRBuilderT.cpp:134-181).  The structure is therefore built from a node adjacency (<= 27 entries
// per node on a structured mesh) instead of from 24x24 equation pairs per element.
#include <algorithm>
#include <cub/cub.cuh>

#include "tb2_internal.h"

namespace tb2 {

static const int kMaxAdj = 160; // candidate buffer per node: up to 20 incident hexes

// ---- equations ----------------------------------------------------------------------------------------------------
__global__ void k_active_flags(int64_t ndof, const unsigned char* __restrict__ bc, int* __restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < ndof) flag[i] = bc[i] ? 0 : 1;
}
__global__ void k_assign_eqnos(int64_t ndof, const int* __restrict__ flag, const int* __restrict__ incl, int* __restrict__ eqnos,
                               int* __restrict__ eq_node)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ndof) return;
    if (flag[i]) {
        eqnos[i] = incl[i];
        eq_node[incl[i] - 1] = (int)i;
    } else
        eqnos[i] = -1; // FieldT::kPrescribed
}

// ---- node adjacency -------------------------------------------------------------------------------------------------
// sorted unique neighbour nodes of node n (incl. n): union of the nodes of its incident elements
__device__ int collect_neighbours(int64_t n, const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ conn,
                                  int64_t stride, int* buf, int& overflow)
{
    int cnt = 0;
    for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; k++) {
        const int64_t e = inc[k] >> 3;
        for (int a = 0; a < 8; a++) {
            const int v = conn[a * stride + e];
            // sorted insert without duplicates
            int lo = 0, hi = cnt;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (buf[mid] < v) lo = mid + 1;
                else hi = mid;
            }
            if (lo < cnt && buf[lo] == v) continue;
            if (cnt >= kMaxAdj) { overflow = 1; continue; }
            for (int q = cnt; q > lo; q--) buf[q] = buf[q - 1];
            buf[lo] = v;
            cnt++;
        }
    }
    return cnt;
}
__global__ void __launch_bounds__(128) k_adj_count(int64_t nn, const int* inc_ptr, const int* inc, const int* conn, int64_t stride,
                                                  int* __restrict__ count, int* overflow)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    int buf[kMaxAdj], ov = 0;
    count[n] = collect_neighbours(n, inc_ptr, inc, conn, stride, buf, ov);
    if (ov) *overflow = 1;
}
// fill adjacency + per-entry column offset; rowlen[n] = number of columns in each row of node n
__global__ void __launch_bounds__(128) k_adj_fill(int64_t nn, const int* inc_ptr, const int* inc, const int* conn, int64_t stride,
                                                 const int* __restrict__ adj_ptr, const int* __restrict__ eqnos, int* __restrict__ adj,
                                                 int* __restrict__ coloff, int* __restrict__ rowlen, long long* __restrict__ node_nnz)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    int buf[kMaxAdj], ov = 0;
    const int cnt = collect_neighbours(n, inc_ptr, inc, conn, stride, buf, ov);
    const int base = adj_ptr[n];
    int off = 0;
    for (int k = 0; k < cnt; k++) {
        const int64_t m = buf[k];
        adj[base + k] = (int)m;
        coloff[base + k] = off;
        off += (eqnos[3 * m] > 0) + (eqnos[3 * m + 1] > 0) + (eqnos[3 * m + 2] > 0);
    }
    rowlen[n] = off;
    const int nact = (eqnos[3 * n] > 0) + (eqnos[3 * n + 1] > 0) + (eqnos[3 * n + 2] > 0);
    node_nnz[n] = (long long)nact * off;
}
__global__ void k_rowptr(int64_t nn, const int* __restrict__ eqnos, const int* __restrict__ rowlen,
                         const long long* __restrict__ node_start, long long* __restrict__ rowptr, int64_t neq, long long nnz)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n == 0) rowptr[neq] = nnz;
    if (n >= nn) return;
    int r = 0;
    for (int i = 0; i < 3; i++) {
        const int eq = eqnos[3 * n + i];
        if (eq > 0) {
            rowptr[eq - 1] = node_start[n] + (long long)r * rowlen[n];
            r++;
        }
    }
}
// one thread per adjacency entry: writes the neighbour's active equations into each active row of the node
__global__ void k_colind(int64_t nn, int64_t nadj, const int* __restrict__ adj_ptr, const int* __restrict__ adj,
                         const int* __restrict__ coloff, const int* __restrict__ eqnos, const long long* __restrict__ rowptr,
                         int* __restrict__ colind)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nadj) return;
    // owning node by binary search in adj_ptr
    int64_t lo = 0, hi = nn;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (adj_ptr[mid + 1] <= k) lo = mid + 1;
        else hi = mid;
    }
    const int64_t n = lo, m = adj[k];
    const int off = coloff[k];
    for (int i = 0; i < 3; i++) {
        const int eq = eqnos[3 * n + i];
        if (eq <= 0) continue;
        long long p = rowptr[eq - 1] + off;
        for (int j = 0; j < 3; j++) {
            const int c = eqnos[3 * m + j];
            if (c > 0) colind[p++] = c - 1;
        }
    }
}
// elem_adjpos[(a*8+b)][e] = index into adj[] of (conn[a] -> conn[b])
__global__ void k_elem_adjpos(int64_t ne, int64_t stride, const int* __restrict__ conn, const int* __restrict__ adj_ptr,
                              const int* __restrict__ adj, int* __restrict__ pos)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t e = t % stride;
    const int ab = (int)(t / stride);
    if (ab >= 64 || e >= ne) return;
    const int na = conn[(ab >> 3) * stride + e], nb = conn[(ab & 7) * stride + e];
    int lo = adj_ptr[na], hi = adj_ptr[na + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (adj[mid] < nb) lo = mid + 1;
        else hi = mid;
    }
    pos[(int64_t)ab * stride + e] = lo;
}

// ---- K6: y = A x (MSRMatrixT::Multx, MSRMatrixT.cpp:385-420) ------------------------------------------------------------
// One warp per ROW GROUP: the <= 3 rows of a node share their column pattern (equation numbering is node-major), so the warp
// reads colind and gathers x once and feeds up to three rows: 8 B/nnz values + 4/3 B/nnz indices instead of 12 B/nnz.
// A matrix without node structure (tb2_matrix_create_csr) uses groups of one row.  Persistent grid: warp w of block b walks
// groups b*W+w, +gridDim*W, ...; the p.Ap partial of a block is accumulated in that fixed order (deterministic).
// grp[g] = first row | (rows-1) << 30.
template <bool WITH_DOT, bool PUBLISH>
__device__ __forceinline__ void spmv_groups(int64_t ngroups, const int* __restrict__ grp, const long long* __restrict__ rowptr,
                                            const int* __restrict__ colind, const double* __restrict__ val, const double* __restrict__ x,
                                            double* __restrict__ y, double* __restrict__ partial, const int* __restrict__ done,
                                            const int* __restrict__ xidx, double* __restrict__ xout, int64_t npublish,
                                            const PeerView* pv, unsigned long long epoch, unsigned* counter)
{
    if (WITH_DOT && done && *done) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
    double dot = 0.0;
    for (int64_t g = blockIdx.x * (int64_t)W + wib; g < ngroups; g += (int64_t)gridDim.x * W) {
        const unsigned packed = (unsigned)__ldg(grp + g);
        const int64_t r0 = packed & 0x3fffffffu;
        const int nr = (int)(packed >> 30) + 1;
        const long long k0 = __ldg(rowptr + r0);
        const int len = (int)(__ldg(rowptr + r0 + 1) - k0);
        const double* v0 = val + k0;
        const int* c0 = colind + k0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (nr == 3 && len <= 96) {
            // the common case (a Q1 hex node couples to <= 27 nodes = 81 columns): all loads of the row group are issued before
            // the first use, so one round trip to HBM covers the whole group instead of one per 32 columns
            const int ka = lane, kb = lane + 32, kc = lane + 64;
            const bool pa = ka < len, pb = kb < len, pc = kc < len;
            const int ca = pa ? __ldcs(c0 + ka) : 0, cb = pb ? __ldcs(c0 + kb) : 0, cc = pc ? __ldcs(c0 + kc) : 0;
            const double a0 = pa ? __ldcs(v0 + ka) : 0.0, a1 = pa ? __ldcs(v0 + len + ka) : 0.0, a2 = pa ? __ldcs(v0 + 2 * len + ka) : 0.0;
            const double b0 = pb ? __ldcs(v0 + kb) : 0.0, b1 = pb ? __ldcs(v0 + len + kb) : 0.0, b2 = pb ? __ldcs(v0 + 2 * len + kb) : 0.0;
            const double d0 = pc ? __ldcs(v0 + kc) : 0.0, d1 = pc ? __ldcs(v0 + len + kc) : 0.0, d2 = pc ? __ldcs(v0 + 2 * len + kc) : 0.0;
            const double xa = pa ? __ldg(x + ca) : 0.0, xb = pb ? __ldg(x + cb) : 0.0, xc = pc ? __ldg(x + cc) : 0.0;
            // same per-lane summation order as the generic loop below: k = lane, lane + 32, lane + 64
            s0 = a0 * xa; s1 = a1 * xa; s2 = a2 * xa;
            s0 += b0 * xb; s1 += b1 * xb; s2 += b2 * xb;
            s0 += d0 * xc; s1 += d1 * xc; s2 += d2 * xc;
        } else if (nr == 3) {
            for (int k = lane; k < len; k += 32) {
                const double xv = __ldg(x + __ldcs(c0 + k));
                s0 += __ldcs(v0 + k) * xv;
                s1 += __ldcs(v0 + len + k) * xv;
                s2 += __ldcs(v0 + 2 * len + k) * xv;
            }
        } else {
            for (int k = lane; k < len; k += 32) {
                const double xv = __ldg(x + __ldcs(c0 + k));
                s0 += __ldcs(v0 + k) * xv;
                if (nr > 1) s1 += __ldcs(v0 + len + k) * xv;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            y[r0] = s0;
            if (nr > 1) y[r0 + 1] = s1;
            if (nr > 2) y[r0 + 2] = s2;
            if (PUBLISH && g < npublish) {
                // a row group of an interface node (they come first in the group list): this rank's partial also goes into its exchange
                // window; the warp that completes the last of them raises the arrival flag in every peer's window
                xout[xidx[r0]] = s0;
                if (nr > 1) xout[xidx[r0 + 1]] = s1;
                if (nr > 2) xout[xidx[r0 + 2]] = s2;
                __threadfence();
                if (atomicAdd(counter, 1u) == (unsigned)(npublish - 1)) {
                    *counter = 0;
                    __threadfence_system();
                    for (int r = 0; r < pv->nranks; r++)
                        if (r != pv->rank) peer_st_release((unsigned long long*)(pv->win[r] + kPeerFlagsOff) + pv->rank, epoch);
                }
            }
            if (WITH_DOT) {
                dot += s0 * x[r0];
                if (nr > 1) dot += s1 * x[r0 + 1];
                if (nr > 2) dot += s2 * x[r0 + 2];
            }
        }
    }
    if (WITH_DOT) {
        __shared__ double sh[8];
        if (lane == 0) sh[wib] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < W; w++) t += sh[w];
            partial[blockIdx.x] = t;
        }
    }
}
template <bool WITH_DOT>
__global__ void __launch_bounds__(256, 6) k_spmv(int64_t ngroups, const int* __restrict__ grp, const long long* __restrict__ rowptr,
                                                 const int* __restrict__ colind, const double* __restrict__ val, const double* __restrict__ x,
                                                 double* __restrict__ y, double* __restrict__ partial, const int* __restrict__ done)
{
    spmv_groups<WITH_DOT, false>(ngroups, grp, rowptr, colind, val, x, y, partial, done, nullptr, nullptr, 0, nullptr, 0ull, nullptr);
}
// The SpMV of a distributed solve over peer memory (tb2_peer.cuh), multiply and exchange in one kernel: the group list holds the
// rows of the interface nodes first; their partials A_loc u go to y AND into this rank's exchange window, and the warp that
// completes the last of them raises the arrival flag in every peer's window while the grid goes on with the interior rows.
// The sharers pull and sum the partials inside the next vector update (k_cg_update<true>), a whole SpMV later.
__global__ void __launch_bounds__(256, 6) k_spmv_publish(int64_t ngroups, int64_t ngroups_if, const int* __restrict__ grp,
                                                         const long long* __restrict__ rowptr, const int* __restrict__ colind,
                                                         const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                         double* __restrict__ partial, const int* __restrict__ done,
                                                         const int* __restrict__ xidx, const PeerView pv, unsigned long long epoch,
                                                         unsigned* counter)
{
    if (*done) { // uniform over the grid and over the ranks (it derives from all-reduced scalars): no rows, but the peers' pull of this
                 // epoch still waits for the arrival flag
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int r = 0; r < pv.nranks; r++)
                if (r != pv.rank) peer_st_release((unsigned long long*)(pv.win[r] + kPeerFlagsOff) + pv.rank, epoch);
        return;
    }
    spmv_groups<true, true>(ngroups, grp, rowptr, colind, val, x, y, partial, done, xidx, peer_data(pv, pv.rank, epoch), ngroups_if, &pv, epoch,
                            counter);
}
// row groups of a mesh-derived matrix: one group per node with active dofs (flag/scan/compact); generic: one per row
__global__ void k_group_flags(int64_t nn, const int* __restrict__ eqnos, int* __restrict__ flag)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < nn) flag[n] = (eqnos[3 * n] > 0 || eqnos[3 * n + 1] > 0 || eqnos[3 * n + 2] > 0) ? 1 : 0;
}
__global__ void k_group_fill(int64_t nn, const int* __restrict__ eqnos, const int* __restrict__ flag, const int* __restrict__ excl,
                             int* __restrict__ grp)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn || !flag[n]) return;
    int first = 0, cnt = 0;
    for (int i = 0; i < 3; i++) {
        const int eq = eqnos[3 * n + i];
        if (eq > 0) {
            if (!cnt) first = eq - 1;
            cnt++;
        }
    }
    grp[excl[n]] = (int)((unsigned)first | ((unsigned)(cnt - 1) << 30));
}
__global__ void k_group_identity(int64_t n, int* __restrict__ grp)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) grp[i] = (int)i;
}

// deterministic two-stage reductions -------------------------------------------------------------------------------
TB2_DEV double block_sum(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t; // valid in thread 0
}

// scalars: [0] rz, [1] pAp, [2] rz_new, [3] rr, [4] alpha, [5] beta, [6] r0norm, [7] rnorm
enum { kRZ = 0, kPAP = 1, kRZNEW = 2, kRR = 3, kALPHA = 4, kBETA = 5, kR0 = 6, kRNORM = 7, kNumScal = 8 };
// control ints: [0] done, [1] iterations, [2] breakdown
struct PcgCtl {
    int done, iters, breakdown, pad;
};

// sum nparts partials (optionally 2 interleaved streams) with one block
__global__ void __launch_bounds__(1024) k_reduce_pap(int nparts, const double* __restrict__ partial, double* __restrict__ scal, PcgCtl* ctl)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) v += partial[i];
    v = block_sum(v, sh);
    if (threadIdx.x == 0) {
        scal[kPAP] = v;
        if (!(v > 0.0)) { ctl->breakdown = 1; ctl->done = 1; scal[kALPHA] = 0.0; }
        else scal[kALPHA] = scal[kRZ] / v;
    }
}
// x += alpha p ; r -= alpha q ; z = dinv r ; partial rz, rr
__global__ void __launch_bounds__(256) k_pcg_update(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl,
                                                   const double* __restrict__ p, const double* __restrict__ q,
                                                   const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                                                   double* __restrict__ z, double* __restrict__ partial,
                                                   const unsigned char* __restrict__ w = nullptr)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    const double alpha = scal[kALPHA];
    double rz = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        const double zi = dinv[i] * ri;
        r[i] = ri;
        z[i] = zi;
        if (!w || w[i]) {
            rz += ri * zi;
            rr += ri * ri;
        }
    }
    rz = block_sum(rz, sh);
    rr = block_sum(rr, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = rz;
        partial[2 * blockIdx.x + 1] = rr;
    }
}
__global__ void __launch_bounds__(1024) k_reduce_rz(int nparts, const double* __restrict__ partial, double* __restrict__ scal, PcgCtl* ctl,
                                                   double rtol, double atol, int max_iter)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) {
        const double rnorm = sqrt(b);
        scal[kBETA] = a / scal[kRZ];
        scal[kRZ] = a;
        scal[kRNORM] = rnorm;
        ctl->iters += 1;
        if (!(rnorm > atol) || !(rnorm > rtol * scal[kR0]) || ctl->iters >= max_iter) ctl->done = 1;
    }
}
// p = z + beta p
__global__ void __launch_bounds__(256) k_pcg_direction(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl,
                                                      const double* __restrict__ z, double* __restrict__ p)
{
    if (ctl->done) return;
    const double beta = scal[kBETA];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}
// setup: dinv from diagonal (DiagonalMatrixT::Factorize semantics); r = b - q ; z = dinv r ; p = z ; partial rz, rr
__global__ void k_extract_dinv(int64_t n, const long long* __restrict__ rowptr, const int* __restrict__ colind, const double* __restrict__ val,
                               double* __restrict__ out, int invert)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long lo = rowptr[i], hi = rowptr[i + 1];
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (colind[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    const double d = (lo < rowptr[i + 1] && colind[lo] == i) ? val[lo] : 0.0;
    out[i] = invert ? (fabs(d) > 1.0e-12 ? 1.0 / d : d) : d;
}
__global__ void __launch_bounds__(256) k_pcg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
                                                 const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ z,
                                                 double* __restrict__ p, double* __restrict__ partial,
                                                 const unsigned char* __restrict__ w = nullptr)
{
    __shared__ double sh[32];
    double rz = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - q[i];
        const double zi = dinv[i] * ri;
        r[i] = ri;
        z[i] = zi;
        p[i] = zi;
        if (!w || w[i]) {
            rz += ri * zi;
            rr += ri * ri;
        }
    }
    rz = block_sum(rz, sh);
    rr = block_sum(rr, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = rz;
        partial[2 * blockIdx.x + 1] = rr;
    }
}
__global__ void __launch_bounds__(1024) k_reduce_init(int nparts, const double* __restrict__ partial, double* __restrict__ scal, PcgCtl* ctl,
                                                     double rtol, double atol, int max_iter)
{
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) {
        const double rnorm = sqrt(b);
        scal[kRZ] = a;
        scal[kR0] = rnorm;
        scal[kRNORM] = rnorm;
        ctl->iters = 0;
        ctl->breakdown = 0;
        ctl->done = (!(rnorm > atol) || !(rnorm > rtol * rnorm) || max_iter <= 0) ? 1 : 0;
    }
}

__global__ void k_eq_gather(int64_t neq, const int* __restrict__ eq_node, const double* __restrict__ nodal, double* __restrict__ v)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < neq) v[i] = nodal[eq_node[i]];
}
__global__ void k_eq_scatter_add(int64_t neq, const int* __restrict__ eq_node, double s, const double* __restrict__ v, double* __restrict__ nodal)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < neq) nodal[eq_node[i]] += s * v[i];
}

// ---- multi-GPU PCG pieces (sub-domain matrices) ------------------------------------------------------------------
// weighted dot: partial[block] = sum_i w_i a_i b_i (w = 1 on equations whose node this rank owns)
__global__ void __launch_bounds__(256) k_wdot(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                             const unsigned char* __restrict__ w, double* __restrict__ partial, const PcgCtl* ctl)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (w[i]) s += a[i] * b[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// red[k] = sum of the k-th interleaved stream of partials (k < nstreams <= 2)
__global__ void __launch_bounds__(1024) k_sum_partials(int nparts, int nstreams, const double* __restrict__ partial, double* __restrict__ red,
                                                      const PcgCtl* ctl)
{
    if (ctl && ctl->done) return;
    __shared__ double sh[32];
    for (int k = 0; k < nstreams; k++) {
        double v = 0.0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) v += partial[nstreams * i + k];
        v = block_sum(v, sh);
        if (threadIdx.x == 0) red[k] = v;
        __syncthreads();
    }
}
__global__ void k_scalars_init(const double* red, double* scal, PcgCtl* ctl, double rtol, double atol, int max_iter)
{
    const double rnorm = sqrt(red[1]);
    scal[kRZ] = red[0];
    scal[kR0] = rnorm;
    scal[kRNORM] = rnorm;
    ctl->iters = 0;
    ctl->breakdown = 0;
    ctl->done = (!(rnorm > atol) || !(rnorm > rtol * rnorm) || max_iter <= 0) ? 1 : 0;
}
__global__ void k_scalars_alpha(const double* red, double* scal, PcgCtl* ctl)
{
    if (ctl->done) return;
    const double v = red[0];
    scal[kPAP] = v;
    if (!(v > 0.0)) { ctl->breakdown = 1; ctl->done = 1; scal[kALPHA] = 0.0; }
    else scal[kALPHA] = scal[kRZ] / v;
}
__global__ void k_scalars_beta(const double* red, double* scal, PcgCtl* ctl, double rtol, double atol, int max_iter)
{
    if (ctl->done) return;
    const double rnorm = sqrt(red[1]);
    scal[kBETA] = red[0] / scal[kRZ];
    scal[kRZ] = red[0];
    scal[kRNORM] = rnorm;
    ctl->iters += 1;
    if (!(rnorm > atol) || !(rnorm > rtol * scal[kR0]) || ctl->iters >= max_iter) ctl->done = 1;
}
__global__ void k_eq_owned(int64_t neq, const int* __restrict__ eq_node, const unsigned char* __restrict__ node_owned, unsigned char* __restrict__ w)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < neq) w[i] = node_owned[eq_node[i] / 3];
}
__global__ void k_invert_diag(int64_t n, double* __restrict__ d)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = d[i];
    d[i] = fabs(x) > 1.0e-12 ? 1.0 / x : x;
}

bool comm_active(tb2_mesh* m);
const unsigned char* comm_owned_mask(tb2_mesh* m);
int comm_allreduce_scalars(tb2_mesh* m, double* d_vals, int n);
int comm_sum_interface_eq(tb2_mesh* m, const int* d_eqnos, double* d_eqvec);
int comm_pack_eq(tb2_mesh* m, const int* d_eqnos, const double* d_eqvec, cudaStream_t st);
int comm_unpack_eq(tb2_mesh* m, const int* d_eqnos, double* d_eqvec, cudaStream_t st);
bool comm_plan(tb2_mesh* m, CommPlan* out);
int comm_allreduce_packed(tb2_mesh* m);

static const int kReduceBlocks = 148 * 8;
static const int kSpmvBlocks = 148 * 8; // upper bound of the persistent SpMV grid (sizes the partial-sum buffer)
// persistent SpMV grid: exactly the CTAs that are resident at once (SM count x occupancy of the kernel), so that no second,
// thinner wave follows the first (r01c: a fixed 8/SM grid ran as 6 + 2 resident CTAs and averaged 53 % of the warp slots)
static unsigned spmv_grid()
{
    static unsigned blocks = 0;
    if (!blocks) {
        int dev = 0, sms = 148, occ = 6, occ_dot = 6;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv<false>, 256, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dot, k_spmv<true>, 256, 0);
        if (occ_dot < occ) occ = occ_dot;
        if (occ < 1) occ = 1;
        blocks = (unsigned)(sms * occ);
        if (blocks > (unsigned)kSpmvBlocks) blocks = kSpmvBlocks;
    }
    return blocks;
}

} // namespace tb2

using namespace tb2;

extern "C" {

int tb2_equations_create(tb2_mesh* m, const uint8_t* h_bc, tb2_equations** out)
{
    TB2_ARG(m && h_bc && out);
    DeviceGuard dg(m->device);
    const int64_t ndof = 3 * m->nn;
    tb2_equations* q = new tb2_equations;
    q->mesh = m;
    DevBuf<unsigned char> bc, tmp;
    DevBuf<int> flag, incl;
    auto fail = [&](int s) { delete q; return s; };
#define Q_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(cuda_fail(_e, #call, __FILE__, __LINE__)); } while (0)
    Q_CUDA(bc.alloc(ndof));
    Q_CUDA(flag.alloc(ndof));
    Q_CUDA(incl.alloc(ndof));
    Q_CUDA(q->eqnos.alloc(ndof));
    Q_CUDA(cudaMemcpyAsync(bc.p, h_bc, ndof, cudaMemcpyHostToDevice, m->stream));
    const int T = 256;
    const unsigned nb = (unsigned)((ndof + T - 1) / T);
    k_active_flags<<<nb, T, 0, m->stream>>>(ndof, bc.p, flag.p);
    size_t tb = 0;
    Q_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, flag.p, incl.p, (int)ndof, m->stream));
    Q_CUDA(tmp.alloc(tb));
    Q_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, flag.p, incl.p, (int)ndof, m->stream));
    int neq = 0;
    Q_CUDA(cudaMemcpyAsync(&neq, incl.p + ndof - 1, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    Q_CUDA(cudaStreamSynchronize(m->stream));
    q->neq = neq;
    Q_CUDA(q->eq_node.alloc(neq > 0 ? neq : 1));
    k_assign_eqnos<<<nb, T, 0, m->stream>>>(ndof, flag.p, incl.p, q->eqnos.p, q->eq_node.p);
    Q_CUDA(cudaGetLastError());
    Q_CUDA(cudaStreamSynchronize(m->stream));
#undef Q_CUDA
    *out = q;
    return TB2_OK;
}
int tb2_equations_destroy(tb2_equations* q)
{
    if (!q) return TB2_OK;
    DeviceGuard dg(q->mesh->device);
    cudaStreamSynchronize(q->mesh->stream);
    delete q;
    return TB2_OK;
}
int tb2_equations_count(const tb2_equations* q, int64_t* neq)
{
    TB2_ARG(q && neq);
    *neq = q->neq;
    return TB2_OK;
}
int tb2_equations_get(const tb2_equations* q, int32_t* h_eqnos)
{
    TB2_ARG(q && h_eqnos);
    DeviceGuard dg(q->mesh->device);
    TB2_CUDA(cudaMemcpy(h_eqnos, q->eqnos.p, 3 * q->mesh->nn * sizeof(int), cudaMemcpyDeviceToHost));
    return TB2_OK;
}
const int32_t* tb2_equations_device(const tb2_equations* q) { return q ? q->eqnos.p : nullptr; }

int tb2_equations_gather(const tb2_equations* q, const double* d_nodal, double* d_eqvec)
{
    TB2_ARG(q && d_nodal && d_eqvec);
    DeviceGuard dg(q->mesh->device);
    if (q->neq) k_eq_gather<<<(unsigned)((q->neq + 255) / 256), 256, 0, q->mesh->stream>>>(q->neq, q->eq_node.p, d_nodal, d_eqvec);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}
int tb2_equations_scatter_add(const tb2_equations* q, double scale, const double* d_eqvec, double* d_nodal)
{
    TB2_ARG(q && d_nodal && d_eqvec);
    DeviceGuard dg(q->mesh->device);
    if (q->neq) k_eq_scatter_add<<<(unsigned)((q->neq + 255) / 256), 256, 0, q->mesh->stream>>>(q->neq, q->eq_node.p, scale, d_eqvec, d_nodal);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int tb2_matrix_create(tb2_equations* q, tb2_matrix** out)
{
    TB2_ARG(q && out && q->neq > 0 && q->neq < (1LL << 30));
    tb2_mesh* m = q->mesh;
    DeviceGuard dg(m->device);
    tb2_matrix* A = new tb2_matrix;
    A->eqs = q;
    A->ctx = m;
    A->neq = q->neq;
    const int64_t nn = m->nn, neq = q->neq;
    auto fail = [&](int s) { delete A; return s; };
#define A_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(cuda_fail(_e, #call, __FILE__, __LINE__)); } while (0)
    DevBuf<int> count, rowlen, overflow;
    DevBuf<long long> node_nnz, node_start;
    DevBuf<unsigned char> tmp;
    A_CUDA(count.alloc(nn + 1));
    A_CUDA(rowlen.alloc(nn));
    A_CUDA(overflow.alloc(1));
    A_CUDA(node_nnz.alloc(nn + 1));
    A_CUDA(node_start.alloc(nn + 1));
    A_CUDA(A->adj_ptr.alloc(nn + 1));
    A_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), m->stream));
    A_CUDA(cudaMemsetAsync(count.p + nn, 0, sizeof(int), m->stream));
    A_CUDA(cudaMemsetAsync(node_nnz.p + nn, 0, sizeof(long long), m->stream));
    const int T = 128;
    const unsigned nbn = (unsigned)((nn + T - 1) / T);
    k_adj_count<<<nbn, T, 0, m->stream>>>(nn, m->inc_ptr.p, m->inc.p, m->conn.p, m->stride, count.p, overflow.p);
    size_t tb = 0, tb2 = 0;
    A_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, count.p, A->adj_ptr.p, (int)(nn + 1), m->stream));
    A_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, node_nnz.p, node_start.p, (int)(nn + 1), m->stream));
    A_CUDA(tmp.alloc(tb > tb2 ? tb : tb2));
    A_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, count.p, A->adj_ptr.p, (int)(nn + 1), m->stream));
    int nadj = 0, ov = 0;
    A_CUDA(cudaMemcpyAsync(&nadj, A->adj_ptr.p + nn, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    A_CUDA(cudaMemcpyAsync(&ov, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    A_CUDA(cudaStreamSynchronize(m->stream));
    if (ov) {
        set_error("a node has more than %d distinct neighbours", kMaxAdj);
        return fail(TB2_ERR_SIZE);
    }
    A->nadj = nadj;
    A_CUDA(A->adj.alloc(nadj));
    A_CUDA(A->adj_coloff.alloc(nadj));
    k_adj_fill<<<nbn, T, 0, m->stream>>>(nn, m->inc_ptr.p, m->inc.p, m->conn.p, m->stride, A->adj_ptr.p, q->eqnos.p, A->adj.p,
                                         A->adj_coloff.p, rowlen.p, node_nnz.p);
    A_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, node_nnz.p, node_start.p, (int)(nn + 1), m->stream));
    long long nnz = 0;
    A_CUDA(cudaMemcpyAsync(&nnz, node_start.p + nn, sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
    A_CUDA(cudaStreamSynchronize(m->stream));
    A->nnz = nnz;
    A_CUDA(A->rowptr.alloc(neq + 1));
    A_CUDA(A->colind.alloc(nnz));
    A_CUDA(A->val.alloc(nnz));
    A_CUDA(cudaMemsetAsync(A->val.p, 0, nnz * sizeof(double), m->stream));
    k_rowptr<<<nbn, T, 0, m->stream>>>(nn, q->eqnos.p, rowlen.p, node_start.p, A->rowptr.p, neq, nnz);
    k_colind<<<(unsigned)((nadj + 255) / 256), 256, 0, m->stream>>>(nn, nadj, A->adj_ptr.p, A->adj.p, A->adj_coloff.p, q->eqnos.p,
                                                                  A->rowptr.p, A->colind.p);
    A_CUDA(A->elem_adjpos.alloc(64 * m->stride));
    k_elem_adjpos<<<(unsigned)((64 * m->stride + 255) / 256), 256, 0, m->stream>>>(m->ne, m->stride, m->conn.p, A->adj_ptr.p, A->adj.p,
                                                                                 A->elem_adjpos.p);
    A_CUDA(A->dinv.alloc(neq));
    A_CUDA(A->r.alloc(neq));
    A_CUDA(A->z.alloc(neq));
    A_CUDA(A->p.alloc(neq));
    A_CUDA(A->q.alloc(neq));
    A_CUDA(A->scal.alloc(kNumScal + 4)); // + PcgCtl
    A_CUDA(A->partial.alloc(2 * kReduceBlocks + kSpmvBlocks + 8));
    {   // row groups: one per node with active dofs
        DevBuf<int> gflag, gexcl;
        A_CUDA(gflag.alloc(nn + 1));
        A_CUDA(gexcl.alloc(nn + 1));
        A_CUDA(cudaMemsetAsync(gflag.p + nn, 0, sizeof(int), m->stream));
        k_group_flags<<<(unsigned)((nn + 255) / 256), 256, 0, m->stream>>>(nn, q->eqnos.p, gflag.p);
        A_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, gflag.p, gexcl.p, (int)(nn + 1), m->stream));
        int ng = 0;
        A_CUDA(cudaMemcpyAsync(&ng, gexcl.p + nn, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        A_CUDA(cudaStreamSynchronize(m->stream));
        A->ngroups = ng;
        A_CUDA(A->grp.alloc(ng > 0 ? ng : 1));
        k_group_fill<<<(unsigned)((nn + 255) / 256), 256, 0, m->stream>>>(nn, q->eqnos.p, gflag.p, gexcl.p, A->grp.p);
    }
    A_CUDA(cudaGetLastError());
    A_CUDA(cudaStreamSynchronize(m->stream));
#undef A_CUDA
    *out = A;
    return TB2_OK;
}
int tb2_matrix_destroy(tb2_matrix* A)
{
    if (!A) return TB2_OK;
    DeviceGuard dg(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    if (A->pcg_exec) cudaGraphExecDestroy(A->pcg_exec);
    tb2_mesh* ctx = A->owns_ctx ? A->ctx : nullptr;
    delete A;
    if (ctx) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
    }
    return TB2_OK;
}

static int alloc_pcg_work(tb2_matrix* A)
{
    const int64_t neq = A->neq;
    TB2_CUDA(A->dinv.alloc(neq));
    TB2_CUDA(A->r.alloc(neq));
    TB2_CUDA(A->z.alloc(neq));
    TB2_CUDA(A->p.alloc(neq));
    TB2_CUDA(A->q.alloc(neq));
    TB2_CUDA(A->scal.alloc(kNumScal + 4));
    TB2_CUDA(A->partial.alloc(2 * kReduceBlocks + kSpmvBlocks + 8));
    if (!A->grp.p) { // no node structure known: one row per group
        TB2_ARG(neq < (1LL << 30));
        TB2_CUDA(A->grp.alloc(neq));
        A->ngroups = neq;
        k_group_identity<<<(unsigned)((neq + 255) / 256), 256, 0, A->ctx->stream>>>(neq, A->grp.p);
        TB2_CUDA(cudaGetLastError());
    }
    return TB2_OK;
}

// a device CSR matrix from host CSR arrays: the path a host-assembled Tahoe matrix (MSRMatrixT-derived plugin) takes
int tb2_matrix_create_csr(int device, int64_t neq, const int64_t* h_rowptr, const int32_t* h_colind, tb2_matrix** out)
{
    TB2_ARG(out && h_rowptr && h_colind && neq > 0);
    int ndev = 0;
    TB2_CUDA(cudaGetDeviceCount(&ndev));
    TB2_ARG(device >= 0 && device < ndev);
    DeviceGuard dg(device);
    tb2_mesh* ctx = new tb2_mesh;
    ctx->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    tb2_matrix* A = new tb2_matrix;
    A->ctx = ctx;
    A->owns_ctx = true;
    A->neq = neq;
    A->nnz = h_rowptr[neq];
    int s = TB2_OK;
    auto chk = [&](cudaError_t err) { if (s == TB2_OK && err != cudaSuccess) s = cuda_fail(err, "tb2_matrix_create_csr", __FILE__, __LINE__); };
    chk(A->rowptr.alloc(neq + 1));
    chk(A->colind.alloc(A->nnz));
    chk(A->val.alloc(A->nnz));
    if (s == TB2_OK) chk(cudaMemcpy(A->rowptr.p, h_rowptr, (neq + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    if (s == TB2_OK) chk(cudaMemcpy(A->colind.p, h_colind, A->nnz * sizeof(int), cudaMemcpyHostToDevice));
    if (s == TB2_OK) chk(cudaMemset(A->val.p, 0, A->nnz * sizeof(double)));
    if (s == TB2_OK) s = alloc_pcg_work(A);
    if (s != TB2_OK) {
        tb2_matrix_destroy(A);
        return s;
    }
    *out = A;
    return TB2_OK;
}
int tb2_matrix_set_values(tb2_matrix* A, const double* h_val)
{
    TB2_ARG(A && h_val);
    DeviceGuard dg(A->ctx->device);
    TB2_CUDA(cudaMemcpyAsync(A->val.p, h_val, A->nnz * sizeof(double), cudaMemcpyHostToDevice, A->ctx->stream));
    TB2_CUDA(cudaStreamSynchronize(A->ctx->stream));
    A->values_zero = false;
    return TB2_OK;
}
int tb2_matrix_nnz(const tb2_matrix* A, int64_t* nnz)
{
    TB2_ARG(A && nnz);
    *nnz = A->nnz;
    return TB2_OK;
}
int tb2_matrix_get_csr(const tb2_matrix* A, int64_t* h_rowptr, int32_t* h_colind, double* h_val)
{
    TB2_ARG(A);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    if (h_rowptr) TB2_CUDA(cudaMemcpy(h_rowptr, A->rowptr.p, (A->neq + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
    if (h_colind) TB2_CUDA(cudaMemcpy(h_colind, A->colind.p, A->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_val) TB2_CUDA(cudaMemcpy(h_val, A->val.p, A->nnz * sizeof(double), cudaMemcpyDeviceToHost));
    return TB2_OK;
}
// MSR index array (MSRMatrixT.h:21-23, MSRBuilderT::SetMSRData): bindx[0..neq] row starts, then off-diagonal columns
int tb2_matrix_get_msr(const tb2_matrix* A, int upper_only, int32_t* h_bindx, int64_t* length)
{
    TB2_ARG(A && length);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    const int64_t neq = A->neq;
    std::vector<long long> rp(neq + 1);
    std::vector<int> ci(A->nnz);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    TB2_CUDA(cudaMemcpy(rp.data(), A->rowptr.p, (neq + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
    TB2_CUDA(cudaMemcpy(ci.data(), A->colind.p, A->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    // export only: a re-indexing of the device structure into the reference's container format
    int64_t pos = neq + 1;
    if (h_bindx) h_bindx[0] = (int32_t)pos;
    for (int64_t r = 0; r < neq; r++) {
        for (long long k = rp[r]; k < rp[r + 1]; k++)
            if (ci[k] != r && (!upper_only || ci[k] > r)) {
                if (h_bindx) h_bindx[pos] = ci[k];
                pos++;
            }
        if (h_bindx) h_bindx[r + 1] = (int32_t)pos;
    }
    *length = pos;
    return TB2_OK;
}
int tb2_matrix_clear(tb2_matrix* A)
{
    TB2_ARG(A);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaMemsetAsync(A->val.p, 0, A->nnz * sizeof(double), m->stream));
    A->values_zero = true;
    return TB2_OK;
}

// val *= s (the integrator's constK on an assembled tangent: eLinearHHTalpha::FormK, eLinearHHTalpha.cpp:30-34)
__global__ void __launch_bounds__(256) k_scale_values(long long n, double s, double* __restrict__ v)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] *= s;
}
int tb2_matrix_scale(tb2_matrix* A, double s)
{
    TB2_ARG(A);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    if (A->nnz == 0) return TB2_OK;
    const long long blocks = (A->nnz + 255) / 256;
    k_scale_values<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, m->stream>>>((long long)A->nnz, s, A->val.p);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int tb2_matrix_multx(tb2_matrix* A, const double* d_x, double* d_y)
{
    TB2_ARG(A && d_x && d_y);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    ProfScope ps(m, kProfSpmv);
    k_spmv<false><<<spmv_grid(), 256, 0, m->stream>>>(A->ngroups, A->grp.p, A->rowptr.p, A->colind.p, A->val.p, d_x, d_y, nullptr, nullptr);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}
int tb2_matrix_multx_host(tb2_matrix* A, const double* h_x, double* h_y)
{
    TB2_ARG(A && h_x && h_y);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaMemcpyAsync(A->p.p, h_x, A->neq * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(tb2_matrix_multx(A, A->p.p, A->q.p));
    TB2_CUDA(cudaMemcpyAsync(h_y, A->q.p, A->neq * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}
int tb2_matrix_copy_diagonal_host(tb2_matrix* A, double* h_diag)
{
    TB2_ARG(A && h_diag);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    TB2_CHECK(tb2_matrix_copy_diagonal(A, A->q.p));
    TB2_CUDA(cudaMemcpyAsync(h_diag, A->q.p, A->neq * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}
int tb2_matrix_copy_diagonal(tb2_matrix* A, double* d_diag)
{
    TB2_ARG(A && d_diag);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    k_extract_dinv<<<(unsigned)((A->neq + 255) / 256), 256, 0, m->stream>>>(A->neq, A->rowptr.p, A->colind.p, A->val.p, d_diag, 0);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

// Jacobi-PCG on an element-partitioned mesh: every rank holds the sub-domain matrix of its own elements (assembled locally,
// never exchanged).  Per iteration: q = A_loc p, ONE interface sum of q (the same packed all-reduce as the force sum), dots
// over owned equations + one all-reduce of 1-2 scalars (SolverT::InnerProduct semantics: each equation counted once).
// b and x must be consistent on all sharers of an interface node (they are when b comes from interface-summed forces).
// ---- distributed Jacobi-PCG on unassembled sub-domain matrices ----------------------------------------------
// Single-reduction (Chronopoulos-Gear) form of the same recurrence as tb2_matrix_pcg:
//     p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; w = A u
//     gamma' = (r, u) ; delta = (w, u) ; beta = gamma'/gamma ; alpha = gamma' / (delta - beta gamma'/alpha)
// so that ONE scalar all-reduce (gamma', |r|^2, delta) per iteration replaces the two of the textbook form (SolverT::InnerProduct,
// SolverT.cpp:854-860, is one all-reduce per dot product).  delta needs no assembled w: sum over ranks of (A_loc u, u) over all
// local rows is (A u, u), so its partial comes out of the local SpMV's epilogue.  The interface sum of w = A_loc u (the packed
// all-reduce, CommManagerT::AllGather's role) is hidden: the rows of interface nodes are multiplied first and their all-reduce
// runs on the communicator's stream beside the SpMV of the interior rows.
__global__ void k_group_interface_flag(int64_t ng, const int* __restrict__ grp, const int* __restrict__ eq_node, const int* __restrict__ node_slot,
                                       unsigned char* __restrict__ flag)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= ng) return;
    const int64_t r0 = (unsigned)grp[g] & 0x3fffffffu;
    flag[g] = node_slot[eq_node[r0] / 3] >= 0 ? 1 : 0;
}
// entry of the exchange buffer (3 slot + dof) of every equation of an interface node, -1 elsewhere
__global__ void k_eq_exchange_index(int64_t neq, const int* __restrict__ eq_node, const int* __restrict__ node_slot, int* __restrict__ xidx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= neq) return;
    const int q = eq_node[i], slot = node_slot[q / 3];
    xidx[i] = slot >= 0 ? 3 * slot + q % 3 : -1;
}
// p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = dinv r ; partials (r,u), (r,r) over owned equations
__global__ void __launch_bounds__(256) k_cg_update(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl,
                                                  const double* __restrict__ dinv, const unsigned char* __restrict__ owned,
                                                  const double* __restrict__ w, double* __restrict__ u, double* __restrict__ p,
                                                  double* __restrict__ s, double* __restrict__ x, double* __restrict__ r,
                                                  double* __restrict__ partial)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    const double alpha = scal[kALPHA], beta = scal[kBETA];
    double ru = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double pi = u[i] + beta * p[i];
        const double si = w[i] + beta * s[i];
        p[i] = pi;
        s[i] = si;
        x[i] += alpha * pi;
        const double ri = r[i] - alpha * si;
        const double ui = dinv[i] * ri;
        r[i] = ri;
        u[i] = ui;
        if (owned[i]) {
            ru += ri * ui;
            rr += ri * ri;
        }
    }
    ru = block_sum(ru, sh);
    rr = block_sum(rr, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = ru;
        partial[2 * blockIdx.x + 1] = rr;
    }
}
// red[0] = (r,u), red[1] = (r,r) from the update's partials, red[2] = (A_loc u, u) from the two SpMV launches' partials
__global__ void __launch_bounds__(1024) k_cg_sum(int nvec, const double* __restrict__ pvec, int nif, const double* __restrict__ pif, int nint,
                                                const double* __restrict__ pint, double* __restrict__ red, const PcgCtl* ctl)
{
    if (ctl && ctl->done) return;
    __shared__ double sh[32];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) { a += pvec[2 * i]; b += pvec[2 * i + 1]; }
    for (int i = threadIdx.x; i < nif; i += blockDim.x) c += pif[i];
    for (int i = threadIdx.x; i < nint; i += blockDim.x) c += pint[i];
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) { red[0] = a; red[1] = b; red[2] = c; }
}
__device__ void cg_scalars(const double* red, double* scal, PcgCtl* ctl, double rtol, double atol, int max_iter, int first)
{
    if (!first && ctl->done) return;
    const double gamma = red[0], rnorm = sqrt(red[1]), delta = red[2];
    double beta = 0.0, den = delta;
    if (first) {
        scal[kR0] = rnorm;
        ctl->iters = 0;
        ctl->breakdown = 0;
        ctl->done = (!(rnorm > atol) || !(rnorm > rtol * rnorm) || max_iter <= 0) ? 1 : 0;
    } else {
        beta = gamma / scal[kRZ];
        den = delta - beta * gamma / scal[kALPHA];
        ctl->iters += 1;
        if (!(rnorm > atol) || !(rnorm > rtol * scal[kR0]) || ctl->iters >= max_iter) ctl->done = 1;
    }
    scal[kRNORM] = rnorm;
    scal[kRZ] = gamma;
    scal[kBETA] = beta;
    scal[kPAP] = den; // = (p, A p) of the next direction
    if (!ctl->done) {
        if (!(den > 0.0)) { ctl->breakdown = 1; ctl->done = 1; scal[kALPHA] = 0.0; }
        else scal[kALPHA] = gamma / den;
    }
}
__global__ void k_cg_scalars(const double* red, double* scal, PcgCtl* ctl, double rtol, double atol, int max_iter, int first)
{
    cg_scalars(red, scal, ctl, rtol, atol, max_iter, first);
}
// the whole scalar step of one iteration in one CTA when the ranks exchange over peer memory (tb2_peer.cuh): this rank's three
// partial sums -> every peer's mailbox -> flags -> wait -> sum over the ranks in rank order (identical bits everywhere) -> alpha, beta
__global__ void __launch_bounds__(1024) k_cg_reduce_peer(int nvec, const double* __restrict__ pvec, int nif, const double* __restrict__ pif,
                                                        int nint, const double* __restrict__ pint, double* __restrict__ scal, PcgCtl* ctl,
                                                        double rtol, double atol, int max_iter, int first, PeerView pv, unsigned long long epoch)
{
    __shared__ double sh[32];
    __shared__ double mine[3];
    const int t = threadIdx.x;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = t; i < nvec; i += blockDim.x) { a += pvec[2 * i]; b += pvec[2 * i + 1]; }
    for (int i = t; i < nif; i += blockDim.x) c += pif[i];
    for (int i = t; i < nint; i += blockDim.x) c += pint[i];
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (t == 0) { mine[0] = a; mine[1] = b; mine[2] = c; }
    __syncthreads();
    if (t < pv.nranks) {
        double* box = (double*)(pv.win[t] + kPeerMailOff) + ((epoch & 1ull) * kMaxPeers + pv.rank) * 4;
        box[0] = mine[0];
        box[1] = mine[1];
        box[2] = mine[2];
        __threadfence_system();
        if (t != pv.rank) peer_st_release((unsigned long long*)(pv.win[t] + kPeerSFlagsOff) + pv.rank, epoch);
    }
    peer_wait(pv, epoch, kPeerSFlagsOff);
    if (t == 0) {
        const double* box = (const double*)(pv.win[pv.rank] + kPeerMailOff) + (epoch & 1ull) * kMaxPeers * 4;
        double red[3] = {0.0, 0.0, 0.0};
        for (int r = 0; r < pv.nranks; r++)
            for (int i = 0; i < 3; i++) red[i] += peer_ld(box + r * 4 + i);
        cg_scalars(red, scal, ctl, rtol, atol, max_iter, first);
    }
}

#ifdef TB2_PCG_TIMELINE // lab builds only (profiles/tools/lab/pcg_timeline.sh): device-side event marks inside 8 iterations
static cudaEvent_t g_tl[8][8];
static int g_tl_it = -1;
#define TL_MARK(k, stream)                                                                     \
    do {                                                                                       \
        if (g_tl_it >= 0 && g_tl_it < 8) {                                                     \
            if (!g_tl[g_tl_it][k]) cudaEventCreate(&g_tl[g_tl_it][k]);                         \
            cudaEventRecord(g_tl[g_tl_it][k], stream);                                         \
        }                                                                                      \
    } while (0)
#else
#define TL_MARK(k, stream) do { } while (0)
#endif

static int pcg_distributed(tb2_matrix* A, const double* d_b, double* d_x, double rtol, double atol, int max_iter, int* iterations,
                           double* final_rnorm)
{
    tb2_mesh* m = A->ctx;
    const int64_t n = A->neq;
    cudaStream_t st = m->stream;
    double* scal = A->scal.p;
    PcgCtl* ctl = (PcgCtl*)(A->scal.p + kNumScal);
    const int* eqnos = A->eqs->eqnos.p;
    CommPlan cp;
    const bool overlap = comm_plan(m, &cp);
    int64_t vb = (n + 255) / 256;
    const unsigned vec_blocks = (unsigned)(vb < kReduceBlocks ? vb : kReduceBlocks);
    const unsigned nb1 = (unsigned)((n + 255) / 256);
    if (!A->eq_owned.p) {
        TB2_CUDA(A->eq_owned.alloc(n));
        k_eq_owned<<<nb1, 256, 0, st>>>(n, A->eqs->eq_node.p, comm_owned_mask(m), A->eq_owned.p);
    }
    if (!A->s.p) TB2_CUDA(A->s.alloc(n));
    if (overlap && !A->grp_split.p) { // row groups of interface nodes first, then the others (stable: deterministic partial sums)
        DevBuf<unsigned char> flag;
        DevBuf<int> nsel;
        TB2_CUDA(flag.alloc(A->ngroups > 0 ? A->ngroups : 1));
        TB2_CUDA(nsel.alloc(1));
        TB2_CUDA(A->grp_split.alloc(A->ngroups > 0 ? A->ngroups : 1));
        k_group_interface_flag<<<(unsigned)((A->ngroups + 255) / 256), 256, 0, st>>>(A->ngroups, A->grp.p, A->eqs->eq_node.p, cp.node_slot, flag.p);
        size_t tmp_bytes = 0;
        TB2_CUDA(cub::DevicePartition::Flagged(nullptr, tmp_bytes, A->grp.p, flag.p, A->grp_split.p, nsel.p, (int)A->ngroups, st));
        DevBuf<unsigned char> tmp;
        TB2_CUDA(tmp.alloc(tmp_bytes));
        TB2_CUDA(cub::DevicePartition::Flagged(tmp.p, tmp_bytes, A->grp.p, flag.p, A->grp_split.p, nsel.p, (int)A->ngroups, st));
        int h_n = 0;
        TB2_CUDA(cudaMemcpyAsync(&h_n, nsel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        TB2_CUDA(cudaStreamSynchronize(st));
        A->ngroups_if = h_n;
        // DevicePartition writes the rejected items from the back in reverse order: put them back in ascending order
        if (A->ngroups > h_n) {
            std::vector<int> h((size_t)(A->ngroups - h_n));
            TB2_CUDA(cudaMemcpy(h.data(), A->grp_split.p + h_n, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
            std::reverse(h.begin(), h.end());
            TB2_CUDA(cudaMemcpy(A->grp_split.p + h_n, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    const unsigned char* owned = A->eq_owned.p;
    double* pvec = A->partial.p;                       // [2 vec_blocks] (r,u), (r,r)
    double* pint = A->partial.p + 2 * kReduceBlocks;   // [spmv blocks] (A_loc u, u) of the interior (or all) row groups
    double* red = A->partial.p + 2 * kReduceBlocks + kSpmvBlocks; // 8 spare doubles at the end of the partial buffer
    if (!A->partial_if.p) TB2_CUDA(A->partial_if.alloc(kSpmvBlocks));
    double* pif = A->partial_if.p;
    const unsigned sg = spmv_grid();
    const int64_t ng_if = overlap ? A->ngroups_if : 0, ng_int = A->ngroups - ng_if;
    unsigned sg_if = (unsigned)((ng_if + 7) / 8);
    if (sg_if > sg) sg_if = sg;
    // w = A_loc u with the interface sum; leaves the (A_loc u, u) partials in pif / pint
    auto multiply = [&](const double* u, double* w, bool with_dot) -> int {
        if (ng_if > 0) {
            // the whole interface lane -- rows of the interface nodes, pack, pull (or all-reduce) -- on the communicator's (high-priority)
            // stream, beside the interior rows on the mesh stream.  Timeline (profiles/r02_summary.md, r02k): the persistent interior grid
            // holds every CTA slot, so the lane's kernels mostly run in the ~35 us after it; an interior grid of short CTAs, one slot
            // per SM left free, or lane kernels cut to the left-over registers hide the lane but cost the interior rows 5-10 %, more
            // than the lane is worth -- the persistent grid stays
            TB2_CUDA(cudaEventRecord(cp.ev_packed, st)); // u is ready
            TB2_CUDA(cudaStreamWaitEvent(cp.stream, cp.ev_packed, 0));
            {
                ProfScope ps(m, kProfSpmv, 1, cp.stream);
                if (with_dot) k_spmv<true><<<sg_if, 256, 0, cp.stream>>>(ng_if, A->grp_split.p, A->rowptr.p, A->colind.p, A->val.p, u, w, pif, &ctl->done);
                else k_spmv<false><<<sg_if, 256, 0, cp.stream>>>(ng_if, A->grp_split.p, A->rowptr.p, A->colind.p, A->val.p, u, w, nullptr, nullptr);
            }
            TL_MARK(2, cp.stream);
            TB2_CHECK(comm_pack_eq(m, eqnos, w, cp.stream));
            TL_MARK(3, cp.stream);
            TB2_CHECK(comm_allreduce_packed(m));
            if (cp.peer) TB2_CHECK(comm_unpack_eq(m, eqnos, w, cp.stream)); // the pull, beside the interior rows (they write other rows of w)
            TL_MARK(4, cp.stream);
            TB2_CUDA(cudaEventRecord(cp.ev_reduced, cp.stream));
        }
        {
            ProfScope ps(m, kProfSpmv, 1, st);
            const int* g = ng_if > 0 ? A->grp_split.p + ng_if : A->grp.p;
            if (with_dot) k_spmv<true><<<sg, 256, 0, st>>>(ng_int, g, A->rowptr.p, A->colind.p, A->val.p, u, w, pint, &ctl->done);
            else k_spmv<false><<<sg, 256, 0, st>>>(ng_int, g, A->rowptr.p, A->colind.p, A->val.p, u, w, nullptr, nullptr);
        }
        TL_MARK(5, st);
        if (ng_if > 0) {
            TB2_CUDA(cudaStreamWaitEvent(st, cp.ev_reduced, 0));
            if (!cp.peer) TB2_CHECK(comm_unpack_eq(m, eqnos, w, st));
        } else
            TB2_CHECK(comm_sum_interface_eq(m, eqnos, w));
        return TB2_OK;
    };
    // Jacobi preconditioner from the ASSEMBLED diagonal
    k_extract_dinv<<<nb1, 256, 0, st>>>(n, A->rowptr.p, A->colind.p, A->val.p, A->dinv.p, 0);
    TB2_CHECK(comm_sum_interface_eq(m, eqnos, A->dinv.p));
    k_invert_diag<<<nb1, 256, 0, st>>>(n, A->dinv.p);
    // r = b - A x0 ; u = M^-1 r ; p = s = 0 ; w = A u ; gamma, |r|, delta -> alpha, beta = 0
    TB2_CHECK(multiply(d_x, A->q.p, false));
    k_pcg_init<<<vec_blocks, 256, 0, st>>>(n, d_b, A->q.p, A->dinv.p, A->r.p, A->z.p, A->p.p, pvec, owned);
    TB2_CUDA(cudaMemsetAsync(A->p.p, 0, n * sizeof(double), st));
    TB2_CUDA(cudaMemsetAsync(A->s.p, 0, n * sizeof(double), st));
    TB2_CUDA(cudaMemsetAsync(ctl, 0, sizeof(PcgCtl), st));
    TB2_CHECK(multiply(A->z.p, A->q.p, true));
    // (r,u), (r,r), (A u, u): this rank's partials, summed over the ranks, alpha and beta -- one CTA over peer memory, else
    // sum + ncclAllReduce + scalar kernel
    auto scalar_step = [&](int first, bool with_pif) -> int {
        const int nif = (with_pif && ng_if > 0) ? (int)sg_if : 0; // partials of a separate interface-row launch
        if (overlap && cp.peer) {
            k_cg_reduce_peer<<<1, 1024, 0, st>>>((int)vec_blocks, pvec, nif, pif, (int)sg, pint, scal, ctl, rtol, atol, max_iter,
                                                first, cp.pv, ++*cp.sepoch);
            return TB2_OK;
        }
        k_cg_sum<<<1, 1024, 0, st>>>((int)vec_blocks, pvec, nif, pif, (int)sg, pint, red, first ? nullptr : ctl);
        TB2_CHECK(comm_allreduce_scalars(m, red, 3));
        k_cg_scalars<<<1, 1, 0, st>>>(red, scal, ctl, rtol, atol, max_iter, first);
        return TB2_OK;
    };
    TB2_CHECK(scalar_step(1, true));
    TB2_CUDA(cudaGetLastError());
    // over peer memory the SpMV publishes the interface rows itself (k_spmv_publish) and the pull runs beside the scalar step
    const bool fused = overlap && cp.peer && ng_if > 0;
    if (fused && !A->eq_xidx.p) {
        TB2_CUDA(A->eq_xidx.alloc(n));
        k_eq_exchange_index<<<nb1, 256, 0, st>>>(n, A->eqs->eq_node.p, cp.node_slot, A->eq_xidx.p);
    }
    PcgCtl h{};
    const int check_every = 8;
    for (int it = 0; it < max_iter;) {
        for (int k = 0; k < check_every && it < max_iter; k++, it++) {
#ifdef TB2_PCG_TIMELINE
            g_tl_it = it - 16;
#endif
            TL_MARK(0, st);
            {
                ProfScope ps(m, kProfPcgVec, 1);
                k_cg_update<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, A->dinv.p, owned, A->q.p, A->z.p, A->p.p, A->s.p, d_x, A->r.p, pvec);
            }
            TL_MARK(1, st);
            if (fused) {
                // one SpMV launch: the rows of the interface nodes come first in the group list and are published from inside the kernel
                // (k_spmv_publish); once it is done the pull of the sharers' partials (communicator's stream) runs beside the scalar step
                ++*cp.epoch;
                {
                    ProfScope ps(m, kProfSpmv, 1, st);
                    k_spmv_publish<<<sg, 256, 0, st>>>(A->ngroups, ng_if, A->grp_split.p, A->rowptr.p, A->colind.p, A->val.p, A->z.p, A->q.p, pint,
                                                      &ctl->done, A->eq_xidx.p, cp.pv, *cp.epoch, cp.counter);
                }
                TL_MARK(5, st);
                TB2_CUDA(cudaEventRecord(cp.ev_packed, st));
                TB2_CUDA(cudaStreamWaitEvent(cp.stream, cp.ev_packed, 0));
                TB2_CHECK(comm_unpack_eq(m, eqnos, A->q.p, cp.stream));
                TL_MARK(4, cp.stream);
                TB2_CUDA(cudaEventRecord(cp.ev_reduced, cp.stream));
            } else
                TB2_CHECK(multiply(A->z.p, A->q.p, true));
            {
                ProfScope ps(m, kProfPcgVec, 2);
                TB2_CHECK(scalar_step(0, !fused));
            }
            if (fused) TB2_CUDA(cudaStreamWaitEvent(st, cp.ev_reduced, 0)); // the next update reads the pulled rows of w
            TL_MARK(6, st);
#ifdef TB2_PCG_TIMELINE
            g_tl_it = -1;
#endif
        }
        TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        TB2_CUDA(cudaStreamSynchronize(st));
        if (h.done) break; // identical on every rank: the flag derives from all-reduced scalars
    }
    TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    double hs[kNumScal];
    TB2_CUDA(cudaMemcpyAsync(hs, scal, sizeof hs, cudaMemcpyDeviceToHost, st));
    TB2_CUDA(cudaStreamSynchronize(st));
#ifdef TB2_PCG_TIMELINE
    if (g_tl[7][6]) {
        const char* names[7] = {"start", "cg_update", "spmv_if", "pack", "pull", "spmv_int", "reduce"};
        double sum[7] = {0};
        for (int i = 0; i < 8; i++)
            for (int k = 1; k < 7; k++) {
                float ms = 0;
                if (g_tl[i][k] && cudaEventQuery(g_tl[i][k]) == cudaSuccess) cudaEventElapsedTime(&ms, g_tl[i][0], g_tl[i][k]);
                sum[k] += ms;
            }
        float per = 0;
        cudaEventElapsedTime(&per, g_tl[0][0], g_tl[7][0]);
        fprintf(stderr, "pcg timeline (us from iteration start, mean of 8; iteration %.1f us):", per / 7 * 1e3);
        for (int k = 1; k < 7; k++) fprintf(stderr, " %s %.1f", names[k], sum[k] / 8 * 1e3);
        fprintf(stderr, "\n");
    }
#endif
    if (iterations) *iterations = h.iters;
    if (final_rnorm) *final_rnorm = hs[kRNORM];
    A->pcg_last_converged = !(hs[kRNORM] > atol) || !(hs[kRNORM] > rtol * hs[kR0]); // the device's own stop test
    A->pcg_last_rel = hs[kR0] > 0.0 ? hs[kRNORM] / hs[kR0] : 0.0;
    if (h.breakdown) {
        set_error("PCG breakdown: p.Ap = %g <= 0 (matrix not positive definite)", hs[kPAP]);
        return TB2_ERR_PCG_BREAKDOWN;
    }
    return TB2_OK;
}

int tb2_matrix_pcg(tb2_matrix* A, const double* d_b, double* d_x, double rtol, double atol, int max_iter, int* iterations,
                   double* final_rnorm)
{
    TB2_ARG(A && d_b && d_x);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    if (A->eqs && comm_active(m)) return pcg_distributed(A, d_b, d_x, rtol, atol, max_iter, iterations, final_rnorm);
    const int64_t n = A->neq;
    cudaStream_t st = m->stream;
    double* scal = A->scal.p;
    PcgCtl* ctl = (PcgCtl*)(A->scal.p + kNumScal);
    const unsigned spmv_blocks = spmv_grid();
    int64_t vb = (n + 255) / 256;
    const unsigned vec_blocks = (unsigned)(vb < kReduceBlocks ? vb : kReduceBlocks);
    {
        ProfScope ps(m, kProfPcgVec, 4);
        k_extract_dinv<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, A->rowptr.p, A->colind.p, A->val.p, A->dinv.p, 1);
        k_spmv<false><<<spmv_blocks, 256, 0, st>>>(A->ngroups, A->grp.p, A->rowptr.p, A->colind.p, A->val.p, d_x, A->q.p, nullptr, nullptr);
        k_pcg_init<<<vec_blocks, 256, 0, st>>>(n, d_b, A->q.p, A->dinv.p, A->r.p, A->z.p, A->p.p, A->partial.p);
        k_reduce_init<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, rtol, atol, max_iter);
    }
    TB2_CUDA(cudaGetLastError());
    PcgCtl h{};
    const int check_every = 16;
    auto enqueue_iteration = [&]() {
        {
            ProfScope ps(m, kProfSpmv);
            k_spmv<true><<<spmv_blocks, 256, 0, st>>>(A->ngroups, A->grp.p, A->rowptr.p, A->colind.p, A->val.p, A->p.p, A->q.p, A->partial.p, &ctl->done);
        }
        ProfScope ps(m, kProfPcgVec, 4);
        k_reduce_pap<<<1, 1024, 0, st>>>((int)spmv_blocks, A->partial.p, scal, ctl);
        k_pcg_update<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, A->p.p, A->q.p, A->dinv.p, d_x, A->r.p, A->z.p, A->partial.p);
        k_reduce_rz<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, rtol, atol, max_iter);
        k_pcg_direction<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, A->z.p, A->p.p);
    };
    // The iteration is launch-bound at the margin (5 launches, ~7 us of gaps each beside 0.44 ms of kernels): 16 iterations are
    // captured once into a CUDA graph and replayed; every kernel is a no-op once the device-side `done` flag is set, so whole
    // batches can always be launched.  Not used while per-launch profiling events are being recorded (TB2_PCG_GRAPH=0 disables).
    static const bool graphs_on = !(getenv("TB2_PCG_GRAPH") && getenv("TB2_PCG_GRAPH")[0] == '0');
    const bool use_graph = graphs_on && !m->prof_on && max_iter >= check_every;
    if (use_graph && (!A->pcg_exec || A->pcg_x != d_x || A->pcg_rtol != rtol || A->pcg_atol != atol || A->pcg_maxit != max_iter)) {
        if (A->pcg_exec) cudaGraphExecDestroy(A->pcg_exec);
        A->pcg_exec = nullptr;
        cudaGraph_t graph = nullptr;
        const uint64_t launches_before = m->launches;
        TB2_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < check_every; k++) enqueue_iteration();
        TB2_CUDA(cudaStreamEndCapture(st, &graph));
        m->launches = launches_before; // nothing ran yet
        TB2_CUDA(cudaGraphInstantiate(&A->pcg_exec, graph, 0));
        cudaGraphDestroy(graph);
        A->pcg_x = d_x;
        A->pcg_rtol = rtol;
        A->pcg_atol = atol;
        A->pcg_maxit = max_iter;
    }
    for (int it = 0; it < max_iter;) {
        if (use_graph) {
            TB2_CUDA(cudaGraphLaunch(A->pcg_exec, st));
            m->launches += 5 * check_every;
            it += check_every;
        } else
            for (int k = 0; k < check_every && it < max_iter; k++, it++) enqueue_iteration();
        TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        TB2_CUDA(cudaStreamSynchronize(st));
        if (h.done) break;
    }
    TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    double hs[kNumScal];
    TB2_CUDA(cudaMemcpyAsync(hs, scal, sizeof hs, cudaMemcpyDeviceToHost, st));
    TB2_CUDA(cudaStreamSynchronize(st));
    if (iterations) *iterations = h.iters;
    if (final_rnorm) *final_rnorm = hs[kRNORM];
    A->pcg_last_converged = !(hs[kRNORM] > atol) || !(hs[kRNORM] > rtol * hs[kR0]); // the device's own stop test (k_reduce_rz)
    A->pcg_last_rel = hs[kR0] > 0.0 ? hs[kRNORM] / hs[kR0] : 0.0;
    if (h.breakdown) {
        set_error("PCG breakdown: p.Ap = %g <= 0 (matrix not positive definite)", hs[kPAP]);
        return TB2_ERR_PCG_BREAKDOWN;
    }
    return TB2_OK;
}

// ---- BiCGStab with the Jacobi preconditioner: the device's linear solve for NON-symmetric tangents ---------------------------
// J2Simo3D::TangentType() is kNonSymmetric (J2Simo3D.cpp:18-21); the reference solves such systems by LU (SolverT.cpp:1108-1109:
// profile_matrix -> CCNSMatrixT) and has a Krylov precedent in NLSolver_NK.cpp:147,175.  van der Vorst's recurrence, right
// preconditioning with M = diag(A) (DiagonalMatrixT::Factorize semantics), two SpMVs per iteration, every scalar on the device,
// deterministic two-stage reductions, kernels no-ops once the device-side done flag is set -- the same machinery as tb2_matrix_pcg.
} // extern "C"
namespace tb2 {
enum { kBiRho = 0, kBiAlpha = 1, kBiOmega = 2, kBiBeta = 3 }; // in scal[kRZ..]: rho, alpha, omega, beta share the PCG scalar block
// p = r + beta (p - omega v) ; y = dinv p
__global__ void __launch_bounds__(256) k_bi_direction(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl,
                                                     const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv,
                                                     double* __restrict__ p, double* __restrict__ y)
{
    if (ctl->done) return;
    const double beta = scal[kBETA], omega = scal[kRR];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double pi = r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        y[i] = dinv[i] * pi;
    }
}
// partial[2b], partial[2b+1] = sum a.b, sum c.d
__global__ void __launch_bounds__(256) k_bi_dot2(int64_t n, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                                const double* __restrict__ d, double* __restrict__ partial, const PcgCtl* __restrict__ ctl)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        s0 += a[i] * b[i];
        s1 += c[i] * d[i];
    }
    s0 = block_sum(s0, sh);
    s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0;
        partial[2 * blockIdx.x + 1] = s1;
    }
}
// stage 0: alpha = rho / (rhat, v).  stage 1: omega = (t, s) / (t, t).  stage 2: rho' = (rhat, r), |r|; beta = (rho'/rho)(alpha/omega)
__global__ void __launch_bounds__(1024) k_bi_scalars(int nparts, const double* __restrict__ partial, double* __restrict__ scal, PcgCtl* ctl, int stage,
                                                    double rtol, double atol, int max_iter)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x != 0) return;
    if (stage == 0) {
        if (a == 0.0) { ctl->breakdown = 1; ctl->done = 1; scal[kALPHA] = 0.0; }
        else scal[kALPHA] = scal[kRZ] / a;
    } else if (stage == 1) {
        scal[kRR] = b > 0.0 ? a / b : 0.0; // omega
    } else {
        const double rnorm = sqrt(b);
        scal[kRNORM] = rnorm;
        ctl->iters += 1;
        if (!(rnorm > atol) || !(rnorm > rtol * scal[kR0]) || ctl->iters >= max_iter) ctl->done = 1;
        else if (scal[kRZ] == 0.0 || scal[kRR] == 0.0) { ctl->breakdown = 1; ctl->done = 1; }
        else {
            scal[kBETA] = (a / scal[kRZ]) * (scal[kALPHA] / scal[kRR]);
            scal[kRZ] = a;
        }
    }
}
// s = r - alpha v ; z = dinv s
__global__ void __launch_bounds__(256) k_bi_half(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl, const double* __restrict__ r,
                                                const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ sv, double* __restrict__ z)
{
    if (ctl->done) return;
    const double alpha = scal[kALPHA];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double si = r[i] - alpha * v[i];
        sv[i] = si;
        z[i] = dinv[i] * si;
    }
}
// x += alpha y + omega z ; r = s - omega t ; partials (rhat, r), (r, r)
__global__ void __launch_bounds__(256) k_bi_update(int64_t n, const double* __restrict__ scal, const PcgCtl* __restrict__ ctl, const double* __restrict__ y,
                                                  const double* __restrict__ z, const double* __restrict__ sv, const double* __restrict__ t,
                                                  const double* __restrict__ rhat, double* __restrict__ x, double* __restrict__ r,
                                                  double* __restrict__ partial)
{
    if (ctl->done) return;
    __shared__ double sh[32];
    const double alpha = scal[kALPHA], omega = scal[kRR];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * y[i] + omega * z[i];
        const double ri = sv[i] - omega * t[i];
        r[i] = ri;
        s0 += rhat[i] * ri;
        s1 += ri * ri;
    }
    s0 = block_sum(s0, sh);
    s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0;
        partial[2 * blockIdx.x + 1] = s1;
    }
}
// r = b - q ; rhat = r ; p = v = 0 ; partials (r, r) twice
__global__ void __launch_bounds__(256) k_bi_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r,
                                                double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ v, double* __restrict__ partial)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - q[i];
        r[i] = ri;
        rhat[i] = ri;
        p[i] = 0.0;
        v[i] = 0.0;
        s += ri * ri;
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = s;
    }
}
__global__ void __launch_bounds__(1024) k_bi_scalars_init(int nparts, const double* __restrict__ partial, double* __restrict__ scal, PcgCtl* ctl,
                                                         double rtol, double atol, int max_iter)
{
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += partial[2 * i];
    a = block_sum(a, sh);
    if (threadIdx.x != 0) return;
    const double rnorm = sqrt(a);
    scal[kRZ] = a;      // rho = (rhat, r) = (r, r)
    scal[kALPHA] = 1.0;
    scal[kRR] = 1.0;    // omega
    scal[kBETA] = 0.0;  // first direction: p = r
    scal[kR0] = rnorm;
    scal[kRNORM] = rnorm;
    ctl->iters = 0;
    ctl->breakdown = 0;
    ctl->done = (!(rnorm > atol) || !(rnorm > rtol * rnorm) || max_iter <= 0) ? 1 : 0;
}
} // namespace tb2
extern "C" {

int tb2_matrix_bicgstab(tb2_matrix* A, const double* d_b, double* d_x, double rtol, double atol, int max_iter, int* iterations,
                        double* final_rnorm)
{
    TB2_ARG(A && d_b && d_x);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    if (A->eqs && comm_active(m)) {
        set_error("tb2_matrix_bicgstab: the partitioned (multi-GPU) form is not implemented");
        return TB2_ERR_ARG;
    }
    const int64_t n = A->neq;
    cudaStream_t st = m->stream;
    double* scal = A->scal.p;
    PcgCtl* ctl = (PcgCtl*)(A->scal.p + kNumScal);
    for (auto* b : {&A->bi_rhat, &A->bi_v, &A->bi_s, &A->bi_t})
        if (!b->p) TB2_CUDA(b->alloc(n));
    double *r = A->r.p, *y = A->z.p, *p = A->p.p, *z = A->q.p, *rhat = A->bi_rhat.p, *v = A->bi_v.p, *sv = A->bi_s.p, *t = A->bi_t.p;
    const unsigned sg = spmv_grid();
    int64_t vb = (n + 255) / 256;
    const unsigned vec_blocks = (unsigned)(vb < kReduceBlocks ? vb : kReduceBlocks);
    auto spmv = [&](const double* x, double* out) {
        ProfScope ps(m, kProfSpmv);
        k_spmv<false><<<sg, 256, 0, st>>>(A->ngroups, A->grp.p, A->rowptr.p, A->colind.p, A->val.p, x, out, nullptr, nullptr);
    };
    {
        ProfScope ps(m, kProfPcgVec, 1);
        k_extract_dinv<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, A->rowptr.p, A->colind.p, A->val.p, A->dinv.p, 1);
    }
    PcgCtl h{};
    double hs[kNumScal];
    double r0_first = -1.0;
    int done_iters = 0;
    const int check_every = 8, max_restarts = 50;
    // A breakdown ((rhat, v), rho or omega vanish: the shadow residual has lost its grip) is answered the standard way: restart
    // from the current iterate with rhat = r.  Tolerances stay relative to the FIRST residual.
    for (int attempt = 0; attempt <= max_restarts; attempt++) {
        {
            ProfScope ps(m, kProfPcgVec, 3);
            spmv(d_x, t);
            k_bi_init<<<vec_blocks, 256, 0, st>>>(n, d_b, t, r, rhat, p, v, A->partial.p);
            k_bi_scalars_init<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, rtol, atol, max_iter - done_iters);
        }
        if (attempt > 0) { // keep the stop test relative to the first residual
            TB2_CUDA(cudaMemcpyAsync(scal + kR0, &r0_first, sizeof(double), cudaMemcpyHostToDevice, st));
        }
        TB2_CUDA(cudaGetLastError());
        for (int it = 0; it < max_iter - done_iters;) {
            for (int k = 0; k < check_every && it < max_iter - done_iters; k++, it++) {
                {
                    ProfScope ps(m, kProfPcgVec, 1);
                    k_bi_direction<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, r, v, A->dinv.p, p, y);
                }
                spmv(y, v);
                {
                    ProfScope ps(m, kProfPcgVec, 3);
                    k_bi_dot2<<<vec_blocks, 256, 0, st>>>(n, rhat, v, rhat, v, A->partial.p, ctl);
                    k_bi_scalars<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, 0, rtol, atol, max_iter - done_iters);
                    k_bi_half<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, r, v, A->dinv.p, sv, z);
                }
                spmv(z, t);
                ProfScope ps(m, kProfPcgVec, 4);
                k_bi_dot2<<<vec_blocks, 256, 0, st>>>(n, t, sv, t, t, A->partial.p, ctl);
                k_bi_scalars<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, 1, rtol, atol, max_iter - done_iters);
                k_bi_update<<<vec_blocks, 256, 0, st>>>(n, scal, ctl, y, z, sv, t, rhat, d_x, r, A->partial.p);
                k_bi_scalars<<<1, 1024, 0, st>>>((int)vec_blocks, A->partial.p, scal, ctl, 2, rtol, atol, max_iter - done_iters);
            }
            TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
            TB2_CUDA(cudaStreamSynchronize(st));
            if (h.done) break;
        }
        TB2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        TB2_CUDA(cudaMemcpyAsync(hs, scal, sizeof hs, cudaMemcpyDeviceToHost, st));
        TB2_CUDA(cudaStreamSynchronize(st));
        if (attempt == 0) r0_first = hs[kR0];
        done_iters += h.iters;
        const bool converged = !(hs[kRNORM] > atol) || !(hs[kRNORM] > rtol * r0_first);
        if (converged || !h.breakdown || done_iters >= max_iter) break;
    }
    h.iters = done_iters;
    hs[kR0] = r0_first;
    if (iterations) *iterations = h.iters;
    if (final_rnorm) *final_rnorm = hs[kRNORM];
    A->pcg_last_converged = !(hs[kRNORM] > atol) || !(hs[kRNORM] > rtol * hs[kR0]);
    A->pcg_last_rel = hs[kR0] > 0.0 ? hs[kRNORM] / hs[kR0] : 0.0;
    if (h.breakdown && !A->pcg_last_converged) {
        set_error("BiCGStab breakdown (rho, omega or (rhat, v) vanished) after %d iterations, |r|/|r0| = %g", h.iters, A->pcg_last_rel);
        return TB2_ERR_PCG_BREAKDOWN;
    }
    return TB2_OK;
}

int tb2_matrix_bicgstab_host(tb2_matrix* A, const double* h_b, double* h_x, double rtol, double atol, int max_iter, int* iterations,
                             double* final_rnorm)
{
    TB2_ARG(A && h_b && h_x);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    DevBuf<double> b, x;
    TB2_CUDA(b.alloc(A->neq));
    TB2_CUDA(x.alloc(A->neq));
    TB2_CUDA(cudaMemcpyAsync(b.p, h_b, A->neq * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(x.p, h_x, A->neq * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    int s = tb2_matrix_bicgstab(A, b.p, x.p, rtol, atol, max_iter, iterations, final_rnorm);
    TB2_CUDA(cudaMemcpyAsync(h_x, x.p, A->neq * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return s;
}

int tb2_matrix_pcg_converged(const tb2_matrix* A, int* converged, double* relative_residual)
{
    TB2_ARG(A);
    if (converged) *converged = A->pcg_last_converged ? 1 : 0;
    if (relative_residual) *relative_residual = A->pcg_last_rel;
    return TB2_OK;
}

int tb2_matrix_pcg_host(tb2_matrix* A, const double* h_b, double* h_x, double rtol, double atol, int max_iter, int* iterations,
                        double* final_rnorm)
{
    TB2_ARG(A && h_b && h_x);
    tb2_mesh* m = A->ctx;
    DeviceGuard dg(m->device);
    DevBuf<double> b, x;
    TB2_CUDA(b.alloc(A->neq));
    TB2_CUDA(x.alloc(A->neq));
    TB2_CUDA(cudaMemcpyAsync(b.p, h_b, A->neq * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(x.p, h_x, A->neq * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    int s = tb2_matrix_pcg(A, b.p, x.p, rtol, atol, max_iter, iterations, final_rnorm);
    TB2_CUDA(cudaMemcpyAsync(h_x, x.p, A->neq * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return s;
}

} // extern "C"
