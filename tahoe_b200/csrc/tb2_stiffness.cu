// tb2_stiffness.cu -- K3: element tangent B^T c B (+ geometric stiffness) assembled in place into the device CSR,
// and the element colouring that makes the assembly free of float atomics.
//
// Replaces SolidElementT::ElementLHSDriver (SolidElementT.cpp:1100-1154), SmallStrainT::FormStiffness (SmallStrainT.cpp:285-324),
// TotalLagrangianT::FormStiffness (TotalLagrangianT.cpp:40-104), UpdatedLagrangianT::FormStiffness (UpdatedLagrangianT.cpp:94-142)
// and MSRMatrixT::Assemble (MSRMatrixT.cpp:66-216).
//
// Work split: 8 threads per element.  Phase 1: thread t evaluates integration point t (Jacobian, F, stress, spatial tangent c,
// spatial dN/dx, w det j) and parks it in shared memory.  Phase 2: thread t owns row-node a = t and accumulates the 3x24 row block
// K[a][:] over the 8 points in registers:  K_ab = sum_ip (B_a^T c) B_b + (dN_a . sigma dN_b) 1.  Both TL and UL use spatial
// gradients and the scale w det j (the reference's TL pushes dN/dX forward with F^-1 and scales by J detJ0: identical).
// Phase 3: each thread adds its 3x24 block to the CSR through the element->adjacency map.  Elements are processed colour by
// colour (no two elements of a colour share a node), so plain read-modify-write is race free and the summation order per
// matrix entry is fixed: colour order.
#include "tb2_internal.h"

namespace tb2 {

int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration);
J2Hist group_hist(tb2_group* g);

// ---- colouring ------------------------------------------------------------------------------------------------------
// colour(e) = smallest colour not used by any lower-numbered element sharing a node with e: the sequential greedy colouring
// in element order (oracle/tahoe_oracle.c: orc_greedy_colouring), evaluated as a dependency wavefront: an element is coloured
// in the round in which all its lower-numbered neighbours are coloured.
__global__ void __launch_bounds__(256) k_colour_round(int64_t ne, int64_t stride, const int* __restrict__ conn, const int* __restrict__ inc_ptr,
                                                     const int* __restrict__ inc, const int* colour_in, int* colour_out, int* remaining,
                                                     int* overflow)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    if (colour_in[e] >= 0) return;
    unsigned long long mask = 0ull;
    for (int a = 0; a < 8; a++) {
        const int n = conn[a * stride + e];
        const int k0 = inc_ptr[n];
        // incidence entries are ascending in element id: walk down from the largest lower-numbered neighbour
        for (int k = inc_ptr[n + 1] - 1; k >= k0; k--) {
            const int64_t f = inc[k] >> 3;
            if (f >= e) continue;
            const int c = colour_in[f];
            if (c < 0) { atomicAdd(remaining, 1); return; } // not ready this round
            mask |= 1ull << c;
        }
    }
    int c = 0;
    while (c < 64 && ((mask >> c) & 1ull)) c++;
    if (c >= 64) { *overflow = 1; c = 63; }
    colour_out[e] = c;
}

struct StiffArgs {
    int64_t ne, stride;
    const int* conn;
    const double* X;
    const double* u;
    const double* ul;
    MatConst mat;
    J2Hist hist;
    int iteration;
    unsigned long long* status;
    // colour slice
    const int* elems; // element ids of this colour
    int64_t count;
    // matrix
    const int* eqnos;         // [nn][3]
    const long long* rowptr;  // [neq+1]
    const int* adj_coloff;    // per adjacency entry
    const int* elem_adjpos;   // [64][stride]
    double* val;
};

static const int kElemsPerBlock = 16;
static const int kIpDoubles = 24 + 36 + 6 + 2; // dN/dx[3][8], c[6][6], sigma[6], scale, pad

template <int FORM, int MAT>
__global__ void __launch_bounds__(128) k_stiffness(const StiffArgs p)
{
    // shared: [ip][field][element-in-block]  (element minor: the 4 elements of a warp hit 4 consecutive words)
    extern __shared__ double sm[];
    const int el = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int64_t idx = blockIdx.x * (int64_t)kElemsPerBlock + el;
    const bool live = idx < p.count;
    const int64_t e = live ? p.elems[idx] : 0;
#define SM(ip, f) sm[((ip)*kIpDoubles + (f)) * kElemsPerBlock + el]
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + e);

    if (live) { // ---- phase 1: integration point t
        Modes cX, cU;
        load_modes(p.X, n, cX);
        load_modes(p.u, n, cU);
        const int ip = t;
        double s0, s1, s2;
        ip_signs(ip, s0, s1, s2);
        double J0[3][3], H[3][3], J0a[3][3], ja[3][3], c[6][6], sig[6];
        mode_gradient(cX, s0, s1, s2, J0);
        mode_gradient(cU, s0, s1, s2, H);
        const double det0 = adj3(J0, J0a);
        int err = det0 <= 0.0 ? kErrBadJacobian : kErrNone;
        double scale;
        if (FORM == kSmallStrain) {
            hooke_moduli(p.mat, c);
#pragma unroll
            for (int I = 0; I < 6; I++) sig[I] = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) ja[i][k] = J0a[i][k];
            scale = det0;
        } else {
            double j[3][3], F[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) j[i][k] = J0[i][k] + H[i][k];
            const double detj = adj3(j, ja);
            if (detj <= 0.0) err = kErrBadJacobian;
            const double rdet0 = 1.0 / det0;
            mul3(j, J0a, F);
            scale3(F, rdet0);
            const double J = detj * rdet0;
            if (MAT == kFDKStV) {
                fdkstv_stress(p.mat, F, J, sig);
                fdkstv_moduli(p.mat, F, J, c);
            } else if (MAT == kSimoIso) {
                double b_bar[6];
                simo_bbar(F, J, b_bar);
                simo_cauchy(p.mat, J, b_bar, sig);
                simo_moduli(p.mat, J, b_bar, c);
            } else if (MAT == kJ2Simo) {
                Modes cL;
                load_modes(p.ul, n, cL);
                double Hl[3][3], Fl[3][3];
                mode_gradient(cL, s0, s1, s2, Hl);
                mul3(Hl, J0a, Fl);
                scale3(Fl, rdet0);
                Fl[0][0] += 1.0; Fl[1][1] += 1.0; Fl[2][2] += 1.0;
                // allocation was settled by the force sweep that tb2_form_stiffness runs first (reference order: FormRHS, then FormLHS)
                int alloc = p.hist.alloc[e];
                const int e2 = j2_eval<true>(p.mat, p.hist, e, ip, alloc, p.iteration, F, Fl, J, sig, c);
                if (e2 > err) err = e2;
            }
            scale = detj;
        }
        if (err) {
            atomicMax(p.status, (unsigned long long)err);
            atomicMin(p.status + 1, (unsigned long long)e);
        }
        // spatial gradients: dN_b/dx_i = sum_k dN_b/dxi_k inv[k][i], inv = adj / det
        const double rs = 1.0 / scale;
        const double RA[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, SA[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, TA[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const double tr = 1.0 + RA[b] * s0 * TB2_G, ts = 1.0 + SA[b] * s1 * TB2_G, tt = 1.0 + TA[b] * s2 * TB2_G;
            const double d0 = 0.125 * RA[b] * ts * tt, d1 = 0.125 * tr * SA[b] * tt, d2 = 0.125 * tr * ts * TA[b];
#pragma unroll
            for (int i = 0; i < 3; i++) SM(ip, i * 8 + b) = (d0 * ja[0][i] + d1 * ja[1][i] + d2 * ja[2][i]) * rs;
        }
#pragma unroll
        for (int I = 0; I < 6; I++)
#pragma unroll
            for (int Jj = 0; Jj < 6; Jj++) SM(ip, 24 + I * 6 + Jj) = c[I][Jj];
#pragma unroll
        for (int I = 0; I < 6; I++) SM(ip, 60 + I) = sig[I];
        SM(ip, 66) = scale;
    }
    __syncthreads();
    if (!live) return;

    // ---- phase 2: row node a = t
    const int a = t;
    double K[3][24];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int q = 0; q < 24; q++) K[i][q] = 0.0;
#pragma unroll 1
    for (int ip = 0; ip < 8; ip++) {
        const double w = SM(ip, 66);
        const double nx = SM(ip, 0 + a) * w, ny = SM(ip, 8 + a) * w, nz = SM(ip, 16 + a) * w;
        double D[3][6]; // w * B_a^T c
#pragma unroll
        for (int Jj = 0; Jj < 6; Jj++) {
            const double c0 = SM(ip, 24 + 0 * 6 + Jj), c1 = SM(ip, 24 + 1 * 6 + Jj), c2 = SM(ip, 24 + 2 * 6 + Jj);
            const double c3 = SM(ip, 24 + 3 * 6 + Jj), c4 = SM(ip, 24 + 4 * 6 + Jj), c5 = SM(ip, 24 + 5 * 6 + Jj);
            D[0][Jj] = nx * c0 + nz * c4 + ny * c5;
            D[1][Jj] = ny * c1 + nz * c3 + nx * c5;
            D[2][Jj] = nz * c2 + ny * c3 + nx * c4;
        }
        double gx = 0.0, gy = 0.0, gz = 0.0; // w * sigma dN_a
        if (FORM != kSmallStrain) {
            const double sg0 = SM(ip, 60), sg1 = SM(ip, 61), sg2 = SM(ip, 62), sg3 = SM(ip, 63), sg4 = SM(ip, 64), sg5 = SM(ip, 65);
            gx = sg0 * nx + sg5 * ny + sg4 * nz;
            gy = sg5 * nx + sg1 * ny + sg3 * nz;
            gz = sg4 * nx + sg3 * ny + sg2 * nz;
        }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const double bx = SM(ip, 0 + b), by = SM(ip, 8 + b), bz = SM(ip, 16 + b);
            const double geo = FORM != kSmallStrain ? gx * bx + gy * by + gz * bz : 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                K[i][3 * b + 0] += D[i][0] * bx + D[i][4] * bz + D[i][5] * by;
                K[i][3 * b + 1] += D[i][1] * by + D[i][3] * bz + D[i][5] * bx;
                K[i][3 * b + 2] += D[i][2] * bz + D[i][3] * by + D[i][4] * bx;
            }
            K[0][3 * b + 0] += geo;
            K[1][3 * b + 1] += geo;
            K[2][3 * b + 2] += geo;
        }
    }
#undef SM
    // ---- phase 3: add to CSR (MSRMatrixT::Assemble: inactive equations are skipped, MSRMatrixT.cpp:118-123)
    const int64_t na = n[a];
    int eqr[3];
#pragma unroll
    for (int i = 0; i < 3; i++) eqr[i] = __ldg(p.eqnos + 3 * na + i);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int64_t nb = n[b];
        const int off = __ldg(p.adj_coloff + __ldg(p.elem_adjpos + (int64_t)(a * 8 + b) * p.stride + e));
        int eqc[3];
#pragma unroll
        for (int j = 0; j < 3; j++) eqc[j] = __ldg(p.eqnos + 3 * nb + j);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (eqr[i] <= 0) continue;
            long long pos = p.rowptr[eqr[i] - 1] + off;
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (eqc[j] > 0) {
                    p.val[pos] += K[i][3 * b + j];
                    pos++;
                }
        }
    }
}

typedef void (*stiff_kernel_t)(const StiffArgs);
static stiff_kernel_t pick_stiff_kernel(int form, int mat)
{
    if (form == kSmallStrain) return mat == kSSKStV ? k_stiffness<kSmallStrain, kSSKStV> : nullptr;
    switch (mat) {
    case kFDKStV: return k_stiffness<kTotalLagrangian, kFDKStV>;
    case kSimoIso: return k_stiffness<kTotalLagrangian, kSimoIso>;
    case kJ2Simo: return k_stiffness<kTotalLagrangian, kJ2Simo>;
    }
    return nullptr;
}

int ensure_colouring(tb2_mesh* m)
{
    if (m->ncolours > 0) return TB2_OK;
    const int64_t ne = m->ne;
    DevBuf<int> col_a, col_b, ctr;
    TB2_CUDA(col_a.alloc(ne));
    TB2_CUDA(col_b.alloc(ne));
    TB2_CUDA(ctr.alloc(2));
    TB2_CUDA(cudaMemsetAsync(col_a.p, 0xff, ne * sizeof(int), m->stream));
    TB2_CUDA(cudaMemsetAsync(col_b.p, 0xff, ne * sizeof(int), m->stream));
    TB2_CUDA(cudaMemsetAsync(ctr.p, 0, 2 * sizeof(int), m->stream));
    const int T = 256;
    const unsigned nb = (unsigned)((ne + T - 1) / T);
    int h[2] = {1, 0};
    // Jacobi-style rounds: read colours of round r-1, write round r, then copy forward (the copy keeps both buffers equal on
    // coloured elements so a round never sees a half-updated neighbour)
    for (int64_t round = 0; round < 8 * ne + 8 && h[0]; round++) {
        TB2_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int), m->stream));
        k_colour_round<<<nb, T, 0, m->stream>>>(ne, m->stride, m->conn.p, m->inc_ptr.p, m->inc.p, col_a.p, col_b.p, ctr.p, ctr.p + 1);
        TB2_CUDA(cudaMemcpyAsync(col_a.p, col_b.p, ne * sizeof(int), cudaMemcpyDeviceToDevice, m->stream));
        if ((round & 15) == 15 || ne < 4096) {
            TB2_CUDA(cudaMemcpyAsync(h, ctr.p, sizeof h, cudaMemcpyDeviceToHost, m->stream));
            TB2_CUDA(cudaStreamSynchronize(m->stream));
        }
    }
    TB2_CUDA(cudaMemcpyAsync(h, ctr.p, sizeof h, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    if (h[1]) {
        set_error("element colouring needs more than 64 colours");
        return TB2_ERR_SIZE;
    }
    m->colour_host.resize(ne);
    TB2_CUDA(cudaMemcpy(m->colour_host.data(), col_a.p, ne * sizeof(int), cudaMemcpyDeviceToHost));
    // bucket elements by colour, ascending element id inside a colour (index bookkeeping of the launch lists)
    int nc = 0;
    for (int64_t e = 0; e < ne; e++) nc = m->colour_host[e] + 1 > nc ? m->colour_host[e] + 1 : nc;
    m->colour_start.assign(nc + 1, 0);
    for (int64_t e = 0; e < ne; e++) m->colour_start[m->colour_host[e] + 1]++;
    for (int c = 0; c < nc; c++) m->colour_start[c + 1] += m->colour_start[c];
    std::vector<int> order(ne);
    std::vector<int64_t> fill(m->colour_start.begin(), m->colour_start.end() - 1);
    for (int64_t e = 0; e < ne; e++) order[fill[m->colour_host[e]]++] = (int)e;
    TB2_CUDA(m->colour_elems.alloc(ne));
    TB2_CUDA(cudaMemcpy(m->colour_elems.p, order.data(), ne * sizeof(int), cudaMemcpyHostToDevice));
    m->ncolours = nc;
    return TB2_OK;
}

} // namespace tb2

using namespace tb2;

extern "C" {

int tb2_mesh_colouring(tb2_mesh* m, int32_t* h_colour, int32_t* num_colours)
{
    TB2_ARG(m);
    DeviceGuard dg(m->device);
    TB2_CHECK(ensure_colouring(m));
    if (h_colour) memcpy(h_colour, m->colour_host.data(), m->ne * sizeof(int32_t));
    if (num_colours) *num_colours = m->ncolours;
    return TB2_OK;
}

int tb2_form_stiffness(tb2_group* g, tb2_matrix* A, const double* d_u, const double* d_ul, int iteration)
{
    TB2_ARG(g && A && d_u);
    tb2_mesh* m = g->mesh;
    TB2_ARG(A->eqs && A->eqs->mesh == m);
    DeviceGuard dg(m->device);
    stiff_kernel_t k = pick_stiff_kernel(g->form, g->mat.kind);
    TB2_ARG(k != nullptr);
    TB2_CHECK(ensure_colouring(m));
    if (g->mat.kind == TB2_J2_SIMO) {
        TB2_ARG(d_ul != nullptr);
        TB2_CHECK(launch_element_forces(g, d_u, d_ul, iteration)); // settles element allocation in the reference's order
    }
    StiffArgs p;
    p.ne = m->ne;
    p.stride = m->stride;
    p.conn = m->conn.p;
    p.X = m->X.p;
    p.u = d_u;
    p.ul = d_ul;
    p.mat = g->mc;
    p.hist = group_hist(g);
    p.iteration = iteration;
    p.status = g->status.p;
    p.eqnos = A->eqs->eqnos.p;
    p.rowptr = A->rowptr.p;
    p.adj_coloff = A->adj_coloff.p;
    p.elem_adjpos = A->elem_adjpos.p;
    p.val = A->val.p;
    const size_t smem = (size_t)8 * kIpDoubles * kElemsPerBlock * sizeof(double);
    TB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int c = 0; c < m->ncolours; c++) {
        p.elems = m->colour_elems.p + m->colour_start[c];
        p.count = m->colour_start[c + 1] - m->colour_start[c];
        if (!p.count) continue;
        ProfScope ps(m, kProfStiffness);
        k<<<(unsigned)((p.count + kElemsPerBlock - 1) / kElemsPerBlock), 128, smem, m->stream>>>(p);
    }
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int tb2_form_stiffness_host(tb2_group* g, tb2_matrix* A, const double* h_u, const double* h_ul, int iteration)
{
    TB2_ARG(g && A && h_u);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    if (!m->stage_a.p) TB2_CUDA(m->stage_a.alloc(3 * m->nn));
    TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_ul) {
        if (!m->stage_c.p) TB2_CUDA(m->stage_c.alloc(3 * m->nn));
        TB2_CUDA(cudaMemcpyAsync(m->stage_c.p, h_ul, bytes, cudaMemcpyHostToDevice, m->stream));
    }
    TB2_CHECK(tb2_form_stiffness(g, A, m->stage_a.p, h_ul ? m->stage_c.p : nullptr, iteration));
    return tb2_group_status(g, nullptr);
}

} // extern "C"
