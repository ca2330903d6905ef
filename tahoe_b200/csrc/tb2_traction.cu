// tb2_traction.cu -- natural_bc tractions on Hex8 facets: ContinuumElementT::ApplyTractionBC
// (ContinuumElementT.cpp:514-665) as two kernels.
//
//   k_traction_cards : one thread per traction card (element facet).  Facet nodes HexahedronT::NodesOnFacet (HexahedronT.cpp:1913-1918),
//                      4-node quad facet shape with the 2x2 rule a hexahedron's facets get (DomainIntegrationT.cpp:117-131), surface
//                      Jacobian |x,r x x,s| on the initial coordinates, optional local frame Q (ParentDomainT.cpp:362-422).
//                      Writes the card's 4x3 nodal forces to a scratch record.
//   k_traction_nodes : one thread per loaded node: sums its card records in ascending card order -- the order in which the
//                      reference assembles the cards (ElementSupportT::AssembleRHS per card) -- so the result does not depend on
//                      the launch shape and needs no float atomics.
//
// The surface is O(N^(2/3)) of the mesh, so this is never a bandwidth question; it is on the device so that a device-resident step
// (tb2_explicit_run, the nonlinear solvers) can take scheduled traction loads without a host round trip.
#include "tb2_internal.h"

#include <algorithm>
#include <vector>

using namespace tb2;

struct tb2_traction {
    tb2_mesh* mesh = nullptr;
    int64_t ncards = 0, nloaded = 0;
    int coord_system = 0;
    tb2::DevBuf<int> elem, facet; // [ncards]
    tb2::DevBuf<int> fnode;       // [ncards][4] global node of each facet node
    tb2::DevBuf<double> tract;    // [ncards][4][3] nodal traction vectors (Traction_CardT::fValues, facet-node order)
    tb2::DevBuf<double> rec;      // [ncards][4][3] card forces of the last evaluation
    tb2::DevBuf<int> node;        // [nloaded] loaded nodes, ascending
    tb2::DevBuf<int> slot_ptr;    // [nloaded+1]
    tb2::DevBuf<int> slot;        // [4*ncards] card*4+a, ascending within a node
    tb2::DevBuf<unsigned long long> status;
};

namespace {

__constant__ int c_facet_nodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};

__global__ void k_facet_nodes(int64_t ncards, int64_t stride, const int* __restrict__ conn, const int* __restrict__ elem,
                              const int* __restrict__ facet, int* __restrict__ fnode)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncards) return;
    for (int a = 0; a < 4; a++) fnode[c * 4 + a] = conn[(int64_t)c_facet_nodes[facet[c]][a] * stride + elem[c]];
}

__global__ void k_traction_cards(int64_t ncards, const int* __restrict__ fnode, const double* __restrict__ X,
                                 const double* __restrict__ tract, int local, double scale, double* __restrict__ rec,
                                 unsigned long long* __restrict__ status)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncards) return;
    const double qr[4] = {-1.0, 1.0, 1.0, -1.0}, qs[4] = {-1.0, -1.0, 1.0, 1.0};
    const double g = 0.57735026918962576451; // 1/sqrt(3)
    double x[4][3], t[4][3], rhs[4][3];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int n = fnode[c * 4 + a];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            x[a][i] = X[(int64_t)n * 3 + i];
            t[a][i] = scale * tract[(c * 4 + a) * 3 + i]; // Traction_CardT::CurrentValue: schedule value times the nodal vectors
            rhs[a][i] = 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double r = g * qr[j], s = g * qs[j];
        double Na[4], m1[3] = {0.0, 0.0, 0.0}, m2[3] = {0.0, 0.0, 0.0}, tip[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const double tr = 1.0 + qr[a] * r, ts = 1.0 + qs[a] * s;
            Na[a] = 0.25 * tr * ts;
            const double dr = 0.25 * qr[a] * ts, ds = 0.25 * tr * qs[a];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                m1[i] += x[a][i] * dr;
                m2[i] += x[a][i] * ds;
                tip[i] += Na[a] * t[a][i];
            }
        }
        double n3[3] = {m1[1] * m2[2] - m1[2] * m2[1], m1[2] * m2[0] - m1[0] * m2[2], m1[0] * m2[1] - m1[1] * m2[0]};
        const double jn = sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
        double tj[3] = {tip[0], tip[1], tip[2]};
        if (local) {
            const double j1 = sqrt(m1[0] * m1[0] + m1[1] * m1[1] + m1[2] * m1[2]);
            if (!(jn > 0.0) || !(j1 > 0.0)) { // ExceptionT::kBadJacobianDet (ParentDomainT.cpp:408,413)
                if (atomicCAS(&status[0], 0ull, (unsigned long long)TB2_ERR_BAD_JACOBIAN) == 0ull) status[1] = (unsigned long long)c;
                return;
            }
            double n1[3], n2[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                n3[i] /= jn;
                n1[i] = m1[i] / j1;
            }
            n2[0] = n3[1] * n1[2] - n3[2] * n1[1];
            n2[1] = n3[2] * n1[0] - n3[0] * n1[2];
            n2[2] = n3[0] * n1[1] - n3[1] * n1[0];
#pragma unroll
            for (int i = 0; i < 3; i++) tj[i] = n1[i] * tip[0] + n2[i] * tip[1] + n3[i] * tip[2];
        }
#pragma unroll
        for (int l = 0; l < 3; l++) {
            const double fact = jn * tj[l];
#pragma unroll
            for (int a = 0; a < 4; a++) rhs[a][l] += fact * Na[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int i = 0; i < 3; i++) rec[(c * 4 + a) * 3 + i] = rhs[a][i];
}

__global__ void k_traction_nodes(int64_t nloaded, const int* __restrict__ node, const int* __restrict__ slot_ptr,
                                 const int* __restrict__ slot, const double* __restrict__ rec, int accumulate, double* __restrict__ f)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nloaded) return;
    const int64_t n = node[k];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (accumulate) {
        s0 = f[n * 3 + 0];
        s1 = f[n * 3 + 1];
        s2 = f[n * 3 + 2];
    }
    for (int q = slot_ptr[k]; q < slot_ptr[k + 1]; q++) {
        const double* r = rec + (int64_t)slot[q] * 3;
        s0 += r[0];
        s1 += r[1];
        s2 += r[2];
    }
    f[n * 3 + 0] = s0;
    f[n * 3 + 1] = s1;
    f[n * 3 + 2] = s2;
}

} // namespace

extern "C" {

int tb2_traction_create(tb2_mesh* m, int64_t ncards, const int32_t* h_elem, const int32_t* h_facet, const double* h_tract,
                        int coord_system, tb2_traction** out)
{
    TB2_ARG(m && out && ncards >= 0 && (ncards == 0 || (h_elem && h_facet && h_tract)));
    TB2_ARG(coord_system == TB2_TRACTION_GLOBAL || coord_system == TB2_TRACTION_LOCAL);
    TB2_ARG(ncards < (int64_t)1 << 29);
    for (int64_t c = 0; c < ncards; c++) {
        if (h_elem[c] < 0 || h_elem[c] >= m->ne || h_facet[c] < 0 || h_facet[c] > 5) {
            tb2::set_error("traction card %lld: element %d facet %d out of range", (long long)c, h_elem[c], h_facet[c]);
            return TB2_ERR_SIZE;
        }
    }
    DeviceGuard dg(m->device);
    tb2_traction* t = new tb2_traction;
    t->mesh = m;
    t->ncards = ncards;
    t->coord_system = coord_system;
    *out = t;
    TB2_CUDA(t->status.alloc(2));
    TB2_CUDA(cudaMemsetAsync(t->status.p, 0, 2 * sizeof(unsigned long long), m->stream));
    if (ncards == 0) return TB2_OK;
    TB2_CUDA(t->elem.alloc(ncards));
    TB2_CUDA(t->facet.alloc(ncards));
    TB2_CUDA(t->fnode.alloc(ncards * 4));
    TB2_CUDA(t->tract.alloc(ncards * 12));
    TB2_CUDA(t->rec.alloc(ncards * 12));
    TB2_CUDA(cudaMemcpyAsync(t->elem.p, h_elem, ncards * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(t->facet.p, h_facet, ncards * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(t->tract.p, h_tract, ncards * 12 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    const int threads = 128;
    const unsigned blocks = (unsigned)((ncards + threads - 1) / threads);
    k_facet_nodes<<<blocks, threads, 0, m->stream>>>(ncards, m->stride, m->conn.p, t->elem.p, t->facet.p, t->fnode.p);
    TB2_CUDA(cudaGetLastError());
    std::vector<int> fnode((size_t)ncards * 4);
    TB2_CUDA(cudaMemcpyAsync(fnode.data(), t->fnode.p, fnode.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    // node -> card records, ascending card order within a node (stable sort of the slots by node)
    std::vector<int> order(fnode.size());
    for (size_t q = 0; q < order.size(); q++) order[q] = (int)q;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fnode[a] < fnode[b]; });
    std::vector<int> node, ptr;
    for (size_t q = 0; q < order.size(); q++) {
        if (q == 0 || fnode[order[q]] != fnode[order[q - 1]]) {
            node.push_back(fnode[order[q]]);
            ptr.push_back((int)q);
        }
    }
    ptr.push_back((int)order.size());
    t->nloaded = (int64_t)node.size();
    TB2_CUDA(t->node.alloc(node.size()));
    TB2_CUDA(t->slot_ptr.alloc(ptr.size()));
    TB2_CUDA(t->slot.alloc(order.size()));
    TB2_CUDA(cudaMemcpy(t->node.p, node.data(), node.size() * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(t->slot_ptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(t->slot.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    return TB2_OK;
}

int tb2_traction_destroy(tb2_traction* t)
{
    if (!t) return TB2_OK;
    DeviceGuard dg(t->mesh->device);
    cudaStreamSynchronize(t->mesh->stream);
    delete t;
    return TB2_OK;
}

int tb2_traction_form(tb2_traction* t, double scale, int accumulate, double* d_f)
{
    TB2_ARG(t && d_f);
    tb2_mesh* m = t->mesh;
    DeviceGuard dg(m->device);
    if (!accumulate) TB2_CUDA(cudaMemsetAsync(d_f, 0, (size_t)m->nn * 3 * sizeof(double), m->stream));
    if (t->ncards == 0) return TB2_OK;
    const int threads = 128;
    k_traction_cards<<<(unsigned)((t->ncards + threads - 1) / threads), threads, 0, m->stream>>>(
        t->ncards, t->fnode.p, m->X.p, t->tract.p, t->coord_system == TB2_TRACTION_LOCAL, scale, t->rec.p, t->status.p);
    TB2_CUDA(cudaGetLastError());
    k_traction_nodes<<<(unsigned)((t->nloaded + threads - 1) / threads), threads, 0, m->stream>>>(
        t->nloaded, t->node.p, t->slot_ptr.p, t->slot.p, t->rec.p, 1, d_f);
    TB2_CUDA(cudaGetLastError());
    if (t->coord_system == TB2_TRACTION_LOCAL) { // the only failure mode: a degenerate facet
        unsigned long long st[2];
        TB2_CUDA(cudaMemcpyAsync(st, t->status.p, sizeof(st), cudaMemcpyDeviceToHost, m->stream));
        TB2_CUDA(cudaStreamSynchronize(m->stream));
        if (st[0]) {
            TB2_CUDA(cudaMemsetAsync(t->status.p, 0, sizeof(st), m->stream));
            tb2::set_error("degenerate facet Jacobian at traction card %llu", st[1]);
            return (int)st[0];
        }
    }
    return TB2_OK;
}

int tb2_traction_form_host(tb2_traction* t, double scale, int accumulate, double* h_f)
{
    TB2_ARG(t && h_f);
    tb2_mesh* m = t->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = (size_t)m->nn * 3 * sizeof(double);
    if (!m->stage_a.p) TB2_CUDA(m->stage_a.alloc((size_t)m->nn * 3));
    if (accumulate) TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_f, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(tb2_traction_form(t, scale, accumulate, m->stage_a.p));
    TB2_CUDA(cudaMemcpyAsync(h_f, m->stage_a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}

} // extern "C"
