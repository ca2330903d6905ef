// tb2_contact.cu -- contact_3D_penalty force on the device: PenaltyContact3DT::RHSDriver (PenaltyContact3DT.cpp:262-500)
// over the list of active striker-facet pairs the host search maintains (Contact3DT::SetActiveInteractions, Contact3DT.cpp:100-175:
// three facet nodes, then the striker).  The search stays host code (it runs at relaxation points, not per step); what runs every
// step -- and in the reference under an OpenMP loop with a critical section around the assembly -- is the force:
//
//   k_contact_pairs : one thread per pair.  Configuration X + constKd u, facet normal n = (x2 - x1) x (x3 - x1) normalised, gap
//                     h = n . (x_s - centroid).  For h < 0: penalty force dphi dh/du with dphi = -K h area (the variation of the normal as
//                     Contact3DT::Set_dn_du, Contact3DT.cpp:176-220), velocity-based regularised Coulomb friction
//                     f_t = -mu |f_n| v_t / sqrt(|v_t|^2 + eps^2) and normal viscous damping f = -c v_n area, each split -1/3 on the
//                     facet nodes and +1 on the striker.  The pair's 4 x 3 record goes to a scratch (zeros when the pair is open);
//                     the number of pairs in contact and the deepest penetration (ContactT::SetTrackingData) are reduced per CTA.
//   k_contact_nodes : one thread per node that appears in a pair: sums its records in ascending pair order -- the order of the
//                     reference's serial loop (its OpenMP variant serialises the same scatter) -- no float atomics, reruns bit-identical.
//
// O(surface) work; it is on the device so that a device-resident step can take contact loads without a host round trip per step.
#include "tb2_internal.h"

#include <algorithm>
#include <vector>

using namespace tb2;

struct tb2_contact {
    tb2_mesh* mesh = nullptr;
    int device = 0;               // copy for tb2_contact_destroy: a host may destroy the mesh (with its element group) first
    double K = 0.0, mu = 0.0, eps = 1.0e-6, visc = 0.0;
    int64_t npairs = 0, ntouched = 0;
    tb2::DevBuf<int> pairs;       // [npairs][4]
    tb2::DevBuf<double> area;     // [npairs] area of the pair's striker (ContactT::fStrikerArea)
    tb2::DevBuf<double> rec;      // [npairs][4][3] pair forces of the last evaluation
    tb2::DevBuf<int> node;        // [ntouched] nodes that appear in a pair, ascending
    tb2::DevBuf<int> slot_ptr;    // [ntouched+1]
    tb2::DevBuf<int> slot;        // [4*npairs] pair*4+a, ascending within a node
    tb2::DevBuf<int> track_n;     // [blocks] pairs in contact per CTA
    tb2::DevBuf<double> track_h;  // [blocks] deepest penetration per CTA
    int track_blocks = 0;
    unsigned long long version = 0; // tb2_contact_set_pairs calls so far
};

namespace {

constexpr int kContactThreads = 128;

__global__ void __launch_bounds__(kContactThreads) k_contact_pairs(int64_t npairs, const int* __restrict__ pairs, const double* __restrict__ area,
                                                                  double K, double mu, double eps, double visc, double constKd,
                                                                  const double* __restrict__ X, const double* __restrict__ u,
                                                                  const double* __restrict__ v, double* __restrict__ rec,
                                                                  int* __restrict__ track_n, double* __restrict__ track_h)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int in_contact = 0;
    double depth = 0.0;
    if (p < npairs) {
        int nd[4];
        double x[4][3], rhs[12];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            nd[a] = pairs[4 * p + a];
#pragma unroll
            for (int i = 0; i < 3; i++) x[a][i] = X[3 * (int64_t)nd[a] + i] + constKd * u[3 * (int64_t)nd[a] + i];
        }
#pragma unroll
        for (int j = 0; j < 12; j++) rhs[j] = 0.0;
        double ea[3], eb[3], n[3], c[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            ea[i] = x[1][i] - x[0][i];
            eb[i] = x[2][i] - x[0][i];
        }
        n[0] = ea[1] * eb[2] - ea[2] * eb[1];
        n[1] = ea[2] * eb[0] - ea[0] * eb[2];
        n[2] = ea[0] * eb[1] - ea[1] * eb[0];
        const double mag = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
        for (int i = 0; i < 3; i++) n[i] /= mag;
#pragma unroll
        for (int i = 0; i < 3; i++) c[i] = x[3][i] - (x[0][i] + x[1][i] + x[2][i]) / 3.0;
        const double h = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
        if (h < 0.0) {
            in_contact = 1;
            depth = h;
            const double dphi = -K * h * area[p];
            // dc/du^T n: -n/3 on the three facet nodes, n on the striker
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int i = 0; i < 3; i++) rhs[3 * a + i] = dphi * (-(1.0 / 3.0) * n[i]);
#pragma unroll
            for (int i = 0; i < 3; i++) rhs[9 + i] = dphi * n[i];
            // variation of the normal: column j of d(a x b)/du is e_i x (edge opposite to facet node a), i = j % 3
            // (Contact3DT::Set_dn_du); the term is -dphi/mag c . (n n^T - 1) dn_j
            const double w[3][3] = {{x[2][0] - x[1][0], x[2][1] - x[1][1], x[2][2] - x[1][2]},   // node 1: x3 - x2
                                    {x[0][0] - x[2][0], x[0][1] - x[2][1], x[0][2] - x[2][2]},   // node 2: x1 - x3
                                    {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]}};  // node 3: x2 - x1
#pragma unroll
            for (int a = 0; a < 3; a++) {
                // dn for the three dofs of facet node a: rows of the skew matrix of w[a], i.e. dn_i = w[a] x e_i
                const double dn[3][3] = {{0.0, w[a][2], -w[a][1]}, {-w[a][2], 0.0, w[a][0]}, {w[a][1], -w[a][0], 0.0}};
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double ndn = n[0] * dn[i][0] + n[1] * dn[i][1] + n[2] * dn[i][2];
                    const double v1 = (n[0] * ndn - dn[i][0]) * c[0] + (n[1] * ndn - dn[i][1]) * c[1] + (n[2] * ndn - dn[i][2]) * c[2];
                    rhs[3 * a + i] += -dphi / mag * v1;
                }
            }
            if (v) {
                double vs[3], vf[3];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    vs[i] = v[3 * (int64_t)nd[3] + i];
                    vf[i] = (v[3 * (int64_t)nd[0] + i] + v[3 * (int64_t)nd[1] + i] + v[3 * (int64_t)nd[2] + i]) / 3.0;
                }
                const double third = 1.0 / 3.0;
                if (mu > 0.0) {
                    double vr[3], vt[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) vr[i] = vs[i] - vf[i];
                    const double vrn = vr[0] * n[0] + vr[1] * n[1] + vr[2] * n[2];
#pragma unroll
                    for (int i = 0; i < 3; i++) vt[i] = vr[i] - vrn * n[i];
                    const double vt2 = vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2];
                    const double s = -mu * fabs(dphi) * (1.0 / sqrt(vt2 + eps * eps));
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        const double ft = s * vt[i];
                        rhs[i] += -ft * third;
                        rhs[3 + i] += -ft * third;
                        rhs[6 + i] += -ft * third;
                        rhs[9 + i] += ft;
                    }
                }
                if (visc > 0.0) {
                    const double vrn = (vs[0] - vf[0]) * n[0] + (vs[1] - vf[1]) * n[1] + (vs[2] - vf[2]) * n[2];
                    const double fv = -visc * vrn * area[p];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        rhs[i] += -fv * n[i] * third;
                        rhs[3 + i] += -fv * n[i] * third;
                        rhs[6 + i] += -fv * n[i] * third;
                        rhs[9 + i] += fv * n[i];
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 12; j++) rec[12 * p + j] = rhs[j];
    }
    // tracking data per CTA (exact: an integer count and a minimum)
    __shared__ int sn[kContactThreads / 32];
    __shared__ double sh[kContactThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cnt = in_contact;
    double dm = depth;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        dm = fmin(dm, __shfl_xor_sync(0xffffffffu, dm, o));
    }
    if (lane == 0) {
        sn[wid] = cnt;
        sh[wid] = dm;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tn = 0;
        double th = 0.0;
        for (int w = 0; w < kContactThreads / 32; w++) {
            tn += sn[w];
            th = fmin(th, sh[w]);
        }
        track_n[blockIdx.x] = tn;
        track_h[blockIdx.x] = th;
    }
}

template <bool ACCUMULATE>
__global__ void k_contact_nodes(int64_t ntouched, const int* __restrict__ node, const int* __restrict__ slot_ptr, const int* __restrict__ slot,
                                const double* __restrict__ rec, double* __restrict__ f)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= ntouched) return;
    const int64_t n = node[k];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (ACCUMULATE) {
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

int contact_launch(tb2_contact* c, double constKd, const double* d_u, const double* d_v, double* d_f, bool accumulate, cudaStream_t st)
{
    tb2_mesh* m = c->mesh;
    if ((c->mu > 0.0 || c->visc > 0.0) && !d_v) {
        tb2::set_error("contact force: friction / viscous damping need the nodal velocities");
        return TB2_ERR_ARG;
    }
    if (c->npairs == 0) return TB2_OK;
    ProfScope ps(m, kProfOther, 2, st);
    k_contact_pairs<<<(unsigned)c->track_blocks, kContactThreads, 0, st>>>(c->npairs, c->pairs.p, c->area.p, c->K, c->mu, c->eps, c->visc, constKd,
                                                                                 m->X.p, d_u, d_v, c->rec.p, c->track_n.p, c->track_h.p);
    TB2_CUDA(cudaGetLastError());
    const int T = 128;
    const unsigned nb = (unsigned)((c->ntouched + T - 1) / T);
    if (accumulate) k_contact_nodes<true><<<nb, T, 0, st>>>(c->ntouched, c->node.p, c->slot_ptr.p, c->slot.p, c->rec.p, d_f);
    else k_contact_nodes<false><<<nb, T, 0, st>>>(c->ntouched, c->node.p, c->slot_ptr.p, c->slot.p, c->rec.p, d_f);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

} // namespace

namespace tb2 {
int contact_form_touched(tb2_contact* c, double constKd, const double* d_u, const double* d_v, double* d_f, cudaStream_t st)
{
    return contact_launch(c, constKd, d_u, d_v, d_f, false, st);
}
unsigned long long contact_version(const tb2_contact* c) { return c->version; }
} // namespace tb2

extern "C" {

int tb2_contact_create(tb2_mesh* m, double penalty_stiffness, double friction_coefficient, double friction_epsilon, double viscous_damping,
                       tb2_contact** out)
{
    TB2_ARG(m && out && penalty_stiffness >= 0.0 && friction_coefficient >= 0.0 && friction_epsilon > 0.0 && viscous_damping >= 0.0);
    tb2_contact* c = new tb2_contact;
    c->mesh = m;
    c->device = m->device;
    c->K = penalty_stiffness;
    c->mu = friction_coefficient;
    c->eps = friction_epsilon;
    c->visc = viscous_damping;
    *out = c;
    return TB2_OK;
}

int tb2_contact_destroy(tb2_contact* c)
{
    if (!c) return TB2_OK;
    DeviceGuard dg(c->device);
    if (cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError(); // not the mesh stream: the mesh may be gone already
    delete c;
    return TB2_OK;
}

int tb2_contact_set_pairs(tb2_contact* c, int64_t npairs, const int32_t* h_pairs, const double* h_area)
{
    TB2_ARG(c && npairs >= 0 && (npairs == 0 || (h_pairs && h_area)));
    TB2_ARG(npairs < (int64_t)1 << 29);
    tb2_mesh* m = c->mesh;
    for (int64_t q = 0; q < 4 * npairs; q++)
        if (h_pairs[q] < 0 || h_pairs[q] >= m->nn) {
            tb2::set_error("contact pair %lld: node %d out of range", (long long)(q / 4), h_pairs[q]);
            return TB2_ERR_SIZE;
        }
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream)); // the previous list may still be in use
    c->npairs = npairs;
    c->ntouched = 0;
    c->version++;
    if (npairs == 0) return TB2_OK;
    // node -> pair records, ascending pair order within a node (stable sort of the slots by node)
    std::vector<int> order((size_t)npairs * 4);
    for (size_t q = 0; q < order.size(); q++) order[q] = (int)q;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_pairs[a] < h_pairs[b]; });
    std::vector<int> node, ptr;
    for (size_t q = 0; q < order.size(); q++)
        if (q == 0 || h_pairs[order[q]] != h_pairs[order[q - 1]]) {
            node.push_back(h_pairs[order[q]]);
            ptr.push_back((int)q);
        }
    ptr.push_back((int)order.size());
    c->ntouched = (int64_t)node.size();
    c->track_blocks = (int)((npairs + kContactThreads - 1) / kContactThreads);
    TB2_CUDA(c->pairs.alloc(npairs * 4));
    TB2_CUDA(c->area.alloc(npairs));
    TB2_CUDA(c->rec.alloc(npairs * 12));
    TB2_CUDA(c->node.alloc(node.size()));
    TB2_CUDA(c->slot_ptr.alloc(ptr.size()));
    TB2_CUDA(c->slot.alloc(order.size()));
    TB2_CUDA(c->track_n.alloc(c->track_blocks));
    TB2_CUDA(c->track_h.alloc(c->track_blocks));
    TB2_CUDA(cudaMemcpy(c->pairs.p, h_pairs, (size_t)npairs * 4 * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(c->area.p, h_area, (size_t)npairs * sizeof(double), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(c->node.p, node.data(), node.size() * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(c->slot_ptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(c->slot.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    return TB2_OK;
}

int tb2_contact_form(tb2_contact* c, double constKd, const double* d_u, const double* d_v, int accumulate, double* d_f)
{
    TB2_ARG(c && d_u && d_f);
    tb2_mesh* m = c->mesh;
    DeviceGuard dg(m->device);
    if (!accumulate) TB2_CUDA(cudaMemsetAsync(d_f, 0, (size_t)m->nn * 3 * sizeof(double), m->stream));
    return contact_launch(c, constKd, d_u, d_v, d_f, true, m->stream);
}

int tb2_contact_form_host(tb2_contact* c, double constKd, const double* h_u, const double* h_v, int accumulate, double* h_f)
{
    TB2_ARG(c && h_u && h_f);
    tb2_mesh* m = c->mesh;
    DeviceGuard dg(m->device);
    const size_t n3 = (size_t)m->nn * 3, bytes = n3 * sizeof(double);
    tb2::DevBuf<double> u, v;
    TB2_CUDA(u.alloc(n3));
    if (h_v) TB2_CUDA(v.alloc(n3));
    if (!m->stage_a.p) TB2_CUDA(m->stage_a.alloc(n3));
    TB2_CUDA(cudaMemcpyAsync(u.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_v) TB2_CUDA(cudaMemcpyAsync(v.p, h_v, bytes, cudaMemcpyHostToDevice, m->stream));
    if (accumulate) TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_f, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(tb2_contact_form(c, constKd, u.p, h_v ? v.p : nullptr, accumulate, m->stage_a.p));
    TB2_CUDA(cudaMemcpyAsync(h_f, m->stage_a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}

int tb2_contact_tracking(tb2_contact* c, int* num_contact, double* h_max)
{
    TB2_ARG(c);
    tb2_mesh* m = c->mesh;
    DeviceGuard dg(m->device);
    int tn = 0;
    double th = 0.0;
    if (c->npairs > 0 && c->track_blocks > 0) {
        std::vector<int> n((size_t)c->track_blocks);
        std::vector<double> h((size_t)c->track_blocks);
        TB2_CUDA(cudaMemcpyAsync(n.data(), c->track_n.p, n.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        TB2_CUDA(cudaMemcpyAsync(h.data(), c->track_h.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        TB2_CUDA(cudaStreamSynchronize(m->stream));
        for (size_t b = 0; b < n.size(); b++) {
            tn += n[b];
            th = h[b] < th ? h[b] : th;
        }
    }
    if (num_contact) *num_contact = tn;
    if (h_max) *h_max = th;
    return TB2_OK;
}

} // extern "C"
