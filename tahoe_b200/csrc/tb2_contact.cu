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

#include <cub/cub.cuh>

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
    // the search (tb2_contact_set_surfaces / tb2_contact_search): triangulated surfaces, strikers, per-node surface membership
    int64_t nfacets = 0, nstrikers = 0;
    tb2::DevBuf<int> facets, facet_surface, strikers, hit;
    tb2::DevBuf<unsigned> node_surfaces; // [nn] bit s: the node belongs to surface s (no self contact per surface)
    tb2::DevBuf<double> striker_area, gap, tile_box;
    tb2::DevBuf<int> flag, pos, keys, keys_sorted, vals, run_len, counters; // the pair list built on the device after a search
    tb2::DevBuf<unsigned char> cub_tmp;
    size_t cub_bytes = 0;
    std::vector<int> h_facets, h_strikers;
    std::vector<double> h_striker_area;
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
    if (n == 0x7fffffff) return; // the run of free strikers' keys at the end of a device-built list
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


// Contact3DT::SetActiveStrikers (Contact3DT.cpp:226-334) with Contact3DT::Intersect (:336-391): one thread per striker walks the facets in
// surface / facet order (tiles of them staged in shared memory by the CTA) and keeps the accepted facet of smallest |h|, the first one
// on a tie.  The reference collects its candidates through a search grid around the facet midpoint (radius 1.65 |mid - x1|); a facet
// is skipped here when the striker lies outside the sphere around that box, which never excludes a striker Intersect would accept.
constexpr int kSearchThreads = 128;
// the box that holds every point a facet of tile t (kSearchThreads consecutive facets) can accept: midpoints +- their reach
__global__ void __launch_bounds__(kSearchThreads) k_facet_tile_bounds(int64_t nfacets, const int* __restrict__ facets, const double* __restrict__ X,
                                                                     const double* __restrict__ u, double* __restrict__ tile_box /*[ntiles][6]*/)
{
    __shared__ double lo[3][kSearchThreads / 32], hi[3][kSearchThreads / 32];
    const int64_t f = blockIdx.x * (int64_t)kSearchThreads + threadIdx.x;
    double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
    if (f < nfacets) {
        double x1[3], m[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int64_t n = facets[3 * f + a];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double c = X[3 * n + i] + u[3 * n + i];
                if (a == 0) x1[i] = c;
                m[i] += c;
            }
        }
        double r2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            m[i] /= 3.0;
            r2 += (m[i] - x1[i]) * (m[i] - x1[i]);
        }
        const double reach = sqrt(3.0 * (1.1 * 1.5) * (1.1 * 1.5) * r2 * 1.0001) * 1.0001;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            bl[i] = m[i] - reach;
            bh[i] = m[i] + reach;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bl[i] = fmin(bl[i], __shfl_xor_sync(0xffffffffu, bl[i], o));
            bh[i] = fmax(bh[i], __shfl_xor_sync(0xffffffffu, bh[i], o));
        }
        if ((threadIdx.x & 31) == 0) {
            lo[i][threadIdx.x >> 5] = bl[i];
            hi[i][threadIdx.x >> 5] = bh[i];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = lo[threadIdx.x][0], b = hi[threadIdx.x][0];
        for (int w = 1; w < kSearchThreads / 32; w++) {
            a = fmin(a, lo[threadIdx.x][w]);
            b = fmax(b, hi[threadIdx.x][w]);
        }
        tile_box[6 * blockIdx.x + threadIdx.x] = a;
        tile_box[6 * blockIdx.x + 3 + threadIdx.x] = b;
    }
}

__global__ void __launch_bounds__(kSearchThreads) k_contact_search(int64_t nstrikers, const int* __restrict__ strikers, int64_t nfacets,
                                                                  const int* __restrict__ facets, const int* __restrict__ facet_surface,
                                                                  const unsigned* __restrict__ node_surfaces, const double* __restrict__ X,
                                                                  const double* __restrict__ u, const double* __restrict__ tile_box,
                                                                  int* __restrict__ hit, double* __restrict__ gap)
{
    __shared__ double fx[kSearchThreads][9];
    __shared__ double fmid[kSearchThreads][4]; // midpoint and squared reach
    __shared__ int fsurf[kSearchThreads];
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = s < nstrikers;
    const int tag = live ? strikers[s] : 0;
    double xs[3] = {0.0, 0.0, 0.0};
    unsigned mine = 0;
    if (live) {
#pragma unroll
        for (int i = 0; i < 3; i++) xs[i] = X[3 * (int64_t)tag + i] + u[3 * (int64_t)tag + i];
        mine = node_surfaces[tag];
    }
    int best_f = -1;
    double best_h = 0.0;
    for (int64_t f0 = 0; f0 < nfacets; f0 += kSearchThreads) {
        const int64_t f = f0 + threadIdx.x;
        {   // a tile none of this CTA's strikers can reach is not even staged (strikers and facets both come in spatial runs)
            const double* box = tile_box + 6 * (f0 / kSearchThreads);
            const int near_tile = live && xs[0] >= box[0] && xs[0] <= box[3] && xs[1] >= box[1] && xs[1] <= box[4] && xs[2] >= box[2] && xs[2] <= box[5];
            if (!__syncthreads_or(near_tile)) continue;
        }
        if (f < nfacets) {
            double m[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int64_t n = facets[3 * f + a];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double c = X[3 * n + i] + u[3 * n + i];
                    fx[threadIdx.x][3 * a + i] = c;
                    m[i] += c;
                }
            }
            double r2 = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                m[i] /= 3.0;
                fmid[threadIdx.x][i] = m[i];
                const double dd = m[i] - fx[threadIdx.x][i];
                r2 += dd * dd;
            }
            fmid[threadIdx.x][3] = 3.0 * (1.1 * 1.5) * (1.1 * 1.5) * r2 * 1.0001; // (sqrt(3) * radius)^2, a hair wider
            fsurf[threadIdx.x] = facet_surface[f];
        }
        __syncthreads();
        const int nt = (int)(nfacets - f0 < kSearchThreads ? nfacets - f0 : kSearchThreads);
        if (!live) continue;
        for (int q = 0; q < nt; q++) {
            if ((mine >> fsurf[q]) & 1u) continue; // no self contact (per surface)
            const double dx = xs[0] - fmid[q][0], dy = xs[1] - fmid[q][1], dz = xs[2] - fmid[q][2];
            if (dx * dx + dy * dy + dz * dz > fmid[q][3]) continue;
            const double* x1 = fx[q];
            const double* x2 = fx[q] + 3;
            const double* x3 = fx[q] + 6;
            double a[3], b[3], c[3], n[3], xsp[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                a[i] = x2[i] - x1[i];
                b[i] = x3[i] - x1[i];
                c[i] = xs[i] - x1[i];
            }
            n[0] = a[1] * b[2] - a[2] * b[1];
            n[1] = a[2] * b[0] - a[0] * b[2];
            n[2] = a[0] * b[1] - a[1] * b[0];
            const double mag = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) n[i] /= mag;
            const double h = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
            if (fabs(h) > sqrt(mag) / 2.0) continue;
#pragma unroll
            for (int i = 0; i < 3; i++) xsp[i] = xs[i] + (-h) * n[i];
            const double area_tol = mag / 50.0;
            bool inside = true;
#pragma unroll
            for (int e = 0; e < 3; e++) {
                const double* p0 = fx[q] + 3 * e;
                const double* p1 = fx[q] + 3 * ((e + 1) % 3);
                double ed[3], xi[3];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    ed[i] = p1[i] - p0[i];
                    xi[i] = xsp[i] - p0[i];
                }
                const double n0 = ed[1] * xi[2] - ed[2] * xi[1], n1 = ed[2] * xi[0] - ed[0] * xi[2], n2 = ed[0] * xi[1] - ed[1] * xi[0];
                if (n[0] * n0 + n[1] * n1 + n[2] * n2 < -area_tol) inside = false;
            }
            if (!inside) continue;
            if (best_f < 0 || fabs(h) < fabs(best_h)) {
                best_f = (int)(f0 + q);
                best_h = h;
            }
        }
    }
    if (live) {
        hit[s] = best_f;
        gap[s] = best_h;
    }
}


// ---- the pair list of a search, built on the device: active strikers compacted in striker order, (node, record slot) keys sorted by
// node (stable radix sort: ascending pair order within a node), run lengths -> the per-node record lists of k_contact_nodes
__global__ void k_hit_flags(int64_t ns, const int* __restrict__ hit, int* __restrict__ flag)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s < ns) flag[s] = hit[s] >= 0 ? 1 : 0;
}
__global__ void k_build_pairs(int64_t ns, const int* __restrict__ hit, const int* __restrict__ flag, const int* __restrict__ pos,
                              const int* __restrict__ facets, const int* __restrict__ strikers, const double* __restrict__ striker_area,
                              int* __restrict__ pairs, double* __restrict__ area, int* __restrict__ keys, int* __restrict__ vals,
                              int* __restrict__ counters)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= ns) return;
    if (s == ns - 1) counters[0] = pos[s] + flag[s]; // number of pairs
    if (!flag[s]) {
#pragma unroll
        for (int a = 0; a < 4; a++) { // keys of free strikers sort to the end
            keys[4 * s + a] = 0x7fffffff;
            vals[4 * s + a] = 0;
        }
        return;
    }
    const int k = pos[s], f = hit[s];
    const int nd[4] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2], strikers[s]};
    area[k] = striker_area[s];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        pairs[4 * k + a] = nd[a];
        keys[4 * s + a] = nd[a];
        vals[4 * s + a] = 4 * k + a;
    }
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
    // the list changes every time a resident run searches: the buffers keep their capacity and the five small arrays go up on the stream
    TB2_CUDA(c->pairs.reserve(npairs * 4));
    TB2_CUDA(c->area.reserve(npairs));
    TB2_CUDA(c->rec.reserve(npairs * 12));
    TB2_CUDA(c->node.reserve(node.size()));
    TB2_CUDA(c->slot_ptr.reserve(ptr.size()));
    TB2_CUDA(c->slot.reserve(order.size()));
    TB2_CUDA(c->track_n.reserve(c->track_blocks));
    TB2_CUDA(c->track_h.reserve(c->track_blocks));
    TB2_CUDA(cudaMemcpyAsync(c->pairs.p, h_pairs, (size_t)npairs * 4 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(c->area.p, h_area, (size_t)npairs * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(c->node.p, node.data(), node.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(c->slot_ptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(c->slot.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream)); // the host vectors go out of scope
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

int tb2_contact_set_surfaces(tb2_contact* c, int64_t nfacets, const int32_t* h_facets, const int32_t* h_facet_surface, int64_t nstrikers,
                             const int32_t* h_strikers, const double* h_striker_area)
{
    TB2_ARG(c && nfacets >= 0 && nstrikers >= 0 && (nfacets == 0 || (h_facets && h_facet_surface)) && (nstrikers == 0 || (h_strikers && h_striker_area)));
    tb2_mesh* m = c->mesh;
    std::vector<unsigned> mask((size_t)m->nn, 0u);
    for (int64_t f = 0; f < nfacets; f++) {
        if (h_facet_surface[f] < 0 || h_facet_surface[f] > 31) {
            tb2::set_error("tb2_contact_set_surfaces: surface %d (at most 32 surfaces per group)", h_facet_surface[f]);
            return TB2_ERR_ARG;
        }
        for (int a = 0; a < 3; a++) {
            const int32_t n = h_facets[3 * f + a];
            if (n < 0 || n >= m->nn) {
                tb2::set_error("tb2_contact_set_surfaces: facet %lld has node %d out of range", (long long)f, n);
                return TB2_ERR_SIZE;
            }
            mask[n] |= 1u << h_facet_surface[f];
        }
    }
    for (int64_t s = 0; s < nstrikers; s++)
        if (h_strikers[s] < 0 || h_strikers[s] >= m->nn) {
            tb2::set_error("tb2_contact_set_surfaces: striker node %d out of range", h_strikers[s]);
            return TB2_ERR_SIZE;
        }
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    c->nfacets = nfacets;
    c->nstrikers = nstrikers;
    c->h_facets.assign(h_facets, h_facets + 3 * nfacets);
    c->h_strikers.assign(h_strikers, h_strikers + nstrikers);
    c->h_striker_area.assign(h_striker_area, h_striker_area + nstrikers);
    TB2_CUDA(c->facets.alloc(3 * (nfacets > 0 ? nfacets : 1)));
    TB2_CUDA(c->facet_surface.alloc(nfacets > 0 ? nfacets : 1));
    TB2_CUDA(c->strikers.alloc(nstrikers > 0 ? nstrikers : 1));
    TB2_CUDA(c->striker_area.alloc(nstrikers > 0 ? nstrikers : 1));
    TB2_CUDA(c->hit.alloc(nstrikers > 0 ? nstrikers : 1));
    TB2_CUDA(c->gap.alloc(nstrikers > 0 ? nstrikers : 1));
    TB2_CUDA(c->node_surfaces.alloc(m->nn));
    TB2_CUDA(c->tile_box.alloc(6 * ((nfacets + kSearchThreads - 1) / kSearchThreads + 1)));
    if (nfacets) TB2_CUDA(cudaMemcpy(c->facets.p, h_facets, (size_t)nfacets * 3 * sizeof(int), cudaMemcpyHostToDevice));
    if (nfacets) TB2_CUDA(cudaMemcpy(c->facet_surface.p, h_facet_surface, (size_t)nfacets * sizeof(int), cudaMemcpyHostToDevice));
    if (nstrikers) TB2_CUDA(cudaMemcpy(c->strikers.p, h_strikers, (size_t)nstrikers * sizeof(int), cudaMemcpyHostToDevice));
    if (nstrikers) TB2_CUDA(cudaMemcpy(c->striker_area.p, h_striker_area, (size_t)nstrikers * sizeof(double), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(c->node_surfaces.p, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    return TB2_OK;
}

// the search on X + u (device array), then the active pairs in striker order become the group's pair list.  The hits of the
// strikers -- O(surface) integers -- pass through the host, which builds the per-node record lists as tb2_contact_set_pairs does.
int tb2_contact_search(tb2_contact* c, const double* d_u, int64_t* npairs_out)
{
    TB2_ARG(c && d_u);
    tb2_mesh* m = c->mesh;
    DeviceGuard dg(m->device);
    if (c->nstrikers == 0 || c->nfacets == 0) {
        if (npairs_out) *npairs_out = 0;
        return tb2_contact_set_pairs(c, 0, nullptr, nullptr);
    }
    cudaStream_t st = m->stream;
    const int64_t ns = c->nstrikers, n4 = 4 * ns;
    const int T = 128;
    const unsigned nbs = (unsigned)((ns + T - 1) / T);
    // buffers of the device-built list: sized by the striker count (every striker can be in one pair), allocated once
    TB2_CUDA(c->flag.reserve(ns));
    TB2_CUDA(c->pos.reserve(ns));
    TB2_CUDA(c->keys.reserve(n4));
    TB2_CUDA(c->keys_sorted.reserve(n4));
    TB2_CUDA(c->vals.reserve(n4));
    TB2_CUDA(c->run_len.reserve(n4 + 1));
    TB2_CUDA(c->counters.reserve(2));
    TB2_CUDA(c->pairs.reserve(n4));
    TB2_CUDA(c->area.reserve(ns));
    TB2_CUDA(c->rec.reserve(12 * ns));
    TB2_CUDA(c->node.reserve(n4));
    TB2_CUDA(c->slot_ptr.reserve(n4 + 1));
    TB2_CUDA(c->slot.reserve(n4));
    TB2_CUDA(c->track_n.reserve((ns + kContactThreads - 1) / kContactThreads));
    TB2_CUDA(c->track_h.reserve((ns + kContactThreads - 1) / kContactThreads));
    {
        size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, c->flag.p, c->pos.p, (int)ns, st);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, c->keys.p, c->keys_sorted.p, c->vals.p, c->slot.p, (int)n4, 0, 32, st);
        cub::DeviceRunLengthEncode::Encode(nullptr, b3, c->keys_sorted.p, c->node.p, c->run_len.p, c->counters.p + 1, (int)n4, st);
        cub::DeviceScan::ExclusiveSum(nullptr, b4, c->run_len.p, c->slot_ptr.p, (int)(n4 + 1), st);
        const size_t need = std::max(std::max(b1, b2), std::max(b3, b4));
        if (need > c->cub_bytes) {
            TB2_CUDA(c->cub_tmp.alloc(need));
            c->cub_bytes = need;
        }
    }
    ProfScope ps(m, kProfOther, 8);
    k_facet_tile_bounds<<<(unsigned)((c->nfacets + kSearchThreads - 1) / kSearchThreads), kSearchThreads, 0, st>>>(c->nfacets, c->facets.p, m->X.p, d_u,
                                                                                                                c->tile_box.p);
    k_contact_search<<<(unsigned)((ns + kSearchThreads - 1) / kSearchThreads), kSearchThreads, 0, st>>>(
        ns, c->strikers.p, c->nfacets, c->facets.p, c->facet_surface.p, c->node_surfaces.p, m->X.p, d_u, c->tile_box.p, c->hit.p, c->gap.p);
    k_hit_flags<<<nbs, T, 0, st>>>(ns, c->hit.p, c->flag.p);
    size_t bytes = c->cub_bytes;
    TB2_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, c->flag.p, c->pos.p, (int)ns, st));
    k_build_pairs<<<nbs, T, 0, st>>>(ns, c->hit.p, c->flag.p, c->pos.p, c->facets.p, c->strikers.p, c->striker_area.p, c->pairs.p, c->area.p, c->keys.p,
                                    c->vals.p, c->counters.p);
    bytes = c->cub_bytes;
    TB2_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, bytes, c->keys.p, c->keys_sorted.p, c->vals.p, c->slot.p, (int)n4, 0, 32, st));
    TB2_CUDA(cudaMemsetAsync(c->run_len.p, 0, (size_t)(n4 + 1) * sizeof(int), st)); // run lengths beyond the last run stay 0 for the scan
    bytes = c->cub_bytes;
    TB2_CUDA(cub::DeviceRunLengthEncode::Encode(c->cub_tmp.p, bytes, c->keys_sorted.p, c->node.p, c->run_len.p, c->counters.p + 1, (int)n4, st));
    bytes = c->cub_bytes;
    TB2_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, c->run_len.p, c->slot_ptr.p, (int)(n4 + 1), st));
    TB2_CUDA(cudaGetLastError());
    int h[2] = {0, 0};
    TB2_CUDA(cudaMemcpyAsync(h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost, st));
    TB2_CUDA(cudaStreamSynchronize(st)); // the one host round trip of a search: two integers
    c->npairs = h[0];
    c->ntouched = h[1]; // includes the run of free strikers' keys (skipped by the node kernel)
    c->track_blocks = (int)((c->npairs + kContactThreads - 1) / kContactThreads);
    c->version++;
    if (npairs_out) *npairs_out = c->npairs;
    return TB2_OK;
}

int tb2_contact_get_pairs(tb2_contact* c, int64_t* npairs, int32_t* h_pairs, double* h_area)
{
    TB2_ARG(c);
    DeviceGuard dg(c->device);
    if (npairs) *npairs = c->npairs;
    if (c->npairs == 0) return TB2_OK;
    if (h_pairs) TB2_CUDA(cudaMemcpy(h_pairs, c->pairs.p, (size_t)c->npairs * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_area) TB2_CUDA(cudaMemcpy(h_area, c->area.p, (size_t)c->npairs * sizeof(double), cudaMemcpyDeviceToHost));
    return TB2_OK;
}

int tb2_contact_has_surfaces(const tb2_contact* c) { return c && c->nfacets > 0 && c->nstrikers > 0 ? 1 : 0; }

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
