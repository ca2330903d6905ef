// tb2_blockplan.h -- host-side construction of the element-block plan of the fused explicit step (tb2_fused_step.cuh).
//
// The reference assembles element forces one element at a time into the global residual (SolverT::AssembleRHS,
// SolverT.cpp:446-477) and then updates every node (nExplicitCD::Corrector, nExplicitCD.cpp:98-139).  On the device the
// same two operations are carried out block by block: the elements are cut into compact blocks of <= kBlockElems elements,
// a block's element forces stay in shared memory, are summed per block-local node in a fixed order, and
//   * "interior" nodes (all their elements lie in this block) are finished on the spot (corrector + next predictor);
//   * "surface" nodes (shared with other blocks) get one partial force per block; the block that delivers the LAST partial of a
//     node (a counter per node) adds all of them in ascending block order and finishes the node.
// The summation order -- ascending lane inside a block, ascending block id across blocks -- is fixed by this plan and does not
// depend on which block arrives last, so reruns are bit-reproducible and no float atomics are needed.
//
// Blocks come from a k-d bisection of the element centroids with exact counts (every leaf but possibly the last holds exactly
// kBlockElems elements, so no FP64 lanes idle), always cutting the longest axis of the current box.  A leaf whose node count
// exceeds kBlockMaxNodes (very elongated leaves of degenerate meshes) is cut again.
//
// Plain C++17, no CUDA: the plan is built once per mesh on the host (threads over sub-trees) and uploaded.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

namespace tb2 {

constexpr int kBlockElems = 128;    // elements per block = threads per compute warpgroup
constexpr int kBlockMaxNodes = 320; // block-local nodes (a 4x4x8 brick has 225)

// One block's record in device memory (32-bit words, 16-byte aligned records of kBlockRecWords words):
//   [0] n_local nodes  [1] n_interior (nodes [0, n_interior) are finished by the block itself)  [2] n_elems  [3] -
//   nodes[kBlockMaxNodes]             global node ids, interior nodes first, each class ascending
//   sbase[kBlockMaxNodes]             surface nodes: first partial slot of the NODE (its slots are contiguous, one per sharing
//                                     block in ascending block order); also the index of the node's arrival counter
//   srank[kBlockMaxNodes] uint16      surface nodes: rank of this block among the sharers | number of arrivals that complete the
//                                     node << 8 (= sharers; sharers + 1 for nodes that must not be finished in the kernel)
//   inc_off[kBlockMaxNodes+2] uint16  CSR offsets of the block-local incidence: the contributions (lane, local node a) to node l
//                                     occupy positions [inc_off[l], inc_off[l+1]) of the force scratch, ascending lane
// epos (separate array, read by the compute lanes): position of lane t's contribution a in that order.
constexpr int kRecHeader = 4;
constexpr int kRecNodes = kRecHeader;
constexpr int kRecSBase = kRecNodes + kBlockMaxNodes;
constexpr int kRecSRank = kRecSBase + kBlockMaxNodes;
constexpr int kRecIncOff = kRecSRank + kBlockMaxNodes / 2;
constexpr int kBlockRecWords = ((kRecIncOff + (kBlockMaxNodes + 2) / 2) + 3) / 4 * 4;

struct BlockPlan {
    int64_t nblocks = 0;
    std::vector<uint32_t> rec;      // [nblocks][kBlockRecWords]
    std::vector<int32_t> bconn;     // [nblocks][8][kBlockElems] global node ids of lane t's element (-1: padding lane)
    std::vector<int32_t> elem;      // [nblocks][kBlockElems] global element id of lane t (-1: padding lane)
    std::vector<uint16_t> epos;     // [nblocks][kBlockElems][8] scratch position of lane t's contribution to its local node a
    int64_t npartial = 0;           // partial-force slots: fpart[npartial][3], arrival counters cnt[npartial]
    std::vector<int32_t> surf_nodes; // global ids of the surface nodes, ascending
    std::vector<int32_t> surf_ptr;   // [nsurf+1]: slots [surf_ptr[s], surf_ptr[s+1]) of surface node s, ascending block id
    std::vector<int32_t> block_of_elem; // [ne] block id (diagnostics / tests)
    int64_t n_interior = 0, max_local = 0, max_share = 0;
};

namespace blockplan_detail {

struct Builder {
    int64_t ne, nn, stride;
    const int32_t* conn; // [8][stride]
    std::vector<float> cx, cy, cz;
    std::vector<int32_t> idx; // element permutation, leaves are contiguous ranges
    std::vector<std::pair<int64_t, int64_t>> leaves;

    const float* axis(int a) const { return a == 0 ? cx.data() : (a == 1 ? cy.data() : cz.data()); }

    int count_nodes(int64_t lo, int64_t hi, std::vector<int32_t>& scratch) const
    {
        scratch.clear();
        for (int64_t q = lo; q < hi; q++)
            for (int a = 0; a < 8; a++) scratch.push_back(conn[(int64_t)a * stride + idx[q]]);
        std::sort(scratch.begin(), scratch.end());
        return (int)(std::unique(scratch.begin(), scratch.end()) - scratch.begin());
    }

    void split(int64_t lo, int64_t hi, std::vector<std::pair<int64_t, int64_t>>& out, std::vector<int32_t>& scratch, int depth_par)
    {
        const int64_t n = hi - lo;
        if (n <= kBlockElems) {
            if (n > 1 && count_nodes(lo, hi, scratch) > kBlockMaxNodes) {
                bisect(lo, hi, lo + n / 2);
                split(lo, lo + n / 2, out, scratch, 0);
                split(lo + n / 2, hi, out, scratch, 0);
                return;
            }
            if (n > 0) out.emplace_back(lo, hi);
            return;
        }
        const int64_t nb = (n + kBlockElems - 1) / kBlockElems;
        const int64_t mid = lo + (nb / 2) * kBlockElems;
        bisect(lo, hi, mid);
        if (depth_par > 0) {
            std::vector<std::pair<int64_t, int64_t>> right;
            std::thread t([&] {
                std::vector<int32_t> s2;
                split(mid, hi, right, s2, depth_par - 1);
            });
            split(lo, mid, out, scratch, depth_par - 1);
            t.join();
            out.insert(out.end(), right.begin(), right.end());
        } else {
            split(lo, mid, out, scratch, 0);
            split(mid, hi, out, scratch, 0);
        }
    }

    // partition idx[lo,hi) so that [lo,mid) holds the elements with the smallest centroid coordinate along the longest axis
    void bisect(int64_t lo, int64_t hi, int64_t mid)
    {
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        for (int64_t q = lo; q < hi; q++) {
            const int32_t e = idx[q];
            const float c[3] = {cx[e], cy[e], cz[e]};
            for (int a = 0; a < 3; a++) {
                mn[a] = std::min(mn[a], c[a]);
                mx[a] = std::max(mx[a], c[a]);
            }
        }
        int ax = 0;
        for (int a = 1; a < 3; a++)
            if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
        const float* c = axis(ax);
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                         [c](int32_t p, int32_t q) { return c[p] < c[q] || (c[p] == c[q] && p < q); });
    }
};

} // namespace blockplan_detail

// conn: SoA [8][stride] element connectivity (0-based), X: [nn][3].  hold (optional, [nn]): nodes that must not be finished inside
// the kernel even when every local sharer has delivered (multi-GPU: nodes shared with other ranks, whose force is still partial).
// first (optional, [ne]): elements flagged 1 are placed in the leading blocks (multi-GPU: the elements touching held nodes, so that
// one short launch over blocks [0, nfirst_blocks) produces everything the interface exchange needs); *nfirst_blocks returns their count.
inline void build_block_plan(int64_t ne, int64_t nn, int64_t stride, const int32_t* conn, const double* X, const unsigned char* hold,
                             const unsigned char* first, BlockPlan& plan, int64_t* nfirst_blocks = nullptr)
{
    using namespace blockplan_detail;
    Builder b;
    b.ne = ne;
    b.nn = nn;
    b.stride = stride;
    b.conn = conn;
    b.cx.resize(ne);
    b.cy.resize(ne);
    b.cz.resize(ne);
    for (int64_t e = 0; e < ne; e++) {
        double c[3] = {0, 0, 0};
        for (int a = 0; a < 8; a++) {
            const double* x = X + 3 * (int64_t)conn[(int64_t)a * stride + e];
            c[0] += x[0];
            c[1] += x[1];
            c[2] += x[2];
        }
        b.cx[e] = (float)(0.125 * c[0]);
        b.cy[e] = (float)(0.125 * c[1]);
        b.cz[e] = (float)(0.125 * c[2]);
    }
    b.idx.resize(ne);
    std::iota(b.idx.begin(), b.idx.end(), 0);
    int64_t n_first = 0;
    if (first) n_first = std::stable_partition(b.idx.begin(), b.idx.end(), [first](int32_t e) { return first[e] != 0; }) - b.idx.begin();
    std::vector<int32_t> scratch;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int depth_par = 0;
    while ((1u << depth_par) < hw && depth_par < 5) depth_par++;
    if (ne < 200000) depth_par = 0;
    std::vector<std::pair<int64_t, int64_t>> leaves_first;
    b.split(0, n_first, leaves_first, scratch, n_first >= 200000 ? depth_par : 0);
    b.split(n_first, ne, b.leaves, scratch, depth_par);
    // elements inside a leaf in ascending global id (the lane order = the summation order inside a block);
    // blocks in ascending order of their first element: neighbouring blocks stay close in launch order
    auto finish = [&](std::vector<std::pair<int64_t, int64_t>>& lv) {
        for (auto& lf : lv) std::sort(b.idx.begin() + lf.first, b.idx.begin() + lf.second);
        std::sort(lv.begin(), lv.end(),
                  [&](const std::pair<int64_t, int64_t>& p, const std::pair<int64_t, int64_t>& q) { return b.idx[p.first] < b.idx[q.first]; });
    };
    finish(leaves_first);
    finish(b.leaves);
    if (nfirst_blocks) *nfirst_blocks = (int64_t)leaves_first.size();
    b.leaves.insert(b.leaves.begin(), leaves_first.begin(), leaves_first.end());
    const int64_t nb = (int64_t)b.leaves.size();

    std::vector<int32_t> node_deg(nn, 0);
    for (int a = 0; a < 8; a++)
        for (int64_t e = 0; e < ne; e++) node_deg[conn[(int64_t)a * stride + e]]++;

    plan.nblocks = nb;
    plan.rec.assign((size_t)nb * kBlockRecWords, 0u);
    plan.bconn.assign((size_t)nb * 8 * kBlockElems, -1);
    plan.elem.assign((size_t)nb * kBlockElems, -1);
    plan.epos.assign((size_t)nb * kBlockElems * 8, 0);
    plan.block_of_elem.assign(ne, -1);
    plan.n_interior = 0;
    plan.max_local = 0;
    plan.max_share = 0;
    const unsigned nthreads = nb > 64 ? hw : 1;
    // pass 1: per block node lists, block-local incidence (needs only local information)
    auto pass1 = [&](int64_t b0, int64_t b1) {
        std::vector<int32_t> nodes, uniq, cnt, order, local_of;
        std::vector<int> fill, pos;
        std::vector<uint16_t> loc(kBlockElems * 8);
        for (int64_t bi = b0; bi < b1; bi++) {
            const int64_t lo = b.leaves[bi].first, hi = b.leaves[bi].second;
            const int nel = (int)(hi - lo);
            uint32_t* r = plan.rec.data() + (size_t)bi * kBlockRecWords;
            nodes.clear();
            for (int64_t q = lo; q < hi; q++)
                for (int a = 0; a < 8; a++) nodes.push_back(conn[(int64_t)a * stride + b.idx[q]]);
            std::sort(nodes.begin(), nodes.end());
            uniq.clear();
            cnt.clear();
            for (size_t k = 0; k < nodes.size();) {
                size_t k2 = k;
                while (k2 < nodes.size() && nodes[k2] == nodes[k]) k2++;
                uniq.push_back(nodes[k]);
                cnt.push_back((int32_t)(k2 - k));
                k = k2;
            }
            const int nl = (int)uniq.size();
            auto is_interior = [&](int k) { return cnt[k] == node_deg[uniq[k]] && !(hold && hold[uniq[k]]); };
            order.clear();
            int nint = 0;
            for (int k = 0; k < nl; k++)
                if (is_interior(k)) {
                    order.push_back(k);
                    nint++;
                }
            for (int k = 0; k < nl; k++)
                if (!is_interior(k)) order.push_back(k);
            local_of.assign(nl, 0);
            for (int k = 0; k < nl; k++) local_of[order[k]] = k;
            r[0] = (uint32_t)nl;
            r[1] = (uint32_t)nint;
            r[2] = (uint32_t)nel;
            for (int k = 0; k < nl; k++) r[kRecNodes + k] = (uint32_t)uniq[order[k]];
            uint16_t* ioff = reinterpret_cast<uint16_t*>(r + kRecIncOff);
            uint16_t* epos = plan.epos.data() + (size_t)bi * kBlockElems * 8;
            fill.assign(nl + 1, 0);
            int32_t* bc = plan.bconn.data() + (size_t)bi * 8 * kBlockElems;
            for (int t = 0; t < nel; t++) {
                const int32_t e = b.idx[lo + t];
                plan.elem[(size_t)bi * kBlockElems + t] = e;
                plan.block_of_elem[e] = (int32_t)bi;
                for (int a = 0; a < 8; a++) {
                    const int32_t n = conn[(int64_t)a * stride + e];
                    bc[a * kBlockElems + t] = n;
                    const int k = (int)(std::lower_bound(uniq.begin(), uniq.end(), n) - uniq.begin());
                    loc[t * 8 + a] = (uint16_t)local_of[k];
                    fill[local_of[k] + 1]++;
                }
            }
            for (int k = 0; k < nl; k++) fill[k + 1] += fill[k];
            for (int k = 0; k <= kBlockMaxNodes + 1; k++) ioff[k] = (uint16_t)fill[std::min(k, nl)];
            pos.assign(fill.begin(), fill.end() - 1);
            for (int t = 0; t < nel; t++)
                for (int a = 0; a < 8; a++) epos[t * 8 + a] = (uint16_t)pos[loc[t * 8 + a]]++;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; t++) th.emplace_back(pass1, nb * t / nthreads, nb * (t + 1) / nthreads);
        for (auto& t : th) t.join();
    }
    // surface nodes: sharers per node, contiguous slots in ascending block order
    std::vector<int32_t> scount(nn, 0);
    for (int64_t bi = 0; bi < nb; bi++) {
        const uint32_t* r = plan.rec.data() + (size_t)bi * kBlockRecWords;
        plan.n_interior += r[1];
        plan.max_local = std::max<int64_t>(plan.max_local, r[0]);
        for (uint32_t k = r[1]; k < r[0]; k++) scount[r[kRecNodes + k]]++;
    }
    plan.surf_nodes.clear();
    plan.surf_ptr.assign(1, 0);
    std::vector<int32_t> base(nn, -1);
    for (int64_t n = 0; n < nn; n++)
        if (scount[n]) {
            base[n] = plan.surf_ptr.back();
            plan.surf_nodes.push_back((int32_t)n);
            plan.surf_ptr.push_back(plan.surf_ptr.back() + scount[n]);
            plan.max_share = std::max<int64_t>(plan.max_share, scount[n]);
        }
    plan.npartial = plan.surf_ptr.back();
    std::vector<int32_t> rank(nn, 0);
    for (int64_t bi = 0; bi < nb; bi++) {
        uint32_t* r = plan.rec.data() + (size_t)bi * kBlockRecWords;
        uint16_t* srank = reinterpret_cast<uint16_t*>(r + kRecSRank);
        for (uint32_t k = r[1]; k < r[0]; k++) {
            const int32_t n = (int32_t)r[kRecNodes + k];
            r[kRecSBase + k] = (uint32_t)base[n];
            int need = scount[n] + ((hold && hold[n]) ? 1 : 0);
            if (need > 255) need = 255; // > 254 sharers: never completes in the kernel; the caller's fallback pass finishes it
            srank[k] = (uint16_t)(std::min(rank[n], 255) | need << 8);
            rank[n]++;
        }
    }
}

} // namespace tb2
