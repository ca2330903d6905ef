// tb2_partition.cu -- element partition of a Hex8 mesh for the one-process-per-GPU path.  Host code, no CUDA calls.
//
// The reference decomposes a model before a parallel run (DecomposeT.cpp:362-592) with a graph partitioner (GraphBaseT::Partition,
// METIS when it is compiled in) and writes per-rank geometry with external / border node lists (PartitionT).  Here the partition is
// by ELEMENTS -- every rank owns a set of elements, nodes on the cuts are replicated -- and what a rank needs is: its elements and
// nodes in local numbering, its interface nodes with their slots in the global interface vector, and which of its nodes it owns for
// dot products.  tb2_partition_rcb assigns elements by recursive coordinate bisection of their centroids (longest extent, split at
// the weighted median, so any rank count works); tb2_partition_part builds one rank's description from any element->rank map, i.e.
// also from a graph partitioner's.  tahoe_b200/mesh.py holds the same two functions for the Python harness; the tests check that
// both give the same arrays.
#include "tb2_internal.h"

#include <algorithm>
#include <numeric>
#include <vector>

using namespace tb2;

namespace {

void rcb_split(std::vector<int64_t>& ids, int64_t lo, int64_t hi, int r0, int nr, const double* cent, int32_t* owner)
{
    if (nr == 1) {
        for (int64_t k = lo; k < hi; k++) owner[ids[k]] = r0;
        return;
    }
    const int nl = nr / 2;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = lo; k < hi; k++)
        for (int i = 0; i < 3; i++) {
            const double c = cent[3 * ids[k] + i];
            mn[i] = c < mn[i] ? c : mn[i];
            mx[i] = c > mx[i] ? c : mx[i];
        }
    int axis = 0;
    for (int i = 1; i < 3; i++)
        if (mx[i] - mn[i] > mx[axis] - mn[axis]) axis = i; // the first of equal extents, as numpy's argmax
    std::sort(ids.begin() + lo, ids.begin() + hi, [&](int64_t a, int64_t b) {
        const double ca = cent[3 * a + axis], cb = cent[3 * b + axis];
        return ca < cb || (ca == cb && a < b); // ties by element id: deterministic
    });
    const int64_t cut = lo + ((hi - lo) * nl) / nr;
    rcb_split(ids, lo, cut, r0, nl, cent, owner);
    rcb_split(ids, cut, hi, r0 + nl, nr - nl, cent, owner);
}

} // namespace

extern "C" {

int tb2_partition_rcb(int64_t nn, int64_t ne, const int32_t* h_conn, const double* h_X, int nparts, int32_t* h_owner)
{
    TB2_ARG(nn > 0 && ne >= 0 && h_conn && h_X && nparts >= 1 && h_owner);
    std::vector<double> cent((size_t)ne * 3);
    for (int64_t e = 0; e < ne; e++)
        for (int i = 0; i < 3; i++) {
            // the 8 nodal coordinates are added in node order, as the harness's numpy mean over that axis does: identical centroids,
            // hence identical cuts
            double x[8];
            for (int a = 0; a < 8; a++) {
                const int32_t n = h_conn[8 * e + a];
                if (n < 0 || n >= nn) {
                    set_error("tb2_partition_rcb: element %lld has node %d out of range", (long long)e, n);
                    return TB2_ERR_SIZE;
                }
                x[a] = h_X[3 * (int64_t)n + i];
            }
            double sum = x[0];
            for (int a = 1; a < 8; a++) sum += x[a];
            cent[3 * e + i] = sum / 8.0;
        }
    std::vector<int64_t> ids((size_t)ne);
    std::iota(ids.begin(), ids.end(), (int64_t)0);
    rcb_split(ids, 0, ne, 0, nparts, cent.data(), h_owner);
    return TB2_OK;
}

// One rank's part of the mesh from an element -> rank map.  Called twice: with the output arrays NULL it returns the sizes.
// Local nodes are numbered by ascending global id, elements keep their global order; an interface node (touched by more than one
// rank) gets the slot of its position among all interface nodes in ascending global id, and is owned by the lowest rank touching it.
int tb2_partition_part(int64_t nn, int64_t ne, const int32_t* h_conn, const int32_t* h_owner, int nparts, int rank, int64_t* num_local_nodes,
                       int64_t* num_local_elements, int64_t* num_interface_nodes, int64_t* num_global_interface_nodes, int64_t* h_node_gid,
                       int64_t* h_elem_gid, int32_t* h_local_conn, int32_t* h_if_nodes, int32_t* h_if_slots, uint8_t* h_node_owned)
{
    TB2_ARG(nn > 0 && ne >= 0 && h_conn && h_owner && nparts >= 1 && nparts <= 64 && rank >= 0 && rank < nparts);
    std::vector<unsigned long long> touch((size_t)nn, 0ull);
    for (int64_t e = 0; e < ne; e++) {
        const int32_t r = h_owner[e];
        if (r < 0 || r >= nparts) {
            set_error("tb2_partition_part: element %lld has owner %d out of range", (long long)e, r);
            return TB2_ERR_SIZE;
        }
        for (int a = 0; a < 8; a++) {
            const int32_t n = h_conn[8 * e + a];
            if (n < 0 || n >= nn) {
                set_error("tb2_partition_part: element %lld has node %d out of range", (long long)e, n);
                return TB2_ERR_SIZE;
            }
            touch[n] |= 1ull << r;
        }
    }
    const unsigned long long me = 1ull << rank;
    int64_t nl = 0, nel = 0, nif = 0, nglob = 0;
    for (int64_t e = 0; e < ne; e++) nel += h_owner[e] == rank;
    std::vector<int32_t> g2l; // filled in the second call only
    const bool fill = h_node_gid && h_elem_gid && h_local_conn && h_if_nodes && h_if_slots && h_node_owned;
    if (fill) g2l.assign((size_t)nn, -1);
    for (int64_t n = 0; n < nn; n++) {
        const unsigned long long t = touch[n];
        const bool shared = (t & (t - 1)) != 0; // more than one bit set
        if (t & me) {
            if (fill) {
                g2l[n] = (int32_t)nl;
                h_node_gid[nl] = n;
                h_node_owned[nl] = (t & (me - 1)) == 0 ? 1 : 0; // no lower rank touches it
                if (shared) {
                    h_if_nodes[nif] = (int32_t)nl;
                    h_if_slots[nif] = (int32_t)nglob;
                }
            }
            nl++;
            nif += shared;
        }
        nglob += shared;
    }
    if (num_local_nodes) *num_local_nodes = nl;
    if (num_local_elements) *num_local_elements = nel;
    if (num_interface_nodes) *num_interface_nodes = nif;
    if (num_global_interface_nodes) *num_global_interface_nodes = nglob;
    if (!fill) return TB2_OK;
    int64_t k = 0;
    for (int64_t e = 0; e < ne; e++) {
        if (h_owner[e] != rank) continue;
        h_elem_gid[k] = e;
        for (int a = 0; a < 8; a++) h_local_conn[8 * k + a] = g2l[h_conn[8 * e + a]];
        k++;
    }
    return TB2_OK;
}

} // extern "C"
