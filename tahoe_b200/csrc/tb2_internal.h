// tb2_internal.h -- host-side handle structs shared by the translation units of libtahoe_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/tahoe_b200.h"
#include "tb2_materials.cuh"
#include "tb2_peer.cuh"

namespace tb2 {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TB2_CUDA(call)                                                    \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return tb2::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)
#define TB2_CHECK(call)            \
    do {                           \
        int _s = (call);           \
        if (_s != TB2_OK) return _s; \
    } while (0)
#define TB2_ARG(cond)                                         \
    do {                                                      \
        if (!(cond)) {                                        \
            tb2::set_error("bad argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
            return TB2_ERR_ARG;                               \
        }                                                     \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    // room for at least count elements, keeping the allocation when it is large enough (contents are not preserved when it grows)
    cudaError_t reserve(size_t count)
    {
        if (p && count <= cap) {
            n = count;
            return cudaSuccess;
        }
        const size_t want = count + count / 2 + 64;
        const cudaError_t e = alloc(want);
        cap = e == cudaSuccess ? want : 0;
        n = count;
        return e;
    }
    size_t cap = 0;
    void release()
    {
        cap = 0;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct Comm; // tb2_comm.cu
// what the overlapped multi-GPU explicit step (tb2_explicit.cu) needs from the communicator
struct CommPlan {
    int64_t n_if = 0, n_glob = 0, nb = 0;
    const int* nodes = nullptr;     // [n_if] local ids of the interface nodes
    const int* slots = nullptr;     // [n_if] their slots in the packed global interface vector
    const int* node_slot = nullptr; // [nn] slot of a node, -1 for nodes private to this rank
    const int* belems = nullptr;    // [nb] elements touching an interface node, ascending
    const unsigned char* belem_flag = nullptr; // [ne] 1 for those elements
    double* packed = nullptr;       // [n_glob][3]
    cudaStream_t stream = nullptr;  // the collective runs here, beside the element sweep
    cudaEvent_t ev_packed = nullptr, ev_reduced = nullptr, ev_done = nullptr;
    // peer-memory exchange (tb2_peer.cuh), when the windows have been imported: no collective library on the data path
    bool peer = false;
    PeerView pv{};
    const unsigned* share_mask = nullptr; // [n_if] sharers of every interface node
    const unsigned* slot_mask = nullptr;  // [n_glob] the same by slot
    unsigned* counter = nullptr;          // last-CTA counter of the publishing kernel
    unsigned long long* epoch = nullptr;  // the communicator's interface-exchange epoch (host side, advanced per exchange)
    unsigned long long* sepoch = nullptr; // ... and its scalar-exchange epoch
};

// per-kernel CUDA-event timing on the mesh stream (bench.py's roofline numbers are measured with these, live, inside the
// timed region) and a count of our own kernel launches
enum { kProfForce = 0, kProfNodeUpdate = 1, kProfPredictor = 2, kProfSpmv = 3, kProfPcgVec = 4, kProfStiffness = 5, kProfComm = 6,
       kProfOther = 7, kProfNumCat = 8 };
struct ProfRec {
    int cat;
    cudaEvent_t a, b;
};

} // namespace tb2

struct tb2_mesh {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t nn = 0, ne = 0, stride = 0; // stride = ne rounded up to 32 (SoA pitch)
    tb2::DevBuf<int> conn;              // [8][stride] SoA connectivity
    tb2::DevBuf<double> X;              // [nn][3]
    tb2::DevBuf<int> inc_ptr;           // [nn+1] node -> incidence range
    tb2::DevBuf<int> inc;               // [8*ne] entries e*8+a, ascending in e within a node
    tb2::DevBuf<int> inc8;              // [nn][8] the first 8 entries of each node (-1 padded): one 32-byte load instead of the
                                        // inc_ptr -> inc chain in the node kernels (a hex-mesh node has <= 8 incident elements
                                        // unless it is an irregular vertex; those continue in inc[])
    tb2::DevBuf<double> fe;             // [24][stride] element force scratch
    tb2::DevBuf<double> stage_a, stage_b, stage_c; // [nn][3] staging for the *_host entry points
    tb2::DevBuf<double> out48;          // [48][stride] per-element nodal output values (tb2_group_nodal_stress), built on first use
    // colouring (built lazily)
    int ncolours = 0;
    std::vector<int32_t> colour_host;   // [ne]
    tb2::DevBuf<int> colour_elems;      // elements sorted by colour
    std::vector<int64_t> colour_start;  // [ncolours+1]
    tb2::Comm* comm = nullptr;
    cudaEvent_t ev_join = nullptr, ev_k5 = nullptr; // the two lanes of the multi-GPU explicit step (tb2_explicit.cu)
    bool prof_on = false;
    std::vector<tb2::ProfRec> prof; // event pool
    size_t prof_used = 0;
    uint64_t launches = 0;
};

namespace tb2 {
// brackets one or more launches of category cat; counts n kernel launches
struct ProfScope {
    tb2_mesh* m;
    int idx = -1;
    cudaStream_t st;
    ProfScope(tb2_mesh* mesh, int cat, int n = 1, cudaStream_t stream = nullptr) : m(mesh), st(stream ? stream : mesh->stream)
    {
        m->launches += (uint64_t)n;
        if (!m->prof_on) return;
        if (m->prof_used == m->prof.size()) { // grow the event pool (events are reused across profile_begin/end)
            ProfRec r;
            r.cat = cat;
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            m->prof.push_back(r);
        }
        idx = (int)m->prof_used++;
        m->prof[idx].cat = cat;
        cudaEventRecord(m->prof[idx].a, st);
    }
    ~ProfScope()
    {
        if (idx >= 0) cudaEventRecord(m->prof[idx].b, st);
    }
};
} // namespace tb2

struct tb2_group {
    tb2_mesh* mesh = nullptr;
    int form = 0;      // kSmallStrain / kTotalLagrangian / kUpdatedLagrangian (TB2_SMALL_STRAIN_BBAR is stored as kSmallStrain + bbar)
    bool bbar = false; // SmallStrainT::kMeanDilBbar
    tb2_material mat{};
    tb2::MatConst mc{};
    // J2 history
    tb2::DevBuf<double> hist;      // [38][8][stride]
    tb2::DevBuf<double> hist_save; // committed copy for ResetStep is not needed: trial fields are recomputed
    tb2::DevBuf<int> hist_flag;    // [8][stride]
    tb2::DevBuf<int> hist_alloc;   // [stride]
    tb2::DevBuf<double> spline;    // cubic_spline hardening table (MatConst::spline)
    tb2::DevBuf<unsigned long long> status; // [0] error code, [1] first bad element
    tb2::DevBuf<double> mass_scale; // [ne] ExplicitElementT::fMassScale (null: no mass scaling)
    tb2::DevBuf<unsigned char> off; // [ne] 1 = ElementCardT::kOFF: the element loops skip it (null: every element is on)
};

struct tb2_equations {
    tb2_mesh* mesh = nullptr;
    int64_t neq = 0;
    tb2::DevBuf<int> eqnos;   // [nn][3], 1-based, -1 prescribed
    tb2::DevBuf<int> eq_node; // [neq] -> nodal dof index 3*n+i
};

struct tb2_matrix {
    tb2_equations* eqs = nullptr; // null for a matrix created from raw CSR arrays (tb2_matrix_create_csr)
    tb2_mesh* ctx = nullptr;      // device + stream + instrumentation context (the mesh, or a private one)
    bool owns_ctx = false;
    int64_t neq = 0, nnz = 0;
    tb2::DevBuf<int> adj_ptr;     // [nn+1] node adjacency (sorted neighbour nodes incl. self)
    tb2::DevBuf<int> adj;         // neighbour node ids
    tb2::DevBuf<int> adj_coloff;  // per adjacency entry: column offset of the neighbour's first active dof inside the node's rows
    tb2::DevBuf<int> elem_adjpos; // [64][stride]: position of node b in node a's adjacency list
    tb2::DevBuf<int> grp;          // SpMV row groups: first row | (rows-1) << 30
    int64_t ngroups = 0;
    tb2::DevBuf<long long> rowptr; // [neq+1]
    tb2::DevBuf<int> colind;      // [nnz]
    tb2::DevBuf<double> val;      // [nnz]
    tb2::DevBuf<double> dinv, r, z, p, q; // PCG work vectors [neq]
    tb2::DevBuf<double> scal;     // reduction scalars
    tb2::DevBuf<double> partial;  // per-block partial sums
    tb2::DevBuf<unsigned char> eq_owned; // multi-GPU: 1 if this rank owns the equation's node (dot products count it once)
    bool values_zero = false;            // val is all zero: set by tb2_matrix_clear, dropped by every assembly / set_values
    tb2::DevBuf<int> eq_xidx;            // multi-GPU over peer memory: entry of the exchange buffer an interface equation maps to, -1 elsewhere
    tb2::DevBuf<double> s, partial_if;   // multi-GPU PCG: s = A p of the single-reduction recurrence; (A u, u) partials of the interface rows
    tb2::DevBuf<double> bi_rhat, bi_v, bi_s, bi_t; // BiCGStab work vectors (tb2_matrix_bicgstab), allocated on first use
    tb2::DevBuf<int> grp_split;          // row groups of interface nodes first, then the others
    int64_t ngroups_if = 0;
    // CUDA graph of 16 PCG iterations (5 launches each) for the solution vector / tolerances it was captured with
    cudaGraphExec_t pcg_exec = nullptr;
    const double* pcg_x = nullptr;
    double pcg_rtol = 0.0, pcg_atol = 0.0;
    int pcg_maxit = 0;
    bool pcg_last_converged = true; // the last tb2_matrix_pcg met rtol / atol (false: it stopped at max_iter)
    double pcg_last_rel = 0.0;      // its |r| / |r0|
    // two-phase assembly (tb2_stiffness.cu): contributions e*64+a*8+b of every node block, ascending in e; the element-matrix
    // scratch of one element chunk; the node range each chunk touches
    int64_t nadj = 0;
    tb2::DevBuf<unsigned> contrib_ptr, contrib;
    tb2::DevBuf<double> ke;        // [k3_chunk][k3_rec] element matrices of the current chunk (k3_rec = 300 packed upper triangle or 576)
    int64_t k3_chunk = 0;
    int k3_rec = 0;
    std::vector<int> k3_nmin, k3_nmax;
};

struct tb2_contact;
namespace tb2 {
// tb2_contact.cu, for the resident explicit step: the pair forces of (u, v) summed into f for the nodes that appear in a pair only
// (f is overwritten there and left alone elsewhere); the version counts tb2_contact_set_pairs calls
int contact_form_touched(tb2_contact* c, double constKd, const double* d_u, const double* d_v, double* d_f, cudaStream_t st);
unsigned long long contact_version(const tb2_contact* c);
} // namespace tb2

struct tb2_explicit {
    tb2_group* group = nullptr;
    int device = 0;                // copy for tb2_explicit_destroy (the group may be gone by then)
    tb2::DevBuf<double> d, v, a, mass, minv, fext, fint, bcval;
    tb2::DevBuf<unsigned char> bccode;
    bool has_fext = false; // fext holds a non-zero entry (an all-zero external force is not read by the node kernel)
    // an attached contact_3D_penalty group (tb2_explicit_attach_contact): its force on the predicted state is re-formed before every sweep
    tb2_contact* contact = nullptr;
    tb2::DevBuf<double> fadd;                     // [nn][3] zero outside the nodes of the current pair list
    unsigned long long contact_version = ~0ull;   // pair-list version fadd was last cleared for
    bool contact_searched = false;                // a group with surfaces: its pair list belongs to the current state
    cudaStream_t stream_aux = nullptr;            // the contact kernels run here, beside the element sweep
    cudaEvent_t ev_state = nullptr, ev_loads = nullptr;
    // tb2_explicit_run_async: displacement snapshots on their way to the host beside the next steps' kernels
    tb2::DevBuf<double> dsnap[2];
    cudaStream_t stream_copy = nullptr;
    cudaEvent_t ev_snap[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    unsigned long long* h_status = nullptr; // pinned [2][2]: the group's status words as of each delivered snapshot
    int tickets = 0;
};
