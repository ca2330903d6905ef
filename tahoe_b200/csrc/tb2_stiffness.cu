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
#include <cub/cub.cuh>

#include <cstdlib>

#include "tb2_internal.h"

int launch_element_mass(tb2_group* g, int mass_type, double constM, int64_t e0, int64_t e1, double* ke); // tb2_elements.cu

namespace tb2 {

int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration);
J2Hist group_hist(tb2_group* g);
int launch_node_gather(tb2_mesh* m, double* d_out, bool per_dof);
int ensure_stage(tb2_mesh* m, int which);

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
    const unsigned char* off; // [ne] 1 = ElementCardT::kOFF: zero element matrix
    // colour slice
    const int* elems; // element ids of this colour
    int64_t count;
    // matrix
    const int* eqnos;         // [nn][3]
    const long long* rowptr;  // [neq+1]
    const int* adj_coloff;    // per adjacency entry
    const int* elem_adjpos;   // [64][stride]
    double* val;
    // two-phase path: element range [e0, e1) of this launch, element-matrix scratch ke[e - e0][kstride = 576] (entry r*24+c),
    // or (diagonal mode) the element-vector scratch fe[24][stride]
    int64_t e0, e1, kstride;
    double* ke;
    double* fe;
};

static const int kElemsPerBlock = 16;
static const int kIpDoubles = 24 + 36 + 6 + 2; // dN/dx[3][8], c[6][6], sigma[6], scale, pad
// upper-triangle packing of a symmetric 24 x 24 element matrix: entry (r, c), r <= c, at r*24 - r(r-1)/2 + (c - r); 300 entries
TB2_DEV int tri24(int r, int c) { return r * 24 - ((r * (r - 1)) >> 1) + (c - r); }
static const int kSymEntries = 300;
static const int kTilePitch = 577;            // element-matrix tile of the two-phase path: [16][577] doubles
static const size_t kElemSmem = (size_t)kElemsPerBlock * kTilePitch * sizeof(double); // >= the integration-point tile (69,632 B)

// ---- phase 1 of K3: one integration point (Jacobian, F, stress, spatial tangent c, spatial dN/dx, w det j) parked in shared
// memory; smcol = this element's column of the [ip][field][element] tile
#define SM(ip, f) smcol[((ip)*kIpDoubles + (f)) * kElemsPerBlock]
template <int FORM, int MAT>
TB2_DEV void stiffness_point(const StiffArgs& p, const int64_t e, const int (&n)[8], const int ip, double* smcol)
{
        Modes cX, cU;
        load_modes(p.X, n, cX);
        load_modes(p.u, n, cU);
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
            if (MAT == kSSKStVBbar) { // B-bar: B_a^T (C - kappa m m^T) B_b here, + kappa vol b_a b_b^T after the point loop (phase 2)
                const double kappa = p.mat.lambda + 2.0 * p.mat.mu * (1.0 / 3.0);
#pragma unroll
                for (int I = 0; I < 3; I++)
#pragma unroll
                    for (int Jj = 0; Jj < 3; Jj++) c[I][Jj] -= kappa;
            }
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
        SM(ip, 66) = (p.off && p.off[e]) ? 0.0 : scale; // ElementCardT::kOFF (SolidElementT.cpp:1116): weight 0, so K_e = 0 exactly
}
#undef SM

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

    if (live) stiffness_point<FORM, MAT>(p, e, n, t, sm + el); // ---- phase 1: integration point t
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


// ---- K3, two-phase form (default) ---------------------------------------------------------------------------------------
// Phase A, k_element_stiffness: the element matrices of a contiguous element chunk go to a scratch that is sized to stay in
// L2, laid out [element][row*24 + col]: a CTA stages its 16 matrices in shared memory and writes one contiguous 73.7 kB block.
// Thread (el = tid & 15, t = tid >> 4): phase 1 evaluates integration point t of element el, phase 2 owns row node a = t.
//   kKSym  symmetric tangents (SSKStV, FDKStV, SimoIso3D): only the 36 node blocks (a, (a+k) & 7), k = 0..4 (k = 4 for a < 4
//          only) are accumulated; each is stored together with its transpose, and the diagonal block from its upper triangle --
//          the element matrix is exactly symmetric, as the reference's MultQTBQ(kUpperOnly) + CopySymmetric
//          (SmallStrainT.cpp:285-324, ElementMatrixT::CopySymmetric) makes it.  29 % fewer FP64 instructions than all 64 blocks;
//          the scratch holds the packed upper triangle (300 entries per element).
//   kKFull all 64 blocks (J2Simo3D: TangentType() == kNonSymmetric, J2Simo3D.cpp:18-21)
//   kKDiag the 24 diagonal entries only, to fe[24][stride]: DiagonalMatrixT::Assemble in kDiagOnly mode
//          (DiagonalMatrixT.cpp:107-113), the preconditioner of PCGSolver_LS
// Phase B, k_assemble_gather: one warp per row node A; the lane of column dof j of neighbour block (A, B) sums the block's
// contributions in ascending element order -- the order of the reference's serial ElementLHSDriver loop -- and adds the 3x3 result to the CSR.  No
// colouring, no atomics; reruns are bit-identical.
enum { kKSym = 0, kKFull = 1, kKDiag = 2 };

template <int FORM, int MAT, int MODE>
__global__ void __launch_bounds__(128, (MODE == kKFull || MAT == kJ2Simo) ? 2 : 3) k_element_stiffness(const StiffArgs p)
{
    extern __shared__ double sm[];
    const int el = threadIdx.x & 15, t = threadIdx.x >> 4;
    const int64_t eg = p.e0 + blockIdx.x * (int64_t)kElemsPerBlock + el;
    const bool live = eg < p.e1;
    const int64_t e = live ? eg : p.e0;
#define SM(ip, f) sm[((ip)*kIpDoubles + (f)) * kElemsPerBlock + el]
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + e);
    if (live) stiffness_point<FORM, MAT>(p, e, n, t, sm + el);
    __syncthreads();

    const int a = t;
    constexpr int NK = MODE == kKSym ? 5 : (MODE == kKFull ? 8 : 1);
    const int nk = MODE == kKSym ? (a < 4 ? 5 : 4) : NK; // uniform per warp (a warp holds two values of a: 2w, 2w+1)
    double K[NK][3][3];
#pragma unroll
    for (int k = 0; k < NK; k++)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) K[k][i][j] = 0.0;
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
        for (int k = 0; k < NK; k++) {
            if (k >= nk) continue;
            const int b = MODE == kKFull ? k : ((a + k) & 7);
            const double bx = SM(ip, 0 + b), by = SM(ip, 8 + b), bz = SM(ip, 16 + b);
            const double geo = FORM != kSmallStrain ? gx * bx + gy * by + gz * bz : 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                K[k][i][0] += D[i][0] * bx + D[i][4] * bz + D[i][5] * by;
                K[k][i][1] += D[i][1] * by + D[i][3] * bz + D[i][5] * bx;
                K[k][i][2] += D[i][2] * bz + D[i][3] * by + D[i][4] * bx;
            }
            K[k][0][0] += geo;
            K[k][1][1] += geo;
            K[k][2][2] += geo;
        }
    }
    if (MAT == kSSKStVBbar) {
        // K-bar_ab = sum_ip w det [B_a^T C B_b - kappa g_a g_b^T] + kappa vol b_a b_b^T   (g = grad N; expand B-bar = B + m (b - g)^T / 3
        // with C m = 3 kappa m, m^T B_b = g_b^T, m^T C m = 9 kappa): the first part was accumulated above with the deviatoric moduli
        double vol = 0.0, ba[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
        for (int ip = 0; ip < 8; ip++) {
            const double w = SM(ip, 66);
            vol += w;
            ba[0] += w * SM(ip, 0 + a); ba[1] += w * SM(ip, 8 + a); ba[2] += w * SM(ip, 16 + a);
        }
        const double kappa = p.mat.lambda + 2.0 * p.mat.mu * (1.0 / 3.0);
        const double f = kappa / vol; // b = (sum w g) / vol: kappa vol b_a b_b^T = kappa / vol (sum w g_a)(sum w g_b)^T
#pragma unroll
        for (int k = 0; k < NK; k++) {
            if (k >= nk) continue;
            const int b = MODE == kKFull ? k : ((a + k) & 7);
            double bb[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
            for (int ip = 0; ip < 8; ip++) {
                const double w = SM(ip, 66);
                bb[0] += w * SM(ip, 0 + b); bb[1] += w * SM(ip, 8 + b); bb[2] += w * SM(ip, 16 + b);
            }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) K[k][i][j] += f * ba[i] * bb[j];
        }
    }
#undef SM
    if (MODE == kKDiag) {
        if (live)
#pragma unroll
            for (int i = 0; i < 3; i++) p.fe[(int64_t)(3 * a + i) * p.stride + e] = K[0][i][i];
        return;
    }
    // The 16 element matrices of this CTA are staged in shared memory and leave as one contiguous block.  Symmetric tangents keep
    // the upper triangle only (300 of 576 entries, pitch 301: the 16 elements of a half-warp fall into distinct banks), which
    // halves the scratch traffic of both phases; J2's non-symmetric matrices keep [row*24 + col] (pitch 577).
    __syncthreads(); // every thread is done reading the integration-point tile
    constexpr int REC = MODE == kKSym ? kSymEntries : 576, PITCH = REC + 1;
    double* tile = sm + el * PITCH;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        if (k >= nk) continue;
        const int b = MODE == kKFull ? k : ((a + k) & 7);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int r = 3 * a + i, c = 3 * b + j;
                if (MODE == kKSym) {
                    if (k == 0 && j < i) continue;                       // diagonal block: its upper triangle only
                    tile[r <= c ? tri24(r, c) : tri24(c, r)] = K[k][i][j]; // block (a, b) with b < a is stored as its transpose
                } else
                    tile[r * 24 + c] = K[k][i][j];
            }
    }
    __syncthreads();
    const int64_t first = p.e0 + blockIdx.x * (int64_t)kElemsPerBlock;
    const int nlive = (int)(p.e1 - first < kElemsPerBlock ? p.e1 - first : kElemsPerBlock);
    double* out = p.ke + (first - p.e0) * REC;
    for (int idx = threadIdx.x; idx < nlive * REC; idx += 128) {
        const int q = idx / REC;
        out[idx] = sm[q * PITCH + (idx - q * REC)];
    }
}

struct GatherArgs {
    int64_t n0, n1;        // row nodes of this launch
    int64_t e0, e1;        // element chunk held by the scratch ke[e - e0][rec]
    int sym;               // 1: records are packed upper triangles (300 entries), 0: full 24 x 24 (576)
    int fresh;             // 1: val is known to be all zero (tb2_matrix_clear and nothing since): the running sums start from 0 unread
    const double* ke;
    const int* adj_ptr;
    const int* adj;
    const int* adj_coloff;
    const unsigned* cptr;    // [nadj+1] contribution ranges of the node blocks
    const unsigned* contrib; // e*64 + a*8 + b, ascending in e within a block
    const int* eqnos;
    const long long* rowptr;
    double* val;
};
static const int kGatherWarps = 8; // row nodes per CTA
__global__ void __launch_bounds__(32 * kGatherWarps) k_assemble_gather(const GatherArgs g)
{
    // One warp per row node A.  Lane slot s = 3 k + j is column dof j of neighbour block k, so the lanes of a warp walk the
    // (contiguous) CSR rows of the node: val is read and written coalesced, and the 24 B (block) / 192 B (whole element row)
    // runs of the [element][row][col] scratch are contiguous too.
    const int lane = threadIdx.x & 31;
    const int64_t A = g.n0 + blockIdx.x * (int64_t)kGatherWarps + (threadIdx.x >> 5);
    if (A >= g.n1) return;
    long long row[3];
    bool any_row = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int eq = __ldg(g.eqnos + 3 * A + i);
        row[i] = eq > 0 ? g.rowptr[eq - 1] : -1; // MSRMatrixT::Assemble skips inactive rows (MSRMatrixT.cpp:118-123)
        any_row |= eq > 0;
    }
    if (!any_row) return;
    const int b0 = __ldg(g.adj_ptr + A), len = __ldg(g.adj_ptr + A + 1) - b0;
    for (int s = lane; s < 3 * len; s += 32) {
        const int k = s / 3, j = s - 3 * k, blk = b0 + k;
        const unsigned c1 = __ldg(g.cptr + blk + 1);
        unsigned c = __ldg(g.cptr + blk);
        while (c < c1 && (int64_t)(__ldg(g.contrib + c) >> 6) < g.e0) c++; // first contribution of this chunk (ascending in e)
        if (c >= c1 || (int64_t)(__ldg(g.contrib + c) >> 6) >= g.e1) continue;
        const int64_t B = __ldg(g.adj + blk);
        const int q0 = __ldg(g.eqnos + 3 * B), q1 = __ldg(g.eqnos + 3 * B + 1), q2 = __ldg(g.eqnos + 3 * B + 2);
        if ((j == 0 ? q0 : (j == 1 ? q1 : q2)) <= 0) continue; // inactive column
        const int col = __ldg(g.adj_coloff + blk) + (j > 0 && q0 > 0) + (j > 1 && q1 > 0);
        // the running sums start from the stored entries, so that chunking never changes the order of the additions:
        // val + k_e1 + k_e2 + ... in ascending element order whatever the chunk boundaries are (a matrix known to be all zero is not read)
        double acc[3];
#pragma unroll
        for (int i = 0; i < 3; i++) acc[i] = (row[i] >= 0 && !g.fresh) ? g.val[row[i] + col] : 0.0;
        // (issuing a block's <= 8 contributions together -- codes first, then all 24 record loads -- was measured: 87 registers and
        // predicated loads for the short blocks made the gather 2 ms slower, r02q)
        for (; c < c1; c++) {
            const unsigned code = __ldg(g.contrib + c);
            const int64_t e = code >> 6;
            if (e >= g.e1) break;
            const int a = (code >> 3) & 7, b = code & 7;
            if (g.sym) {
                const double* rec = g.ke + (e - g.e0) * kSymEntries;
                const int cc = 3 * b + j;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const int r = 3 * a + i;
                    acc[i] += rec[r <= cc ? tri24(r, cc) : tri24(cc, r)];
                }
            } else {
                const double* src = g.ke + (e - g.e0) * 576 + (3 * a) * 24 + 3 * b + j;
#pragma unroll
                for (int i = 0; i < 3; i++) acc[i] += src[i * 24];
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (row[i] >= 0) g.val[row[i] + col] = acc[i];
    }
}

// contribution lists of the node blocks: thread per row node walks its incident (element, local node a) pairs in ascending
// element order (the incidence is sorted) and appends e*64 + a*8 + b to block (A, conn[b]) -- every block belongs to one thread
template <bool FILL>
__global__ void __launch_bounds__(128) k_contrib(int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                const int* __restrict__ elem_adjpos, int64_t stride, const unsigned* __restrict__ cptr,
                                                unsigned* __restrict__ cursor, unsigned* __restrict__ contrib)
{
    const int64_t A = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (A >= nn) return;
    for (int k = inc_ptr[A]; k < inc_ptr[A + 1]; k++) {
        const int ent = inc[k];
        const int64_t e = ent >> 3;
        const int a = ent & 7;
        for (int b = 0; b < 8; b++) {
            const int pos = elem_adjpos[(int64_t)(a * 8 + b) * stride + e];
            const unsigned c = cursor[pos];
            cursor[pos] = c + 1;
            if (FILL) contrib[cptr[pos] + c] = (unsigned)(e * 64 + a * 8 + b);
        }
    }
}
// smallest / largest node of every element chunk (chunk is a multiple of 32, so a warp never straddles two chunks)
__global__ void __launch_bounds__(256) k_chunk_node_range(int64_t ne, int64_t stride, const int* __restrict__ conn, int64_t chunk,
                                                         int* __restrict__ nmin, int* __restrict__ nmax)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int lo = 0x7fffffff, hi = -1;
    if (e < ne)
        for (int a = 0; a < 8; a++) {
            const int n = conn[a * stride + e];
            lo = n < lo ? n : lo;
            hi = n > hi ? n : hi;
        }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && hi >= 0) {
        const int64_t c = (e & ~31LL) / chunk;
        atomicMin(nmin + c, lo);
        atomicMax(nmax + c, hi);
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


typedef void (*elem_kernel_t)(const StiffArgs);
static elem_kernel_t pick_elem_kernel(int form, int mat, bool diag, bool bbar = false)
{
    if (form == kSmallStrain) {
        if (mat != kSSKStV) return nullptr;
        if (bbar) return diag ? k_element_stiffness<kSmallStrain, kSSKStVBbar, kKDiag> : k_element_stiffness<kSmallStrain, kSSKStVBbar, kKSym>;
        return diag ? k_element_stiffness<kSmallStrain, kSSKStV, kKDiag> : k_element_stiffness<kSmallStrain, kSSKStV, kKSym>;
    }
    switch (mat) { // UpdatedLagrangianT shares the finite-strain body
    case kFDKStV: return diag ? k_element_stiffness<kTotalLagrangian, kFDKStV, kKDiag> : k_element_stiffness<kTotalLagrangian, kFDKStV, kKSym>;
    case kSimoIso: return diag ? k_element_stiffness<kTotalLagrangian, kSimoIso, kKDiag> : k_element_stiffness<kTotalLagrangian, kSimoIso, kKSym>;
    case kJ2Simo: return diag ? k_element_stiffness<kTotalLagrangian, kJ2Simo, kKDiag> : k_element_stiffness<kTotalLagrangian, kJ2Simo, kKFull>;
    }
    return nullptr;
}

static StiffArgs stiff_args(tb2_group* g, const double* d_u, const double* d_ul, int iteration)
{
    tb2_mesh* m = g->mesh;
    StiffArgs p{};
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
    p.off = g->off.p;
    return p;
}

// Contribution lists, chunk size and per-chunk node ranges of the two-phase assembly (once per matrix).
// r01e measurements on 1M elements: L2-sized chunks (2 waves, 65 MB) 5.2 ms, 8 waves 4.2 ms, one 4.6 GB chunk 3.6 ms -- launch
// tails and revisiting the boundary rows cost more than the DRAM round trip of the scratch; running the gather of chunk c on a
// second stream beside the element kernel of chunk c + 1 gained nothing either (4.1 ms with 4 chunks: the element kernel owns
// the whole register file, the gather only runs in its tail).  So: the largest chunk that fits the scratch budget, one stream.
static int ensure_gather_plan(tb2_matrix* A, elem_kernel_t k, int rec)
{
    // every element kernel that assembles into this matrix needs the opt-in shared-memory size (several groups -- one per material --
    // may share a matrix, each with its own template instance)
    TB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kElemSmem));
    if (A->k3_chunk > 0 && A->k3_rec >= rec) return TB2_OK; // (a later group with a larger record re-plans)
    tb2_mesh* m = A->eqs->mesh;
    const int64_t ne = m->ne, nn = m->nn;
    if (ne >= (1LL << 26)) {
        set_error("two-phase assembly packs the element id in 26 bits: %lld elements on one GPU", (long long)ne);
        return TB2_ERR_SIZE;
    }
    const size_t smem = kElemSmem;
    TB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ---- contribution lists
    DevBuf<unsigned> cursor;
    DevBuf<unsigned char> tmp;
    TB2_CUDA(cursor.alloc(A->nadj + 1));
    TB2_CUDA(A->contrib_ptr.alloc(A->nadj + 1));
    TB2_CUDA(A->contrib.alloc(64 * ne));
    TB2_CUDA(cudaMemsetAsync(cursor.p, 0, (A->nadj + 1) * sizeof(unsigned), m->stream));
    const unsigned nbn = (unsigned)((nn + 127) / 128);
    k_contrib<false><<<nbn, 128, 0, m->stream>>>(nn, m->inc_ptr.p, m->inc.p, A->elem_adjpos.p, m->stride, nullptr, cursor.p, nullptr);
    size_t tb = 0;
    TB2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cursor.p, A->contrib_ptr.p, (int)(A->nadj + 1), m->stream));
    TB2_CUDA(tmp.alloc(tb));
    TB2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cursor.p, A->contrib_ptr.p, (int)(A->nadj + 1), m->stream));
    TB2_CUDA(cudaMemsetAsync(cursor.p, 0, (A->nadj + 1) * sizeof(unsigned), m->stream));
    k_contrib<true><<<nbn, 128, 0, m->stream>>>(nn, m->inc_ptr.p, m->inc.p, A->elem_adjpos.p, m->stride, A->contrib_ptr.p, cursor.p, A->contrib.p);
    TB2_CUDA(cudaGetLastError());
    // ---- chunk size: the fewest chunks whose scratch stays under 6 GB, each a whole number of waves of the element kernel
    // (TB2_K3_CHUNK overrides)
    int sms = 148, occ = 2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, smem);
    if (occ < 1) occ = 1;
    const int64_t wave = (int64_t)sms * occ * kElemsPerBlock;
    const int64_t budget = (int64_t)(6.0e9 / (rec * sizeof(double)));
    const int64_t nch = (ne + budget - 1) / budget;
    int64_t chunk = ((ne + nch - 1) / nch + wave - 1) / wave * wave;
    bool forced = false; // an explicit TB2_K3_CHUNK is taken as is (tests: many chunks on a shuffled mesh)
    if (const char* s = getenv("TB2_K3_CHUNK"))
        if (atoll(s) > 0) { chunk = atoll(s); forced = true; }
    const int64_t ne_pad = (ne + 31) & ~31LL;
    DevBuf<int> d_min, d_max;
    for (;;) {
        chunk = (chunk + 31) & ~31LL;
        if (chunk > ne_pad) chunk = ne_pad;
        const int nchunks = (int)((ne + chunk - 1) / chunk);
        TB2_CUDA(d_min.alloc(nchunks));
        TB2_CUDA(d_max.alloc(nchunks));
        TB2_CUDA(cudaMemsetAsync(d_min.p, 0x7f, nchunks * sizeof(int), m->stream));
        TB2_CUDA(cudaMemsetAsync(d_max.p, 0xff, nchunks * sizeof(int), m->stream));
        k_chunk_node_range<<<(unsigned)((ne + 255) / 256), 256, 0, m->stream>>>(ne, m->stride, m->conn.p, chunk, d_min.p, d_max.p);
        A->k3_nmin.resize(nchunks);
        A->k3_nmax.resize(nchunks);
        TB2_CUDA(cudaMemcpyAsync(A->k3_nmin.data(), d_min.p, nchunks * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        TB2_CUDA(cudaMemcpyAsync(A->k3_nmax.data(), d_max.p, nchunks * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        TB2_CUDA(cudaStreamSynchronize(m->stream));
        int64_t swept = 0;
        for (int c = 0; c < nchunks; c++) swept += (int64_t)A->k3_nmax[c] - A->k3_nmin[c] + 1;
        if (forced || swept <= 4 * nn + 4096 || chunk >= ne_pad || chunk * 4 > budget) break;
        chunk *= 4;
    }
    TB2_CUDA(A->ke.alloc((size_t)rec * chunk));
    A->k3_rec = rec;
    A->k3_chunk = chunk;
    return TB2_OK;
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

// the colour-by-colour read-modify-write form (TB2_K3_COLOURED=1): kept as the cross-check of the two-phase default
static int form_stiffness_coloured(tb2_group* g, tb2_matrix* A, const double* d_u, const double* d_ul, int iteration)
{
    TB2_ARG(g && A && d_u);
    tb2_mesh* m = g->mesh;
    TB2_ARG(A->eqs && A->eqs->mesh == m);
    DeviceGuard dg(m->device);
    stiff_kernel_t k = pick_stiff_kernel(g->form, g->mat.kind);
    TB2_ARG(k != nullptr);
    if (g->bbar) {
        set_error("the coloured assembly form (TB2_K3_COLOURED=1) has no B-bar variant");
        return TB2_ERR_ARG;
    }
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
    p.off = g->off.p;
    p.eqnos = A->eqs->eqnos.p;
    p.rowptr = A->rowptr.p;
    p.adj_coloff = A->adj_coloff.p;
    p.elem_adjpos = A->elem_adjpos.p;
    p.val = A->val.p;
    A->values_zero = false;
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


int tb2_form_stiffness(tb2_group* g, tb2_matrix* A, const double* d_u, const double* d_ul, int iteration)
{
    TB2_ARG(g && A && d_u);
    tb2_mesh* m = g->mesh;
    TB2_ARG(A->eqs && A->eqs->mesh == m);
    const char* env_coloured = getenv("TB2_K3_COLOURED"); // read per call: the tests flip it to cross-check the two forms
    const bool coloured = env_coloured && env_coloured[0] == '1';
    if (coloured) return form_stiffness_coloured(g, A, d_u, d_ul, iteration);
    DeviceGuard dg(m->device);
    elem_kernel_t k = pick_elem_kernel(g->form, g->mat.kind, false, g->bbar);
    TB2_ARG(k != nullptr);
    TB2_CHECK(ensure_gather_plan(A, k, g->mat.kind == TB2_J2_SIMO ? 576 : kSymEntries));
    if (g->mat.kind == TB2_J2_SIMO) {
        TB2_ARG(d_ul != nullptr);
        TB2_CHECK(launch_element_forces(g, d_u, d_ul, iteration)); // settles element allocation in the reference's order
    }
    StiffArgs p = stiff_args(g, d_u, d_ul, iteration);
    p.ke = A->ke.p;
    p.kstride = 576;
    GatherArgs ga;
    ga.sym = g->mat.kind == TB2_J2_SIMO ? 0 : 1; // must match the MODE pick_elem_kernel chose (kKFull for J2, kKSym otherwise)
    ga.ke = A->ke.p;
    ga.adj_ptr = A->adj_ptr.p;
    ga.adj = A->adj.p;
    ga.adj_coloff = A->adj_coloff.p;
    ga.cptr = A->contrib_ptr.p;
    ga.contrib = A->contrib.p;
    ga.eqnos = A->eqs->eqnos.p;
    ga.rowptr = A->rowptr.p;
    ga.val = A->val.p;
    const size_t smem = kElemSmem;
    const int nchunks = (int)A->k3_nmin.size();
    for (int c = 0; c < nchunks; c++) {
        p.e0 = (int64_t)c * A->k3_chunk;
        p.e1 = p.e0 + A->k3_chunk < m->ne ? p.e0 + A->k3_chunk : m->ne;
        ga.e0 = p.e0;
        ga.e1 = p.e1;
        ga.n0 = A->k3_nmin[c];
        ga.n1 = (int64_t)A->k3_nmax[c] + 1;
        ga.fresh = (A->values_zero && c == 0) ? 1 : 0;
        ProfScope ps(m, kProfStiffness, 2);
        k<<<(unsigned)((p.e1 - p.e0 + kElemsPerBlock - 1) / kElemsPerBlock), 128, smem, m->stream>>>(p);
        k_assemble_gather<<<(unsigned)((ga.n1 - ga.n0 + kGatherWarps - 1) / kGatherWarps), 32 * kGatherWarps, 0, m->stream>>>(ga);
    }
    A->values_zero = false;
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int tb2_form_stiffness_diagonal(tb2_group* g, const double* d_u, const double* d_ul, int iteration, double* d_diag)
{
    TB2_ARG(g && d_u && d_diag);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    elem_kernel_t k = pick_elem_kernel(g->form, g->mat.kind, true, g->bbar);
    TB2_ARG(k != nullptr);
    const size_t smem = kElemSmem;
    TB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (g->mat.kind == TB2_J2_SIMO) {
        TB2_ARG(d_ul != nullptr);
        TB2_CHECK(launch_element_forces(g, d_u, d_ul, iteration));
    }
    StiffArgs p = stiff_args(g, d_u, d_ul, iteration);
    p.e0 = 0;
    p.e1 = m->ne;
    p.fe = m->fe.p;
    {
        ProfScope ps(m, kProfStiffness);
        k<<<(unsigned)((m->ne + kElemsPerBlock - 1) / kElemsPerBlock), 128, smem, m->stream>>>(p);
    }
    TB2_CUDA(cudaGetLastError());
    return launch_node_gather(m, d_diag, true);
}

int tb2_form_stiffness_diagonal_host(tb2_group* g, const double* h_u, const double* h_ul, int iteration, double* h_diag)
{
    TB2_ARG(g && h_u && h_diag);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    for (int w = 0; w < 3; w++) TB2_CHECK(ensure_stage(m, w));
    TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_ul) TB2_CUDA(cudaMemcpyAsync(m->stage_c.p, h_ul, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(tb2_form_stiffness_diagonal(g, m->stage_a.p, h_ul ? m->stage_c.p : nullptr, iteration, m->stage_b.p));
    TB2_CUDA(cudaMemcpyAsync(h_diag, m->stage_b.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return tb2_group_status(g, nullptr);
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

// a16 with formM (SolidElementT::ElementLHSDriver, SolidElementT.cpp:1100-1154 -> ContinuumElementT::FormMass): A += constM * M through
// the same two-phase path as the tangent: packed-upper-triangle element records, then the per-node gather in element order
int tb2_form_mass(tb2_group* g, tb2_matrix* A, int mass_type, double constM)
{
    TB2_ARG(g && A && (mass_type == TB2_MASS_CONSISTENT || mass_type == TB2_MASS_LUMPED));
    tb2_mesh* m = g->mesh;
    TB2_ARG(A->eqs && A->eqs->mesh == m);
    DeviceGuard dg(m->device);
    elem_kernel_t k = pick_elem_kernel(g->form, g->mat.kind, false, g->bbar);
    TB2_ARG(k != nullptr);
    TB2_CHECK(ensure_gather_plan(A, k, g->mat.kind == TB2_J2_SIMO ? 576 : kSymEntries));
    GatherArgs ga;
    ga.sym = 1;
    ga.ke = A->ke.p;
    ga.adj_ptr = A->adj_ptr.p;
    ga.adj = A->adj.p;
    ga.adj_coloff = A->adj_coloff.p;
    ga.cptr = A->contrib_ptr.p;
    ga.contrib = A->contrib.p;
    ga.eqnos = A->eqs->eqnos.p;
    ga.rowptr = A->rowptr.p;
    ga.val = A->val.p;
    const int nchunks = (int)A->k3_nmin.size();
    for (int c = 0; c < nchunks; c++) {
        ga.e0 = (int64_t)c * A->k3_chunk;
        ga.e1 = ga.e0 + A->k3_chunk < m->ne ? ga.e0 + A->k3_chunk : m->ne;
        ga.n0 = A->k3_nmin[c];
        ga.n1 = (int64_t)A->k3_nmax[c] + 1;
        ga.fresh = (A->values_zero && c == 0) ? 1 : 0;
        ProfScope ps(m, kProfStiffness, 2);
        TB2_CHECK(launch_element_mass(g, mass_type, constM, ga.e0, ga.e1, A->ke.p));
        k_assemble_gather<<<(unsigned)((ga.n1 - ga.n0 + kGatherWarps - 1) / kGatherWarps), 32 * kGatherWarps, 0, m->stream>>>(ga);
    }
    A->values_zero = false;
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

} // extern "C"
