// tb2_elements.cu -- K1 (internal force), K4 (lumped mass), the deterministic node gather, and the element-group API.
//
// K1 replaces SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295).  One thread owns one element and walks its 8
// integration points; the integration-point loop itself lives in tb2_force_core.cuh (trilinear-mode form: no shape-function
// table, 27-36 FMA per 3x3 instead of 72).
//
// Scatter.  Threads write the 24 element values to an SoA scratch fe[24][stride] (coalesced); a node kernel then sums
// each node's <= 8 contributions in ascending element order -- the order of the reference's serial assembly
// (SolverT::AssembleRHS, SolverT.cpp:446-477) -- so there are no float atomics and reruns are bit-reproducible.
//
// Round 2 measured the alternatives to this two-kernel scheme on B200 (profiles/r02_summary.md, profiles/tools/k1_lab.cu):
// forces kept in shared memory with a block-local ordered sum and in-kernel node updates -- as one CTA per element block and as
// a persistent warp-specialised kernel with setmaxnreg -- and a two-stream slab pipeline with a register-lean sweep.  All were
// slower than sweep + node kernel back to back: the node work costs more when it is done block-locally (8-byte scattered
// accesses, release atomics and completion reads) than as one streaming pass at 92 % of the HBM peak.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tb2_force_core.cuh"
#include "tb2_internal.h"
#include "tb2_node_update.cuh"

namespace tb2 {

struct ElemArgs {
    int64_t e_begin, ne, stride; // elements [e_begin, ne), or entries [e_begin, ne) of elist
    const int* elist; // optional element index list (multi-GPU: the elements touching partition-interface nodes)
    const unsigned char* skip; // optional [ne] flags: elements a range launch leaves to the index-list launch
    const unsigned char* off;  // optional [ne] flags: ElementCardT::kOFF elements contribute nothing (their scratch rows are zeroed)
    const int* conn;  // [8][stride]
    const double* X;  // [nn][3]
    const double* u;  // [nn][3]
    const double* ul; // [nn][3] (J2) or null
    double* fe;       // [24][stride]
    MatConst mat;
    J2Hist hist;
    int iteration;
    unsigned long long* status; // [0] = max error code, [1] = min failing element
};

TB2_DEV void report(const ElemArgs& p, int err, int64_t e)
{
    atomicMax(p.status, (unsigned long long)err);
    atomicMin(p.status + 1, (unsigned long long)e);
}

// ElementCardT::kOFF (SolidElementT.cpp:1177, ElementRHSDriver: "if (CurrentElement().Flag() != ElementCardT::kOFF)"): the element
// contributes nothing; its 24 scratch rows are zeroed so that the node gather adds zeros
TB2_DEV bool element_is_off(const ElemArgs& p, const int64_t e)
{
    if (!(p.off && p.off[e])) return false;
#pragma unroll
    for (int r = 0; r < 24; r++) p.fe[(int64_t)r * p.stride + e] = 0.0;
    return true;
}

// element of thread t of the launch, or -1 if the thread has nothing to do (out of range, left to another launch, kOFF)
TB2_DEV int64_t element_of_thread(const ElemArgs& p)
{
    const int64_t t = p.e_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= p.ne) return -1;
    const int64_t e = p.elist ? (int64_t)__ldg(p.elist + t) : t;
    if (p.skip && p.skip[e]) return -1;
    if (element_is_off(p, e)) return -1;
    return e;
}
TB2_DEV void store_element_forces(const ElemArgs& p, const int64_t e, const Modes& A)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double f[8];
        modes_to_nodes(A, i, f);
#pragma unroll
        for (int a = 0; a < 8; a++) p.fe[(int64_t)(3 * a + i) * p.stride + e] = f[a];
    }
}

// K1, general form: one thread = one element, modes in registers.  MINB resident CTAs of 128 threads set the register budget.
template <int FORM, int MAT, int MINB>
__global__ void __launch_bounds__(128, MINB) k_internal_force(const ElemArgs p)
{
    const int64_t e = element_of_thread(p);
    if (e < 0) return;
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + e);
    Modes cX, cU, cL, A;
    load_modes(p.X, n, cX);
    load_modes(p.u, n, cU);
    if (MAT == kJ2Simo) load_modes(p.ul, n, cL);
    if (FORM != kSmallStrain) {
#pragma unroll
        for (int k = 0; k < 7; k++)
#pragma unroll
            for (int i = 0; i < 3; i++) cU.m[k][i] += cX.m[k][i]; // finite strain: modes of x = X + u, so that j = dx/dxi directly
    }
    ForceCtx fc;
    fc.mat = p.mat;
    fc.hist = p.hist;
    fc.e = e;
    fc.stride = p.stride;
    fc.iteration = p.iteration;
    const int err = force_modes<FORM, MAT>(fc, RegModes(cX), RegModes(cU), cL, A);
    if (err) report(p, err, e);
    store_element_forces(p, e, A);
}

// K1 of the finite-strain Neo-Hookean laws (SimoIso3D, ExplNeoHookeanT: the explicit headline path).  The 42 read-only mode
// coefficients of X and x = X + u live in private shared-memory columns [coefficient][thread] (conflict-free), the integration
// points are taken in pairs (tb2_force_core.cuh: force_modes_neo_pairs), and only the 21 force modes stay in registers across
// the loop: 160 registers, no spills, 3 CTAs/SM.  B200, 10^6 elements (profiles/r02_summary.md): 149 us against 168 us for the
// one-point-per-iteration register form (1934 instead of 2174 FP64 instructions per element, two dependency chains per warp).
template <int MAT>
__global__ void __launch_bounds__(128, 3) k_internal_force_neo(const ElemArgs p)
{
    __shared__ double sX[21 * 128], sx[21 * 128];
    const int64_t e = element_of_thread(p);
    if (e < 0) return;
    const int tid = threadIdx.x;
    {
        int n[8];
#pragma unroll
        for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + e);
        Modes cX, cU;
        load_modes(p.X, n, cX);
        load_modes(p.u, n, cU);
#pragma unroll
        for (int k = 0; k < 7; k++)
#pragma unroll
            for (int i = 0; i < 3; i++) {
                sX[(3 * k + i) * 128 + tid] = cX.m[k][i];
                sx[(3 * k + i) * 128 + tid] = cX.m[k][i] + cU.m[k][i];
            }
    }
    asm volatile("" ::: "memory"); // the mode loads of the loop are volatile asms the compiler must keep behind these stores
    Modes A;
    const int err = force_modes_neo_pairs<MAT>(p.mat, SmemModes(sX + tid, 128), SmemModes(sx + tid, 128), A);
    if (err) report(p, err, e);
    store_element_forces(p, e, A);
}

// ---- nodal stress output ---------------------------------------------------------------------------------
// SolidElementT::ComputeOutput, iNodalStress (SolidElementT.cpp:1352-1840): Cauchy stress at the 8 points, extrapolated to the
// element's nodes with HexahedronT::SetExtrapolation's matrix E[a][ip] = (1 + sqrt3 s_a . s_ip) / 8 (HexahedronT.cpp:2099-2150),
// then averaged over the elements at each node (GroupAverageT).  With S0 = sum_ip sigma and S_d = sum_ip s_ip,d sigma the
// extrapolation is (S0 + sqrt3 (s_a,0 S_1 + s_a,1 S_2 + s_a,2 S_3)) / 8: 4 x 6 accumulators instead of an 8 x 8 matrix product.
template <int FORM, int MAT>
__global__ void __launch_bounds__(128) k_nodal_stress(const ElemArgs p, double* __restrict__ out48)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= p.ne) return;
    if (p.off && p.off[e]) return; // SolidElementT::ComputeOutput skips kOFF elements (SolidElementT.cpp:1450); the averaging does too
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + e);
    Modes cX, cU, cL;
    load_modes(p.X, n, cX);
    load_modes(p.u, n, cU);
    int alloc = 0;
    if (MAT == kJ2Simo) { // history material: F of the last converged step and the element's history, as K1 reads them
        load_modes(p.ul, n, cL);
        alloc = p.hist.alloc[e];
    }
    double theta_bar = 0.0;
    if (MAT == kSSKStVBbar) { // see force_modes (tb2_force_core.cuh)
        double num = 0.0, vol = 0.0;
#pragma unroll 1
        for (int ip = 0; ip < 8; ip++) {
            double s0, s1, s2, J0[3][3], H[3][3], J0a[3][3];
            ip_signs(ip, s0, s1, s2);
            mode_gradient(cX, s0, s1, s2, J0);
            mode_gradient(cU, s0, s1, s2, H);
            vol += adj3(J0, J0a);
#pragma unroll
            for (int i = 0; i < 3; i++) num += H[i][0] * J0a[0][i] + H[i][1] * J0a[1][i] + H[i][2] * J0a[2][i];
        }
        theta_bar = num / vol;
    }
    double S[4][6];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int I = 0; I < 6; I++) S[k][I] = 0.0;
    int err = kErrNone;
#pragma unroll 1
    for (int ip = 0; ip < 8; ip++) {
        double s0, s1, s2, J0[3][3], H[3][3], J0a[3][3], g[3][3], sig[6];
        ip_signs(ip, s0, s1, s2);
        mode_gradient(cX, s0, s1, s2, J0);
        mode_gradient(cU, s0, s1, s2, H);
        const double det0 = adj3(J0, J0a);
        if (det0 <= 0.0) err = kErrBadJacobian;
        const double rdet0 = 1.0 / det0;
        mul3(H, J0a, g);
        scale3(g, rdet0); // grad_X u
        if (FORM == kSmallStrain) {
            double eps[6] = {g[0][0], g[1][1], g[2][2], 0.5 * (g[1][2] + g[2][1]), 0.5 * (g[0][2] + g[2][0]), 0.5 * (g[0][1] + g[1][0])};
            if (MAT == kSSKStVBbar) {
                const double corr = (theta_bar - (eps[0] + eps[1] + eps[2])) * (1.0 / 3.0);
                eps[0] += corr; eps[1] += corr; eps[2] += corr;
            }
            hooke_stress(p.mat, eps, sig);
        } else {
            g[0][0] += 1.0; g[1][1] += 1.0; g[2][2] += 1.0; // F
            const double J = det3(g);
            if (J <= 0.0) err = kErrBadJacobian;
            if (MAT == kFDKStV) fdkstv_stress(p.mat, g, J, sig);
            else if (MAT == kJ2Simo) { // J2Simo3D::s_ij as in the residual sweep; the trial fields it rewrites hold the same values
                double Hl[3][3], Fl[3][3], c[6][6];
                mode_gradient(cL, s0, s1, s2, Hl);
                mul3(Hl, J0a, Fl);
                scale3(Fl, rdet0);
                Fl[0][0] += 1.0; Fl[1][1] += 1.0; Fl[2][2] += 1.0;
                const int e2 = j2_eval<false>(p.mat, p.hist, e, ip, alloc, p.iteration, g, Fl, J, sig, c);
                if (e2) err = e2 > err ? e2 : err;
            } else {
                double b_bar[6];
                simo_bbar(g, J, b_bar);
                simo_cauchy(p.mat, J, b_bar, sig);
            }
        }
#pragma unroll
        for (int I = 0; I < 6; I++) {
            S[0][I] += sig[I];
            S[1][I] += s0 * sig[I];
            S[2][I] += s1 * sig[I];
            S[3][I] += s2 * sig[I];
        }
    }
    if (err) report(p, err, e);
    const double r3 = 1.7320508075688772935;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        double s0, s1, s2;
        ip_signs(a, s0, s1, s2); // node a sits in the corner of integration point a
#pragma unroll
        for (int I = 0; I < 6; I++)
            out48[(int64_t)(6 * a + I) * p.stride + e] = 0.125 * (S[0][I] + r3 * (s0 * S[1][I] + s1 * S[2][I] + s2 * S[3][I]));
    }
}
// GroupAverageT::AssembleAverage + Average: sum of the incident ACTIVE elements' nodal values (ascending element order) times
// 1 / their count; a node with no active element keeps zeros (GroupAverageT.cpp:190-205 divides only where the count is positive)
__global__ void __launch_bounds__(256) k_node_average6(int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                      const double* __restrict__ out48, int64_t stride,
                                                      const unsigned char* __restrict__ off, double* __restrict__ out)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    const int k0 = inc_ptr[n], k1 = inc_ptr[n + 1];
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int count = 0;
    for (int k = k0; k < k1; k++) {
        const int ent = __ldg(inc + k);
        const int64_t e = ent >> 3;
        const int a = ent & 7;
        if (off && off[e]) continue;
        count++;
#pragma unroll
        for (int I = 0; I < 6; I++) acc[I] += __ldg(out48 + (int64_t)(6 * a + I) * stride + e);
    }
    const double s = count > 0 ? 1.0 / (double)count : 0.0;
#pragma unroll
    for (int I = 0; I < 6; I++) out[6 * n + I] = acc[I] * s;
}

// K4: ContinuumElementT::FormMass, kLumpedMass branch (ContinuumElementT.cpp:767-842).  me[a] -> fe[a][stride]
__global__ void __launch_bounds__(128) k_lumped_mass(int64_t ne, int64_t stride, const int* __restrict__ conn,
                                                    const double* __restrict__ X, double density, double* __restrict__ fe,
                                                    unsigned long long* status, const double* __restrict__ mass_scale = nullptr,
                                                    const unsigned char* __restrict__ off = nullptr)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    if (mass_scale) density *= mass_scale[e]; // ExplicitElementT::LHSDriver: FormMass with density * fMassScale[e]
    if (off && off[e]) { // ElementCardT::kOFF (SolidElementT.cpp:1116): no mass contribution
#pragma unroll
        for (int a = 0; a < 8; a++) fe[(int64_t)a * stride + e] = 0.0;
        return;
    }
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(conn + a * stride + e);
    Modes cX;
    load_modes(X, n, cX);
    const double RA[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, SA[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, TA[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    double nee[8], dsum = 0.0, totmas = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++) nee[a] = 0.0;
#pragma unroll
    for (int ip = 0; ip < 8; ip++) {
        double J0[3][3];
        mode_gradient(cX, RA[ip], SA[ip], TA[ip], J0);
        const double det0 = det3(J0);
        if (det0 <= 0.0) {
            atomicMax(status, (unsigned long long)kErrBadJacobian);
            atomicMin(status + 1, (unsigned long long)e);
        }
        const double temp1 = density * det0; // weight = 1
        totmas += temp1;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const double Na = 0.125 * (1.0 + RA[a] * RA[ip] * TB2_G) * (1.0 + SA[a] * SA[ip] * TB2_G) * (1.0 + TA[a] * TA[ip] * TB2_G);
            const double temp2 = temp1 * Na * Na;
            dsum += temp2;
            nee[a] += temp2;
        }
    }
    const double diagmass = totmas / dsum;
#pragma unroll
    for (int a = 0; a < 8; a++) fe[(int64_t)a * stride + e] = diagmass * nee[a];
}

// ---- inertia (a2 FormMa, a16 FormMass): ContinuumElementT::FormMass / FormMa (ContinuumElementT.cpp:678-866, 868-1002) on the
// reference configuration.  mass_type 1 = kConsistentMass (M_ab = sum_ip rho w detJ0 N_a N_b on equal dofs), 2 = kLumpedMass (the
// HRZ diagonal of K4).  One thread per element; `rho` carries the integrator constant (constM or constMa).
struct ElemMass {
    double M[8][8]; // consistent: upper triangle a <= b; lumped: diagonal only
};
TB2_DEV bool element_mass(const int* __restrict__ conn, const double* __restrict__ X, int64_t stride, int64_t e, double rho, int mass_type,
                          ElemMass& out, int (&n)[8])
{
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(conn + a * stride + e);
    Modes cX;
    load_modes(X, n, cX);
    const double RA[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, SA[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, TA[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) out.M[a][b] = 0.0;
    double dsum = 0.0, totmas = 0.0;
    bool ok = true;
#pragma unroll
    for (int ip = 0; ip < 8; ip++) {
        double J0[3][3];
        mode_gradient(cX, RA[ip], SA[ip], TA[ip], J0);
        const double det0 = det3(J0);
        ok = ok && det0 > 0.0;
        const double temp = rho * det0; // weight = 1
        totmas += temp;
        double Na[8];
#pragma unroll
        for (int a = 0; a < 8; a++)
            Na[a] = 0.125 * (1.0 + RA[a] * RA[ip] * TB2_G) * (1.0 + SA[a] * SA[ip] * TB2_G) * (1.0 + TA[a] * TA[ip] * TB2_G);
        if (mass_type == 1) {
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = a; b < 8; b++) out.M[a][b] += temp * Na[a] * Na[b];
        } else {
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const double temp2 = temp * Na[a] * Na[a];
                dsum += temp2;
                out.M[a][a] += temp2;
            }
        }
    }
    if (mass_type != 1) {
        const double diagmass = totmas / dsum;
#pragma unroll
        for (int a = 0; a < 8; a++) out.M[a][a] *= diagmass;
    }
    return ok;
}

// f_e = M_e a_e -> fe[24][stride] (then the K1 node gather)
__global__ void __launch_bounds__(128) k_inertial_force(int64_t ne, int64_t stride, const int* __restrict__ conn, const double* __restrict__ X,
                                                       const double* __restrict__ acc, double rho, int mass_type,
                                                       const double* __restrict__ mass_scale, const unsigned char* __restrict__ off,
                                                       double* __restrict__ fe, unsigned long long* status)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    if (off && off[e]) {
#pragma unroll
        for (int r = 0; r < 24; r++) fe[(int64_t)r * stride + e] = 0.0;
        return;
    }
    if (mass_scale) rho *= mass_scale[e];
    ElemMass em;
    int n[8];
    if (!element_mass(conn, X, stride, e, rho, mass_type, em, n)) {
        atomicMax(status, (unsigned long long)kErrBadJacobian);
        atomicMin(status + 1, (unsigned long long)e);
    }
    double a[8][3];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int i = 0; i < 3; i++) a[b][i] = __ldg(acc + (int64_t)n[b] * 3 + i);
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 8; b++) s += (r <= b ? em.M[r][b] : em.M[b][r]) * a[b][i];
            fe[(int64_t)(3 * r + i) * stride + e] = s;
        }
}

// element mass matrices of [e0, e1) as packed upper triangles ke[e - e0][300] (entry tri24(r, c)), the record form of the two-phase
// assembly (tb2_stiffness.cu): only entries on equal dofs are non-zero
TB2_DEV int mass_tri24(int r, int c) { return r * 24 - ((r * (r - 1)) >> 1) + (c - r); }
__global__ void __launch_bounds__(128) k_element_mass(int64_t e0, int64_t e1, int64_t stride, const int* __restrict__ conn,
                                                     const double* __restrict__ X, double rho, int mass_type,
                                                     const double* __restrict__ mass_scale, const unsigned char* __restrict__ off,
                                                     double* __restrict__ ke, unsigned long long* status)
{
    const int64_t e = e0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= e1) return;
    double* rec = ke + (e - e0) * 300;
    for (int q = 0; q < 300; q++) rec[q] = 0.0;
    if (off && off[e]) return;
    if (mass_scale) rho *= mass_scale[e];
    ElemMass em;
    int n[8];
    if (!element_mass(conn, X, stride, e, rho, mass_type, em, n)) {
        atomicMax(status, (unsigned long long)kErrBadJacobian);
        atomicMin(status + 1, (unsigned long long)e);
    }
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = a; b < 8; b++) {
            if (mass_type != 1 && b != a) continue;
#pragma unroll
            for (int i = 0; i < 3; i++) rec[mass_tri24(3 * a + i, 3 * b + i)] = em.M[a][b];
        }
}

// node gather: out[n][i] = sum over incident (e,a), ascending e, of fe[rows(a,i)][e].
// PER_DOF: rows = 3a+i (forces); else rows = a for all three dofs (lumped mass: same value on the 3 dofs of a node)
template <bool PER_DOF>
__global__ void __launch_bounds__(256) k_node_gather(int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                    const double* __restrict__ fe, int64_t stride, double* __restrict__ out)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    const int k0 = inc_ptr[n], k1 = inc_ptr[n + 1];
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int k = k0; k < k1; k++) {
        const int ent = __ldg(inc + k);
        const int64_t e = ent >> 3;
        const int a = ent & 7;
        if (PER_DOF) {
            f0 += __ldg(fe + (int64_t)(3 * a) * stride + e);
            f1 += __ldg(fe + (int64_t)(3 * a + 1) * stride + e);
            f2 += __ldg(fe + (int64_t)(3 * a + 2) * stride + e);
        } else {
            const double m = __ldg(fe + (int64_t)a * stride + e);
            f0 += m; f1 += m; f2 += m;
        }
    }
    out[3 * n] = f0;
    out[3 * n + 1] = f1;
    out[3 * n + 2] = f2;
}

// ExplicitElementT::ComputeStableTimeStep / ApplyMassScaling (ExplicitElementT.cpp:404-478, 492-571): per element dt = h / c with
// h = cbrt(|d1 . (d2 x d3)| / 6) of the diagonals 0-6, 1-7, 3-5; scale = (target / dt)^2 where dt < target, else 1
__global__ void __launch_bounds__(128) k_explicit_solid_dt(int64_t ne, int64_t stride, const int* __restrict__ conn, const double* __restrict__ X,
                                                          double c, double target, double* __restrict__ dt_out, double* __restrict__ scale)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int pa[3] = {6, 7, 5}, pb[3] = {0, 1, 3};
    double d[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int64_t na = conn[pa[k] * stride + e], nb = conn[pb[k] * stride + e];
#pragma unroll
        for (int i = 0; i < 3; i++) d[k][i] = X[3 * na + i] - X[3 * nb + i];
    }
    const double vol = fabs(d[0][0] * (d[1][1] * d[2][2] - d[1][2] * d[2][1]) - d[0][1] * (d[1][0] * d[2][2] - d[1][2] * d[2][0]) +
                            d[0][2] * (d[1][0] * d[2][1] - d[1][1] * d[2][0])) / 6.0;
    const double dt = cbrt(vol) / c;
    if (dt_out) dt_out[e] = dt;
    if (scale) {
        const double alpha = target / dt;
        scale[e] = dt < target ? alpha * alpha : 1.0;
    }
}
// ExplJ2PlasticityT::InitializeHistory: F_n = 1
__global__ void k_expl_j2_init(int64_t ne, int64_t stride, double* __restrict__ hist)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t e = t % stride;
    const int ip = (int)(t / stride);
    if (ip >= 8 || e >= ne) return;
    double* h = hist + (int64_t)(ip * 16) * stride + e;
    h[0] = 1.0; h[4 * stride] = 1.0; h[8 * stride] = 1.0;
}

__global__ void k_j2_close_step(int64_t ne, J2Hist h, double mu)
{
    // J2SimoC0HardeningT::Update (J2SimoC0HardeningT.cpp:341-384), allocated elements only; one thread per (e, ip)
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t e = t % h.stride;
    const int ip = (int)(t / h.stride);
    if (ip >= 8 || e >= ne || !h.alloc[e]) return;
    double bt[6], bbt[6];
    hist_load6(h, e, ip, kHBBarTrial, bt);
    hist_load6(h, e, ip, kHBetaBarTrial, bbt);
    int& flag = h.flag[(int64_t)ip * h.stride + e];
    if (flag == kJ2Plastic) {
        flag = kJ2Elastic;
        const double dgamma = hist(h, e, ip, kHInternal + kDGamma), mbb = hist(h, e, ip, kHInternal + kMuBarBar);
        const double k = 2.0 * mbb * dgamma / mu;
        hist(h, e, ip, kHInternal + kAlpha) += TB2_SQRT23 * dgamma;
        double n[6];
        hist_load6(h, e, ip, kHUnitNorm, n);
#pragma unroll
        for (int I = 0; I < 6; I++) bt[I] += -k * n[I];
    }
    hist_store6(h, e, ip, kHBBar, bt);
    hist_store6(h, e, ip, kHBetaBar, bbt);
}
__global__ void k_j2_reset_step(int64_t ne, J2Hist h)
{
    // J2SimoC0HardeningT::Reset (:387-407)
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t e = t % h.stride;
    const int ip = (int)(t / h.stride);
    if (ip >= 8 || e >= ne || !h.alloc[e]) return;
    h.flag[(int64_t)ip * h.stride + e] = kJ2Elastic;
    hist(h, e, ip, kHInternal + kDGamma) = 0.0;
}

typedef void (*force_kernel_t)(const ElemArgs);
// Register budgets (resident CTAs of 128 threads per SM) from the ncu studies: the small-strain and history materials run at 2
// (<= 255 registers, no spills); the finite-strain Neo-Hookean laws use the shared-memory pair kernel.
static force_kernel_t pick_force_kernel(int form, int mat, bool bbar = false)
{
    if (form == kSmallStrain) {
        if (mat != TB2_SSKSTV) return nullptr;
        return bbar ? k_internal_force<kSmallStrain, kSSKStVBbar, 2> : k_internal_force<kSmallStrain, kSSKStV, 2>;
    }
    // UpdatedLagrangianT shares the finite-strain body (tb2_force_core.cuh)
    switch (mat) {
    case TB2_FDKSTV: return k_internal_force<kTotalLagrangian, kFDKStV, 2>;
    case TB2_SIMO_ISO: return k_internal_force_neo<kSimoIso>;
    case TB2_J2_SIMO: return k_internal_force<kTotalLagrangian, kJ2Simo, 2>;
    case TB2_EXPL_NEO_HOOKEAN: return k_internal_force_neo<kExplNeo>;
    case TB2_EXPL_J2: return k_internal_force<kTotalLagrangian, kExplJ2, 2>;
    }
    return nullptr;
}

J2Hist group_hist(tb2_group* g)
{
    J2Hist h;
    h.data = g->hist.p;
    h.flag = g->hist_flag.p;
    h.alloc = g->hist_alloc.p;
    h.stride = g->mesh->stride;
    return h;
}

int launch_element_forces_range(tb2_group* g, const double* d_u, const double* d_ul, int iteration, int64_t e0, int64_t e1, cudaStream_t st,
                                const int* d_elist = nullptr, const unsigned char* d_skip = nullptr);

// element sweep only: fe scratch <- element forces (used by the fused explicit path too)
int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration)
{
    return launch_element_forces_range(g, d_u, d_ul, iteration, 0, g->mesh->ne, g->mesh->stream);
}

// elements [e0, e1) on stream st (the slab pipeline of tb2_explicit.cu launches the sweep in chunks)
// (with d_elist: entries [e0, e1) of that element index list; with d_skip: flagged elements are left out)
int launch_element_forces_range(tb2_group* g, const double* d_u, const double* d_ul, int iteration, int64_t e0, int64_t e1, cudaStream_t st,
                                const int* d_elist, const unsigned char* d_skip)
{
    tb2_mesh* m = g->mesh;
    force_kernel_t k = pick_force_kernel(g->form, g->mat.kind, g->bbar);
    if (!k) {
        set_error("formulation %d does not support material %d", g->form, g->mat.kind);
        return TB2_ERR_ARG;
    }
    if (g->mat.kind == TB2_J2_SIMO && !d_ul) {
        set_error("J2Simo3D needs the last converged displacement (u_last)");
        return TB2_ERR_ARG;
    }
    ElemArgs p;
    p.e_begin = e0;
    p.ne = e1;
    p.stride = m->stride;
    p.elist = d_elist;
    p.skip = d_skip;
    p.conn = m->conn.p;
    p.X = m->X.p;
    p.u = d_u;
    p.ul = d_ul;
    p.fe = m->fe.p;
    p.off = g->off.p;
    p.mat = g->mc;
    p.hist = group_hist(g);
    p.iteration = iteration;
    p.status = g->status.p;
    if (e1 <= e0) return TB2_OK;
    const int T = 128;
    const unsigned grid = (unsigned)((e1 - e0 + T - 1) / T);
    {
        ProfScope ps(m, kProfForce, 1, st);
        k<<<grid, T, 0, st>>>(p);
    }
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int launch_node_gather(tb2_mesh* m, double* d_out, bool per_dof)
{
    const int T = 256;
    const unsigned nb = (unsigned)((m->nn + T - 1) / T);
    ProfScope ps(m, kProfNodeUpdate);
    if (per_dof) k_node_gather<true><<<nb, T, 0, m->stream>>>(m->nn, m->inc_ptr.p, m->inc.p, m->fe.p, m->stride, d_out);
    else k_node_gather<false><<<nb, T, 0, m->stream>>>(m->nn, m->inc_ptr.p, m->inc.p, m->fe.p, m->stride, d_out);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int ensure_stage(tb2_mesh* m, int which)
{
    DevBuf<double>& b = which == 0 ? m->stage_a : (which == 1 ? m->stage_b : m->stage_c);
    if (!b.p) TB2_CUDA(b.alloc(3 * m->nn));
    return TB2_OK;
}

} // namespace tb2

using namespace tb2;

// CubicSplineT::SetSpline (toolbox/src/C1functions/CubicSplineT.cpp:254-324): second derivatives at the knots from the tridiagonal
// continuity system (free_run: zero end curvature; parabolic: end curvature equal to the neighbour's), then one cubic per interval
// about its left knot, plus one row before the first and after the last knot that continue the curve with its end slope/curvature.
// table = [knot_x[n] | (n+1) x 4 coefficients]
static bool spline_table(const tb2_material& mat, std::vector<double>& table)
{
    const int n = mat.num_knots;
    if (n < 3 || n > TB2_MAX_KNOTS) return false;
    if (mat.spline_fixity != TB2_SPLINE_PARABOLIC && mat.spline_fixity != TB2_SPLINE_FREE_RUN) return false;
    const double* x = mat.knot_x;
    const double* y = mat.knot_y;
    std::vector<double> h(n - 1), m2(n, 0.0), lo(n), di(n), up(n);
    for (int i = 0; i + 1 < n; i++) {
        h[i] = x[i + 1] - x[i];
        if (!(h[i] > 0.0)) return false;
    }
    const int neq = n - 2;
    for (int i = 0; i < neq; i++) {
        lo[i] = h[i] / 6.0;
        di[i] = (h[i] + h[i + 1]) / 3.0;
        up[i] = h[i + 1] / 6.0;
        m2[i + 1] = (y[i + 2] - y[i + 1]) / h[i + 1] - (y[i + 1] - y[i]) / h[i];
    }
    if (mat.spline_fixity == TB2_SPLINE_PARABOLIC) {
        di[0] += h[0] / 6.0;
        di[neq - 1] += h[neq - 1] / 6.0;
    }
    for (int i = 1; i < neq; i++) { // Thomas algorithm in TriDiagdMatrixT::LinearSolve's operation order
        const double factor = lo[i] / di[i - 1];
        di[i] -= up[i - 1] * factor;
        m2[i + 1] -= m2[i] * factor;
    }
    m2[neq] /= di[neq - 1];
    for (int i = neq - 2; i >= 0; i--) m2[i + 1] = (m2[i + 1] - up[i] * m2[i + 2]) / di[i];
    if (mat.spline_fixity == TB2_SPLINE_PARABOLIC) {
        m2[0] = m2[1];
        m2[n - 1] = m2[n - 2];
    } else
        m2[0] = m2[n - 1] = 0.0;
    table.assign((size_t)n + 4 * (n + 1), 0.0);
    for (int i = 0; i < n; i++) table[i] = x[i];
    double* c = table.data() + n;
    for (int j = 1; j < n; j++) {
        const int i = j - 1;
        c[4 * j + 0] = y[i];
        c[4 * j + 1] = -h[i] * (2.0 * m2[i] + m2[i + 1]) / 6.0 + (y[i + 1] - y[i]) / h[i];
        c[4 * j + 2] = m2[i] / 2.0;
        c[4 * j + 3] = (m2[i + 1] - m2[i]) / (6.0 * h[i]);
    }
    c[0] = c[4];
    c[1] = c[5];
    c[2] = c[6];
    c[3] = 0.0;
    c[4 * n + 0] = y[n - 1];
    c[4 * n + 1] = h[n - 2] * (m2[n - 2] + 2.0 * m2[n - 1]) / 6.0 + (y[n - 1] - y[n - 2]) / h[n - 2];
    c[4 * n + 2] = m2[n - 1] / 2.0;
    c[4 * n + 3] = 0.0;
    return true;
}

extern "C" {

int tb2_group_create(tb2_mesh* mesh, int form, const tb2_material* mat, tb2_group** out)
{
    TB2_ARG(mesh && mat && out);
    TB2_ARG(form >= 0 && form <= 3 && mat->kind >= 0 && mat->kind <= 5);
    const bool bbar = form == TB2_SMALL_STRAIN_BBAR;
    if (bbar) form = TB2_SMALL_STRAIN;
    if ((form == TB2_SMALL_STRAIN) != (mat->kind == TB2_SSKSTV)) {
        // SSSolidMatT materials go with SmallStrainT, FSSolidMatT materials with FiniteStrainT (MaterialListT checks)
        set_error("material %d is not valid for formulation %d", mat->kind, form);
        return TB2_ERR_ARG;
    }
    DeviceGuard dg(mesh->device);
    tb2_group* g = new tb2_group;
    g->mesh = mesh;
    g->form = form;
    g->bbar = bbar;
    g->mat = *mat;
    g->mc.mu = mat->mu;
    g->mc.lambda = mat->lambda;
    g->mc.kappa = mat->kappa;
    g->mc.hard_kind = mat->hard_kind;
    for (int i = 0; i < 4; i++) g->mc.hard[i] = mat->hard[i];
    g->mc.nknots = 0;
    g->mc.spline = nullptr;
    cudaError_t e = g->status.alloc(2);
    if (e == cudaSuccess && mat->kind == TB2_J2_SIMO && mat->hard_kind == TB2_HARD_CUBIC_SPLINE) {
        std::vector<double> table;
        if (!spline_table(*mat, table)) {
            delete g;
            set_error("cubic_spline hardening needs 3..%d knots in ascending order and a valid fixity", (int)TB2_MAX_KNOTS);
            return TB2_ERR_ARG;
        }
        e = g->spline.alloc(table.size());
        if (e == cudaSuccess) e = cudaMemcpy(g->spline.p, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice);
        g->mc.nknots = mat->num_knots;
        g->mc.spline = g->spline.p;
    }
    if (e == cudaSuccess && mat->kind == TB2_EXPL_J2) { // [ip][16][stride], F_n = 1
        e = g->hist.alloc((size_t)16 * 8 * mesh->stride);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->hist.p, 0, g->hist.n * sizeof(double), mesh->stream);
        if (e == cudaSuccess) {
            k_expl_j2_init<<<(unsigned)((8 * mesh->stride + 255) / 256), 256, 0, mesh->stream>>>(mesh->ne, mesh->stride, g->hist.p);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess && mat->kind == TB2_J2_SIMO) {
        e = g->hist.alloc((size_t)kHNumDouble * 8 * mesh->stride);
        if (e == cudaSuccess) e = g->hist_flag.alloc(8 * mesh->stride);
        if (e == cudaSuccess) e = g->hist_alloc.alloc(mesh->stride);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->hist.p, 0, g->hist.n * sizeof(double), mesh->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->hist_flag.p, 0, g->hist_flag.n * sizeof(int), mesh->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->hist_alloc.p, 0, g->hist_alloc.n * sizeof(int), mesh->stream);
    }
    const unsigned long long init[2] = {0ull, ~0ull};
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->status.p, init, sizeof init, cudaMemcpyHostToDevice, mesh->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(mesh->stream);
    if (e != cudaSuccess) {
        delete g;
        return cuda_fail(e, "group allocation", __FILE__, __LINE__);
    }
    *out = g;
    return TB2_OK;
}

int tb2_group_destroy(tb2_group* g)
{
    if (!g) return TB2_OK;
    DeviceGuard dg(g->mesh->device);
    cudaStreamSynchronize(g->mesh->stream);
    delete g;
    return TB2_OK;
}

int tb2_form_internal_force(tb2_group* g, const double* d_u, const double* d_ul, int iteration, double* d_f)
{
    TB2_ARG(g && d_u && d_f);
    DeviceGuard dg(g->mesh->device);
    TB2_CHECK(launch_element_forces(g, d_u, d_ul, iteration));
    return launch_node_gather(g->mesh, d_f, true);
}

int tb2_group_status(tb2_group* g, int64_t* bad_element)
{
    TB2_ARG(g);
    DeviceGuard dg(g->mesh->device);
    unsigned long long st[2];
    TB2_CUDA(cudaMemcpyAsync(st, g->status.p, sizeof st, cudaMemcpyDeviceToHost, g->mesh->stream));
    TB2_CUDA(cudaStreamSynchronize(g->mesh->stream));
    if (bad_element) *bad_element = st[0] ? (int64_t)st[1] : -1;
    if (st[0]) {
        const unsigned long long init[2] = {0ull, ~0ull};
        TB2_CUDA(cudaMemcpyAsync(g->status.p, init, sizeof init, cudaMemcpyHostToDevice, g->mesh->stream));
        TB2_CUDA(cudaStreamSynchronize(g->mesh->stream));
        set_error(st[0] == kErrBadJacobian ? "non-positive Jacobian determinant in element %lld" : "J2 local iteration failed in element %lld",
                  (long long)st[1]);
    }
    return st[0] == kErrBadJacobian ? TB2_ERR_BAD_JACOBIAN : (st[0] == kErrJ2Local ? TB2_ERR_J2_LOCAL : TB2_OK);
}

int tb2_form_internal_force_host(tb2_group* g, const double* h_u, const double* h_ul, int iteration, double* h_f)
{
    TB2_ARG(g && h_u && h_f);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    TB2_CHECK(ensure_stage(m, 0));
    TB2_CHECK(ensure_stage(m, 1));
    TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_ul) {
        TB2_CHECK(ensure_stage(m, 2));
        TB2_CUDA(cudaMemcpyAsync(m->stage_c.p, h_ul, bytes, cudaMemcpyHostToDevice, m->stream));
    }
    TB2_CHECK(tb2_form_internal_force(g, m->stage_a.p, h_ul ? m->stage_c.p : nullptr, iteration, m->stage_b.p));
    TB2_CUDA(cudaMemcpyAsync(h_f, m->stage_b.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    return tb2_group_status(g, nullptr);
}

// ---- <explicit_solid> extras ------------------------------------------------------------------------------
static int explicit_solid_dt(tb2_group* g, double target, std::vector<double>* h_dt, bool want_scale)
{
    tb2_mesh* m = g->mesh;
    const double c = sqrt((g->mat.kappa + 4.0 * g->mat.mu / 3.0) / g->mat.density); // ExplicitMaterialT::WaveSpeed
    DevBuf<double> dt;
    if (h_dt) TB2_CUDA(dt.alloc(m->ne));
    if (want_scale && !g->mass_scale.p) TB2_CUDA(g->mass_scale.alloc(m->ne));
    k_explicit_solid_dt<<<(unsigned)((m->ne + 127) / 128), 128, 0, m->stream>>>(m->ne, m->stride, m->conn.p, m->X.p, c, target, dt.p,
                                                                               want_scale ? g->mass_scale.p : nullptr);
    m->launches++;
    TB2_CUDA(cudaGetLastError());
    if (h_dt) {
        h_dt->resize(m->ne);
        TB2_CUDA(cudaMemcpyAsync(h_dt->data(), dt.p, m->ne * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}

int tb2_group_set_element_status(tb2_group* g, const uint8_t* h_off)
{
    TB2_ARG(g);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    if (!h_off) {
        g->off.release();
        return TB2_OK;
    }
    if (!g->off.p) TB2_CUDA(g->off.alloc(m->stride));
    TB2_CUDA(cudaMemset(g->off.p, 0, m->stride));
    TB2_CUDA(cudaMemcpy(g->off.p, h_off, m->ne, cudaMemcpyHostToDevice));
    return TB2_OK;
}

int tb2_group_stable_time_step(tb2_group* g, double* dt)
{
    TB2_ARG(g && dt);
    DeviceGuard dg(g->mesh->device);
    std::vector<double> h;
    TB2_CHECK(explicit_solid_dt(g, 0.0, &h, false));
    double dt_min = 1.0e30; // the reference's start value
    for (double v : h) dt_min = v < dt_min ? v : dt_min;
    *dt = dt_min;
    return TB2_OK;
}

int tb2_group_set_mass_scaling(tb2_group* g, double target_dt, double scale_factor, int64_t* num_scaled, double* max_factor, double* h_scale)
{
    TB2_ARG(g && target_dt > 0.0 && scale_factor > 0.0);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CHECK(explicit_solid_dt(g, target_dt * scale_factor, nullptr, true));
    std::vector<double> sc(m->ne);
    TB2_CUDA(cudaMemcpy(sc.data(), g->mass_scale.p, m->ne * sizeof(double), cudaMemcpyDeviceToHost));
    int64_t n = 0;
    double mx = 1.0;
    for (double v : sc) {
        n += v != 1.0;
        mx = v > mx ? v : mx;
    }
    if (num_scaled) *num_scaled = n;
    if (max_factor) *max_factor = mx;
    if (h_scale) memcpy(h_scale, sc.data(), m->ne * sizeof(double));
    return TB2_OK;
}

int tb2_group_get_explicit_history(tb2_group* g, double* h_hist)
{
    TB2_ARG(g && h_hist && g->mat.kind == TB2_EXPL_J2);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    TB2_CUDA(cudaMemcpy2D(h_hist, m->ne * sizeof(double), g->hist.p, m->stride * sizeof(double), m->ne * sizeof(double), 128,
                          cudaMemcpyDeviceToHost));
    return TB2_OK;
}

int tb2_group_nodal_stress_at(tb2_group* g, const double* d_u, const double* d_ul, int iteration, double* d_stress)
{
    TB2_ARG(g && d_u && d_stress);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    void (*k)(const ElemArgs, double*) = nullptr;
    if (g->form == TB2_SMALL_STRAIN && g->mat.kind == TB2_SSKSTV)
        k = g->bbar ? k_nodal_stress<kSmallStrain, kSSKStVBbar> : k_nodal_stress<kSmallStrain, kSSKStV>;
    else if (g->form != TB2_SMALL_STRAIN && g->mat.kind == TB2_FDKSTV) k = k_nodal_stress<kTotalLagrangian, kFDKStV>;
    else if (g->form != TB2_SMALL_STRAIN && g->mat.kind == TB2_SIMO_ISO) k = k_nodal_stress<kTotalLagrangian, kSimoIso>;
    else if (g->form != TB2_SMALL_STRAIN && g->mat.kind == TB2_J2_SIMO && d_ul) k = k_nodal_stress<kTotalLagrangian, kJ2Simo>;
    if (!k) {
        set_error("nodal stress output is implemented for SSKStV, FDKStV, SimoIso3D and (with the last displacement) J2Simo3D (material %d)", g->mat.kind);
        return TB2_ERR_ARG;
    }
    if (!m->out48.p) TB2_CUDA(m->out48.alloc((size_t)48 * m->stride));
    ElemArgs p{};
    p.e_begin = 0;
    p.ne = m->ne;
    p.stride = m->stride;
    p.conn = m->conn.p;
    p.X = m->X.p;
    p.u = d_u;
    p.ul = d_ul;
    p.iteration = iteration;
    p.hist = group_hist(g);
    p.mat = g->mc;
    p.status = g->status.p;
    p.off = g->off.p;
    ProfScope ps(m, kProfOther, 2);
    k<<<(unsigned)((m->ne + 127) / 128), 128, 0, m->stream>>>(p, m->out48.p);
    k_node_average6<<<(unsigned)((m->nn + 255) / 256), 256, 0, m->stream>>>(m->nn, m->inc_ptr.p, m->inc.p, m->out48.p, m->stride, g->off.p, d_stress);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

int tb2_group_nodal_stress(tb2_group* g, const double* d_u, double* d_stress)
{
    return tb2_group_nodal_stress_at(g, d_u, nullptr, 0, d_stress);
}

int tb2_group_nodal_stress_at_host(tb2_group* g, const double* h_u, const double* h_ul, int iteration, double* h_stress)
{
    TB2_ARG(g && h_u && h_stress);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CHECK(ensure_stage(m, 0));
    DevBuf<double> out;
    TB2_CUDA(out.alloc(6 * m->nn));
    TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_u, 3 * m->nn * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    if (h_ul) {
        TB2_CHECK(ensure_stage(m, 2));
        TB2_CUDA(cudaMemcpyAsync(m->stage_c.p, h_ul, 3 * m->nn * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    }
    TB2_CHECK(tb2_group_nodal_stress_at(g, m->stage_a.p, h_ul ? m->stage_c.p : nullptr, iteration, out.p));
    TB2_CUDA(cudaMemcpyAsync(h_stress, out.p, 6 * m->nn * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return tb2_group_status(g, nullptr);
}

int tb2_group_nodal_stress_host(tb2_group* g, const double* h_u, double* h_stress)
{
    return tb2_group_nodal_stress_at_host(g, h_u, nullptr, 0, h_stress);
}

int tb2_form_lumped_mass(tb2_group* g, double* d_mass)
{
    TB2_ARG(g && d_mass);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const int T = 128;
    ProfScope ps(m, kProfOther);
    k_lumped_mass<<<(unsigned)((m->ne + T - 1) / T), T, 0, m->stream>>>(m->ne, m->stride, m->conn.p, m->X.p, g->mat.density, m->fe.p,
                                                                       g->status.p, g->mass_scale.p, g->off.p);
    TB2_CUDA(cudaGetLastError());
    return launch_node_gather(m, d_mass, false);
}

int tb2_form_lumped_mass_host(tb2_group* g, double* h_mass)
{
    TB2_ARG(g && h_mass);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CHECK(ensure_stage(m, 1));
    TB2_CHECK(tb2_form_lumped_mass(g, m->stage_b.p));
    TB2_CUDA(cudaMemcpyAsync(h_mass, m->stage_b.p, 3 * m->nn * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    return tb2_group_status(g, nullptr);
}

// a2: the inertia term of SolidElementT::ElementRHSDriver (SolidElementT.cpp:1243-1265): d_f = scale * M a
int tb2_form_inertial_force(tb2_group* g, int mass_type, double scale, const double* d_acc, double* d_f)
{
    TB2_ARG(g && d_acc && d_f && (mass_type == TB2_MASS_CONSISTENT || mass_type == TB2_MASS_LUMPED));
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const int T = 128;
    ProfScope ps(m, kProfOther);
    k_inertial_force<<<(unsigned)((m->ne + T - 1) / T), T, 0, m->stream>>>(m->ne, m->stride, m->conn.p, m->X.p, d_acc, scale * g->mat.density,
                                                                          mass_type, g->mass_scale.p, g->off.p, m->fe.p, g->status.p);
    TB2_CUDA(cudaGetLastError());
    return launch_node_gather(m, d_f, true);
}

int tb2_form_inertial_force_host(tb2_group* g, int mass_type, double scale, const double* h_acc, double* h_f)
{
    TB2_ARG(g && h_acc && h_f);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CHECK(ensure_stage(m, 0));
    TB2_CHECK(ensure_stage(m, 1));
    const size_t bytes = 3 * m->nn * sizeof(double);
    TB2_CUDA(cudaMemcpyAsync(m->stage_a.p, h_acc, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CHECK(tb2_form_inertial_force(g, mass_type, scale, m->stage_a.p, m->stage_b.p));
    TB2_CUDA(cudaMemcpyAsync(h_f, m->stage_b.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    return tb2_group_status(g, nullptr);
}

} // extern "C"

// element mass records of [e0, e1) for the two-phase assembly (tb2_form_mass in tb2_stiffness.cu)
int launch_element_mass(tb2_group* g, int mass_type, double constM, int64_t e0, int64_t e1, double* ke)
{
    tb2_mesh* m = g->mesh;
    const int T = 128;
    k_element_mass<<<(unsigned)((e1 - e0 + T - 1) / T), T, 0, m->stream>>>(e0, e1, m->stride, m->conn.p, m->X.p, constM * g->mat.density,
                                                                        mass_type, g->mass_scale.p, g->off.p, ke, g->status.p);
    return cudaGetLastError() == cudaSuccess ? TB2_OK : TB2_ERR_CUDA;
}

extern "C" {

int tb2_group_close_step(tb2_group* g)
{
    TB2_ARG(g);
    if (g->mat.kind != TB2_J2_SIMO) return TB2_OK;
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const int T = 256;
    k_j2_close_step<<<(unsigned)((8 * m->stride + T - 1) / T), T, 0, m->stream>>>(m->ne, group_hist(g), g->mat.mu);
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}
int tb2_group_reset_step(tb2_group* g)
{
    TB2_ARG(g);
    if (g->mat.kind != TB2_J2_SIMO) return TB2_OK;
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    const int T = 256;
    k_j2_reset_step<<<(unsigned)((8 * m->stride + T - 1) / T), T, 0, m->stream>>>(m->ne, group_hist(g));
    TB2_CUDA(cudaGetLastError());
    return TB2_OK;
}

// layout conversion between the device SoA history and the reference's per-element block
// [b_bar 8x6 | unit_norm 8x6 | beta_bar 8x6 | b_bar_trial 8x6 | beta_bar_trial 8x6 | internal 8x8] (J2SimoC0HardeningT.cpp:429-452)
int tb2_group_get_history(tb2_group* g, double* h_data, int32_t* h_flags, int32_t* h_alloc)
{
    TB2_ARG(g && g->mat.kind == TB2_J2_SIMO);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    const int64_t S = m->stride, ne = m->ne;
    if (h_data) {
        std::vector<double> raw((size_t)kHNumDouble * 8 * S);
        TB2_CUDA(cudaMemcpy(raw.data(), g->hist.p, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < ne; e++) {
            double* out = h_data + e * (5 * 48 + 64);
            for (int blk = 0; blk < 5; blk++)
                for (int ip = 0; ip < 8; ip++)
                    for (int I = 0; I < 6; I++) out[blk * 48 + ip * 6 + I] = raw[((size_t)(blk * 6 + I) * 8 + ip) * S + e];
            for (int ip = 0; ip < 8; ip++)
                for (int k = 0; k < 8; k++) out[240 + ip * 8 + k] = raw[((size_t)(kHInternal + k) * 8 + ip) * S + e];
        }
    }
    if (h_flags) {
        std::vector<int> raw((size_t)8 * S);
        TB2_CUDA(cudaMemcpy(raw.data(), g->hist_flag.p, raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < ne; e++)
            for (int ip = 0; ip < 8; ip++) h_flags[e * 8 + ip] = raw[(size_t)ip * S + e];
    }
    if (h_alloc) TB2_CUDA(cudaMemcpy(h_alloc, g->hist_alloc.p, ne * sizeof(int), cudaMemcpyDeviceToHost));
    return TB2_OK;
}
int tb2_group_set_history(tb2_group* g, const double* h_data, const int32_t* h_flags, const int32_t* h_alloc)
{
    TB2_ARG(g && g->mat.kind == TB2_J2_SIMO && h_data && h_flags && h_alloc);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    const int64_t S = m->stride, ne = m->ne;
    std::vector<double> raw((size_t)kHNumDouble * 8 * S, 0.0);
    std::vector<int> rawf((size_t)8 * S, 0);
    for (int64_t e = 0; e < ne; e++) {
        const double* in = h_data + e * (5 * 48 + 64);
        for (int blk = 0; blk < 5; blk++)
            for (int ip = 0; ip < 8; ip++)
                for (int I = 0; I < 6; I++) raw[((size_t)(blk * 6 + I) * 8 + ip) * S + e] = in[blk * 48 + ip * 6 + I];
        for (int ip = 0; ip < 8; ip++) {
            for (int k = 0; k < 8; k++) raw[((size_t)(kHInternal + k) * 8 + ip) * S + e] = in[240 + ip * 8 + k];
            rawf[(size_t)ip * S + e] = h_flags[e * 8 + ip];
        }
    }
    TB2_CUDA(cudaMemcpy(g->hist.p, raw.data(), raw.size() * sizeof(double), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(g->hist_flag.p, rawf.data(), rawf.size() * sizeof(int), cudaMemcpyHostToDevice));
    TB2_CUDA(cudaMemcpy(g->hist_alloc.p, h_alloc, ne * sizeof(int), cudaMemcpyHostToDevice));
    return TB2_OK;
}

} // extern "C"
