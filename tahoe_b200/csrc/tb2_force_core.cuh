// tb2_force_core.cuh -- the integration-point loop of the internal-force sweep (K1), shared by the stand-alone sweep
// (tb2_elements.cu: residuals of the implicit solvers, output) and the block-fused explicit step (tb2_block_step.cuh).
//
// Replaces the body of SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295): SetGlobalShape + FormKd of
// SmallStrainT (SmallStrainT.cpp:255-282,327-401), TotalLagrangianT (TotalLagrangianT.cpp:107-144) and UpdatedLagrangianT
// (UpdatedLagrangianT.cpp:145-171), one element per thread, in the trilinear-mode form of tb2_math.cuh.
//
// Finite strain.  With j = dx/dxi = J0 + du/dxi (J0 = dX/dxi):   F = j J0^-1,  det F = det j / det J0, and both
//   TotalLagrangianT::FormKd  (f_a = sum_ip w detJ0 J (sigma F^-T) dN_a/dX) and
//   UpdatedLagrangianT::FormKd (f_a = sum_ip w det j  sigma dN_a/dx)
// reduce to  f_a = sum_ip w sigma adj(j)^T dN_a/dxi  (adj = det * inverse): the two reference classes are two roundings of the
// same integral, so one body serves both and needs no reciprocal of det j.
#pragma once
#include "tb2_materials.cuh"

namespace tb2 {

// what the material laws need besides the deformation: constants, J2 history (indexed by the global element id), Newton iteration
struct ForceCtx {
    MatConst mat;
    J2Hist hist;
    int64_t e;      // global element id
    int64_t stride; // SoA pitch of the history arrays
    int iteration;
};

// mode coefficients held in registers ...
struct RegModes {
    const Modes& c;
    TB2_DEV explicit RegModes(const Modes& m) : c(m) {}
    TB2_DEV double operator()(int k, int i) const { return c.m[k][i]; }
};
// ... or in a private shared-memory column [3k+i][pitch] (read-only inside the integration-point loop).  The load is an
// asm volatile so that the compiler neither hoists the 21 values out of the loop nor keeps them in registers across iterations:
// the point of the layout is the 42 registers it frees.
struct SmemModes {
    unsigned addr; // shared-window byte address of this thread's entry of row 0
    unsigned pitch_bytes;
    TB2_DEV SmemModes(const double* col, int pitch) : addr((unsigned)__cvta_generic_to_shared(col)), pitch_bytes((unsigned)pitch * 8u) {}
    TB2_DEV double operator()(int k, int i) const
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + (unsigned)(3 * k + i) * pitch_bytes));
        return v;
    }
};

// d v_i / d xi_k at the integration point with signs (s0,s1,s2): D[i][k] (see mode_gradient in tb2_math.cuh)
template <class M>
TB2_DEV void mode_gradient_of(const M& c, double s0, double s1, double s2, double (&D)[3][3])
{
    const double s12 = s1 * s2, s02 = s0 * s2, s01 = s0 * s1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double m2 = c(2, i), m4 = c(4, i), m5 = c(5, i), m6 = c(6, i);
        D[i][0] = c(0, i) + s1 * m2 + s2 * m4 + s12 * m6;
        D[i][1] = c(1, i) + s0 * m2 + s2 * m5 + s02 * m6;
        D[i][2] = c(3, i) + s0 * m4 + s1 * m5 + s01 * m6;
    }
}

// One integration point of the finite-strain Neo-Hookean laws (SimoIso3D / ExplNeoHookeanT) from J0 = dX/dxi and j = dx/dxi:
// G = w det(j) sigma j^-T (see force_modes for the algebra).  Returns false on a non-positive Jacobian.
template <int MAT>
TB2_DEV bool neo_point(const MatConst& mat, const double (&J0)[3][3], const double (&j)[3][3], double (&G)[3][3])
{
    double J0a[3][3], ja[3][3], M0[6], N[3][3];
    const double det0 = adj3(J0, J0a);
    sym_fft(J0a, M0); // adj(J0) adj(J0)^T
    const double detj = adj3(j, ja);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        N[i][0] = j[i][0] * M0[0] + j[i][1] * M0[5] + j[i][2] * M0[4];
        N[i][1] = j[i][0] * M0[5] + j[i][1] * M0[1] + j[i][2] * M0[3];
        N[i][2] = j[i][0] * M0[4] + j[i][1] * M0[3] + j[i][2] * M0[2];
    }
    double al, q;
    if (MAT == kSimoIso) {
        double trb = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) trb += N[i][0] * j[i][0] + N[i][1] * j[i][1] + N[i][2] * j[i][2];
        const double dj2 = detj * detj;
        const double t = rcbrt(dj2 * dj2 * detj * det0);
        const double rdd = (t * t) * t * (dj2 * dj2);
        const double sc = mat.mu * t;
        const double pr = 0.5 * mat.kappa * (dj2 - det0 * det0) * rdd;
        al = sc * detj;
        q = pr - sc * trb * (1.0 / 3.0);
    } else {
        const double r = 1.0 / (det0 * detj);
        const double rJ = det0 * det0 * r;
        al = mat.mu * detj * r;
        q = mat.kappa * (1.0 - rJ) - mat.mu * rJ;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) G[i][k] = al * N[i][k] + q * ja[k][i];
    return det0 > 0.0 && detj > 0.0;
}

// The integration-point loop of the finite-strain Neo-Hookean laws, two points per iteration: the pair (s0 = -1, +1) at fixed
// (s1, s2).  The pair shares the mode loads and the s0-independent part of both gradients
//   d/dxi_0 = m0 + s1 m2 + s2 m4 + s1 s2 m6,   d/dxi_1 = (m1 + s2 m5) + s0 (m2 + s2 m6),   d/dxi_2 = (m3 + s1 m5) + s0 (m4 + s1 m6)
// (11 instead of 18 FP64 instructions per component and pair), the force-mode accumulation works on G+ + G- and G+ - G-
// (18 instead of 24), and the two points give every warp two independent dependency chains.
template <int MAT, class MX, class MU>
TB2_DEV int force_modes_neo_pairs(const MatConst& mat, const MX& cX, const MU& cx, Modes& A)
{
#pragma unroll
    for (int k = 0; k < 7; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) A.m[k][i] = 0.0;
    bool ok = true;
#pragma unroll 1
    for (int pr = 0; pr < 4; pr++) {
        const double s1 = (pr & 1) ? 1.0 : -1.0, s2 = (pr & 2) ? 1.0 : -1.0, s12 = s1 * s2;
        double J0m[3][3], J0p[3][3], jm[3][3], jp[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            {
                const double m2 = cX(2, i), m4 = cX(4, i), m5 = cX(5, i), m6 = cX(6, i);
                const double c0 = cX(0, i) + s1 * m2 + s2 * m4 + s12 * m6;
                const double P = cX(1, i) + s2 * m5, Q = m2 + s2 * m6, R = cX(3, i) + s1 * m5, S = m4 + s1 * m6;
                J0m[i][0] = c0; J0p[i][0] = c0;
                J0m[i][1] = P - Q; J0p[i][1] = P + Q;
                J0m[i][2] = R - S; J0p[i][2] = R + S;
            }
            {
                const double m2 = cx(2, i), m4 = cx(4, i), m5 = cx(5, i), m6 = cx(6, i);
                const double c0 = cx(0, i) + s1 * m2 + s2 * m4 + s12 * m6;
                const double P = cx(1, i) + s2 * m5, Q = m2 + s2 * m6, R = cx(3, i) + s1 * m5, S = m4 + s1 * m6;
                jm[i][0] = c0; jp[i][0] = c0;
                jm[i][1] = P - Q; jp[i][1] = P + Q;
                jm[i][2] = R - S; jp[i][2] = R + S;
            }
        }
        double Gm[3][3], Gp[3][3];
        ok = neo_point<MAT>(mat, J0m, jm, Gm) && ok;
        ok = neo_point<MAT>(mat, J0p, jp, Gp) && ok;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double S0 = Gp[i][0] + Gm[i][0], S1 = Gp[i][1] + Gm[i][1], S2 = Gp[i][2] + Gm[i][2];
            const double D1 = Gp[i][1] - Gm[i][1], D2 = Gp[i][2] - Gm[i][2];
            A.m[0][i] += S0;
            A.m[1][i] += S1;
            A.m[3][i] += S2;
            A.m[2][i] += fma(s1, S0, D1);
            A.m[4][i] += fma(s2, S0, D2);
            A.m[5][i] = fma(s1, S2, fma(s2, S1, A.m[5][i]));
            A.m[6][i] = fma(s1, D2, fma(s2, D1, fma(s12, S0, A.m[6][i])));
        }
    }
    return ok ? kErrNone : kErrBadJacobian;
}

// the integration-point loop: A <- force modes of one element.  cX: modes of X; cU: modes of u (small strain) or of x = X + u
// (finite strain); cL: modes of u_last (J2Simo3D only).  Returns the element's error code (kErrNone / kErrBadJacobian / ...).
template <int FORM, int MAT, class MX, class MU>
TB2_DEV int force_modes(const ForceCtx& p, const MX& cX, const MU& rU, const Modes& cL, Modes& A)
{
#pragma unroll
    for (int k = 0; k < 7; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) A.m[k][i] = 0.0;
    int alloc = 0, err = kErrNone;
    if (MAT == kJ2Simo) alloc = p.hist.alloc[p.e];
    // Mean-dilatation B-bar (SmallStrainT.cpp:337-374).  B-bar_a = B_a + 1/3 m (b_a - grad N_a)^T with b_a the volume average of
    // grad N_a (Hughes 4.5.23), so the strain at a point is eps + 1/3 (theta_bar - tr eps) 1 with
    //   theta_bar = sum_a b_a . u_a = sum_ip w det0 tr(grad u) / sum_ip w det0 = sum_ip tr(H adj(J0)) / sum_ip det0,
    // and, tr(sigma) = 3 kappa theta_bar being the same at every point for the isotropic linear material, the B-bar^T sigma
    // integral equals the plain B^T sigma integral of those stresses (the (b_a - grad N_a) tr(sigma)/3 terms cancel in the sum).
    double theta_bar = 0.0;
    if (MAT == kSSKStVBbar) {
        double num = 0.0, vol = 0.0;
#pragma unroll 1
        for (int ip = 0; ip < 8; ip++) {
            double s0, s1, s2, J0[3][3], H[3][3], J0a[3][3];
            ip_signs(ip, s0, s1, s2);
            mode_gradient_of(cX, s0, s1, s2, J0);
            mode_gradient_of(rU, s0, s1, s2, H);
            vol += adj3(J0, J0a);
#pragma unroll
            for (int i = 0; i < 3; i++) num += H[i][0] * J0a[0][i] + H[i][1] * J0a[1][i] + H[i][2] * J0a[2][i];
        }
        theta_bar = num / vol;
    }

#pragma unroll 1
    for (int ip = 0; ip < 8; ip++) {
        double s0, s1, s2;
        ip_signs(ip, s0, s1, s2);
        double J0[3][3], H[3][3], J0a[3][3], G[3][3], S[3][3];
        mode_gradient_of(cX, s0, s1, s2, J0);
        mode_gradient_of(rU, s0, s1, s2, H); // small strain: du/dxi; finite strain: j = dx/dxi
        const double det0 = adj3(J0, J0a);
        if (det0 <= 0.0) err = kErrBadJacobian; // ParentDomainT::ComputeDNa, ParentDomainT.cpp:451
        double sig[6];
        if (FORM == kSmallStrain) {
            // SmallStrainT::SetGlobalShape (SmallStrainT.cpp:327-401): eps = sym(grad_X u), grad_X u = H J0^-1
            const double rdet0 = 1.0 / det0;
            double g[3][3], eps[6];
            mul3(H, J0a, g);
            eps[0] = g[0][0] * rdet0;
            eps[1] = g[1][1] * rdet0;
            eps[2] = g[2][2] * rdet0;
            eps[3] = 0.5 * (g[1][2] + g[2][1]) * rdet0;
            eps[4] = 0.5 * (g[0][2] + g[2][0]) * rdet0;
            eps[5] = 0.5 * (g[0][1] + g[1][0]) * rdet0;
            if (MAT == kSSKStVBbar) {
                const double corr = (theta_bar - (eps[0] + eps[1] + eps[2])) * (1.0 / 3.0);
                eps[0] += corr; eps[1] += corr; eps[2] += corr;
            }
            hooke_stress(p.mat, eps, sig);
            sym_to_mat(sig, S);
            mul3_abt(S, J0a, G); // G = w detJ0 sigma J0^-T
        } else {
            double ja[3][3], F[3][3];
            const double (&j)[3][3] = H;
            const double detj = adj3(j, ja);
            if (detj <= 0.0) err = kErrBadJacobian; // TotalLagrangianT.cpp:127-128 / current-configuration ComputeDNa
            if (MAT == kSimoIso || MAT == kExplNeo) {
                // SimoIso3D::s_ij (SimoIso3D.cpp:127-136) folded into the force integrand.  With F' = det0 F = j adj(J0) and
                // cof(j) = adj(j)^T:  b' cof(j) = F' F'^T cof(j) = det(j) F' adj(J0)^T = det(j) j M0,  M0 = adj(J0) adj(J0)^T, so
                //   G = w det(j) sigma j^-T = sc det(j) (j M0) + (U'(J) - sc tr(b')/3) cof(j),   tr(b') = (j M0) : j,
                //   sc = (mu/J) J^(-2/3) / det0^2 = mu (det(j)^5 det0)^(-1/3):  one reciprocal cube root and no division;
                //   1/(det0 det j) = t^3 det(j)^4 with t = sc/mu supplies 1/J for U'(J) = kappa/2 (J - 1/J) (SimoIso3D.h:93-96).
                // ExplNeoHookeanT (ExplNeoHookeanT.cpp:79-111), sigma = (mu/J)(b - 1) + kappa (1 - 1/J) 1, folds the same way:
                //   G = (mu/det0) (j M0) + (kappa (1 - 1/J) - mu/J) cof(j)
                double M0[6], N[3][3];
                sym_fft(J0a, M0); // adj(J0) adj(J0)^T
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    N[i][0] = j[i][0] * M0[0] + j[i][1] * M0[5] + j[i][2] * M0[4];
                    N[i][1] = j[i][0] * M0[5] + j[i][1] * M0[1] + j[i][2] * M0[3];
                    N[i][2] = j[i][0] * M0[4] + j[i][1] * M0[3] + j[i][2] * M0[2];
                }
                double al, q;
                if (MAT == kSimoIso) {
                    double trb = 0.0;
#pragma unroll
                    for (int i = 0; i < 3; i++) trb += N[i][0] * j[i][0] + N[i][1] * j[i][1] + N[i][2] * j[i][2];
                    const double dj2 = detj * detj;
                    const double t = rcbrt(dj2 * dj2 * detj * det0);
                    const double rdd = (t * t) * t * (dj2 * dj2);               // 1 / (det0 det j)
                    const double sc = p.mat.mu * t;
                    const double pr = 0.5 * p.mat.kappa * (dj2 - det0 * det0) * rdd; // U'(J) = kappa/2 (J - 1/J)
                    al = sc * detj;
                    q = pr - sc * trb * (1.0 / 3.0);
                } else {
                    const double r = 1.0 / (det0 * detj); // one reciprocal: 1/det0 = det(j) r, 1/J = det0^2 r
                    const double rJ = det0 * det0 * r;
                    al = p.mat.mu * detj * r;
                    q = p.mat.kappa * (1.0 - rJ) - p.mat.mu * rJ;
                }
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int k = 0; k < 3; k++) G[i][k] = al * N[i][k] + q * ja[k][i];
                mode_accumulate(A, s0, s1, s2, G);
                continue;
            }
            const double rdet0 = 1.0 / det0;
            mul3(j, J0a, F); // = det0 * F
            const double J = detj * rdet0;
            scale3(F, rdet0);
            if (MAT == kFDKStV)
                fdkstv_stress(p.mat, F, J, sig);
            else if (MAT == kExplJ2)
                expl_j2_stress(p.mat, p.hist.data + (int64_t)(ip * 16) * p.stride + p.e, p.stride, F, sig);
            else if (MAT == kJ2Simo) {
                double Hl[3][3], Fl[3][3], c[6][6];
                mode_gradient(cL, s0, s1, s2, Hl);
                mul3(Hl, J0a, Fl);
                scale3(Fl, rdet0);
                Fl[0][0] += 1.0; Fl[1][1] += 1.0; Fl[2][2] += 1.0; // FiniteStrainT::SetGlobalShape, FiniteStrainT.cpp:267-304
                const int e2 = j2_eval<false>(p.mat, p.hist, p.e, ip, alloc, p.iteration, F, Fl, J, sig, c);
                if (e2) err = e2 > err ? e2 : err;
            }
            sym_to_mat(sig, S);
            mul3_abt(S, ja, G); // G = w det(j) sigma j^-T
        }
        mode_accumulate(A, s0, s1, s2, G);
    }
    return err;
}

} // namespace tb2
