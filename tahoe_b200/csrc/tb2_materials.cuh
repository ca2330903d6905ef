// tb2_materials.cuh -- device constitutive laws of the hot path (a11-a14).
// Each function returns the Cauchy stress (and, when asked, the spatial tangent in reduced 6x6 form with
// Tahoe's index order 11,22,33,23,13,12) for one integration point, entirely in registers.
#pragma once
#include "tb2_math.cuh"

namespace tb2 {

enum { kSSKStV = 0, kFDKStV = 1, kSimoIso = 2, kJ2Simo = 3,
       kSSKStVBbar = 4, /* kernel-template tag only: SSKStV under SmallStrainT's mean-dilatation B-bar */
       kExplNeo = 5, kExplJ2 = 6 /* <explicit_solid> materials (public kinds TB2_EXPL_NEO_HOOKEAN = 4, TB2_EXPL_J2 = 5) */ };
enum { kSmallStrain = 0, kTotalLagrangian = 1, kUpdatedLagrangian = 2 };
enum { kHardLinear = 0, kHardLinearExp = 1, kHardPowerLaw = 2, kHardCubicSpline = 3 };
enum { kErrNone = 0, kErrBadJacobian = 1, kErrJ2Local = 2 };
// J2SimoC0HardeningT.h:33-36
enum { kJ2NotInit = -1, kJ2Plastic = 0, kJ2Elastic = 1 };
// J2SimoC0HardeningT internal variable slots
enum { kAlpha = 0, kStressNorm = 1, kDGamma = 2, kFTrial = 3, kMuBar = 4, kMuBarBar = 5, kDetFTot = 6, kHeatIncr = 7 };
// history field offsets (doubles) in the order of J2SimoC0HardeningT::LoadData (J2SimoC0HardeningT.cpp:429-452)
enum { kHBBar = 0, kHUnitNorm = 6, kHBetaBar = 12, kHBBarTrial = 18, kHBetaBarTrial = 24, kHInternal = 30, kHNumDouble = 38 };

struct MatConst {
    double mu, lambda, kappa;
    int hard_kind;
    double hard[4];
    // cubic_spline hardening: [knot_x[nknots] | (nknots+1) rows of 4 coefficients] in global memory (a few hundred bytes, read
    // through the read-only path by the yielding points of a J2 group only)
    int nknots;
    const double* spline;
};

// J2 history, SoA: data[(field*8 + ip)*stride + e], flag[ip*stride + e], alloc[e]
struct J2Hist {
    double* data;
    int* flag;
    int* alloc;
    int64_t stride;
};

#define TB2_SQRT23 0.81649658092772603273 /* J2SimoC0HardeningT.cpp:11 */
#define TB2_YIELD_TOL 1.0e-10             /* J2SimoC0HardeningT.cpp:12 */

// ---- Hooke: HookeanMatT::HookeanStress (HookeanMatT.cpp:103-108) with IsotropicT::ComputeModuli (IsotropicT.cpp:153-169).
// e holds tensor shear components; C has mu on the shear diagonal and A_ijkl_B_kl doubles them (dSymMatrixT.cpp:856-887).
TB2_DEV void hooke_stress(const MatConst& m, const double (&e)[6], double (&s)[6])
{
    const double lt = m.lambda * (e[0] + e[1] + e[2]), m2 = 2.0 * m.mu;
    s[0] = lt + m2 * e[0];
    s[1] = lt + m2 * e[1];
    s[2] = lt + m2 * e[2];
    s[3] = m2 * e[3];
    s[4] = m2 * e[4];
    s[5] = m2 * e[5];
}
TB2_DEV void hooke_moduli(const MatConst& m, double (&c)[6][6])
{
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) c[i][j] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j] = m.lambda;
        c[i][i] = m.lambda + 2.0 * m.mu;
        c[i + 3][i + 3] = m.mu;
    }
}

// ---- <explicit_solid> materials ---------------------------------------------------------------------------
// ExplNeoHookeanT::ComputeStress3D (elements/explicit/materials/ExplNeoHookeanT.cpp:79-111): sigma = (mu/J)(b - 1) + kappa (J - 1)/J 1
TB2_DEV void expl_neo_stress(const MatConst& m, const double (&F)[3][3], double J, double (&sig)[6])
{
    double b[6];
    sym_fft(F, b);
    const double rJ = 1.0 / J, muJ = m.mu * rJ, pres = m.kappa * (J - 1.0) * rJ;
    sig[0] = muJ * (b[0] - 1.0) + pres;
    sig[1] = muJ * (b[1] - 1.0) + pres;
    sig[2] = muJ * (b[2] - 1.0) + pres;
    sig[3] = muJ * b[3];
    sig[4] = muJ * b[4];
    sig[5] = muJ * b[5];
}
// ExplJ2PlasticityT::ComputeStress3D (materials/ExplJ2PlasticityT.cpp:87-310): Hughes-Winget incrementally objective J2 with
// linear isotropic hardening (hard[0] = sigma_Y, hard[1] = H).  h points at this (element, ip)'s 16 history values, pitch
// `stride` doubles: F_n (row-major), sigma_n (Voigt), eps_p -- read and overwritten on EVERY evaluation, as the reference does.
TB2_DEV void expl_j2_stress(const MatConst& m, double* __restrict__ h, int64_t stride, const double (&F)[3][3], double (&sig)[6])
{
    double Fn[3][3], Fa[3][3], f[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) Fn[i][k] = h[(int64_t)(3 * i + k) * stride];
    const double rdet = 1.0 / adj3(Fn, Fa); // adj3 returns det and the adjugate: Fn^-1 = adj / det
    mul3(F, Fa, f);
    scale3(f, rdet); // f_rel = F F_n^-1
    const double de11 = f[0][0] - 1.0, de22 = f[1][1] - 1.0, de33 = f[2][2] - 1.0, de23 = 0.5 * (f[1][2] + f[2][1]),
                 de13 = 0.5 * (f[0][2] + f[2][0]), de12 = 0.5 * (f[0][1] + f[1][0]);
    const double w23 = 0.25 * (f[1][2] - f[2][1]), w13 = 0.25 * (f[0][2] - f[2][0]), w12 = 0.25 * (f[0][1] - f[1][0]);
    // Cayley transform Q = (1 - W)^-1 (1 + W)
    const double B[3][3] = {{1.0, w12, -w13}, {-w12, 1.0, w23}, {w13, -w23, 1.0}}, A[3][3] = {{1.0, -w12, w13}, {w12, 1.0, -w23}, {-w13, w23, 1.0}};
    double Ba[3][3], Q[3][3];
    const double rB = 1.0 / adj3(B, Ba);
    mul3(Ba, A, Q);
    scale3(Q, rB);
    double sn[6], sr[6];
#pragma unroll
    for (int I = 0; I < 6; I++) sn[I] = h[(int64_t)(9 + I) * stride];
    sym_qsqt(Q, sn, sr); // Q sigma_n Q^T
    const double lam = m.kappa - 2.0 * m.mu * (1.0 / 3.0), lamTr = lam * (de11 + de22 + de33), mu2 = 2.0 * m.mu;
    const double st11 = sr[0] + lamTr + mu2 * de11, st22 = sr[1] + lamTr + mu2 * de22, st33 = sr[2] + lamTr + mu2 * de33;
    const double st23 = sr[3] + mu2 * de23, st13 = sr[4] + mu2 * de13, st12 = sr[5] + mu2 * de12;
    const double pm = (st11 + st22 + st33) * (1.0 / 3.0);
    const double s11 = st11 - pm, s22 = st22 - pm, s33 = st33 - pm;
    const double eps_p = h[(int64_t)15 * stride];
    const double q = sqrt(1.5 * (s11 * s11 + s22 * s22 + s33 * s33 + 2.0 * (st23 * st23 + st13 * st13 + st12 * st12)));
    const double phi = q - (m.hard[0] + m.hard[1] * eps_p);
    double factor = 1.0, new_eps = eps_p;
    if (phi > 0.0) { // radial return
        const double dlam = phi / (3.0 * m.mu + m.hard[1]);
        factor = 1.0 - 3.0 * m.mu * dlam / q;
        new_eps = eps_p + dlam;
    }
    sig[0] = s11 * factor + pm;
    sig[1] = s22 * factor + pm;
    sig[2] = s33 * factor + pm;
    sig[3] = st23 * factor;
    sig[4] = st13 * factor;
    sig[5] = st12 * factor;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) h[(int64_t)(3 * i + k) * stride] = F[i][k];
#pragma unroll
    for (int I = 0; I < 6; I++) h[(int64_t)(9 + I) * stride] = sig[I];
    h[(int64_t)15 * stride] = new_eps;
}

// ---- FDKStV: FDHookeanMatT::s_ij (FDHookeanMatT.cpp:39-55): E = (F^T F - 1)/2, S = C:E, sigma = F S F^T / J
TB2_DEV void fdkstv_stress(const MatConst& m, const double (&F)[3][3], double J, double (&sig)[6])
{
    double C[6], E[6], S[6];
    sym_ftf(F, C);
    E[0] = 0.5 * (C[0] - 1.0); E[1] = 0.5 * (C[1] - 1.0); E[2] = 0.5 * (C[2] - 1.0);
    E[3] = 0.5 * C[3]; E[4] = 0.5 * C[4]; E[5] = 0.5 * C[5];
    hooke_stress(m, E, S);
    sym_qsqt(F, S, sig);
    const double rJ = 1.0 / J;
#pragma unroll
    for (int I = 0; I < 6; I++) sig[I] *= rJ;
}
// FDHookeanMatT::c_ijkl (FDHookeanMatT.cpp:30-37): push forward of the isotropic C with F (TensorTransformT::FFFFC_3D,
// TensorTransformT.cpp:102-153).  For C = lambda 1x1 + 2 mu I_sym:  c_ijkl = (lambda b_ij b_kl + mu (b_ik b_jl + b_il b_jk)) / J
TB2_DEV void fdkstv_moduli(const MatConst& m, const double (&F)[3][3], double J, double (&c)[6][6])
{
    double b6[6], b[3][3];
    sym_fft(F, b6);
    sym_to_mat(b6, b);
    const double rJ = 1.0 / J;
    const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
    for (int A = 0; A < 6; A++)
#pragma unroll
        for (int B = 0; B < 6; B++) {
            const int i = VI[A], j = VJ[A], k = VI[B], l = VJ[B];
            c[A][B] = (m.lambda * b[i][j] * b[k][l] + m.mu * (b[i][k] * b[j][l] + b[i][l] * b[j][k])) * rJ;
        }
}

// ---- SimoIso3D (SimoIso3D.h:88-101, SimoIso3D.cpp:102-136)
TB2_DEV double simo_dU(const MatConst& m, double J) { return 0.5 * m.kappa * (J - 1.0 / J); }
TB2_DEV double simo_ddU(const MatConst& m, double J) { return 0.5 * m.kappa * (1.0 + 1.0 / (J * J)); }
// sigma = (mu/J) dev(b_bar) + U'(J) 1
TB2_DEV void simo_cauchy(const MatConst& m, double J, const double (&b_bar)[6], double (&sig)[6])
{
    const double mJ = m.mu / J;
#pragma unroll
    for (int I = 0; I < 6; I++) sig[I] = mJ * b_bar[I];
    sym_dev(sig);
    const double p = simo_dU(m, J);
    sig[0] += p; sig[1] += p; sig[2] += p;
}
// c = (U' + J U'') 1x1 - 2 U' I + 2 mu_bar I_dev - (4/3J) sym(dev(mu b_bar) x 1)
TB2_DEV void simo_moduli(const MatConst& m, double J, const double (&b_bar)[6], double (&c)[6][6])
{
    const double du = simo_dU(m, J), ddu = simo_ddU(m, J);
    const double mu_bar = m.mu * sym_trace(b_bar) / (J * 3.0);
    double s[6];
#pragma unroll
    for (int I = 0; I < 6; I++) s[I] = m.mu * b_bar[I];
    sym_dev(s);
    const double k43 = 4.0 / (J * 3.0);
#pragma unroll
    for (int A = 0; A < 6; A++)
#pragma unroll
        for (int B = 0; B < 6; B++) {
            const double oA = A < 3 ? 1.0 : 0.0, oB = B < 3 ? 1.0 : 0.0;
            const double IxI = oA * oB;
            const double I4 = (A == B) ? (A < 3 ? 1.0 : 0.5) : 0.0;
            const double Dev = I4 - IxI * (1.0 / 3.0);
            const double symo = 0.5 * (s[A] * oB + oA * s[B]);
            c[A][B] = (du + J * ddu) * IxI - 2.0 * du * I4 + 2.0 * mu_bar * Dev - k43 * symo;
        }
}
// SimoIso3D::s_ij (SimoIso3D.cpp:36-53): b = F F^T, b_bar = J^(-2/3) b
TB2_DEV void simo_bbar(const double (&F)[3][3], double J, double (&b_bar)[6])
{
    sym_fft(F, b_bar);
    const double r = rcbrt(J), sc = r * r;
#pragma unroll
    for (int I = 0; I < 6; I++) b_bar[I] *= sc;
}

// ---- J2 hardening K(alpha), K'(alpha) (J2_C0HardeningT.h:68-69): LinearT, LinearExponentialT, PowerLawT (PowerLawT.cpp:28-37),
// CubicSplineT (CubicSplineT.cpp:162-182 with dRangeArrayT::Range, dRangeArrayT.cpp:66-87)
TB2_DEV const double* j2_spline_row(const MatConst& m, double x, double& dx)
{
    const double* kx = m.spline;
    int i = 0;
    if (!(x < __ldg(kx))) {
        int lower = 0, upper = m.nknots;
        do {
            const int dex = (lower + upper) / 2;
            if (x > __ldg(kx + dex)) lower = dex;
            else upper = dex;
        } while (upper > lower + 1);
        i = upper;
    }
    dx = x - __ldg(kx + (i == 0 ? 0 : i - 1));
    return kx + m.nknots + 4 * i;
}
TB2_DEV double j2_K(const MatConst& m, double a)
{
    if (m.hard_kind == kHardLinear) return m.hard[0] * a + m.hard[1];
    if (m.hard_kind == kHardPowerLaw) return m.hard[0] * pow(m.hard[1] + m.hard[2] * a, m.hard[3]);
    if (m.hard_kind == kHardCubicSpline) {
        double dx;
        const double* c = j2_spline_row(m, a, dx);
        return __ldg(c) + __ldg(c + 1) * dx + __ldg(c + 2) * dx * dx + __ldg(c + 3) * dx * dx * dx;
    }
    return m.hard[0] + m.hard[1] * a + m.hard[2] * (1.0 - exp(-a / m.hard[3]));
}
TB2_DEV double j2_dK(const MatConst& m, double a)
{
    if (m.hard_kind == kHardLinear) return m.hard[0];
    if (m.hard_kind == kHardPowerLaw) return m.hard[0] * m.hard[2] * m.hard[3] * pow(m.hard[1] + m.hard[2] * a, m.hard[3] - 1.0);
    if (m.hard_kind == kHardCubicSpline) {
        double dx;
        const double* c = j2_spline_row(m, a, dx);
        return __ldg(c + 1) + 2.0 * __ldg(c + 2) * dx + 3.0 * __ldg(c + 3) * dx * dx;
    }
    return m.hard[1] + m.hard[2] * exp(-a / m.hard[3]) / m.hard[3];
}

TB2_DEV double& hist(const J2Hist& h, int64_t e, int ip, int field) { return h.data[((int64_t)(field * 8 + ip)) * h.stride + e]; }
TB2_DEV void hist_load6(const J2Hist& h, int64_t e, int ip, int field, double (&v)[6])
{
#pragma unroll
    for (int I = 0; I < 6; I++) v[I] = hist(h, e, ip, field + I);
}
TB2_DEV void hist_store6(const J2Hist& h, int64_t e, int ip, int field, const double (&v)[6])
{
#pragma unroll
    for (int I = 0; I < 6; I++) hist(h, e, ip, field + I) = v[I];
}

// J2SimoC0HardeningT::TrialElasticState on an allocated element (J2SimoC0HardeningT.cpp:42-88, InitIntermediate :409-426)
TB2_DEV void j2_trial_allocated(const J2Hist& h, int64_t e, int ip, const double (&F)[3][3], const double (&frel)[3][3], double J,
                                double (&b_tr)[6], double (&beta_tr)[6], double& trace_beta_tr)
{
    double b_bar[6], beta_bar[6];
    int& flag = h.flag[(int64_t)ip * h.stride + e];
    if (flag == kJ2NotInit) {
        double ad[3][3], Fn[3][3];
        const double d = adj3(frel, ad);
        scale3(ad, 1.0 / d);
        mul3(ad, F, Fn);
        sym_fft(Fn, b_bar);
        const double sc = rcbrt(sym_det(b_bar));
#pragma unroll
        for (int I = 0; I < 6; I++) { b_bar[I] *= sc; beta_bar[I] = 0.0; }
        hist_store6(h, e, ip, kHBBar, b_bar);
        hist_store6(h, e, ip, kHBetaBar, beta_bar);
        flag = kJ2Elastic;
    } else {
        hist_load6(h, e, ip, kHBBar, b_bar);
        hist_load6(h, e, ip, kHBetaBar, beta_bar);
    }
    double fbar[3][3];
    const double sc = rcbrt(det3(frel));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) fbar[i][j] = sc * frel[i][j];
    sym_qsqt(fbar, b_bar, b_tr);
    sym_qsqt(fbar, beta_bar, beta_tr);
    trace_beta_tr = sym_trace(beta_tr);
    const double t3 = trace_beta_tr / 3.0;
    beta_tr[0] -= t3; beta_tr[1] -= t3; beta_tr[2] -= t3;
    hist(h, e, ip, kHInternal + kDetFTot) = J;
    hist_store6(h, e, ip, kHBBarTrial, b_tr);
    hist_store6(h, e, ip, kHBetaBarTrial, beta_tr);
}

// J2Simo3D::s_ij / c_ijkl (J2Simo3D.cpp:41-105) + J2SimoC0HardeningT::{PlasticLoading :91-144, StressCorrection :148-253,
// ModuliCorrection :260-309, AllocateElement :312-333}.  alloc is the element's IsAllocated flag held by the calling
// thread (one thread owns the element and walks its integration points in order, as the reference does).
template <bool WITH_MODULI>
TB2_DEV int j2_eval(const MatConst& m, const J2Hist& h, int64_t e, int ip, int& alloc, int iteration, const double (&F)[3][3],
                    const double (&Fl)[3][3], double J, double (&sig)[6], double (&c)[6][6])
{
    const double mu = m.mu;
    double frel[3][3];
    {
        double ad[3][3];
        const double d = adj3(Fl, ad);
        scale3(ad, 1.0 / d);
        mul3(F, ad, frel); // J2Simo3D::ComputeGradients :253-261
    }
    double b_tr[6], beta_tr[6], trace_beta_tr = 0.0;
    if (alloc)
        j2_trial_allocated(h, e, ip, F, frel, J, b_tr, beta_tr, trace_beta_tr);
    else {
        simo_bbar(F, J, b_tr);
#pragma unroll
        for (int I = 0; I < 6; I++) beta_tr[I] = 0.0;
    }
    simo_cauchy(m, J, b_tr, sig);
    if (iteration > -1) { // the first iteration of a step is elastic (J2Simo3D.cpp:83-84)
        if (!alloc) {
            double rel[6];
#pragma unroll
            for (int I = 0; I < 6; I++) rel[I] = b_tr[I];
            sym_dev(rel);
#pragma unroll
            for (int I = 0; I < 6; I++) rel[I] = mu * rel[I];
            const double f = sqrt(sym_norm2(rel)) - TB2_SQRT23 * j2_K(m, 0.0);
            if (f > TB2_YIELD_TOL) {
                // AllocateElement: every integration point of the element, data = 0, flags = kNotInit
                alloc = 1;
                h.alloc[e] = 1;
                for (int q = 0; q < 8; q++) {
                    for (int fld = 0; fld < kHNumDouble; fld++) hist(h, e, q, fld) = 0.0;
                    h.flag[(int64_t)q * h.stride + e] = kJ2NotInit;
                }
                j2_trial_allocated(h, e, ip, F, frel, J, b_tr, beta_tr, trace_beta_tr);
            }
        }
        if (alloc) {
            double rel[6];
#pragma unroll
            for (int I = 0; I < 6; I++) rel[I] = b_tr[I];
            sym_dev(rel);
#pragma unroll
            for (int I = 0; I < 6; I++) rel[I] = mu * rel[I] - beta_tr[I];
            const double alpha = hist(h, e, ip, kHInternal + kAlpha);
            const double stressnorm = sqrt(sym_norm2(rel));
            const double ftrial = stressnorm - TB2_SQRT23 * j2_K(m, alpha);
            const double mu_bar = mu * sym_trace(b_tr) / 3.0;
            const double mbb = mu_bar - trace_beta_tr / 3.0;
            double n[6];
#pragma unroll
            for (int I = 0; I < 6; I++) n[I] = rel[I] / stressnorm;
            hist(h, e, ip, kHInternal + kStressNorm) = stressnorm;
            hist(h, e, ip, kHInternal + kFTrial) = ftrial;
            hist(h, e, ip, kHInternal + kMuBar) = mu_bar;
            hist(h, e, ip, kHInternal + kMuBarBar) = mbb;
            hist_store6(h, e, ip, kHUnitNorm, n);
            double heat = 0.0;
            if (ftrial > TB2_YIELD_TOL) {
                h.flag[(int64_t)ip * h.stride + e] = kJ2Plastic;
                double dgamma;
                if (m.hard_kind == kHardLinear)
                    dgamma = ftrial / (2.0 * mbb) / (1.0 + (j2_dK(m, alpha) / 3.0 / mbb));
                else {
                    const double x_tr = ftrial + TB2_SQRT23 * j2_K(m, alpha);
                    double f_hat = -ftrial;
                    const double k = 2.0 * mbb;
                    dgamma = 0.0;
                    int count = 0;
                    const int max_iteration = 15;
                    while (fabs(f_hat) > TB2_YIELD_TOL && ++count <= max_iteration) {
                        const double df_hat = 2.0 * j2_dK(m, alpha + TB2_SQRT23 * dgamma) / 3.0 + k;
                        if (df_hat < 1.0e-12) return kErrJ2Local;
                        dgamma -= f_hat / df_hat;
                        f_hat = TB2_SQRT23 * j2_K(m, alpha + TB2_SQRT23 * dgamma) - x_tr + k * dgamma;
                    }
                    if (count == max_iteration) return kErrJ2Local;
                }
                hist(h, e, ip, kHInternal + kDGamma) = dgamma;
                const double k2 = -2.0 * mbb * dgamma / J;
#pragma unroll
                for (int I = 0; I < 6; I++) sig[I] += k2 * n[I];
                heat = 0.9 * dgamma * j2_K(m, alpha + TB2_SQRT23 * dgamma) / J;
            } else
                h.flag[(int64_t)ip * h.stride + e] = kJ2Elastic;
            hist(h, e, ip, kHInternal + kHeatIncr) = heat;
        }
    }
    if (WITH_MODULI) {
        simo_moduli(m, J, b_tr, c);
        if (alloc && h.flag[(int64_t)ip * h.stride + e] == kJ2Plastic) { // ModuliCorrection
            double n[6];
            hist_load6(h, e, ip, kHUnitNorm, n);
            const double stressnorm = hist(h, e, ip, kHInternal + kStressNorm), dgamma = hist(h, e, ip, kHInternal + kDGamma);
            const double alpha = hist(h, e, ip, kHInternal + kAlpha), mb = hist(h, e, ip, kHInternal + kMuBar);
            const double mbb = hist(h, e, ip, kHInternal + kMuBarBar), detF = hist(h, e, ip, kHInternal + kDetFTot);
            const double f0 = 2.0 * mb * dgamma / stressnorm;
            const double d0 = 1.0 + j2_dK(m, alpha) / 3.0 / mbb;
            const double f1 = 1.0 / d0 - f0;
            const double d1 = 2.0 * mbb * f1 - (4.0 / 3.0) * dgamma * (1.0 / d0 - 1.0);
            const double d2 = 2.0 * stressnorm * f1;
            double N[3][3], NN[3][3], nn2[6];
            sym_to_mat(n, N);
            mul3(N, N, NN);
            nn2[0] = NN[0][0]; nn2[1] = NN[1][1]; nn2[2] = NN[2][2]; nn2[3] = NN[1][2]; nn2[4] = NN[0][2]; nn2[5] = NN[0][1];
            sym_dev(nn2);
            const double rdet = 1.0 / detF;
#pragma unroll
            for (int A = 0; A < 6; A++)
#pragma unroll
                for (int B = 0; B < 6; B++) {
                    const double oA = A < 3 ? 1.0 : 0.0, oB = B < 3 ? 1.0 : 0.0;
                    const double I4 = (A == B) ? (A < 3 ? 1.0 : 0.5) : 0.0;
                    const double Dev = I4 - oA * oB * (1.0 / 3.0);
                    const double corr = -2.0 * mbb * f0 * Dev + f0 * (4.0 / 3.0) * stressnorm * 0.5 * (n[A] * oB + oA * n[B]) -
                                        d1 * n[A] * n[B] - d2 * n[A] * nn2[B];
                    c[A][B] += corr * rdet;
                }
        }
    }
    return kErrNone;
}

} // namespace tb2
