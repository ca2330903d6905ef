/* oracle/tahoe_oracle.c -- TEST INFRASTRUCTURE ONLY (see tahoe_oracle.h).
 *
 * Plain-C restatement of the reference algorithm for the Hex8 hot path.
 * Every function cites the reference file:line (relative to /root/reference)
 * it follows, including loop / summation order where the reference fixes one.
 * Written for clarity, not speed: this is the checker, never the product.
 */
#include "tahoe_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SQRT23 0.81649658092772603273 /* sqrt(2/3), J2SimoC0HardeningT.cpp:11 */
static const double kYieldTol = 1.0e-10; /* J2SimoC0HardeningT.cpp:12 */

/* ------------------------------------------------------------------ */
/* small dense helpers (column-major 3x3: A[i+3j])                     */
/* ------------------------------------------------------------------ */
static double det3(const double* A)
{
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[3] * (A[1] * A[8] - A[2] * A[7]) + A[6] * (A[1] * A[5] - A[2] * A[4]);
}
static void inv3(const double* A, double det, double* B)
{
    double r = 1.0 / det;
    B[0] = (A[4] * A[8] - A[5] * A[7]) * r;
    B[1] = -(A[1] * A[8] - A[2] * A[7]) * r;
    B[2] = (A[1] * A[5] - A[2] * A[4]) * r;
    B[3] = -(A[3] * A[8] - A[5] * A[6]) * r;
    B[4] = (A[0] * A[8] - A[2] * A[6]) * r;
    B[5] = -(A[0] * A[5] - A[2] * A[3]) * r;
    B[6] = (A[3] * A[7] - A[4] * A[6]) * r;
    B[7] = -(A[0] * A[7] - A[1] * A[6]) * r;
    B[8] = (A[0] * A[4] - A[1] * A[3]) * r;
}
static void mul3(const double* A, const double* B, double* C) /* C = A B */
{
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += A[i + 3 * k] * B[k + 3 * j];
            C[i + 3 * j] = s;
        }
}
/* symmetric (11,22,33,23,13,12) <-> full */
static const int VI[6] = {0, 1, 2, 1, 0, 0};
static const int VJ[6] = {0, 1, 2, 2, 2, 1};
static void sym_to_mat(const double* s, double* A)
{
    A[0] = s[0]; A[4] = s[1]; A[8] = s[2];
    A[5] = A[7] = s[3]; A[2] = A[6] = s[4]; A[1] = A[3] = s[5];
}
static double sym_det(const double* s)
{
    double A[9];
    sym_to_mat(s, A);
    return det3(A);
}
static double sym_trace(const double* s) { return s[0] + s[1] + s[2]; }
static void sym_dev(double* s)
{
    double p = sym_trace(s) / 3.0;
    s[0] -= p; s[1] -= p; s[2] -= p;
}
static double sym_scalar_product(const double* s) /* dSymMatrixT::ScalarProduct: s:s */
{
    return s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
}
/* B = Q A Q^T with A symmetric (dSymMatrixT::MultQBQT) */
static void sym_QAQT(const double* Q, const double* a, double* b)
{
    double A[9], T[9], R[9];
    sym_to_mat(a, A);
    mul3(Q, A, T);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += T[i + 3 * k] * Q[j + 3 * k];
            R[i + 3 * j] = s;
        }
    for (int I = 0; I < 6; I++) b[I] = R[VI[I] + 3 * VJ[I]];
}
/* b = F F^T, FSSolidMatT::Compute_b (FSSolidMatT.cpp:402-434) */
static void compute_b(const double* f, double* a)
{
    a[0] = f[0] * f[0] + f[3] * f[3] + f[6] * f[6];
    a[1] = f[1] * f[1] + f[4] * f[4] + f[7] * f[7];
    a[2] = f[2] * f[2] + f[5] * f[5] + f[8] * f[8];
    a[3] = f[1] * f[2] + f[4] * f[5] + f[7] * f[8];
    a[4] = f[0] * f[2] + f[3] * f[5] + f[6] * f[8];
    a[5] = f[0] * f[1] + f[3] * f[4] + f[6] * f[7];
}

/* ------------------------------------------------------------------ */
/* materials                                                           */
/* ------------------------------------------------------------------ */
/* IsotropicT::Set_E_nu (materials/primitives/IsotropicT.cpp:32-45) */
void orc_material_from_E_nu(orc_material_t* m, int kind, double E, double nu, double density)
{
    memset(m, 0, sizeof(*m));
    m->kind = kind;
    m->mu = 0.5 * E / (1.0 + nu);
    m->lambda = 2.0 * m->mu * nu / (1.0 - 2.0 * nu);
    m->kappa = m->lambda + 2.0 / 3.0 * m->mu;
    m->density = density;
}

/* IsotropicT::ComputeModuli (IsotropicT.cpp:153-169): reduced-index C, row-major/symmetric here */
static void hooke_moduli(const orc_material_t* m, double C[6][6])
{
    memset(C, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) C[i][j] = m->lambda;
        C[i][i] = m->lambda + 2.0 * m->mu;
        C[i + 3][i + 3] = m->mu;
    }
}
/* HookeanMatT::HookeanStress (Hookean/HookeanMatT.cpp:103-108) -> dSymMatrixT::A_ijkl_B_kl
 * (dSymMatrixT.cpp:856-887): explicit factor 2 on the shear columns */
static void hooke_stress(const orc_material_t* m, const double* e, double* s)
{
    double C[6][6];
    hooke_moduli(m, C);
    for (int I = 0; I < 6; I++)
        s[I] = C[I][0] * e[0] + C[I][1] * e[1] + C[I][2] * e[2] + 2.0 * (C[I][3] * e[3] + C[I][4] * e[4] + C[I][5] * e[5]);
}

/* FDKStV / FDHookeanMatT::s_ij (Hookean/FDHookeanMatT.cpp:39-55):
 * E = (F^T F - 1)/2 (FSSolidMatT::Compute_E, FSSolidMatT.cpp:472-505), S = C:E, sigma = F S F^T / J */
static void fdkstv_stress(const orc_material_t* m, const double* F, double* sig)
{
    const double* f = F;
    double e[6], S[6];
    e[0] = (f[0] * f[0] + f[1] * f[1] + f[2] * f[2] - 1.0) * 0.5;
    e[1] = (f[3] * f[3] + f[4] * f[4] + f[5] * f[5] - 1.0) * 0.5;
    e[2] = (f[6] * f[6] + f[7] * f[7] + f[8] * f[8] - 1.0) * 0.5;
    e[3] = (f[3] * f[6] + f[4] * f[7] + f[5] * f[8]) * 0.5;
    e[4] = (f[0] * f[6] + f[1] * f[7] + f[2] * f[8]) * 0.5;
    e[5] = (f[0] * f[3] + f[1] * f[4] + f[2] * f[5]) * 0.5;
    hooke_stress(m, e, S);
    sym_QAQT(F, S, sig);
    double J = det3(F);
    for (int I = 0; I < 6; I++) sig[I] /= J;
}
/* FDHookeanMatT::c_ijkl (FDHookeanMatT.cpp:30-37) + TensorTransformT::FFFFC_3D
 * (primitives/TensorTransformT.cpp:102-153): c_ijkl = F_iI F_jJ F_kK F_lL C_IJKL / J */
static void fdkstv_moduli(const orc_material_t* m, const double* F, double c[6][6])
{
    double C[6][6];
    hooke_moduli(m, C);
    /* full 4th-order tensor from the reduced form (minor symmetries) */
    static const int V[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
    double J = det3(F);
    for (int A = 0; A < 6; A++)
        for (int B = 0; B < 6; B++) {
            int i = VI[A], j = VJ[A], k = VI[B], l = VJ[B];
            double s = 0.0;
            for (int I = 0; I < 3; I++)
                for (int Jj = 0; Jj < 3; Jj++)
                    for (int K = 0; K < 3; K++)
                        for (int L = 0; L < 3; L++)
                            s += F[i + 3 * I] * F[j + 3 * Jj] * F[k + 3 * K] * F[l + 3 * L] * C[V[I][Jj]][V[K][L]];
            c[A][B] = s / J;
        }
}

/* SimoIso3D::dU, ddU (materials/Simo/SimoIso3D.h:93-101) */
static double simo_dU(const orc_material_t* m, double J) { return 0.5 * m->kappa * (J - 1.0 / J); }
static double simo_ddU(const orc_material_t* m, double J) { return 0.5 * m->kappa * (1.0 + 1.0 / (J * J)); }

/* SimoIso3D::ComputeCauchy (SimoIso3D.cpp:127-136) */
static void simo_cauchy(const orc_material_t* m, double J, const double* b_bar, double* sig)
{
    for (int I = 0; I < 6; I++) sig[I] = m->mu / J * b_bar[I];
    sym_dev(sig);
    double p = simo_dU(m, J);
    sig[0] += p; sig[1] += p; sig[2] += p;
}
/* SimoIso3D::ComputeModuli (SimoIso3D.cpp:102-125) with the fixed forms of
 * TakeParameterList :74-97 (ReducedIndexI / ReducedIndexDeviatoric, dMatrixT.cpp:328-385) */
static void simo_moduli(const orc_material_t* m, double J, const double* b_bar, double c[6][6])
{
    double du = simo_dU(m, J), ddu = simo_ddU(m, J);
    double mu_bar = m->mu * sym_trace(b_bar) / (J * 3.0);
    double s[6];
    for (int I = 0; I < 6; I++) s[I] = m->mu * b_bar[I];
    sym_dev(s);
    static const double one[6] = {1, 1, 1, 0, 0, 0};
    for (int A = 0; A < 6; A++)
        for (int B = 0; B < 6; B++) {
            double IxI = one[A] * one[B];
            double I4 = (A == B) ? (A < 3 ? 1.0 : 0.5) : 0.0;
            double Dev = I4 - IxI / 3.0;
            double symo = 0.5 * (s[A] * one[B] + one[A] * s[B]); /* Outer + dMatrixT::Symmetrize */
            c[A][B] = (du + J * ddu) * IxI - 2.0 * du * I4 + 2.0 * mu_bar * Dev - 4.0 / (J * 3.0) * symo;
        }
}
/* SimoIso3D::s_ij (SimoIso3D.cpp:36-53) */
static int simo_eval(const orc_material_t* m, const double* F, double* sig, double c[6][6])
{
    double b[6], b_bar[6];
    compute_b(F, b);
    double J = sym_det(b);
    if (J <= 0.0) return ORC_BAD_JACOBIAN;
    J = sqrt(J);
    double sc = pow(J, -2.0 / 3.0);
    for (int I = 0; I < 6; I++) b_bar[I] = sc * b[I];
    if (sig) simo_cauchy(m, J, b_bar, sig);
    if (c) simo_moduli(m, J, b_bar, c);
    return ORC_OK;
}

/* CubicSplineT (toolbox/src/C1functions/CubicSplineT.cpp): natural / parabolic-run-out cubic spline through the knots, stored as
 * nknots+1 rows of (a0..a3) about the left knot of each interval; rows 0 and nknots extend the curve beyond the ends (:303-323) */
int orc_material_set_spline(orc_material_t* m, int n, const double* x, const double* y, int fixity)
{
    if (n < 3 || n > ORC_MAX_KNOTS || (fixity != 0 && fixity != 1)) return -1;
    double dxi[ORC_MAX_KNOTS], ddy[ORC_MAX_KNOTS], L[ORC_MAX_KNOTS], D[ORC_MAX_KNOTS], R[ORC_MAX_KNOTS];
    for (int i = 1; i < n; i++) dxi[i - 1] = x[i] - x[i - 1];
    const int neq = n - 2;
    double* rhs = ddy + 1;
    for (int i = 0; i < neq; i++) { /* :271-277 */
        L[i] = dxi[i] / 6.0;
        D[i] = (dxi[i] + dxi[i + 1]) / 3.0;
        R[i] = dxi[i + 1] / 6.0;
        rhs[i] = ((y[i + 2] - y[i + 1]) / dxi[i + 1]) - ((y[i + 1] - y[i]) / dxi[i]);
    }
    if (fixity == 0) { /* kParabolic :280-284 */
        D[0] += dxi[0] / 6.0;
        D[neq - 1] += dxi[neq - 1] / 6.0;
    }
    /* TriDiagdMatrixT::LinearSolve (TriDiagdMatrixT.cpp:25-80) */
    for (int i = 1; i < neq; i++) {
        const double factor = L[i] / D[i - 1];
        D[i] -= R[i - 1] * factor;
        rhs[i] -= rhs[i - 1] * factor;
    }
    rhs[neq - 1] /= D[neq - 1];
    for (int i = neq - 2; i >= 0; i--) rhs[i] = (rhs[i] - R[i] * rhs[i + 1]) / D[i];
    if (fixity == 1) ddy[0] = ddy[neq + 1] = 0.0;
    else {
        ddy[0] = ddy[1];
        ddy[neq + 1] = ddy[neq];
    }
    m->nknots = n;
    for (int i = 0; i < n; i++) m->knot_x[i] = x[i];
    double* c = m->spline;
    for (int j = 1; j < n; j++) { /* :303-313 */
        const int i = j - 1;
        const double dx = dxi[i];
        c[j * 4 + 0] = y[i];
        c[j * 4 + 1] = -dx * (2.0 * ddy[i] + ddy[i + 1]) / 6.0 + (y[i + 1] - y[i]) / dx;
        c[j * 4 + 2] = ddy[i] / 2.0;
        c[j * 4 + 3] = (ddy[i + 1] - ddy[i]) / (6.0 * dx);
    }
    c[0] = c[4]; c[1] = c[5]; c[2] = c[6]; c[3] = 0.0; /* extensions :316-323 */
    const int dex = n - 1;
    c[(dex + 1) * 4 + 0] = y[dex];
    c[(dex + 1) * 4 + 1] = dxi[dex - 1] * (ddy[dex - 1] + 2.0 * ddy[dex]) / 6.0 + (y[dex] - y[dex - 1]) / dxi[dex - 1];
    c[(dex + 1) * 4 + 2] = ddy[dex] / 2.0;
    c[(dex + 1) * 4 + 3] = 0.0;
    return 0;
}

/* dRangeArrayT::Range (toolbox/src/abc/other/dRangeArrayT.cpp:66-87) and the interval-local abscissa of CubicSplineT::function */
static const double* spline_row(const orc_material_t* m, double x, double* dx)
{
    int i = 0;
    if (!(x < m->knot_x[0])) {
        int lower = 0, upper = m->nknots;
        do {
            const int dex = (lower + upper) / 2;
            if (x > m->knot_x[dex]) lower = dex;
            else upper = dex;
        } while (upper > lower + 1);
        i = upper;
    }
    *dx = (i == 0) ? x - m->knot_x[0] : x - m->knot_x[i - 1];
    return m->spline + 4 * i;
}

/* J2 hardening K(alpha), K'(alpha) (J2_C0HardeningT.h:68-69; C1functions/LinearT.h:71,
 * LinearExponentialT.cpp:48-57, PowerLawT.cpp:28-37, CubicSplineT.cpp:162-182) */
static double j2_K(const orc_material_t* m, double a)
{
    if (m->hard_kind == ORC_HARD_LINEAR) return m->hard[0] * a + m->hard[1];
    if (m->hard_kind == ORC_HARD_POWER_LAW) return m->hard[0] * pow(m->hard[1] + m->hard[2] * a, m->hard[3]);
    if (m->hard_kind == ORC_HARD_CUBIC_SPLINE) {
        double dx;
        const double* c = spline_row(m, a, &dx);
        return c[0] + c[1] * dx + c[2] * dx * dx + c[3] * dx * dx * dx;
    }
    return m->hard[0] + m->hard[1] * a + m->hard[2] * (1.0 - exp(-a / m->hard[3]));
}
static double j2_dK(const orc_material_t* m, double a)
{
    if (m->hard_kind == ORC_HARD_LINEAR) return m->hard[0];
    if (m->hard_kind == ORC_HARD_POWER_LAW) return m->hard[0] * m->hard[2] * m->hard[3] * pow(m->hard[1] + m->hard[2] * a, m->hard[3] - 1.0);
    if (m->hard_kind == ORC_HARD_CUBIC_SPLINE) {
        double dx;
        const double* c = spline_row(m, a, &dx);
        return c[1] + 2.0 * c[2] * dx + 3.0 * c[3] * dx * dx;
    }
    return m->hard[1] + m->hard[2] * exp(-a / m->hard[3]) / m->hard[3];
}
void orc_hardening(const orc_material_t* m, double alpha, double* K, double* dK)
{
    *K = j2_K(m, alpha);
    *dK = j2_dK(m, alpha);
}
enum { kalpha = 0, kstressnorm = 1, kdgamma = 2, kftrial = 3, kmu_bar = 4, kmu_bar_bar = 5, kDetF_tot = 6, kHeatIncr = 7 };

/* J2Simo3D::s_ij / c_ijkl (plasticity_J2/J2Simo3D.cpp:41-105) with
 * J2SimoC0HardeningT::{TrialElasticState :42-88, PlasticLoading :91-144, StressCorrection :148-253,
 * ModuliCorrection :260-309, InitIntermediate :409-426}.  dH = 0 (J2_C0HardeningT.h:50-51). */
static int j2_eval(const orc_material_t* m, const double* F, const double* Flast, orc_j2_ip_t* ips, int ip,
                   int* alloc, int iteration, double* sig, double c[6][6])
{
    orc_j2_ip_t* st = &ips[ip];
    double mu = m->mu;
    double Flinv[9], frel[9];
    inv3(Flast, det3(Flast), Flinv);
    mul3(F, Flinv, frel); /* J2Simo3D::ComputeGradients :253-261 */
    double J = det3(F);
    double b_tr[6], beta_tr[6], trace_beta_tr = 0.0;

    /* TrialElasticState */
    for (int pass = 0; pass < 2; pass++) {
        if (*alloc) {
            if (st->flag == ORC_J2_NOTINIT) { /* InitIntermediate */
                double frinv[9], Fn[9];
                inv3(frel, det3(frel), frinv);
                mul3(frinv, F, Fn);
                compute_b(Fn, st->b_bar);
                double sc = pow(sym_det(st->b_bar), -1.0 / 3.0);
                for (int I = 0; I < 6; I++) { st->b_bar[I] *= sc; st->beta_bar[I] = 0.0; }
                st->flag = ORC_J2_ELASTIC;
            }
            double fbar[9], sc = pow(det3(frel), -1.0 / 3.0);
            for (int i = 0; i < 9; i++) fbar[i] = sc * frel[i];
            sym_QAQT(fbar, st->b_bar, b_tr);
            sym_QAQT(fbar, st->beta_bar, beta_tr);
            trace_beta_tr = sym_trace(beta_tr);
            beta_tr[0] -= trace_beta_tr / 3.0; beta_tr[1] -= trace_beta_tr / 3.0; beta_tr[2] -= trace_beta_tr / 3.0;
            st->internal[kDetF_tot] = J;
            memcpy(st->b_bar_trial, b_tr, sizeof b_tr);
            memcpy(st->beta_bar_trial, beta_tr, sizeof beta_tr);
        } else {
            compute_b(F, b_tr);
            double sc = pow(J, -2.0 / 3.0);
            for (int I = 0; I < 6; I++) { b_tr[I] *= sc; beta_tr[I] = 0.0; }
            trace_beta_tr = 0.0;
        }
        if (pass == 1 || !sig) break;

        /* s_ij: elastic stress then return map */
        simo_cauchy(m, J, b_tr, sig);
        if (!(iteration > -1)) break; /* 1st iteration is elastic, J2Simo3D.cpp:83-84 */

        /* PlasticLoading */
        double rel[6];
        memcpy(rel, b_tr, sizeof rel);
        sym_dev(rel);
        for (int I = 0; I < 6; I++) rel[I] = mu * rel[I] - beta_tr[I];
        if (!*alloc) {
            double f = sqrt(sym_scalar_product(rel)) - SQRT23 * j2_K(m, 0.0);
            if (!(f > kYieldTol)) break;
            /* AllocateElement :312-333: all 8 IPs, flags = kNotInit, data = 0 */
            *alloc = 1;
            for (int q = 0; q < 8; q++) { memset(&ips[q], 0, sizeof(orc_j2_ip_t)); ips[q].flag = ORC_J2_NOTINIT; }
            continue; /* redo trial state on the allocated element, then PlasticLoading again */
        }
        break;
    }
    if (sig && iteration > -1 && *alloc) {
        /* PlasticLoading on an allocated element :101-143 */
        double rel[6];
        memcpy(rel, b_tr, sizeof rel);
        sym_dev(rel);
        for (int I = 0; I < 6; I++) rel[I] = mu * rel[I] - beta_tr[I];
        double* in = st->internal;
        in[kstressnorm] = sqrt(sym_scalar_product(rel));
        in[kftrial] = in[kstressnorm] - SQRT23 * j2_K(m, in[kalpha]);
        in[kmu_bar] = mu * sym_trace(b_tr) / 3.0;
        in[kmu_bar_bar] = in[kmu_bar] - trace_beta_tr / 3.0;
        in[kHeatIncr] = 0.0;
        for (int I = 0; I < 6; I++) st->unit_norm[I] = rel[I] / in[kstressnorm];
        if (in[kftrial] > kYieldTol) {
            st->flag = ORC_J2_PLASTIC;
            /* StressCorrection :148-253 */
            double alpha = in[kalpha], mbb = in[kmu_bar_bar], dgamma;
            if (m->hard_kind == ORC_HARD_LINEAR)
                dgamma = in[kftrial] / (2.0 * mbb) / (1.0 + (j2_dK(m, alpha) / 3.0 / mbb));
            else {
                double x_tr = in[kftrial] + SQRT23 * j2_K(m, alpha);
                double f_hat = -in[kftrial];
                double k = 2.0 * mbb;
                dgamma = 0.0;
                int count = 0, max_iteration = 15;
                while (fabs(f_hat) > kYieldTol && ++count <= max_iteration) {
                    double df_hat = 2.0 * j2_dK(m, alpha + SQRT23 * dgamma) / 3.0 + k;
                    if (df_hat < 1.0e-12) return ORC_J2_LOCAL_FAIL;
                    dgamma -= f_hat / df_hat;
                    f_hat = SQRT23 * j2_K(m, alpha + SQRT23 * dgamma) - x_tr + k * dgamma;
                }
                if (count == max_iteration) return ORC_J2_LOCAL_FAIL;
            }
            in[kdgamma] = dgamma;
            for (int I = 0; I < 6; I++) sig[I] += -2.0 * mbb * dgamma / in[kDetF_tot] * st->unit_norm[I];
            in[kHeatIncr] = 0.9 * dgamma * j2_K(m, alpha + SQRT23 * dgamma) / in[kDetF_tot];
        } else {
            st->flag = ORC_J2_ELASTIC; /* StressCorrection is not reached: dgamma keeps its old value */
        }
    }
    if (c) {
        simo_moduli(m, J, b_tr, c);
        if (*alloc && st->flag == ORC_J2_PLASTIC) { /* ModuliCorrection :260-309 */
            const double* in = st->internal;
            const double* n = st->unit_norm;
            double stressnorm = in[kstressnorm], dgamma = in[kdgamma], alpha = in[kalpha];
            double mb = in[kmu_bar], mbb = in[kmu_bar_bar];
            double f0 = 2.0 * mb * dgamma / stressnorm;
            double d0 = 1.0 + j2_dK(m, alpha) / 3.0 / mbb;
            double f1 = 1.0 / d0 - f0;
            double d1 = 2.0 * mbb * f1 - (4.0 / 3.0) * dgamma * (1.0 / d0 - 1.0);
            double d2 = 2.0 * stressnorm * f1;
            double N[9], NN[9], nn2[6];
            sym_to_mat(n, N);
            mul3(N, N, NN);
            for (int I = 0; I < 6; I++) nn2[I] = NN[VI[I] + 3 * VJ[I]];
            sym_dev(nn2);
            static const double one[6] = {1, 1, 1, 0, 0, 0};
            for (int A = 0; A < 6; A++)
                for (int B = 0; B < 6; B++) {
                    double I4 = (A == B) ? (A < 3 ? 1.0 : 0.5) : 0.0;
                    double Dev = I4 - one[A] * one[B] / 3.0;
                    double corr = -2.0 * mbb * f0 * Dev + f0 * (4.0 / 3.0) * stressnorm * 0.5 * (n[A] * one[B] + one[A] * n[B]) -
                                  d1 * n[A] * n[B] - d2 * n[A] * nn2[B];
                    c[A][B] += corr / in[kDetF_tot];
                }
        }
    }
    return ORC_OK;
}
/* J2SimoC0HardeningT::Update (J2SimoC0HardeningT.cpp:341-384), called from J2Simo3D::UpdateHistory
 * for allocated elements only */
void orc_j2_update(const orc_material_t* m, orc_j2_ip_t* j2)
{
    for (int ip = 0; ip < 8; ip++) {
        orc_j2_ip_t* st = &j2[ip];
        memcpy(st->b_bar, st->b_bar_trial, sizeof st->b_bar);
        memcpy(st->beta_bar, st->beta_bar_trial, sizeof st->beta_bar);
        if (st->flag == ORC_J2_PLASTIC) {
            st->flag = ORC_J2_ELASTIC;
            double dgamma = st->internal[kdgamma], mbb = st->internal[kmu_bar_bar];
            double k = 2.0 * mbb * dgamma / m->mu;
            st->internal[kalpha] += SQRT23 * dgamma;
            for (int I = 0; I < 6; I++) st->b_bar[I] += -k * st->unit_norm[I];
        }
    }
}
/* J2SimoC0HardeningT::Reset (:387-407) */
void orc_j2_reset(orc_j2_ip_t* j2)
{
    for (int ip = 0; ip < 8; ip++) { j2[ip].flag = ORC_J2_ELASTIC; j2[ip].internal[kdgamma] = 0.0; }
}

/* ------------------------------------------------------------------ */
/* geometry                                                            */
/* ------------------------------------------------------------------ */
/* HexahedronT.cpp:23-25 vertex signs; :421-430 N and dN/dxi; SetLocalShape :1492-1669
 * (8 points: g = 1/sqrt3, points in node order, weights 1) */
static const double RA[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
static const double SA[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
static const double TA[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
void orc_hex8_parent(double Na[8][8], double DNa[8][3][8], double w[8])
{
    double g = 1.0 / sqrt(3.0);
    for (int ip = 0; ip < 8; ip++) {
        double r = g * RA[ip], s = g * SA[ip], t = g * TA[ip];
        w[ip] = 1.0;
        for (int a = 0; a < 8; a++) {
            double tr = 1.0 + RA[a] * r, ts = 1.0 + SA[a] * s, tt = 1.0 + TA[a] * t;
            Na[ip][a] = 0.125 * tr * ts * tt;
            DNa[ip][0][a] = 0.125 * RA[a] * ts * tt;
            DNa[ip][1][a] = 0.125 * tr * SA[a] * tt;
            DNa[ip][2][a] = 0.125 * tr * ts * TA[a];
        }
    }
}
/* ParentDomainT::Jacobian (toolbox/src/geometry/ParentDomainT.cpp:99-189, node loop a=0..7) and
 * ComputeDNa (:426-510): J_ij = sum_a X_a,i dN_a/dxi_j ; dN/dX = J^-T dN/dxi */
int orc_hex8_shape(const double X[8][3], double dNdX[8][3][8], double det[8])
{
    double Na[8][8], DNa[8][3][8], w[8];
    orc_hex8_parent(Na, DNa, w);
    for (int ip = 0; ip < 8; ip++) {
        double Jm[9] = {0}, Ji[9];
        for (int a = 0; a < 8; a++)
            for (int j = 0; j < 3; j++)
                for (int i = 0; i < 3; i++) Jm[i + 3 * j] += X[a][i] * DNa[ip][j][a];
        det[ip] = det3(Jm);
        if (det[ip] <= 0.0) return ORC_BAD_JACOBIAN;
        inv3(Jm, det[ip], Ji);
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++)
                dNdX[ip][i][a] = Ji[0 + 3 * i] * DNa[ip][0][a] + Ji[1 + 3 * i] * DNa[ip][1][a] + Ji[2 + 3 * i] * DNa[ip][2][a];
    }
    return ORC_OK;
}
/* ShapeFunctionT::GradU (ShapeFunctionT.h:384-388): G_ij = sum_a u_a,i dN_a/dX_j (column-major) */
static void grad_u(const double u[8][3], double dN[3][8], double* G)
{
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
            double s = 0.0;
            for (int a = 0; a < 8; a++) s += u[a][i] * dN[j][a];
            G[i + 3 * j] = s;
        }
}
/* f += scale * B^T sigma with Set_B of SolidElementT.cpp:854-920 */
static void add_BT_sigma(double scale, double dN[3][8], const double* s, double* fe)
{
    for (int a = 0; a < 8; a++) {
        double nx = dN[0][a], ny = dN[1][a], nz = dN[2][a];
        fe[3 * a + 0] += scale * (nx * s[0] + nz * s[4] + ny * s[5]);
        fe[3 * a + 1] += scale * (ny * s[1] + nz * s[3] + nx * s[5]);
        fe[3 * a + 2] += scale * (nz * s[2] + ny * s[3] + nx * s[4]);
    }
}
/* 6x24 B from dN (Hughes 2.8.21, SolidElementT::Set_B) */
static void set_B(double dN[3][8], double B[6][24])
{
    memset(B, 0, 6 * 24 * sizeof(double));
    for (int a = 0; a < 8; a++) {
        double nx = dN[0][a], ny = dN[1][a], nz = dN[2][a];
        B[0][3 * a + 0] = nx; B[4][3 * a + 0] = nz; B[5][3 * a + 0] = ny;
        B[1][3 * a + 1] = ny; B[3][3 * a + 1] = nz; B[5][3 * a + 1] = nx;
        B[2][3 * a + 2] = nz; B[3][3 * a + 2] = ny; B[4][3 * a + 2] = nx;
    }
}

/* mean shape-function gradient, Hughes (4.5.23): SmallStrainT::SetMeanGradient (SmallStrainT.cpp:404-421) */
static void mean_gradient(double dN[8][3][8], const double* det, double mg[3][8])
{
    double vol = 0.0;
    for (int ip = 0; ip < 8; ip++) vol += det[ip]; /* unit weights */
    memset(mg, 0, 24 * sizeof(double));
    for (int ip = 0; ip < 8; ip++)
        for (int d = 0; d < 3; d++)
            for (int a = 0; a < 8; a++) mg[d][a] += det[ip] / vol * dN[ip][d][a];
}
/* B-bar of Hughes (4.5.11-16): SolidElementT::Set_B_bar, 3D branch (SolidElementT.cpp:1003-1043) */
static void set_B_bar(double dN[3][8], double mg[3][8], double B[6][24])
{
    for (int a = 0; a < 8; a++) {
        double nx = dN[0][a], ny = dN[1][a], nz = dN[2][a];
        double fx = (mg[0][a] - nx) / 3.0, fy = (mg[1][a] - ny) / 3.0, fz = (mg[2][a] - nz) / 3.0;
        B[0][3 * a + 0] = nx + fx; B[1][3 * a + 0] = fx; B[2][3 * a + 0] = fx; B[3][3 * a + 0] = 0.0; B[4][3 * a + 0] = nz; B[5][3 * a + 0] = ny;
        B[0][3 * a + 1] = fy; B[1][3 * a + 1] = ny + fy; B[2][3 * a + 1] = fy; B[3][3 * a + 1] = nz; B[4][3 * a + 1] = 0.0; B[5][3 * a + 1] = nx;
        B[0][3 * a + 2] = fz; B[1][3 * a + 2] = fz; B[2][3 * a + 2] = nz + fz; B[3][3 * a + 2] = ny; B[4][3 * a + 2] = nx; B[5][3 * a + 2] = 0.0;
    }
}

/* finite-strain stress / modulus dispatch: FSSolidMatT s_ij, c_ijkl */
static int fs_material(const orc_material_t* m, const double* F, const double* Fl, orc_j2_ip_t* j2, int ip, int* alloc,
                       int iteration, double* sig, double c[6][6])
{
    switch (m->kind) {
    case ORC_FDKSTV:
        if (sig) fdkstv_stress(m, F, sig);
        if (c) fdkstv_moduli(m, F, c);
        return ORC_OK;
    case ORC_SIMO_ISO: return simo_eval(m, F, sig, c);
    case ORC_J2_SIMO: return j2_eval(m, F, Fl, j2, ip, alloc, iteration, sig, c);
    }
    return ORC_BAD_JACOBIAN;
}

/* K1. SmallStrainT::SetGlobalShape/FormKd (SmallStrainT.cpp:327-401,255-282);
 * FiniteStrainT::SetGlobalShape (FiniteStrainT.cpp:267-304); TotalLagrangianT::FormKd
 * (TotalLagrangianT.cpp:107-144); UpdatedLagrangianT::SetGlobalShape/FormKd (:83-91,145-171). */
int orc_element_force(int form, const orc_material_t* m, const double X[8][3], const double u[8][3],
                      const double u_last[8][3], orc_j2_ip_t* j2, int* alloc, int iteration, double fe[24])
{
    double dN[8][3][8], det[8], dNc[8][3][8], detc[8];
    static const double zero[8][3] = {{0}};
    int dummy_alloc = 0;
    if (!alloc) alloc = &dummy_alloc;
    if (!u_last) u_last = zero;
    memset(fe, 0, 24 * sizeof(double));
    int err = orc_hex8_shape(X, dN, det);
    if (err) return err;
    if (form == ORC_UPDATED_LAGRANGIAN) {
        double x[8][3];
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) x[a][i] = X[a][i] + u[a][i];
        err = orc_hex8_shape(x, dNc, detc);
        if (err) return err;
    }
    if (form == ORC_SMALL_STRAIN_BBAR) { /* SmallStrainT::SetGlobalShape B-bar branch (:340-374) + FormKd (:255-282) */
        double mg[3][8], B[6][24];
        mean_gradient(dN, det, mg);
        for (int ip = 0; ip < 8; ip++) {
            double e[6], sig[6];
            set_B_bar(dN[ip], mg, B);
            for (int I = 0; I < 6; I++) { /* fB.Multx(u) then ScaleOffDiagonal(0.5) */
                double t = 0.0;
                for (int a = 0; a < 8; a++)
                    for (int i = 0; i < 3; i++) t += B[I][3 * a + i] * u[a][i];
                e[I] = I < 3 ? t : 0.5 * t;
            }
            hooke_stress(m, e, sig);
            for (int r = 0; r < 24; r++) { /* fB.MultTx(s_ij) */
                double t = 0.0;
                for (int I = 0; I < 6; I++) t += B[I][r] * sig[I];
                fe[r] += det[ip] * t;
            }
        }
        return ORC_OK;
    }
    for (int ip = 0; ip < 8; ip++) {
        double G[9], sig[6];
        grad_u(u, dN[ip], G);
        if (form == ORC_SMALL_STRAIN) {
            double e[6]; /* dSymMatrixT::Symmetrize */
            e[0] = G[0]; e[1] = G[4]; e[2] = G[8];
            e[3] = 0.5 * (G[5] + G[7]); e[4] = 0.5 * (G[2] + G[6]); e[5] = 0.5 * (G[1] + G[3]);
            hooke_stress(m, e, sig);
            add_BT_sigma(det[ip], dN[ip], sig, fe);
            continue;
        }
        double F[9], Fl[9];
        memcpy(F, G, sizeof F);
        F[0] += 1.0; F[4] += 1.0; F[8] += 1.0;
        grad_u(u_last, dN[ip], Fl);
        Fl[0] += 1.0; Fl[4] += 1.0; Fl[8] += 1.0;
        err = fs_material(m, F, Fl, j2, ip, alloc, iteration, sig, NULL);
        if (err) return err;
        if (form == ORC_UPDATED_LAGRANGIAN) {
            add_BT_sigma(detc[ip], dNc[ip], sig, fe);
        } else { /* total Lagrangian: P/J = sigma F^-T ; f_a,i += J w det (P/J)_iJ dN_a/dX_J */
            double J = det3(F);
            if (J <= 0.0) return ORC_BAD_JACOBIAN;
            double Fi[9], S[9], P[9];
            inv3(F, J, Fi);
            sym_to_mat(sig, S);
            for (int jj = 0; jj < 3; jj++)
                for (int i = 0; i < 3; i++) {
                    double s = 0.0;
                    for (int k = 0; k < 3; k++) s += S[i + 3 * k] * Fi[jj + 3 * k]; /* MultABT */
                    P[i + 3 * jj] = s;
                }
            double sc = J * det[ip];
            for (int a = 0; a < 8; a++)
                for (int i = 0; i < 3; i++)
                    fe[3 * a + i] += sc * (P[i] * dN[ip][0][a] + P[i + 3] * dN[ip][1][a] + P[i + 6] * dN[ip][2][a]);
        }
    }
    return ORC_OK;
}

/* K3. SmallStrainT::FormStiffness (SmallStrainT.cpp:285-324); TotalLagrangianT::FormStiffness
 * (TotalLagrangianT.cpp:40-104); UpdatedLagrangianT::FormStiffness (:94-142);
 * nMatrixT::MultQTBQ (nMatrixT.h:1406-1500); dMatrixT::Expand (dMatrixT.cpp:690-727). */
int orc_element_stiffness(int form, const orc_material_t* m, const double X[8][3], const double u[8][3],
                          const double u_last[8][3], orc_j2_ip_t* j2, int* alloc, int iteration, double Ke[576])
{
    double dN[8][3][8], det[8], dNc[8][3][8], detc[8];
    static const double zero[8][3] = {{0}};
    int dummy_alloc = 0;
    if (!alloc) alloc = &dummy_alloc;
    if (!u_last) u_last = zero;
    memset(Ke, 0, 576 * sizeof(double));
    int err = orc_hex8_shape(X, dN, det);
    if (err) return err;
    if (form == ORC_UPDATED_LAGRANGIAN) {
        double x[8][3];
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) x[a][i] = X[a][i] + u[a][i];
        err = orc_hex8_shape(x, dNc, detc);
        if (err) return err;
    }
    double kg[8][8]; /* stress stiffness */
    memset(kg, 0, sizeof kg);
    double mg[3][8];
    if (form == ORC_SMALL_STRAIN_BBAR) mean_gradient(dN, det, mg);
    for (int ip = 0; ip < 8; ip++) {
        double c[6][6], sig[6], B[6][24], scale;
        double(*dNx)[8]; /* spatial gradients used for B */
        double dNp[3][8];
        if (form == ORC_SMALL_STRAIN || form == ORC_SMALL_STRAIN_BBAR) {
            hooke_moduli(m, c);
            dNx = dN[ip];
            scale = det[ip];
        } else {
            double F[9], Fl[9];
            grad_u(u, dN[ip], F);
            F[0] += 1.0; F[4] += 1.0; F[8] += 1.0;
            grad_u(u_last, dN[ip], Fl);
            Fl[0] += 1.0; Fl[4] += 1.0; Fl[8] += 1.0;
            err = fs_material(m, F, Fl, j2, ip, alloc, iteration, sig, c);
            if (err) return err;
            if (form == ORC_UPDATED_LAGRANGIAN) {
                dNx = dNc[ip];
                scale = detc[ip];
            } else { /* push dN/dX forward with F^-1: dN/dx_i = F^-1_Ji dN/dX_J */
                double J = det3(F), Fi[9];
                inv3(F, J, Fi);
                for (int a = 0; a < 8; a++)
                    for (int i = 0; i < 3; i++)
                        dNp[i][a] = Fi[0 + 3 * i] * dN[ip][0][a] + Fi[1 + 3 * i] * dN[ip][1][a] + Fi[2 + 3 * i] * dN[ip][2][a];
                dNx = dNp;
                scale = det[ip] * J;
            }
            double S[9];
            sym_to_mat(sig, S);
            for (int a = 0; a < 8; a++)
                for (int b = 0; b < 8; b++) {
                    double s = 0.0;
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++) s += dNx[i][a] * S[i + 3 * j] * dNx[j][b];
                    kg[a][b] += scale * s;
                }
        }
        if (form == ORC_SMALL_STRAIN_BBAR) set_B_bar(dNx, mg, B);
        else set_B(dNx, B);
        for (int r = 0; r < 24; r++)
            for (int cc = 0; cc < 24; cc++) {
                double s = 0.0;
                for (int I = 0; I < 6; I++) {
                    double t = 0.0;
                    for (int Jj = 0; Jj < 6; Jj++) t += c[I][Jj] * B[Jj][cc];
                    s += B[I][r] * t;
                }
                Ke[r + 24 * cc] += scale * s;
            }
    }
    if (form != ORC_SMALL_STRAIN && form != ORC_SMALL_STRAIN_BBAR)
        for (int a = 0; a < 8; a++)
            for (int b = 0; b < 8; b++)
                for (int i = 0; i < 3; i++) Ke[(3 * a + i) + 24 * (3 * b + i)] += kg[a][b];
    return ORC_OK;
}

/* K4. ContinuumElementT::FormMass kLumpedMass (continuum/common/ContinuumElementT.cpp:767-842) */
int orc_element_lumped_mass(double density, const double X[8][3], double me[8])
{
    double Na[8][8], DNa[8][3][8], w[8], dN[8][3][8], det[8];
    orc_hex8_parent(Na, DNa, w);
    int err = orc_hex8_shape(X, dN, det);
    if (err) return err;
    double dsum = 0.0, totmas = 0.0, nee[8] = {0};
    for (int ip = 0; ip < 8; ip++) {
        double temp1 = density * w[ip] * det[ip];
        totmas += temp1;
        for (int a = 0; a < 8; a++) {
            double temp2 = temp1 * Na[ip][a] * Na[ip][a];
            dsum += temp2;
            nee[a] += temp2;
        }
    }
    double diagmass = totmas / dsum;
    for (int a = 0; a < 8; a++) me[a] = diagmass * nee[a];
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* mesh sweeps: SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295), element order, then
 * SolverT::AssembleRHS (solvers/SolverT.cpp:446-477) adds in that order                               */
/* ------------------------------------------------------------------ */
static void gather(const int32_t* c, const double* A, double out[8][3])
{
    for (int a = 0; a < 8; a++)
        for (int i = 0; i < 3; i++) out[a][i] = A[3 * (int64_t)c[a] + i];
}
int orc_internal_force(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X,
                       const double* u, const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* f,
                       int64_t* bad_elem)
{
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], ue[8][3], ul[8][3], fe[24];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        gather(c, u, ue);
        if (u_last) gather(c, u_last, ul);
        int err = orc_element_force(form, m, Xe, ue, u_last ? ul : NULL, j2 ? j2 + 8 * e : NULL, alloc ? alloc + e : NULL,
                                    iteration, fe);
        if (err) {
            if (bad_elem) *bad_elem = e;
            return err;
        }
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) f[3 * (int64_t)c[a] + i] += fe[3 * a + i];
    }
    return ORC_OK;
}
int orc_lumped_mass(double density, int64_t ne, const int32_t* conn, const double* X, double* mass)
{
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], me[8];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        int err = orc_element_lumped_mass(density, Xe, me);
        if (err) return err;
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) mass[3 * (int64_t)c[a] + i] += me[a];
    }
    return ORC_OK;
}

/* a24. FieldT::InitEquations (nodes/FieldT.cpp:635-659) + NodeManagerT::SetEquationNumbers
 * (nodes/NodeManagerT.cpp:712-767): node-major, dof-minor, prescribed = kPrescribed (-1) */
int64_t orc_set_equation_numbers(int64_t nn, const uint8_t* bc, int32_t* eqnos)
{
    int64_t num_eq = 0;
    for (int64_t i = 0; i < nn * 3; i++) eqnos[i] = bc[i] ? -1 : (int32_t)(++num_eq);
    return num_eq;
}

/* a22. GraphT::MakeGraph(active, add_self=true, upper_only) (toolbox/src/graph/GraphT.cpp:376-488):
 * every pair of active equations of an element is an edge; MSRBuilderT::GenerateSuperLU
 * (MSRBuilderT.cpp:216-244) / GenerateMSR (:134-181) sort each row ascending. */
static int cmp_i32(const void* a, const void* b)
{
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}
typedef struct { int64_t* ptr; int32_t* col; } rows_t;
static rows_t build_rows(int64_t ne, const int32_t* conn, int64_t nn, const int32_t* eqnos, int64_t neq, int upper_only)
{
    /* node -> elements */
    int64_t* nptr = calloc(nn + 1, sizeof(int64_t));
    for (int64_t i = 0; i < ne * 8; i++) nptr[conn[i] + 1]++;
    for (int64_t n = 0; n < nn; n++) nptr[n + 1] += nptr[n];
    int64_t* fill = malloc(nn * sizeof(int64_t));
    memcpy(fill, nptr, nn * sizeof(int64_t));
    int64_t* nel = malloc(ne * 8 * sizeof(int64_t));
    for (int64_t e = 0; e < ne; e++)
        for (int a = 0; a < 8; a++) nel[fill[conn[8 * e + a]]++] = e;
    rows_t R;
    R.ptr = calloc(neq + 1, sizeof(int64_t));
    int64_t cap = 1024, len = 0;
    R.col = malloc(cap * sizeof(int32_t));
    int32_t* tmp = malloc(27 * 3 * 8 * sizeof(int32_t));
    for (int64_t n = 0; n < nn; n++) {
        /* candidate columns: all active eqs of all nodes of all elements at node n */
        int nt = 0;
        for (int64_t k = nptr[n]; k < nptr[n + 1]; k++)
            for (int a = 0; a < 8; a++)
                for (int i = 0; i < 3; i++) {
                    int32_t q = eqnos[3 * (int64_t)conn[8 * nel[k] + a] + i];
                    if (q > 0) tmp[nt++] = q - 1;
                }
        qsort(tmp, nt, sizeof(int32_t), cmp_i32);
        int nu = 0;
        for (int k = 0; k < nt; k++)
            if (nu == 0 || tmp[k] != tmp[nu - 1]) tmp[nu++] = tmp[k];
        for (int i = 0; i < 3; i++) {
            int32_t r = eqnos[3 * n + i];
            if (r <= 0) continue;
            r -= 1;
            if (len + nu > cap) { while (len + nu > cap) cap *= 2; R.col = realloc(R.col, cap * sizeof(int32_t)); }
            int64_t start = len;
            for (int k = 0; k < nu; k++)
                if (!upper_only || tmp[k] >= r) R.col[len++] = tmp[k];
            R.ptr[r + 1] = len - start;
        }
    }
    /* rows were emitted in equation order because numbering is node-major */
    for (int64_t r = 0; r < neq; r++) R.ptr[r + 1] += R.ptr[r];
    free(nptr); free(fill); free(nel); free(tmp);
    return R;
}
int64_t orc_csr_structure(int64_t ne, const int32_t* conn, int64_t nn, const int32_t* eqnos, int64_t neq, int upper_only,
                          int64_t* rowptr, int32_t* colind)
{
    rows_t R = build_rows(ne, conn, nn, eqnos, neq, upper_only);
    int64_t nnz = R.ptr[neq];
    if (rowptr) memcpy(rowptr, R.ptr, (neq + 1) * sizeof(int64_t));
    if (colind) memcpy(colind, R.col, nnz * sizeof(int32_t));
    free(R.ptr); free(R.col);
    return nnz;
}
int64_t orc_msr_structure(int64_t ne, const int32_t* conn, int64_t nn, const int32_t* eqnos, int64_t neq, int upper_only,
                          int32_t* bindx)
{
    rows_t R = build_rows(ne, conn, nn, eqnos, neq, upper_only);
    int64_t nnz = R.ptr[neq];
    int64_t total = nnz - neq + neq + 1;
    if (bindx) {
        int64_t pos = neq + 1;
        bindx[0] = (int32_t)pos;
        for (int64_t r = 0; r < neq; r++) {
            for (int64_t k = R.ptr[r]; k < R.ptr[r + 1]; k++)
                if (R.col[k] != r) bindx[pos++] = R.col[k];
            bindx[r + 1] = (int32_t)pos;
        }
    }
    free(R.ptr); free(R.col);
    return total;
}

/* greedy colouring, elements visited in order, smallest colour not used by any
 * element sharing a node.  No reference counterpart. */
int orc_greedy_colouring(int64_t ne, const int32_t* conn, int64_t nn, int32_t* colour)
{
    uint64_t* used = calloc(nn, sizeof(uint64_t)); /* bitmask of colours present at each node (<= 64 colours) */
    int ncol = 0;
    for (int64_t e = 0; e < ne; e++) {
        uint64_t mask = 0;
        for (int a = 0; a < 8; a++) mask |= used[conn[8 * e + a]];
        int c = 0;
        while (c < 64 && (mask >> c) & 1) c++;
        if (c >= 64) { free(used); return -1; }
        colour[e] = c;
        if (c + 1 > ncol) ncol = c + 1;
        for (int a = 0; a < 8; a++) used[conn[8 * e + a]] |= (uint64_t)1 << c;
    }
    free(used);
    return ncol;
}

/* SolidElementT::ElementLHSDriver (SolidElementT.cpp:1100-1154) + MSRMatrixT::Assemble
 * (MSRMatrixT.cpp:66-216) restated on the full CSR: val[pos(eq_r, eq_c)] += Ke[r][c] in element order */
int orc_assemble_stiffness(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X,
                           const double* u, const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration,
                           const int32_t* eqnos, int64_t neq, const int64_t* rowptr, const int32_t* colind, double* val)
{
    (void)neq;
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], ue[8][3], ul[8][3], Ke[576];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        gather(c, u, ue);
        if (u_last) gather(c, u_last, ul);
        int err = orc_element_stiffness(form, m, Xe, ue, u_last ? ul : NULL, j2 ? j2 + 8 * e : NULL,
                                        alloc ? alloc + e : NULL, iteration, Ke);
        if (err) return err;
        int32_t eq[24];
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) eq[3 * a + i] = eqnos[3 * (int64_t)c[a] + i];
        for (int r = 0; r < 24; r++) {
            if (eq[r] <= 0) continue;
            int64_t row = eq[r] - 1;
            for (int cc = 0; cc < 24; cc++) {
                if (eq[cc] <= 0) continue;
                int32_t col = eq[cc] - 1;
                int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
                while (lo < hi) { int64_t mid = (lo + hi) / 2; if (colind[mid] < col) lo = mid + 1; else hi = mid; }
                if (colind[lo] != col) return -1;
                val[lo] += Ke[r + 24 * cc];
            }
        }
    }
    return ORC_OK;
}

/* K6. MSRMatrixT::Multx (primitives/globalmatrix/MSRMatrixT.cpp:385-420) on the CSR form */
void orc_csr_spmv(int64_t n, const int64_t* rowptr, const int32_t* colind, const double* val, const double* x, double* y)
{
    for (int64_t r = 0; r < n; r++) {
        double s = 0.0;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) s += val[k] * x[colind[k]];
        y[r] = s;
    }
}
static double dot(int64_t n, const double* a, const double* b) /* nArrayT::Dot (nArrayT.h:988-998) */
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
/* K6-K8: linear CG with Jacobi preconditioner = DiagonalMatrixT::Factorize/BackSubstitute
 * (DiagonalMatrixT.cpp:267-323: reciprocal, |m| < 1e-12 skipped) applied to r.
 * Textbook PCG (Aztec-style AZ_cg + AZ_Jacobi, the precedent AztecMatrixT.h:18-75 names). */
int orc_pcg_jacobi(int64_t n, const int64_t* rowptr, const int32_t* colind, const double* val, const double* b, double* x,
                   double rtol, double atol, int max_iter, double* final_rnorm)
{
    double *r = malloc(n * 8), *z = malloc(n * 8), *p = malloc(n * 8), *q = malloc(n * 8), *dinv = malloc(n * 8);
    for (int64_t i = 0; i < n; i++) {
        double d = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
            if (colind[k] == i) d = val[k];
        dinv[i] = fabs(d) > 1.0e-12 ? 1.0 / d : d;
    }
    orc_csr_spmv(n, rowptr, colind, val, x, q);
    for (int64_t i = 0; i < n; i++) { r[i] = b[i] - q[i]; z[i] = dinv[i] * r[i]; p[i] = z[i]; }
    double rz = dot(n, r, z), rnorm = sqrt(dot(n, r, r)), r0 = rnorm;
    int it = 0;
    while (it < max_iter && rnorm > atol && rnorm > rtol * r0) {
        orc_csr_spmv(n, rowptr, colind, val, p, q);
        double alpha = rz / dot(n, p, q);
        for (int64_t i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; z[i] = dinv[i] * r[i]; }
        double rz_new = dot(n, r, z);
        rnorm = sqrt(dot(n, r, r));
        double beta = rz_new / rz;
        rz = rz_new;
        for (int64_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        it++;
    }
    if (final_rnorm) *final_rnorm = rnorm;
    free(r); free(z); free(p); free(q); free(dinv);
    return it;
}

/* a19. nExplicitCD::Predictor (integrators/explicitCD/nExplicitCD.cpp:72-96, constants :214-223)
 * then ConsistentKBC (:20-69) for kFix / kDsp cards, as FieldT::InitStep orders them (FieldT.cpp:326-387) */
void orc_cd_predictor(int64_t ndof, double dt, double* d, double* v, double* a, const uint8_t* bc, const double* bcval)
{
    double dpred_v = dt, dpred_a = 0.5 * dt * dt, vpred_a = 0.5 * dt;
    for (int64_t i = 0; i < ndof; i++) {
        d[i] += dpred_v * v[i] + dpred_a * a[i];
        v[i] += vpred_a * a[i];
        a[i] = 0.0;
        if (bc && bc[i] == 1) { d[i] = 0.0; v[i] = 0.0; a[i] = 0.0; }
        else if (bc && bc[i] == 2) { d[i] = bcval[i]; a[i] = 0.0; }
    }
}
/* LinearSolver::Solve (solvers/LinearSolver.cpp:37-101): update = M^-1 R (DiagonalMatrixT.cpp:267-323),
 * FieldT::AssembleUpdate (FieldT.cpp:531-556: prescribed dofs get 0), nExplicitCD::Corrector (:98-139) */
void orc_cd_corrector(int64_t ndof, double dt, double* v, double* a, const double* R, const double* mass, const uint8_t* bc)
{
    double vcorr_a = 0.5 * dt;
    for (int64_t i = 0; i < ndof; i++) {
        double upd = 0.0;
        if (!(bc && bc[i])) {
            double minv = fabs(mass[i]) > 1.0e-12 ? 1.0 / mass[i] : mass[i];
            upd = R[i] * minv;
        }
        v[i] += vcorr_a * upd;
        a[i] += upd;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * a21: PCGSolver_LS (solvers/PCGSolver_LS.cpp) inside NLSolver::Solve (solvers/NLSolver.cpp:57-263), with
 * <diagonal_matrix/> as the matrix type: DiagonalMatrixT in kDiagOnly mode (SolverT.cpp:1097-1102) = diag K(u),
 * re-formed every `restart` iterations (fReformTangentIterations = fRestart, PCGSolver_LS.cpp:97).
 * ------------------------------------------------------------------------------------------------------------------ */

/* SolidElementT::ElementLHSDriver assembled by DiagonalMatrixT::Assemble, kDiagOnly (DiagonalMatrixT.cpp:107-113) */
int orc_stiffness_diagonal(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, const double* u,
                           const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* diag /*[nn][3] accumulated*/)
{
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], ue[8][3], ul[8][3], Ke[576];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        gather(c, u, ue);
        if (u_last) gather(c, u_last, ul);
        int err = orc_element_stiffness(form, m, Xe, ue, u_last ? ul : NULL, j2 ? j2 + 8 * e : NULL, alloc ? alloc + e : NULL,
                                        iteration, Ke);
        if (err) return err;
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) diag[3 * (int64_t)c[a] + i] += Ke[(3 * a + i) * 25];
    }
    return ORC_OK;
}

typedef struct {
    int form;
    const orc_material_t* m;
    int64_t ne, nn, neq;
    const int32_t *conn, *eqnos;
    const double *X, *u_last, *fext;
    orc_j2_ip_t* j2;
    int* alloc;
    double* u;     /* [nn][3] current displacement */
    double* work;  /* [nn][3] */
    int iteration; /* SolverT::fNumIteration */
} nlpcg_sys_t;

/* FEManagerT::FormRHS: nodal forces first (NodeManagerT::FormRHS), then the element group (-fint); active equations only */
static int nlpcg_form_rhs(nlpcg_sys_t* s, double* R)
{
    memset(s->work, 0, sizeof(double) * 3 * s->nn);
    int err = orc_internal_force(s->form, s->m, s->ne, s->conn, s->X, s->u, s->u_last, s->j2, s->alloc, s->iteration, s->work, NULL);
    if (err) return err;
    for (int64_t k = 0; k < 3 * s->nn; k++)
        if (s->eqnos[k] > 0) R[s->eqnos[k] - 1] = s->fext[k] - s->work[k];
    return ORC_OK;
}
/* FEManagerT::Update -> FieldT::AssembleUpdate (FieldT.cpp:531-556): u[active] += update */
static void nlpcg_update(nlpcg_sys_t* s, const double* upd)
{
    for (int64_t k = 0; k < 3 * s->nn; k++)
        if (s->eqnos[k] > 0) s->u[k] += upd[s->eqnos[k] - 1];
}
/* FormLHS into the cleared DiagonalMatrixT, then Factorize (DiagonalMatrixT.cpp:267-310: reciprocal unless |m| <= kSmall) */
static int nlpcg_form_preconditioner(nlpcg_sys_t* s, double* minv)
{
    memset(s->work, 0, sizeof(double) * 3 * s->nn);
    int err = orc_stiffness_diagonal(s->form, s->m, s->ne, s->conn, s->X, s->u, s->u_last, s->j2, s->alloc, s->iteration, s->work);
    if (err) return err;
    for (int64_t k = 0; k < 3 * s->nn; k++)
        if (s->eqnos[k] > 0) {
            const double d = s->work[k];
            minv[s->eqnos[k] - 1] = fabs(d) > 1.0e-12 ? 1.0 / d : d;
        }
    return ORC_OK;
}

/* PCGSolver_LS::GValue (PCGSolver_LS.cpp:351-371) */
static int nlpcg_gvalue(nlpcg_sys_t* s, const double* update, double step, double* s_current, double* R, double* scratch, double* G)
{
    const double ds = step - *s_current;
    for (int64_t i = 0; i < s->neq; i++) scratch[i] = update[i] * ds;
    *s_current = step;
    nlpcg_update(s, scratch);
    int err = nlpcg_form_rhs(s, R);
    if (err) return err;
    *G = dot(s->neq, update, R);
    return ORC_OK;
}

int orc_nlpcg_solve(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X, double* u,
                    const double* u_last, orc_j2_ip_t* j2, int* alloc, const int32_t* eqnos, int64_t neq, const double* fext,
                    const orc_nlpcg_params_t* prm, int* iterations, double* error_out, double* error0_out)
{
    nlpcg_sys_t s = {form, m, ne, nn, neq, conn, eqnos, X, u_last, fext, j2, alloc, u, NULL, -1};
    s.work = malloc(sizeof(double) * 3 * nn);
    double *R = malloc(8 * neq), *Rres = malloc(8 * neq), *R_last = malloc(8 * neq), *u_lastdir = malloc(8 * neq),
           *diffR = malloc(8 * neq), *minv = malloc(8 * neq), *upd = malloc(8 * neq), *scratch = malloc(8 * neq);
    double(*search)[2] = malloc(sizeof(double) * 2 * (prm->ls_iterations > 2 ? prm->ls_iterations : 2));
    int status = ORC_NLPCG_CONTINUE, err = ORC_OK;
    int num_iterations = 0, tan_iterations = 0, restart_count = -1;
    const int reform = prm->restart; /* fReformTangentIterations = fRestart */
    double error = 0.0, error0 = 0.0;

#define EXIT_ITERATION(iter)                                                                                     \
    do {                                                                                                          \
        if ((iter) == -1) {                                                                                       \
            error0 = error;                                                                                       \
            status = error0 < prm->abs_tol ? ORC_NLPCG_CONVERGED : ORC_NLPCG_CONTINUE;                            \
        } else {                                                                                                  \
            const double rel = error / error0;                                                                    \
            if (rel > prm->div_tol) status = ORC_NLPCG_FAILED;                                                    \
            else if ((iter) < prm->min_iterations - 1) status = ORC_NLPCG_CONTINUE;                               \
            else if (rel < prm->rel_tol || error < prm->abs_tol) status = ORC_NLPCG_CONVERGED;                    \
            else if ((iter) >= prm->max_iterations) status = ORC_NLPCG_FAILED;                                    \
            else status = ORC_NLPCG_CONTINUE;                                                                     \
        }                                                                                                         \
    } while (0)

    /* NLSolver.cpp:75-86: first residual, fNumIteration = -1 */
    if ((err = nlpcg_form_rhs(&s, R))) goto done;
    error = sqrt(dot(neq, R, R));
    EXIT_ITERATION(s.iteration);
    while (status == ORC_NLPCG_CONTINUE) {
        num_iterations++;
        tan_iterations++;
        int lhs_update = 0;
        if (num_iterations == 1 || tan_iterations >= reform) { lhs_update = 1; tan_iterations = 0; }
        if (lhs_update && (err = nlpcg_form_preconditioner(&s, minv))) goto done;

        /* ---- PCGSolver_LS::Iterate (:107-118): fR = fRHS; CGSearch; Update(fRHS, &fR) */
        memcpy(Rres, R, 8 * neq);
        restart_count++;
        if (restart_count == 0 || restart_count == prm->restart) { /* CGSearch :152-165 */
            memcpy(R_last, R, 8 * neq);
            for (int64_t i = 0; i < neq; i++) R[i] *= minv[i];
            memcpy(u_lastdir, R, 8 * neq);
            restart_count = 0;
        } else { /* :166-201, Bertsekas (6.32) */
            for (int64_t i = 0; i < neq; i++) diffR[i] = (R[i] - R_last[i]) * minv[i];
            double beta = dot(neq, R, diffR);
            for (int64_t i = 0; i < neq; i++) diffR[i] = R_last[i] * minv[i];
            double denominator = dot(neq, R_last, diffR);
            if (fabs(denominator) < 1.0e-24) { denominator = 1.0; beta = 0.0; }
            else beta /= denominator;
            memcpy(R_last, R, 8 * neq);
            for (int64_t i = 0; i < neq; i++) R[i] = R[i] * minv[i] + beta * u_lastdir[i];
            memcpy(u_lastdir, R, 8 * neq);
        }
        /* ---- PCGSolver_LS::Update (:213-348) */
        if (prm->ls_iterations == 0) {
            nlpcg_update(&s, R);
        } else {
            memcpy(upd, R, 8 * neq);
            double s_current = 0.0, s_a = 0.0, G_a = dot(neq, upd, Rres), s_b = 1.0, G_b, G_new = 0.0;
            search[0][0] = s_a; search[0][1] = G_a;
            if ((err = nlpcg_gvalue(&s, upd, s_b, &s_current, R, scratch, &G_b))) goto done;
            search[1][0] = s_b; search[1][1] = G_b;
            const double G_0 = fabs(G_a) > fabs(G_b) ? G_b : G_a;
            int count = 2, give_up = 0;
            do {
                const double mm = (G_a - G_b) / (s_a - s_b);
                const double bb = G_b - mm * s_b;
                double s_new = -bb / mm;
                if (s_new > prm->max_step || s_new < 0.0) {
                    give_up = 1;
                    if (s_new > prm->max_step) {
                        s_new = prm->max_step;
                        if ((err = nlpcg_gvalue(&s, upd, s_new, &s_current, R, scratch, &G_new))) goto done;
                        search[count][0] = s_new; search[count][1] = G_new;
                        count++;
                    }
                    break;
                }
                if ((err = nlpcg_gvalue(&s, upd, s_new, &s_current, R, scratch, &G_new))) goto done;
                search[count][0] = s_new; search[count][1] = G_new;
                if (fabs(G_a) > fabs(G_new) && fabs(G_a) > fabs(G_b)) { G_a = G_new; s_a = s_new; give_up = 0; }
                else if (fabs(G_b) > fabs(G_new) && fabs(G_b) > fabs(G_a)) { G_b = G_new; s_b = s_new; give_up = 0; }
                else if (G_b * G_a > 0) {
                    if (G_a * G_new < 0) { G_a = G_new; s_a = s_new; }
                    else if (G_b * G_new < 0) { G_b = G_new; s_b = s_new; }
                    else give_up = 1;
                } else give_up = 1;
                if (++count >= prm->ls_iterations) give_up = 1;
            } while (fabs(G_new) > prm->abs_tol && fabs(G_new / G_0) > prm->ls_tolerance && !give_up);
            if (give_up) { /* best step on fail (:316-340) */
                double s_best = fabs(search[0][0]), G_best = fabs(search[0][1]);
                int best = 0;
                for (int i = 1; i < count; i++) {
                    const double s_test = fabs(search[i][0]), G_test = fabs(search[i][1]);
                    if (fabs(s_best) < 1.0e-12 || (s_test > 1.0e-12 && G_test < G_best)) { s_best = s_test; G_best = G_test; best = i; }
                }
                double G_dummy;
                if ((err = nlpcg_gvalue(&s, upd, search[best][0], &s_current, R, scratch, &G_dummy))) goto done;
            }
        }
        s.iteration++; /* fNumIteration++ (NLSolver.cpp:172) */
        /* NLSolver.cpp:174-195: residual at the updated state */
        if ((err = nlpcg_form_rhs(&s, R))) goto done;
        error = sqrt(dot(neq, R, R));
        if (getenv("ORC_NLPCG_TRACE")) fprintf(stderr, "%d: Relative error = %e\n", s.iteration, error / error0);
        EXIT_ITERATION(s.iteration);
        if (prm->solve_max_iterations >= 0 && num_iterations >= prm->solve_max_iterations) break;
    }
#undef EXIT_ITERATION
done:
    if (iterations) *iterations = s.iteration;
    if (error_out) *error_out = error;
    if (error0_out) *error0_out = error0;
    free(s.work); free(R); free(Rres); free(R_last); free(u_lastdir); free(diffR); free(minv); free(upd); free(scratch); free(search);
    return err ? -err : status;
}

/* ------------------------------------------------------------------------------------------------------------------
 * <explicit_solid> on Hex8 -- ExplicitElementT::BatchedInternalForce (elements/explicit/ExplicitElementT.cpp:
 * 649-993) with Hex8KernelT::ComputeIPData (kernels/Hex8KernelT.cpp:13-79: 2x2x2 Gauss, unit weights), ExplNeoHookeanT::
 * ComputeStress3D (materials/ExplNeoHookeanT.cpp:79-111) or ExplJ2PlasticityT::ComputeStress3D (materials/ExplJ2PlasticityT.cpp:
 * 87-310), CFL estimate (:404-478) and fixed mass scaling (:492-571, LHSDriver :576-618).
 * ------------------------------------------------------------------------------------------------------------------ */
static double xs_ip_data(int ip, const double x[8][3], double dN[3][8])
{
    static const double g = 0.5773502691896258;
    static const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    const double xi = sx[ip] * g, eta = sy[ip] * g, mu = sz[ip] * g;
    double dxi[8], deta[8], dmu[8];
    for (int n = 0; n < 8; n++) {
        dxi[n] = 0.125 * sx[n] * (1.0 + sy[n] * eta) * (1.0 + sz[n] * mu);
        deta[n] = 0.125 * sy[n] * (1.0 + sx[n] * xi) * (1.0 + sz[n] * mu);
        dmu[n] = 0.125 * sz[n] * (1.0 + sx[n] * xi) * (1.0 + sy[n] * eta);
    }
    double J[3][3] = {{0.0}};
    for (int n = 0; n < 8; n++)
        for (int r = 0; r < 3; r++) {
            J[r][0] += x[n][r] * dxi[n];
            J[r][1] += x[n][r] * deta[n];
            J[r][2] += x[n][r] * dmu[n];
        }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const double inv = 1.0 / det;
    double Ji[3][3];
    Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * inv;
    Ji[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]) * inv;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
    Ji[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]) * inv;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
    Ji[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]) * inv;
    Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * inv;
    Ji[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]) * inv;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
    for (int n = 0; n < 8; n++)
        for (int c = 0; c < 3; c++) dN[c][n] = Ji[0][c] * dxi[n] + Ji[1][c] * deta[n] + Ji[2][c] * dmu[n];
    return det;
}

/* F row-major F[3*i+j]; sig Voigt 11,22,33,23,13,12 */
static void xs_neo_hookean(const orc_material_t* m, const double* F, double* sig)
{
    const double J = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
    const double b11 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2], b22 = F[3] * F[3] + F[4] * F[4] + F[5] * F[5],
                 b33 = F[6] * F[6] + F[7] * F[7] + F[8] * F[8], b12 = F[0] * F[3] + F[1] * F[4] + F[2] * F[5],
                 b13 = F[0] * F[6] + F[1] * F[7] + F[2] * F[8], b23 = F[3] * F[6] + F[4] * F[7] + F[5] * F[8];
    const double muJ = m->mu / J, pres = m->kappa * (J - 1.0) / J;
    sig[0] = muJ * (b11 - 1.0) + pres;
    sig[1] = muJ * (b22 - 1.0) + pres;
    sig[2] = muJ * (b33 - 1.0) + pres;
    sig[3] = muJ * b23;
    sig[4] = muJ * b13;
    sig[5] = muJ * b12;
}

static void inv3_rowmajor(const double* A, double* Ai)
{
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    const double id = 1.0 / det;
    Ai[0] = (A[4] * A[8] - A[5] * A[7]) * id;
    Ai[1] = -(A[1] * A[8] - A[2] * A[7]) * id;
    Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    Ai[3] = -(A[3] * A[8] - A[5] * A[6]) * id;
    Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id;
    Ai[5] = -(A[0] * A[5] - A[2] * A[3]) * id;
    Ai[6] = (A[3] * A[7] - A[4] * A[6]) * id;
    Ai[7] = -(A[0] * A[7] - A[1] * A[6]) * id;
    Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
static void mul3_rowmajor(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
/* Hughes-Winget J2 with linear isotropic hardening; h[16] = F_n (9), sigma_n (6), eps_p; updated in place on every call */
static void xs_j2(const orc_material_t* m, const double* F, double* h, double* sig)
{
    const double mu = m->mu, lam = m->kappa - 2.0 * m->mu / 3.0, sigY0 = m->hard[0], H = m->hard[1];
    double Fi[9], f[9];
    inv3_rowmajor(h, Fi);
    mul3_rowmajor(F, Fi, f);
    const double de11 = f[0] - 1.0, de22 = f[4] - 1.0, de33 = f[8] - 1.0, de23 = 0.5 * (f[5] + f[7]), de13 = 0.5 * (f[2] + f[6]),
                 de12 = 0.5 * (f[1] + f[3]);
    const double w23 = 0.25 * (f[5] - f[7]), w13 = 0.25 * (f[2] - f[6]), w12 = 0.25 * (f[1] - f[3]);
    const double B[9] = {1.0, w12, -w13, -w12, 1.0, w23, w13, -w23, 1.0}, A[9] = {1.0, -w12, w13, w12, 1.0, -w23, -w13, w23, 1.0};
    double Binv[9], Q[9];
    inv3_rowmajor(B, Binv);
    mul3_rowmajor(Binv, A, Q);
    const double sn11 = h[9], sn22 = h[10], sn33 = h[11], sn23 = h[12], sn13 = h[13], sn12 = h[14];
    const double T11 = Q[0] * sn11 + Q[1] * sn12 + Q[2] * sn13, T12 = Q[0] * sn12 + Q[1] * sn22 + Q[2] * sn23,
                 T13 = Q[0] * sn13 + Q[1] * sn23 + Q[2] * sn33, T21 = Q[3] * sn11 + Q[4] * sn12 + Q[5] * sn13,
                 T22 = Q[3] * sn12 + Q[4] * sn22 + Q[5] * sn23, T23 = Q[3] * sn13 + Q[4] * sn23 + Q[5] * sn33,
                 T31 = Q[6] * sn11 + Q[7] * sn12 + Q[8] * sn13, T32 = Q[6] * sn12 + Q[7] * sn22 + Q[8] * sn23,
                 T33 = Q[6] * sn13 + Q[7] * sn23 + Q[8] * sn33;
    const double sr11 = T11 * Q[0] + T12 * Q[1] + T13 * Q[2], sr22 = T21 * Q[3] + T22 * Q[4] + T23 * Q[5],
                 sr33 = T31 * Q[6] + T32 * Q[7] + T33 * Q[8], sr23 = T21 * Q[6] + T22 * Q[7] + T23 * Q[8],
                 sr13 = T11 * Q[6] + T12 * Q[7] + T13 * Q[8], sr12 = T11 * Q[3] + T12 * Q[4] + T13 * Q[5];
    const double lamTr = lam * (de11 + de22 + de33);
    const double st11 = sr11 + lamTr + 2.0 * mu * de11, st22 = sr22 + lamTr + 2.0 * mu * de22, st33 = sr33 + lamTr + 2.0 * mu * de33,
                 st23 = sr23 + 2.0 * mu * de23, st13 = sr13 + 2.0 * mu * de13, st12 = sr12 + 2.0 * mu * de12;
    const double p = (st11 + st22 + st33) / 3.0;
    const double s11 = st11 - p, s22 = st22 - p, s33 = st33 - p;
    const double eps_p = h[15], sig_Y = sigY0 + H * eps_p;
    const double s2 = s11 * s11 + s22 * s22 + s33 * s33 + 2.0 * (st23 * st23 + st13 * st13 + st12 * st12);
    const double q = sqrt(1.5 * s2), phi = q - sig_Y;
    double new_eps_p = eps_p;
    if (phi > 0.0) { /* radial return */
        const double dlam = phi / (3.0 * mu + H);
        const double factor = 1.0 - 3.0 * mu * dlam / q;
        sig[0] = s11 * factor + p;
        sig[1] = s22 * factor + p;
        sig[2] = s33 * factor + p;
        sig[3] = st23 * factor;
        sig[4] = st13 * factor;
        sig[5] = st12 * factor;
        new_eps_p = eps_p + dlam;
    } else {
        sig[0] = s11 + p;
        sig[1] = s22 + p;
        sig[2] = s33 + p;
        sig[3] = st23;
        sig[4] = st13;
        sig[5] = st12;
    }
    for (int k = 0; k < 9; k++) h[k] = F[k];
    for (int k = 0; k < 6; k++) h[9 + k] = sig[k];
    h[15] = new_eps_p;
}

void orc_explicit_solid_init_history(int64_t ne, double* hist)
{
    memset(hist, 0, sizeof(double) * 128 * ne); /* ExplJ2PlasticityT::InitializeHistory: F_n = I */
    for (int64_t k = 0; k < 8 * ne; k++) hist[16 * k + 0] = hist[16 * k + 4] = hist[16 * k + 8] = 1.0;
}

int orc_explicit_solid_force(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, const double* u,
                             double* hist /*[ne][8][16], EXPL_J2 only*/, double* f /*[nn][3] accumulated: +B^T sigma*/)
{
    for (int64_t e = 0; e < ne; e++) {
        const int32_t* c = conn + 8 * e;
        double xr[8][3], xc[8][3], fe[8][3];
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) {
                xr[a][i] = X[3 * (int64_t)c[a] + i];
                xc[a][i] = xr[a][i] + u[3 * (int64_t)c[a] + i]; /* ElementSupportT::CurrentCoordinates */
                fe[a][i] = 0.0;
            }
        for (int ip = 0; ip < 8; ip++) {
            double dNX[3][8], dNx[3][8], F[9] = {0.0}, sig[6];
            xs_ip_data(ip, xr, dNX);
            const double det = xs_ip_data(ip, xc, dNx);
            for (int n = 0; n < 8; n++)
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) F[3 * i + j] += xc[n][i] * dNX[j][n];
            if (m->kind == ORC_EXPL_J2) xs_j2(m, F, hist + 16 * (8 * e + ip), sig);
            else xs_neo_hookean(m, F, sig);
            for (int n = 0; n < 8; n++) {
                fe[n][0] += (dNx[0][n] * sig[0] + dNx[1][n] * sig[5] + dNx[2][n] * sig[4]) * det;
                fe[n][1] += (dNx[0][n] * sig[5] + dNx[1][n] * sig[1] + dNx[2][n] * sig[3]) * det;
                fe[n][2] += (dNx[0][n] * sig[4] + dNx[1][n] * sig[3] + dNx[2][n] * sig[2]) * det;
            }
        }
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) f[3 * (int64_t)c[a] + i] += fe[a][i];
    }
    return ORC_OK;
}

/* characteristic length of ExplicitElementT::ComputeStableTimeStep / ApplyMassScaling: cbrt of the three-diagonal volume estimate */
static double xs_char_length(const int32_t* c, const double* X)
{
    double d[3][3];
    static const int pa[3] = {6, 7, 5}, pb[3] = {0, 1, 3};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) d[k][i] = X[3 * (int64_t)c[pa[k]] + i] - X[3 * (int64_t)c[pb[k]] + i];
    const double vol = fabs(d[0][0] * (d[1][1] * d[2][2] - d[1][2] * d[2][1]) - d[0][1] * (d[1][0] * d[2][2] - d[1][2] * d[2][0]) +
                            d[0][2] * (d[1][0] * d[2][1] - d[1][1] * d[2][0])) / 6.0;
    return cbrt(vol);
}
double orc_explicit_solid_stable_dt(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X)
{
    const double c = sqrt((m->kappa + 4.0 * m->mu / 3.0) / m->density);
    double dt_min = 1.0e30;
    for (int64_t e = 0; e < ne; e++) {
        const double dt = xs_char_length(conn + 8 * e, X) / c;
        if (dt < dt_min) dt_min = dt;
    }
    return dt_min;
}
void orc_explicit_solid_mass_scale(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, double target_dt,
                                   double scale_factor, double* scale)
{
    const double c = sqrt((m->kappa + 4.0 * m->mu / 3.0) / m->density), target = target_dt * scale_factor;
    for (int64_t e = 0; e < ne; e++) {
        const double dt = xs_char_length(conn + 8 * e, X) / c;
        const double alpha = target / dt;
        scale[e] = dt < target ? alpha * alpha : 1.0;
    }
}
/* ExplicitElementT::LHSDriver (:576-618): FormMass with density * fMassScale[e] */
int orc_lumped_mass_scaled(double density, int64_t ne, const int32_t* conn, const double* X, const double* scale, double* mass)
{
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], me[8];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        int err = orc_element_lumped_mass(density * (scale ? scale[e] : 1.0), Xe, me);
        if (err) return err;
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) mass[3 * (int64_t)c[a] + i] += me[a];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------------------------
 * nodal stress output.  SolidElementT::ComputeOutput (SolidElementT.cpp:1352-1840), iNodalStress branch:
 * Cauchy stress at the 8 points -> ShapeFunctionT::Extrapolate = ParentDomainT::NodalValues (ParentDomainT.cpp:1954-1998) with the
 * smoothing matrix of HexahedronT::SetExtrapolation (HexahedronT.cpp:2099-2150: E[a][ip] = (1 + sqrt3 s_a.s_ip)/8) ->
 * ElementSupportT::AssembleAverage / GroupAverageT::Average (toolbox/src/misc/GroupAverageT.cpp:40-49, 190-205).
 * ------------------------------------------------------------------------------------------------------------------ */
static int nodal_stress_core(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X, const double* u,
                             const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* out /*[nn][6]*/)
{
    static const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    const double sqrt3 = sqrt(3.0);
    int* count = calloc(nn, sizeof(int));
    memset(out, 0, sizeof(double) * 6 * nn);
    for (int64_t e = 0; e < ne; e++) {
        const int32_t* c = conn + 8 * e;
        double Xe[8][3], ue[8][3], dN[8][3][8], det[8], nodal[8][6], mg[3][8];
        double ule[8][3];
        gather(c, X, Xe);
        gather(c, u, ue);
        if (u_last) gather(c, u_last, ule);
        int err = orc_hex8_shape(Xe, dN, det);
        if (err) { free(count); return err; }
        if (form == ORC_SMALL_STRAIN_BBAR) mean_gradient(dN, det, mg);
        memset(nodal, 0, sizeof nodal);
        for (int ip = 0; ip < 8; ip++) {
            double G[9], sig[6];
            grad_u(ue, dN[ip], G);
            if (form == ORC_SMALL_STRAIN || form == ORC_SMALL_STRAIN_BBAR) {
                double eps[6];
                if (form == ORC_SMALL_STRAIN_BBAR) {
                    double B[6][24];
                    set_B_bar(dN[ip], mg, B);
                    for (int I = 0; I < 6; I++) {
                        double t = 0.0;
                        for (int a = 0; a < 8; a++)
                            for (int i = 0; i < 3; i++) t += B[I][3 * a + i] * ue[a][i];
                        eps[I] = I < 3 ? t : 0.5 * t;
                    }
                } else {
                    eps[0] = G[0]; eps[1] = G[4]; eps[2] = G[8];
                    eps[3] = 0.5 * (G[5] + G[7]); eps[4] = 0.5 * (G[2] + G[6]); eps[5] = 0.5 * (G[1] + G[3]);
                }
                hooke_stress(m, eps, sig);
            } else {
                double F[9], Fl[9];
                memcpy(F, G, sizeof F);
                F[0] += 1.0; F[4] += 1.0; F[8] += 1.0;
                memcpy(Fl, F, sizeof Fl);
                if (u_last) { /* J2: F of the last converged step (FiniteStrainT::SetGlobalShape) and the element's history */
                    grad_u(ule, dN[ip], Fl);
                    Fl[0] += 1.0; Fl[4] += 1.0; Fl[8] += 1.0;
                }
                err = fs_material(m, F, Fl, j2 ? j2 + 8 * e : NULL, ip, alloc ? alloc + e : NULL, iteration, sig, NULL);
                if (err) { free(count); return err; }
            }
            for (int a = 0; a < 8; a++) {
                const double E = 0.125 * (1.0 + sqrt3 * (sx[a] * sx[ip] + sy[a] * sy[ip] + sz[a] * sz[ip]));
                for (int I = 0; I < 6; I++) nodal[a][I] += E * sig[I];
            }
        }
        for (int a = 0; a < 8; a++) {
            count[c[a]]++;
            for (int I = 0; I < 6; I++) out[6 * (int64_t)c[a] + I] += nodal[a][I];
        }
    }
    for (int64_t n = 0; n < nn; n++)
        if (count[n] > 0) {
            const double s = 1.0 / count[n];
            for (int I = 0; I < 6; I++) out[6 * n + I] *= s;
        }
    free(count);
    return ORC_OK;
}

int orc_nodal_stress(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X, const double* u,
                     double* out)
{
    return nodal_stress_core(form, m, ne, conn, nn, X, u, NULL, NULL, NULL, 0, out);
}
/* the same output for a history material (J2Simo3D): SolidElementT::ComputeOutput evaluates s_ij at every integration point with the
 * element's history and the last converged displacement, BEFORE the step's history update (FEManagerT::CloseStep :639-645) */
int orc_nodal_stress_history(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X,
                             const double* u, const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* out)
{
    return nodal_stress_core(form, m, ne, conn, nn, X, u, u_last, j2, alloc, iteration, out);
}

/* one evaluation of the <explicit_solid> laws for known-answer tests (the reference's tests/materials/test_ExplJ2Plasticity.cpp calls
 * ExplJ2PlasticityT::ComputeStress3D the same way): F row-major, h[16] updated in place (EXPL_J2 only), sig Voigt 11,22,33,23,13,12 */
void orc_explicit_material_stress(const orc_material_t* m, const double* F, double* h, double* sig)
{
    if (m->kind == ORC_EXPL_J2) xs_j2(m, F, h, sig);
    else xs_neo_hookean(m, F, sig);
}

/* ---- contact_3D_penalty ------------------------------------------------------------------------------------- */
/* Contact3DT::Set_dn_du (Contact3DT.cpp:176-220): d(a x b)/du for a = x2 - x1, b = x3 - x1, as a 3 x 12 matrix stored by columns
 * (dMatrixT is column-major); the striker's columns are zero */
static void contact_dn_du(const double* x1, const double* x2, const double* x3, double dn[12][3])
{
    const double col[12][3] = {
        {0, -x2[2] + x3[2], x2[1] - x3[1]}, {x2[2] - x3[2], 0, -x2[0] + x3[0]}, {-x2[1] + x3[1], x2[0] - x3[0], 0},
        {0, x1[2] - x3[2], -x1[1] + x3[1]}, {-x1[2] + x3[2], 0, x1[0] - x3[0]}, {x1[1] - x3[1], -x1[0] + x3[0], 0},
        {0, -x1[2] + x2[2], x1[1] - x2[1]}, {x1[2] - x2[2], 0, -x1[0] + x2[0]}, {-x1[1] + x2[1], x1[0] - x2[0], 0},
        {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    memcpy(dn, col, sizeof col);
}

/* Contact3DT::Intersect (Contact3DT.cpp:336-391) */
static int contact_intersect(const double* x1, const double* x2, const double* x3, const double* xs, double* h_out)
{
    double a[3], b[3], c[3], n[3], xsp[3];
    for (int i = 0; i < 3; i++) { a[i] = x2[i] - x1[i]; b[i] = x3[i] - x1[i]; c[i] = xs[i] - x1[i]; }
    n[0] = a[1] * b[2] - a[2] * b[1];
    n[1] = a[2] * b[0] - a[0] * b[2];
    n[2] = a[0] * b[1] - a[1] * b[0];
    const double mag = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int i = 0; i < 3; i++) n[i] /= mag;
    const double h = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
    *h_out = h;
    for (int i = 0; i < 3; i++) xsp[i] = 1.0 * xs[i] + (-h) * n[i];
    const double dist_tol = sqrt(mag) / 2.0, area_tol = mag / 50.0;
    if (fabs(h) > dist_tol) return 0;
    const double* corner[3] = {x1, x2, x3};
    const double* next[3] = {x2, x3, x1};
    for (int e = 0; e < 3; e++) { /* edges 1-2, 2-3, 3-1: the projection must lie on the inner side of each */
        double edge[3], xis[3], ni[3];
        for (int i = 0; i < 3; i++) { edge[i] = next[e][i] - corner[e][i]; xis[i] = xsp[i] - corner[e][i]; }
        ni[0] = edge[1] * xis[2] - edge[2] * xis[1];
        ni[1] = edge[2] * xis[0] - edge[0] * xis[2];
        ni[2] = edge[0] * xis[1] - edge[1] * xis[0];
        if (n[0] * ni[0] + n[1] * ni[1] + n[2] * ni[2] < -area_tol) return 0;
    }
    return 1;
}

int orc_contact_search(int64_t nfacets, const int32_t* facets, const int32_t* facet_surface, int64_t nstrikers, const int32_t* strikers,
                       int64_t nn, const double* x, int32_t* hit, double* gap)
{
    (void)nn;
    int nsurf = 0;
    for (int64_t f = 0; f < nfacets; f++) nsurf = facet_surface[f] + 1 > nsurf ? facet_surface[f] + 1 : nsurf;
    int nactive = 0;
    for (int64_t s = 0; s < nstrikers; s++) {
        const int32_t tag = strikers[s];
        hit[s] = -1;
        double best = 0.0;
        for (int64_t f = 0; f < nfacets; f++) { /* surfaces in order, facets in order: the reference's loops */
            /* no self contact per surface: surface.HasValue(strikertag) */
            int self = 0;
            for (int64_t q = 0; q < nfacets && !self; q++)
                if (facet_surface[q] == facet_surface[f] && (facets[3 * q] == tag || facets[3 * q + 1] == tag || facets[3 * q + 2] == tag)) self = 1;
            if (self) continue;
            double h;
            if (!contact_intersect(x + 3 * (int64_t)facets[3 * f], x + 3 * (int64_t)facets[3 * f + 1], x + 3 * (int64_t)facets[3 * f + 2],
                                   x + 3 * (int64_t)tag, &h))
                continue;
            if (hit[s] < 0 || fabs(h) < fabs(best)) { /* first time to a facet, or a closer projection (:300-318) */
                hit[s] = (int32_t)f;
                best = h;
            }
        }
        if (gap) gap[s] = best;
        nactive += hit[s] >= 0;
    }
    (void)nsurf;
    return nactive;
}

int orc_contact_force(int64_t npairs, const int32_t* pairs, const double* area, double K, double mu, double eps, double visc, double constKd,
                      int64_t nn, const double* X, const double* u, const double* v, double* f, double* h_max_out)
{
    int num_contact = 0;
    double h_max = 0.0;
    (void)nn;
    for (int64_t p = 0; p < npairs; p++) {
        const int32_t* nd = pairs + 4 * p;
        double x[4][3]; /* PenaltyContact3DT.cpp:297-311: X + constKd u of facet nodes 1-3 and the striker */
        for (int a = 0; a < 4; a++)
            for (int i = 0; i < 3; i++) x[a][i] = X[3 * (int64_t)nd[a] + i] + constKd * u[3 * (int64_t)nd[a] + i];
        double a_[3], b_[3], n[3], c[3];
        for (int i = 0; i < 3; i++) { a_[i] = x[1][i] - x[0][i]; b_[i] = x[2][i] - x[0][i]; }
        n[0] = a_[1] * b_[2] - a_[2] * b_[1];
        n[1] = a_[2] * b_[0] - a_[0] * b_[2];
        n[2] = a_[0] * b_[1] - a_[1] * b_[0];
        const double mag = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int i = 0; i < 3; i++) n[i] /= mag;
        for (int i = 0; i < 3; i++) c[i] = x[3][i] - (x[0][i] + x[1][i] + x[2][i]) / 3.0;
        const double h = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
        if (!(h < 0.0)) continue; /* :325 */
        num_contact++;
        if (h < h_max) h_max = h;
        const double dphi = -K * h * area[p]; /* :332-333 */
        double rhs[12];
        /* fdc_du^T n (:136-172: -1/3 on the facet nodes, 1 on the striker), scaled by dphi */
        for (int a = 0; a < 3; a++)
            for (int i = 0; i < 3; i++) rhs[3 * a + i] = dphi * (-(1.0 / 3.0) * n[i]);
        for (int i = 0; i < 3; i++) rhs[9 + i] = dphi * n[i];
        /* d_normal (:339-344): V1_j = c . (n n^T - 1) dn_du[:, j]; RHS += -dphi / mag V1 */
        double dn[12][3];
        contact_dn_du(x[0], x[1], x[2], dn);
        for (int j = 0; j < 12; j++) {
            const double nd_ = n[0] * dn[j][0] + n[1] * dn[j][1] + n[2] * dn[j][2];
            double v1 = 0.0;
            for (int i = 0; i < 3; i++) v1 += (n[i] * nd_ - dn[j][i]) * c[i];
            rhs[j] += -dphi / mag * v1;
        }
        if (v) {
            double vs[3], vf[3];
            for (int i = 0; i < 3; i++) {
                vs[i] = v[3 * (int64_t)nd[3] + i];
                vf[i] = (v[3 * (int64_t)nd[0] + i] + v[3 * (int64_t)nd[1] + i] + v[3 * (int64_t)nd[2] + i]) / 3.0;
            }
            if (mu > 0.0) { /* explicit branch, :375-426: f_t = -mu |f_n| v_t / sqrt(|v_t|^2 + eps^2) */
                double vr[3], vt[3];
                for (int i = 0; i < 3; i++) vr[i] = vs[i] - vf[i];
                const double vrn = vr[0] * n[0] + vr[1] * n[1] + vr[2] * n[2];
                for (int i = 0; i < 3; i++) vt[i] = vr[i] - vrn * n[i];
                const double vt2 = vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2];
                const double inv_smooth = 1.0 / sqrt(vt2 + eps * eps);
                const double s = -mu * fabs(dphi) * inv_smooth;
                const double third = 1.0 / 3.0;
                for (int i = 0; i < 3; i++) {
                    const double ft = s * vt[i];
                    rhs[i] += -ft * third;
                    rhs[3 + i] += -ft * third;
                    rhs[6 + i] += -ft * third;
                    rhs[9 + i] += ft;
                }
            }
            if (visc > 0.0) { /* :463-488: f = -c v_n area along n */
                const double vrn = (vs[0] - vf[0]) * n[0] + (vs[1] - vf[1]) * n[1] + (vs[2] - vf[2]) * n[2];
                const double fv = -visc * vrn * area[p];
                const double third = 1.0 / 3.0;
                for (int i = 0; i < 3; i++) {
                    rhs[i] += -fv * n[i] * third;
                    rhs[3 + i] += -fv * n[i] * third;
                    rhs[6 + i] += -fv * n[i] * third;
                    rhs[9 + i] += fv * n[i];
                }
            }
        }
        for (int a = 0; a < 4; a++) /* SolverT::AssembleRHS, pair by pair */
            for (int i = 0; i < 3; i++) f[3 * (int64_t)nd[a] + i] += rhs[3 * a + i];
    }
    if (h_max_out) *h_max_out = h_max;
    return num_contact;
}

/* ---- natural_bc tractions: ContinuumElementT::ApplyTractionBC (ContinuumElementT.cpp:514-665) on Hex8 facets ------------------
 * facet nodes HexahedronT::NodesOnFacet (HexahedronT.cpp:1913-1918); 4-node quad facet shape with the 2x2 rule the hexahedron gets
 * (DomainIntegrationT.cpp:117-131; QuadT.cpp:75-81,379-389, weights 1); surface Jacobian |x,r x x,s| on the INITIAL coordinates
 * (LocalArrayT::kInitCoords, :533) and, for coordinate_system="local", Q = [x,r/|x,r| , n x t1 , n] (ParentDomainT.cpp:362-422).
 * tract[card][4][3] are the nodal traction vectors (Traction_CardT, already in facet-node order), scale = schedule value. */
int orc_traction_force(int64_t ncards, const int32_t* elem, const int32_t* facet, const int32_t* conn, const double* X,
                       const double* tract, int coord_system, double scale, double* f)
{
    static const int fn[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};
    static const double qr[4] = {-1.0, 1.0, 1.0, -1.0}, qs[4] = {-1.0, -1.0, 1.0, 1.0};
    const double g = 1.0 / sqrt(3.0);
    for (int64_t c = 0; c < ncards; c++) {
        if (facet[c] < 0 || facet[c] > 5) return -1;
        int node[4];
        double x[4][3], t[4][3], rhs[4][3];
        for (int a = 0; a < 4; a++) {
            node[a] = conn[elem[c] * 8 + fn[facet[c]][a]];
            for (int i = 0; i < 3; i++) {
                x[a][i] = X[node[a] * 3 + i];
                t[a][i] = scale * tract[(c * 4 + a) * 3 + i];
                rhs[a][i] = 0.0;
            }
        }
        for (int j = 0; j < 4; j++) {
            const double r = g * qr[j], s = g * qs[j];
            double Na[4], m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0}, tip[3] = {0, 0, 0};
            for (int a = 0; a < 4; a++) {
                const double tr = 1.0 + qr[a] * r, ts = 1.0 + qs[a] * s;
                Na[a] = 0.25 * tr * ts;
                const double dr = 0.25 * qr[a] * ts, ds = 0.25 * tr * qs[a];
                for (int i = 0; i < 3; i++) {
                    m1[i] += x[a][i] * dr;
                    m2[i] += x[a][i] * ds;
                    tip[i] += Na[a] * t[a][i];
                }
            }
            double n3[3] = {m1[1] * m2[2] - m1[2] * m2[1], m1[2] * m2[0] - m1[0] * m2[2], m1[0] * m2[1] - m1[1] * m2[0]};
            const double jn = sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
            double tj[3] = {tip[0], tip[1], tip[2]};
            if (coord_system == 1) { /* Traction_CardT::kLocal */
                const double j1 = sqrt(m1[0] * m1[0] + m1[1] * m1[1] + m1[2] * m1[2]);
                if (jn <= 0.0 || j1 <= 0.0) return ORC_BAD_JACOBIAN;
                double n1[3], n2[3];
                for (int i = 0; i < 3; i++) {
                    n3[i] /= jn;
                    n1[i] = m1[i] / j1;
                }
                n2[0] = n3[1] * n1[2] - n3[2] * n1[1];
                n2[1] = n3[2] * n1[0] - n3[0] * n1[2];
                n2[2] = n3[0] * n1[1] - n3[1] * n1[0];
                for (int i = 0; i < 3; i++) tj[i] = n1[i] * tip[0] + n2[i] * tip[1] + n3[i] * tip[2];
            }
            for (int l = 0; l < 3; l++) {
                const double fact = jn * tj[l]; /* weights are 1 */
                for (int a = 0; a < 4; a++) rhs[a][l] += fact * Na[a];
            }
        }
        for (int a = 0; a < 4; a++)
            for (int i = 0; i < 3; i++) f[node[a] * 3 + i] += rhs[a][i];
    }
    return ORC_OK;
}

/* ---- inertia: ContinuumElementT::FormMass / FormMa (ContinuumElementT.cpp:678-866, 868-1002) -----------------------------------
 * mass_type 1 = kConsistentMass: M_(a i)(b j) = delta_ij sum_ip rho w detJ0 N_a N_b (reference configuration);
 * mass_type 2 = kLumpedMass: the HRZ diagonal of orc_element_lumped_mass.  Element matrix is column-major 24 x 24 like Ke. */
static int element_mass(double density, int mass_type, const double X[8][3], double Me[576])
{
    for (int i = 0; i < 576; i++) Me[i] = 0.0;
    if (mass_type == 2) {
        double me[8];
        int err = orc_element_lumped_mass(density, X, me);
        if (err) return err;
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) Me[(3 * a + i) * 25] = me[a];
        return ORC_OK;
    }
    double Na[8][8], DNa[8][3][8], w[8], dN[8][3][8], det[8];
    orc_hex8_parent(Na, DNa, w);
    int err = orc_hex8_shape(X, dN, det);
    if (err) return err;
    for (int ip = 0; ip < 8; ip++) {
        const double temp = density * w[ip] * det[ip];
        for (int a = 0; a < 8; a++)
            for (int b = 0; b < 8; b++)
                for (int i = 0; i < 3; i++) Me[(3 * a + i) + 24 * (3 * b + i)] += temp * Na[ip][a] * Na[ip][b];
    }
    return ORC_OK;
}
/* f[nn][3] += scale * sum_e M_e a_e, accumulated element by element as SolidElementT::ElementRHSDriver does (:1243-1265):
 * consistent: per ip the interpolated acceleration times rho w detJ0 N_a (:940-966); lumped: diagonal times nodal value (:969-997) */
int orc_inertial_force(double density, int mass_type, int64_t ne, const int32_t* conn, const double* X, const double* acc,
                       double scale, double* f)
{
    double Na[8][8], DNa[8][3][8], w[8];
    orc_hex8_parent(Na, DNa, w);
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], ae[8][3], fe[24] = {0};
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        gather(c, acc, ae);
        if (mass_type == 2) {
            double me[8];
            int err = orc_element_lumped_mass(scale * density, Xe, me);
            if (err) return err;
            for (int a = 0; a < 8; a++)
                for (int i = 0; i < 3; i++) fe[3 * a + i] += ae[a][i] * me[a];
        } else {
            double dN[8][3][8], det[8];
            int err = orc_hex8_shape(Xe, dN, det);
            if (err) return err;
            for (int ip = 0; ip < 8; ip++) {
                double aip[3] = {0, 0, 0};
                for (int a = 0; a < 8; a++)
                    for (int i = 0; i < 3; i++) aip[i] += Na[ip][a] * ae[a][i];
                const double temp = scale * density * w[ip] * det[ip];
                for (int a = 0; a < 8; a++) {
                    const double temp2 = temp * Na[ip][a];
                    for (int i = 0; i < 3; i++) fe[3 * a + i] += temp2 * aip[i];
                }
            }
        }
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) f[3 * (int64_t)c[a] + i] += fe[3 * a + i];
    }
    return ORC_OK;
}
/* val += constM * M on the CSR of orc_csr_structure (SolidElementT::ElementLHSDriver with formM, SolidElementT.cpp:1100-1154) */
int orc_assemble_mass(double density, int mass_type, double constM, int64_t ne, const int32_t* conn, const double* X,
                      const int32_t* eqnos, const int64_t* rowptr, const int32_t* colind, double* val)
{
    for (int64_t e = 0; e < ne; e++) {
        double Xe[8][3], Me[576];
        const int32_t* c = conn + 8 * e;
        gather(c, X, Xe);
        int err = element_mass(constM * density, mass_type, Xe, Me);
        if (err) return err;
        int32_t eq[24];
        for (int a = 0; a < 8; a++)
            for (int i = 0; i < 3; i++) eq[3 * a + i] = eqnos[3 * (int64_t)c[a] + i];
        for (int r = 0; r < 24; r++) {
            if (eq[r] <= 0) continue;
            const int64_t row = eq[r] - 1;
            for (int cc = 0; cc < 24; cc++) {
                if (eq[cc] <= 0 || Me[r + 24 * cc] == 0.0) continue;
                const int32_t col = eq[cc] - 1;
                int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
                while (lo < hi) { int64_t mid = (lo + hi) / 2; if (colind[mid] < col) lo = mid + 1; else hi = mid; }
                if (colind[lo] != col) return -1;
                val[lo] += Me[r + 24 * cc];
            }
        }
    }
    return ORC_OK;
}
