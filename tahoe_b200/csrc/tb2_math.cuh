// tb2_math.cuh -- register-resident 3x3 / symmetric-tensor helpers and the trilinear "mode" form of the
// Hex8 parent domain.  Everything is __forceinline__ with compile-time indices so that nvcc keeps the
// small arrays in registers (check with -Xptxas -v: no local memory).
//
// Hex8 in mode form.  Tahoe evaluates, per integration point, J = sum_a X_a (x) dN_a/dxi with the 8x3 table
// of HexahedronT.cpp:421-430 (ParentDomainT::Jacobian, ParentDomainT.cpp:99-189): 72 FMA per 3x3 and the
// table has to live somewhere.  Because N_a = (1+xi_a xi)(1+eta_a eta)(1+zeta_a zeta)/8, any nodal field
// v_a is   v(xi) = c0 + c1 xi + c2 eta + c4 zeta + c3 xi eta + c5 xi zeta + c6 eta zeta + c7 xi eta zeta
// with c = (Walsh-Hadamard transform of v over the 8 vertices)/8, so
//     dv/dxi   = c1 + c3 eta + c5 zeta + c7 eta zeta      (3 FMA instead of 8)
// and the B^T sigma accumulation is the transposed operation: accumulate 7 mode coefficients per
// component over the integration points, one inverse transform per element at the end.  Same algebra as
// the reference, different summation order (differences ~1e-16 relative; parity bar is 1e-10).
#pragma once
#include <cstdint>

#define TB2_DEV __device__ __forceinline__
#define TB2_G 0.57735026918962576451 /* 1/sqrt(3): HexahedronT.cpp:1540-1547 */

namespace tb2 {

// ---------------------------------------------------------------- 3x3 (row-major a[i][j])
TB2_DEV double det3(const double (&a)[3][3])
{
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
// adjugate: adj(a) = det(a) * inverse(a); returns det
TB2_DEV double adj3(const double (&a)[3][3], double (&c)[3][3])
{
    c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    c[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    c[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    c[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    c[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    c[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    c[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    return a[0][0] * c[0][0] + a[0][1] * c[1][0] + a[0][2] * c[2][0];
}
TB2_DEV void scale3(double (&a)[3][3], double s)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) a[i][j] *= s;
}
// c = a b
TB2_DEV void mul3(const double (&a)[3][3], const double (&b)[3][3], double (&c)[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
// c = a b^T
TB2_DEV void mul3_abt(const double (&a)[3][3], const double (&b)[3][3], double (&c)[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j] = a[i][0] * b[j][0] + a[i][1] * b[j][1] + a[i][2] * b[j][2];
}

// ---------------------------------------------------------------- symmetric tensors, order 11,22,33,23,13,12
TB2_DEV void sym_to_mat(const double (&s)[6], double (&a)[3][3])
{
    a[0][0] = s[0]; a[1][1] = s[1]; a[2][2] = s[2];
    a[1][2] = a[2][1] = s[3];
    a[0][2] = a[2][0] = s[4];
    a[0][1] = a[1][0] = s[5];
}
TB2_DEV double sym_trace(const double (&s)[6]) { return s[0] + s[1] + s[2]; }
TB2_DEV void sym_dev(double (&s)[6])
{
    double p = (s[0] + s[1] + s[2]) * (1.0 / 3.0);
    s[0] -= p; s[1] -= p; s[2] -= p;
}
TB2_DEV double sym_det(const double (&s)[6])
{
    return s[0] * (s[1] * s[2] - s[3] * s[3]) - s[5] * (s[5] * s[2] - s[3] * s[4]) + s[4] * (s[5] * s[3] - s[1] * s[4]);
}
TB2_DEV double sym_norm2(const double (&s)[6]) // s:s (dSymMatrixT::ScalarProduct)
{
    return s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
}
// b = f f^T  (FSSolidMatT::Compute_b, FSSolidMatT.cpp:402-434)
TB2_DEV void sym_fft(const double (&f)[3][3], double (&b)[6])
{
    b[0] = f[0][0] * f[0][0] + f[0][1] * f[0][1] + f[0][2] * f[0][2];
    b[1] = f[1][0] * f[1][0] + f[1][1] * f[1][1] + f[1][2] * f[1][2];
    b[2] = f[2][0] * f[2][0] + f[2][1] * f[2][1] + f[2][2] * f[2][2];
    b[3] = f[1][0] * f[2][0] + f[1][1] * f[2][1] + f[1][2] * f[2][2];
    b[4] = f[0][0] * f[2][0] + f[0][1] * f[2][1] + f[0][2] * f[2][2];
    b[5] = f[0][0] * f[1][0] + f[0][1] * f[1][1] + f[0][2] * f[1][2];
}
// c = f^T f
TB2_DEV void sym_ftf(const double (&f)[3][3], double (&c)[6])
{
    c[0] = f[0][0] * f[0][0] + f[1][0] * f[1][0] + f[2][0] * f[2][0];
    c[1] = f[0][1] * f[0][1] + f[1][1] * f[1][1] + f[2][1] * f[2][1];
    c[2] = f[0][2] * f[0][2] + f[1][2] * f[1][2] + f[2][2] * f[2][2];
    c[3] = f[0][1] * f[0][2] + f[1][1] * f[1][2] + f[2][1] * f[2][2];
    c[4] = f[0][0] * f[0][2] + f[1][0] * f[1][2] + f[2][0] * f[2][2];
    c[5] = f[0][0] * f[0][1] + f[1][0] * f[1][1] + f[2][0] * f[2][1];
}
// r = q s q^T, s symmetric (dSymMatrixT::MultQBQT)
TB2_DEV void sym_qsqt(const double (&q)[3][3], const double (&s)[6], double (&r)[6])
{
    double t[3][3]; // t = q S
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[i][0] = q[i][0] * s[0] + q[i][1] * s[5] + q[i][2] * s[4];
        t[i][1] = q[i][0] * s[5] + q[i][1] * s[1] + q[i][2] * s[3];
        t[i][2] = q[i][0] * s[4] + q[i][1] * s[3] + q[i][2] * s[2];
    }
    r[0] = t[0][0] * q[0][0] + t[0][1] * q[0][1] + t[0][2] * q[0][2];
    r[1] = t[1][0] * q[1][0] + t[1][1] * q[1][1] + t[1][2] * q[1][2];
    r[2] = t[2][0] * q[2][0] + t[2][1] * q[2][1] + t[2][2] * q[2][2];
    r[3] = t[1][0] * q[2][0] + t[1][1] * q[2][1] + t[1][2] * q[2][2];
    r[4] = t[0][0] * q[2][0] + t[0][1] * q[2][1] + t[0][2] * q[2][2];
    r[5] = t[0][0] * q[1][0] + t[0][1] * q[1][1] + t[0][2] * q[1][2];
}

// ---------------------------------------------------------------- Hex8 modes
// Tahoe node a -> lexicographic vertex index (bit0 = xi sign, bit1 = eta sign, bit2 = zeta sign); HexahedronT.cpp:23-25
__device__ __constant__ const int kLexOfNode[8] = {0, 1, 3, 2, 4, 5, 7, 6};

// in-place Walsh-Hadamard butterfly over the 8 lexicographic vertex values: v[bits] <- sum_idx sign(idx,bits) v[idx]
TB2_DEV void wht8(double (&v)[8])
{
#pragma unroll
    for (int h = 1; h < 8; h <<= 1)
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (!(i & h)) {
                double lo = v[i], hi = v[i | h];
                v[i] = lo + hi;
                v[i | h] = hi - lo;
            }
}
// synthesis: v[idx] <- sum_bits sign(idx,bits) v[bits]
TB2_DEV void iwht8(double (&v)[8])
{
#pragma unroll
    for (int h = 1; h < 8; h <<= 1)
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (!(i & h)) {
                double lo = v[i], hi = v[i | h];
                v[i] = lo - hi;
                v[i | h] = lo + hi;
            }
}

// scaled modes of one nodal vector field: m[k][i], k = bits-1 (bits 1..7), i = component.
// m1,m2,m4 carry 1/8, m3,m5,m6 carry g/8, m7 carries g^2/8 so that at the integration point with signs
// (s0,s1,s2):  d v_i/d xi = m1 + s1 m3 + s2 m5 + s1 s2 m7   etc.
struct Modes {
    double m[7][3];
};

// gather the 8 nodes of an element (node ids n[] in Tahoe order) from a [node][3] array and transform
TB2_DEV void load_modes(const double* __restrict__ field, const int (&n)[8], Modes& out)
{
    const double s1 = 0.125, sg = 0.125 * TB2_G, sgg = 0.125 * TB2_G * TB2_G;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double v[8];
        v[0] = __ldg(field + 3 * (int64_t)n[0] + i);
        v[1] = __ldg(field + 3 * (int64_t)n[1] + i);
        v[3] = __ldg(field + 3 * (int64_t)n[2] + i);
        v[2] = __ldg(field + 3 * (int64_t)n[3] + i);
        v[4] = __ldg(field + 3 * (int64_t)n[4] + i);
        v[5] = __ldg(field + 3 * (int64_t)n[5] + i);
        v[7] = __ldg(field + 3 * (int64_t)n[6] + i);
        v[6] = __ldg(field + 3 * (int64_t)n[7] + i);
        wht8(v);
        out.m[0][i] = v[1] * s1;
        out.m[1][i] = v[2] * s1;
        out.m[2][i] = v[3] * sg;
        out.m[3][i] = v[4] * s1;
        out.m[4][i] = v[5] * sg;
        out.m[5][i] = v[6] * sg;
        out.m[6][i] = v[7] * sgg;
    }
}

// d v_i / d xi_k at the integration point with signs (s0,s1,s2) = (+-1): D[i][k]
TB2_DEV void mode_gradient(const Modes& c, double s0, double s1, double s2, double (&D)[3][3])
{
    const double s12 = s1 * s2, s02 = s0 * s2, s01 = s0 * s1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        D[i][0] = c.m[0][i] + s1 * c.m[2][i] + s2 * c.m[4][i] + s12 * c.m[6][i];
        D[i][1] = c.m[1][i] + s0 * c.m[2][i] + s2 * c.m[5][i] + s02 * c.m[6][i];
        D[i][2] = c.m[3][i] + s0 * c.m[4][i] + s1 * c.m[5][i] + s01 * c.m[6][i];
    }
}

// accumulate G[i][k] (= weight * det * sigma . dxi_k/dx) into force modes: transposed mode_gradient
TB2_DEV void mode_accumulate(Modes& A, double s0, double s1, double s2, const double (&G)[3][3])
{
    const double s12 = s1 * s2, s02 = s0 * s2, s01 = s0 * s1;
    // explicit FMA chains: one instruction per term (a sum of products added to the accumulator costs one more)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        A.m[0][i] += G[i][0];
        A.m[1][i] += G[i][1];
        A.m[3][i] += G[i][2];
        A.m[2][i] = fma(s0, G[i][1], fma(s1, G[i][0], A.m[2][i]));
        A.m[4][i] = fma(s0, G[i][2], fma(s2, G[i][0], A.m[4][i]));
        A.m[5][i] = fma(s1, G[i][2], fma(s2, G[i][1], A.m[5][i]));
        A.m[6][i] = fma(s01, G[i][2], fma(s02, G[i][1], fma(s12, G[i][0], A.m[6][i])));
    }
}

// nodal values from accumulated force modes: f[a][i] in Tahoe node order
TB2_DEV void modes_to_nodes(const Modes& A, int i, double (&f)[8])
{
    const double s1 = 0.125, sg = 0.125 * TB2_G, sgg = 0.125 * TB2_G * TB2_G;
    double v[8];
    v[0] = 0.0;
    v[1] = A.m[0][i] * s1;
    v[2] = A.m[1][i] * s1;
    v[3] = A.m[2][i] * sg;
    v[4] = A.m[3][i] * s1;
    v[5] = A.m[4][i] * sg;
    v[6] = A.m[5][i] * sg;
    v[7] = A.m[6][i] * sgg;
    iwht8(v);
    f[0] = v[0]; f[1] = v[1]; f[2] = v[3]; f[3] = v[2];
    f[4] = v[4]; f[5] = v[5]; f[6] = v[7]; f[7] = v[6];
}

// integration-point signs in Tahoe's IP order (= node order, HexahedronT.cpp:1540-1547)
TB2_DEV void ip_signs(int ip, double& s0, double& s1, double& s2)
{
    // ra = -,+,+,-,-,+,+,-   sa = -,-,+,+,-,-,+,+   ta = -,-,-,-,+,+,+,+
    s0 = (((ip + 1) >> 1) & 1) ? 1.0 : -1.0;
    s1 = ((ip >> 1) & 1) ? 1.0 : -1.0;
    s2 = ((ip >> 2) & 1) ? 1.0 : -1.0;
}

} // namespace tb2
