// tb2_node_update.cuh -- the node kernel body of the explicit step (K5).
#pragma once
#include "../../include/tahoe_b200.h"
#include "tb2_math.cuh"

namespace tb2 {

// nExplicitCD::Predictor (nExplicitCD.cpp:72-96) / Corrector (:98-139) with explicit roundings, so that every kernel that
// carries them produces bit-identical fields
TB2_DEV void cd_predict(double dt, double& d, double& v, double a)
{
    d = __fma_rn(dt, v, d);
    d = __fma_rn(__dmul_rn(__dmul_rn(0.5, dt), dt), a, d);
    v = __fma_rn(__dmul_rn(0.5, dt), a, v);
}
TB2_DEV void cd_correct(double dt, double& v, double& a, double upd)
{
    v = __fma_rn(__dmul_rn(0.5, dt), upd, v);
    a = __dadd_rn(a, upd);
}

struct StepConsts {
    double dt, fext_scale, next_value_scale;
};

// One nodal dof from its assembled internal force f: R = s fext - f (LinearSolver::Solve), upd = M^-1 R on free dofs
// (DiagonalMatrixT.cpp:313-323, FieldT::AssembleUpdate: prescribed dofs get 0), corrector; NEXT_PREDICTOR: the next step's
// predictor and ConsistentKBC (nExplicitCD.cpp:20-69).  The acceleration the predictor left behind is 0 on every dof
// (nExplicitCD.cpp:93), so it is not read; after a fused next predictor it is 0 again and not written.
template <bool NEXT_PREDICTOR>
TB2_DEV void cd_update_dof(const StepConsts& u, const unsigned char c, const double f, const double fx, const double mi, const double bcv,
                           double& di, double& vi, double& ai)
{
    const double R = __dsub_rn(fx, f); // fx: the dof's external load of this step (nodal_load)
    const double upd = c ? 0.0 : __dmul_rn(R, mi);
    ai = 0.0;
    cd_correct(u.dt, vi, ai, upd);
    if (NEXT_PREDICTOR) {
        cd_predict(u.dt, di, vi, ai);
        ai = 0.0;
        if (c == TB2_BC_FIX) { di = 0.0; vi = 0.0; }
        else if (c == TB2_BC_DSP) di = u.next_value_scale * bcv;
    }
}

struct NodeArrays {
    const double* fext; // null: no external force
    const double* fadd; // null, or a second, unscaled load re-formed every step (contact forces: residual contributions of other groups)
    const double* minv;
    const unsigned char* code;
    const double* bcval;
    double* d;
    double* v;
    double* a;
    double* fint;
};

// the external load of dof q in this step: s fext (+ fadd).  Without the second array the expression is the one it always was.
TB2_DEV double nodal_load(const StepConsts& sc, const NodeArrays& na, const int64_t q)
{
    const double fx = na.fext ? __dmul_rn(sc.fext_scale, na.fext[q]) : 0.0;
    return na.fadd ? __dadd_rn(fx, na.fadd[q]) : fx;
}

// one node: gather fint (GATHER) or take it from fint[] (multi-GPU: the interface-summed force), then the update above.
// GATHER reads the node's incidence from the fixed-width table inc8 (one 32-byte load).  Per node and step the steady state (NEXT_PREDICTOR) moves
// d, v in and out, 1/m, the boundary codes and fext (when there is one) in: a stays 0 and fint is written by the last step only.
template <bool GATHER, bool NEXT_PREDICTOR>
TB2_DEV void cd_node_update_one(const int64_t n, const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int4* __restrict__ inc8,
                                const double* __restrict__ fe, int64_t stride, const StepConsts& sc, const NodeArrays& na)
{
    double f[3] = {0.0, 0.0, 0.0};
    if (GATHER) {
        int ent[8];
        const int4 lo = __ldg(inc8 + 2 * n), hi = __ldg(inc8 + 2 * n + 1);
        ent[0] = lo.x; ent[1] = lo.y; ent[2] = lo.z; ent[3] = lo.w;
        ent[4] = hi.x; ent[5] = hi.y; ent[6] = hi.z; ent[7] = hi.w;
        // ascending-element order = the reference's serial assembly order (SolverT::AssembleRHS, SolverT.cpp:446-477).  The loop is
        // left to the compiler under the kernel's 40-register budget (6 CTAs of 256 threads per SM): twice the resident warps moved
        // the kernel from 0.71 to 0.86 of the HBM peak, which issuing all 24 gathers before the first use (74 registers) did not
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (ent[q] >= 0) {
                const int64_t e = ent[q] >> 3;
                const int a3 = 3 * (ent[q] & 7);
                f[0] += __ldg(fe + (int64_t)(a3)*stride + e);
                f[1] += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
                f[2] += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
            }
        if (ent[7] >= 0) // an irregular vertex with more than 8 incident elements: the rest of its list
            for (int k = inc_ptr[n] + 8, k1 = inc_ptr[n + 1]; k < k1; k++) {
                const int en = __ldg(inc + k);
                const int64_t e = en >> 3;
                const int a3 = 3 * (en & 7);
                f[0] += __ldg(fe + (int64_t)(a3)*stride + e);
                f[1] += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
                f[2] += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
            }
    } else {
        f[0] = na.fint[3 * n];
        f[1] = na.fint[3 * n + 1];
        f[2] = na.fint[3 * n + 2];
    }
    // nodal operands after the gather: 40 registers, 6 CTAs of 256 threads per SM -- the resident warps, not the loads one thread
    // keeps in flight, hide the latency of this HBM-bound kernel
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        const unsigned char c = na.code[q];
        double di = NEXT_PREDICTOR ? na.d[q] : 0.0, vi = na.v[q], ai;
        const double bcv = (NEXT_PREDICTOR && c == TB2_BC_DSP) ? na.bcval[q] : 0.0;
        cd_update_dof<NEXT_PREDICTOR>(sc, c, f[i], nodal_load(sc, na, q), na.minv[q], bcv, di, vi, ai);
        if (NEXT_PREDICTOR) na.d[q] = di;
        else {
            na.a[q] = ai;
            if (GATHER) na.fint[q] = f[i];
        }
        na.v[q] = vi;
    }
}

} // namespace tb2
