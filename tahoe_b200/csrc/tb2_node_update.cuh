// tb2_node_update.cuh -- the node kernel body of the explicit step (K5), shared by the stand-alone kernel (tb2_explicit.cu) and
// the fused element + node launches of the slab pipeline (tb2_elements.cu).
#pragma once
#include "tb2_math.cuh"

namespace tb2 {

// nExplicitCD::Predictor (nExplicitCD.cpp:72-96) / Corrector (:98-139) with explicit roundings, so that the stand-alone and
// the fused kernels produce bit-identical fields
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

// one node: gather fint, R = s*fext - fint, upd = minv*R on free dofs, corrector; optionally the next predictor (see k_cd_node_update)
template <bool GATHER, bool NEXT_PREDICTOR>
TB2_DEV void cd_node_update_one(const int64_t n, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                const int4* __restrict__ inc8, const double* __restrict__ fe, int64_t stride, double dt,
                                double fext_scale, double next_value_scale, const double* __restrict__ fext,
                                const double* __restrict__ minv, const unsigned char* __restrict__ code,
                                const double* __restrict__ bcval, double* __restrict__ d,
                                double* __restrict__ v, double* __restrict__ a, double* __restrict__ fint,
                                const int* __restrict__ skip_slot)
{
    if (skip_slot && skip_slot[n] >= 0) return;
    double f[3] = {0.0, 0.0, 0.0};
    // nodal operands first: independent of the gather, in flight while it resolves its two dependent round trips
    double fx[3], mi[3], vv[3], aa[3], dd[3] = {0.0, 0.0, 0.0};
    unsigned char cc[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        cc[i] = code[q];
        fx[i] = fext[q];
        mi[i] = minv[q];
        vv[i] = v[q];
        aa[i] = a[q];
        if (NEXT_PREDICTOR) dd[i] = d[q];
    }
    if (GATHER) {
        int ent[8];
        double g[8][3];
        const int4 lo = __ldg(inc8 + 2 * n), hi = __ldg(inc8 + 2 * n + 1);
        ent[0] = lo.x; ent[1] = lo.y; ent[2] = lo.z; ent[3] = lo.w;
        ent[4] = hi.x; ent[5] = hi.y; ent[6] = hi.z; ent[7] = hi.w;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const bool on = ent[q] >= 0;
            const int64_t e = ent[q] >> 3;
            const int a3 = 3 * (ent[q] & 7);
            g[q][0] = on ? __ldg(fe + (int64_t)(a3)*stride + e) : 0.0;
            g[q][1] = on ? __ldg(fe + (int64_t)(a3 + 1) * stride + e) : 0.0;
            g[q][2] = on ? __ldg(fe + (int64_t)(a3 + 2) * stride + e) : 0.0;
        }
        // ascending-element order = the reference's serial assembly order (SolverT::AssembleRHS, SolverT.cpp:446-477)
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (ent[q] >= 0) {
                f[0] += g[q][0];
                f[1] += g[q][1];
                f[2] += g[q][2];
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
        f[0] = fint[3 * n];
        f[1] = fint[3 * n + 1];
        f[2] = fint[3 * n + 2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        const unsigned char c = cc[i];
        const double R = __dsub_rn(__dmul_rn(fext_scale, fx[i]), f[i]);
        const double upd = c ? 0.0 : __dmul_rn(R, mi[i]);
        double vi = vv[i], ai = aa[i];
        cd_correct(dt, vi, ai, upd);
        if (GATHER) fint[q] = f[i];
        if (NEXT_PREDICTOR) {
            double di = dd[i];
            cd_predict(dt, di, vi, ai);
            ai = 0.0;
            if (c == TB2_BC_FIX) { di = 0.0; vi = 0.0; }
            else if (c == TB2_BC_DSP) di = next_value_scale * bcval[q];
            d[q] = di;
        }
        v[q] = vi;
        a[q] = ai;
    }
}

// arguments of the node part of a fused launch
struct NodeArgs {
    int64_t n0, n1; // node range updated by this launch (empty: element work only)
    const int* inc_ptr;
    const int* inc;
    const int4* inc8;
    const double* fe;
    int64_t stride;
    double dt, fext_scale, next_value_scale;
    const double* fext;
    const double* minv;
    const unsigned char* code;
    const double* bcval;
    double* d;
    double* v;
    double* a;
    double* fint;
    const int* skip_slot;
    int next_predictor;
};

} // namespace tb2
