// tb2_fused_step.cuh -- the fused explicit step: element forces, their deterministic assembly and the central-difference node
// update in ONE persistent, warp-specialised kernel per step.
//
// Reference path (a2, a15, a18, a19): SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295) ->
// SolverT::AssembleRHS (SolverT.cpp:446-477) -> DiagonalMatrixT::Solve (DiagonalMatrixT.cpp:267-323) -> FieldT::AssembleUpdate
// (FieldT.cpp:531-556) -> nExplicitCD::Corrector / Predictor (nExplicitCD.cpp:72-139).
//
// Why this shape.  The element sweep is FP64-pipe bound (2174 FP64 instructions per element; ~16 T DFMA/s measured on B200) and
// the node update is HBM bound; round 1 ran them as two kernels that could not share an SM (the sweep's 3 x 128 x 168 registers
// are the whole register file), moved 192 B/element of force scratch to HBM and back, and paid sweep + update per step.  Here
// one CTA per SM runs
//   * three COMPUTE warpgroups (setmaxnreg 160): each owns one element block of the plan (tb2_blockplan.h) at a time -- gather,
//     trilinear modes (the 21 read-only modes of X in a private shared-memory column, which is what fits the body into 160
//     registers without spills), integration-point loop (tb2_force_core.cuh), 24 forces into a shared scratch [24][128] -- and
//     never waits for anything but its own gather;
//   * one HELPER warpgroup (setmaxnreg 32..56) that does everything latency- or bandwidth-bound for the three of them in turn:
//     block record in (cp.async, one block ahead), ordered sum of the block's forces per block-local node, corrector + next
//     predictor of the interior nodes, one partial force per surface node, and -- for the block that delivers a surface node's
//     last partial (counter per node) -- the ordered sum of that node's partials and its update.
// The element forces never leave the SM and the FP64 pipe never idles behind the HBM-bound work; DRAM traffic per element-update
// drops to block connectivity + record (~80 B), X, d, v, 1/m in and d, v out.
//
// Determinism: lanes are summed in ascending order inside a block, partials in ascending block order, whoever arrives last.
// Hand-off compute <-> helper: two mbarriers per warpgroup (full: 128 compute arrivals, free: 128 helper arrivals).  Every warp
// polls the phase on its own, so the four warps of a warpgroup are never re-aligned at a block boundary (a CTA-style barrier
// there cost 12 % in the lab: the fast warps idle until the slowest one arrives); the scratch is single-buffered because the
// helper needs ~2 us per block and a warpgroup ~10 us.
#pragma once
#include "../../../include/tahoe_b200.h"
#include "tb2_blockplan.h"
#include "tb2_force_core.cuh"
#include "tb2_node_update.cuh"

namespace tb2 {

struct FusedArgs {
    const uint32_t* rec;  // [nblocks][kBlockRecWords]
    const int32_t* bconn; // [nblocks][8][kBlockElems]
    const int32_t* elem;  // [nblocks][kBlockElems]
    const uint16_t* epos; // [nblocks][kBlockElems][8]
    int64_t b_begin, b_end; // block range of this launch
    const double* X;      // [nn][3]
    double* d;            // [nn][3] displacement (predicted; read by the gather, advanced by the fused next predictor)
    double* v;
    double* a;
    double* fint;
    const double* minv;
    const double* fext; // null: no external force
    const double* bcval;
    const unsigned char* code;
    double* fpart;            // [npartial][3] partial forces of the surface nodes
    int* cnt;                 // [npartial] arrival counters (index = first slot of the node), zero between launches
    const unsigned char* off; // [ne] ElementCardT::kOFF flags or null
    MatConst mat;
    J2Hist hist;
    int64_t stride;
    double dt, fext_scale, next_value_scale;
    int next_predictor; // 1: another step follows -- fuse its predictor (then a stays 0 in memory and fint is not written)
    unsigned long long* status;
};

constexpr int kFsThreads = 4 * kBlockElems; // 3 compute warpgroups + 1 helper warpgroup
#ifndef TB2_FS_STAGGER_NS
#define TB2_FS_STAGGER_NS 3000
#endif
#ifndef TB2_FS_REGS_COMPUTE
#define TB2_FS_REGS_COMPUTE 152
#endif
constexpr int kFsRegsCompute = TB2_FS_REGS_COMPUTE, kFsRegsHelper = (65536 - 3 * kBlockElems * kFsRegsCompute) / kBlockElems / 8 * 8;
constexpr int kFsBars = 64;                        // 6 mbarriers
constexpr int kFsScratch = 3 * 8 * kBlockElems * 8; // [3][1024] forces in incidence order
constexpr int kFsModes = 21 * kBlockElems * 8;
constexpr int kFsRec = kBlockRecWords * 4;
constexpr int kFsPerWg = kFsScratch + kFsModes + 2 * kFsRec;
constexpr int kFsSmem = kFsBars + 3 * kFsPerWg;

// the update of one nodal dof from its assembled internal force (cd_node_update_one of round 1, same roundings)
struct DofUpdate {
    double dt, fext_scale, next_value_scale;
};
template <bool NEXT_PREDICTOR>
TB2_DEV void cd_update_dof(const DofUpdate& u, const unsigned char c, const double f, const double fx, const double mi, const double bcv,
                           double& di, double& vi, double& ai)
{
    const double R = __dsub_rn(__dmul_rn(u.fext_scale, fx), f);
    const double upd = c ? 0.0 : __dmul_rn(R, mi);
    ai = 0.0; // the predictor left a = 0 (nExplicitCD.cpp:93)
    cd_correct(u.dt, vi, ai, upd);
    if (NEXT_PREDICTOR) {
        cd_predict(u.dt, di, vi, ai);
        ai = 0.0;
        if (c == TB2_BC_FIX) { di = 0.0; vi = 0.0; }
        else if (c == TB2_BC_DSP) di = u.next_value_scale * bcv;
    }
}

template <int N> TB2_DEV void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N> TB2_DEV void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
TB2_DEV void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
TB2_DEV void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
TB2_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TB2_DEV void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
TB2_DEV void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
TB2_DEV void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}

// gather the 8 nodes of an element from a [node][3] array that other CTAs update in this launch (plain loads, no .nc path)
TB2_DEV void load_modes_coherent(const double* field, const int (&n)[8], Modes& out)
{
    const double s1 = 0.125, sg = 0.125 * TB2_G, sgg = 0.125 * TB2_G * TB2_G;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double v[8];
        v[0] = field[3 * (int64_t)n[0] + i];
        v[1] = field[3 * (int64_t)n[1] + i];
        v[3] = field[3 * (int64_t)n[2] + i];
        v[2] = field[3 * (int64_t)n[3] + i];
        v[4] = field[3 * (int64_t)n[4] + i];
        v[5] = field[3 * (int64_t)n[5] + i];
        v[7] = field[3 * (int64_t)n[6] + i];
        v[6] = field[3 * (int64_t)n[7] + i];
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

enum { kBarHelper = 1 }; // named barrier of the helper warpgroup (0 = __syncthreads)

// the nodal operands of one node's update
struct NodeOps {
    double mi[3], vv[3], dd[3], fx[3];
    unsigned char c[3];
};
TB2_DEV void fs_load_ops(const FusedArgs& p, const int64_t n3, NodeOps& o)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o.c[i] = p.code[n3 + i];
        o.mi[i] = p.minv[n3 + i];
        o.vv[i] = p.v[n3 + i];
        o.dd[i] = p.next_predictor ? p.d[n3 + i] : 0.0;
        o.fx[i] = p.fext ? p.fext[n3 + i] : 0.0;
    }
}
// finish node n3/3 from its assembled force f
TB2_DEV void fs_finish_node(const FusedArgs& p, const DofUpdate& du, const int64_t n3, const NodeOps& o, const double (&f)[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double vi = o.vv[i], di = o.dd[i], ai;
        const double bcv = o.c[i] == TB2_BC_DSP ? p.bcval[n3 + i] : 0.0;
        if (p.next_predictor) {
            cd_update_dof<true>(du, o.c[i], f[i], o.fx[i], o.mi[i], bcv, di, vi, ai);
            p.d[n3 + i] = di;
        } else {
            cd_update_dof<false>(du, o.c[i], f[i], o.fx[i], o.mi[i], bcv, di, vi, ai);
            p.a[n3 + i] = ai;
            p.fint[n3 + i] = f[i];
        }
        p.v[n3 + i] = vi;
    }
}

// DBG (lab only): 1 = helper skips the node work, 3 = compute warpgroups alone
template <int FORM, int MAT, int DBG = 0>
__global__ void __launch_bounds__(kFsThreads, 1) k_fused_step(const FusedArgs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wg = threadIdx.x >> 7, tid = threadIdx.x & (kBlockElems - 1);
    const int64_t nsm = gridDim.x;
    const unsigned bars = (unsigned)__cvta_generic_to_shared(smem_raw); // full[w] at 8 w, free[w] at 24 + 8 w
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; q++) mbar_init(bars + 8 * q, kBlockElems);
    }
    __syncthreads();
    // block k of warpgroup w on this SM: three neighbouring blocks per SM and round
    auto block_of = [&](int k, int w) -> int64_t { return p.b_begin + ((int64_t)k * nsm + blockIdx.x) * 3 + w; };
    auto wg_base = [&](int w) -> unsigned char* { return smem_raw + kFsBars + w * kFsPerWg; };

    if (wg < 3) {
        // ---------------------------------------------------------------- compute warpgroup
        reg_inc<kFsRegsCompute>();
        double* se = reinterpret_cast<double*>(wg_base(wg));
        double* sXm = reinterpret_cast<double*>(wg_base(wg) + kFsScratch);
        // The three warpgroups do identical work; started together they would gather together and leave the FP64 pipe idle during
        // every gather phase.  A third of a block period of head start per warpgroup keeps two of them computing at any time.
        if (TB2_FS_STAGGER_NS > 0 && wg > 0) {
            const long long t0 = clock64();
            while (clock64() - t0 < (long long)wg * (TB2_FS_STAGGER_NS * 2LL)) {} // ~2 cycles per ns
        }
        for (int k = 0;; k++) {
            const int64_t b = block_of(k, wg);
            if (b >= p.b_end) break;
            int n[8];
            {
                const int32_t* bc = p.bconn + (DBG == 5 ? (b % 3) + 300 : b) * (8 * kBlockElems) + tid;
#pragma unroll
                for (int a = 0; a < 8; a++) n[a] = __ldg(bc + a * kBlockElems);
            }
            bool active = n[0] >= 0;
            int64_t e = 0;
            if (p.off || MAT == kExplJ2 || MAT == kJ2Simo) {
                e = __ldg(p.elem + b * kBlockElems + tid);
                if (active && p.off && p.off[e]) active = false;
            }
            Modes A;
            int err = kErrNone;
            if (active) {
                Modes cU;
                {
                    Modes cX;
                    load_modes(p.X, n, cX);
                    load_modes_coherent(p.d, n, cU);
#pragma unroll
                    for (int k2 = 0; k2 < 7; k2++)
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            sXm[(3 * k2 + i) * kBlockElems + tid] = cX.m[k2][i]; // private column: no barrier needed
                            if (FORM != kSmallStrain) cU.m[k2][i] += cX.m[k2][i]; // modes of x = X + u
                        }
                }
                ForceCtx fc;
                fc.mat = p.mat;
                fc.hist = p.hist;
                fc.e = e;
                fc.stride = p.stride;
                fc.iteration = 0;
                err = force_modes<FORM, MAT>(fc, SmemModes(sXm + tid, kBlockElems), RegModes(cU), cU, A);
            }
            if (err) {
                if (!(p.off || MAT == kExplJ2 || MAT == kJ2Simo)) e = __ldg(p.elem + b * kBlockElems + tid);
                atomicMax(p.status, (unsigned long long)err);
                atomicMin(p.status + 1, (unsigned long long)e);
            }
            // scratch positions of this lane's 8 contributions (incidence order of the block)
            int pos[8];
            {
                // asm volatile: issued here, after the integration-point loop (whose shared-memory loads are volatile asms too), not
                // hoisted above it where the four registers would be spilled
                uint4 q;
                asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                             : "l"(p.epos + ((DBG == 5 ? (b % 3) + 300 : b) * kBlockElems + tid) * 8));
                pos[0] = q.x & 0xFFFF; pos[1] = q.x >> 16; pos[2] = q.y & 0xFFFF; pos[3] = q.y >> 16;
                pos[4] = q.z & 0xFFFF; pos[5] = q.z >> 16; pos[6] = q.w & 0xFFFF; pos[7] = q.w >> 16;
            }
            if (k > 0 && DBG < 3) mbar_wait(bars + 24 + 8 * wg, (k - 1) & 1); // the helper is done with the previous block's forces
            if (active) {
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    double f[8];
                    modes_to_nodes(A, i, f);
#pragma unroll
                    for (int a = 0; a < 8; a++) se[i * (8 * kBlockElems) + pos[a]] = f[a];
                }
            } else if (n[0] >= 0) { // ElementCardT::kOFF: the element contributes zeros
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int a = 0; a < 8; a++) se[i * (8 * kBlockElems) + pos[a]] = 0.0;
            }
            if (DBG < 3) mbar_arrive(bars + 8 * wg);
        }
        return;
    }

    // -------------------------------------------------------------------- helper warpgroup
    reg_dec<kFsRegsHelper>();
    if (DBG >= 3) return;
    DofUpdate du;
    du.dt = p.dt;
    du.fext_scale = p.fext_scale;
    du.next_value_scale = p.next_value_scale;
    auto srec_of = [&](int w, int buf) -> uint32_t* { return reinterpret_cast<uint32_t*>(wg_base(w) + kFsScratch + kFsModes + buf * kFsRec); };
    auto fetch_rec = [&](int64_t b, uint32_t* dst) {
        const uint4* g = reinterpret_cast<const uint4*>(p.rec + b * (int64_t)kBlockRecWords);
        for (int q = tid; q < kBlockRecWords / 4; q += kBlockElems) cp_async16(reinterpret_cast<uint4*>(dst) + q, g + q);
    };
#pragma unroll 1
    for (int w = 0; w < 3; w++)
        if (block_of(0, w) < p.b_end) fetch_rec(block_of(0, w), srec_of(w, 0));
    cp_async_wait_all();
    bar_sync(kBarHelper, kBlockElems);
#pragma unroll 1
    for (int k = 0;; k++) {
        if (block_of(k, 0) >= p.b_end) break;
#pragma unroll 1
        for (int w = 0; w < 3; w++) {
            const int64_t b = block_of(k, w);
            if (b >= p.b_end) break;
            if (block_of(k + 1, w) < p.b_end) fetch_rec(block_of(k + 1, w), srec_of(w, (k + 1) & 1));
            const uint32_t* srec = srec_of(w, k & 1);
            const double* se = reinterpret_cast<const double*>(wg_base(w));
            const int nl = (int)srec[0], nint = (int)srec[1];
            const uint16_t* ioff = reinterpret_cast<const uint16_t*>(srec + kRecIncOff);
            const uint16_t* srank = reinterpret_cast<const uint16_t*>(srec + kRecSRank);
            // operands of this thread's first interior node: in flight while the warpgroup is still computing
            NodeOps ops0;
            const bool own0 = tid < nint && DBG != 1;
            if (own0) fs_load_ops(p, 3 * (int64_t)srec[kRecNodes + tid], ops0);
            mbar_wait(bars + 8 * w, k & 1); // the warpgroup's forces are in the scratch
            int nsurf = 0;
            for (int l = tid; l < (DBG == 1 ? 0 : nl); l += kBlockElems) {
                double f[3] = {0.0, 0.0, 0.0};
                const int o1 = ioff[l + 1];
                for (int o = ioff[l]; o < o1; o++) {
                    f[0] += se[o];
                    f[1] += se[8 * kBlockElems + o];
                    f[2] += se[16 * kBlockElems + o];
                }
                const int64_t n3 = 3 * (int64_t)srec[kRecNodes + l];
                if (l < nint) {
                    if (l == tid) fs_finish_node(p, du, n3, ops0, f);
                    else {
                        NodeOps ops;
                        fs_load_ops(p, n3, ops);
                        fs_finish_node(p, du, n3, ops, f);
                    }
                } else {
                    double* out = p.fpart + 3 * ((int64_t)srec[kRecSBase + l] + (srank[l] & 0xFF));
                    out[0] = f[0];
                    out[1] = f[1];
                    out[2] = f[2];
                    nsurf++;
                }
            }
            mbar_arrive(bars + 24 + 8 * w); // the scratch may be overwritten
            if (nsurf) {
                // Arrivals are RELEASE atomics (MEMBAR.ALL.GPU + ATOMG: this thread's partials are visible before its arrival is
                // counted).  No acquire fence anywhere: __threadfence() / acquire emit CCTL.IVALL, which throws away the L1 lines the
                // compute warpgroups gather from; the completing thread reads the partials with L2-coherent loads (ld.cg) instead.
                // All arrivals of this thread's surface nodes first (independent round trips), then the completions.
                int old[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int l = tid + j * kBlockElems;
                    old[j] = -1;
                    if (l >= nint && l < nl)
                        asm volatile("atom.add.release.gpu.global.s32 %0, [%1], 1;" : "=r"(old[j]) : "l"(p.cnt + srec[kRecSBase + l]) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int l = tid + j * kBlockElems;
                    if (old[j] < 0) continue;
                    const int need = srank[l] >> 8;
                    if (old[j] != need - 1) continue;
                    // last arrival: every sharer's partial is in L2 (their release preceded our atomic)
                    const int64_t base = srec[kRecSBase + l], n3 = 3 * (int64_t)srec[kRecNodes + l];
                    p.cnt[base] = 0;
                    NodeOps ops;
                    fs_load_ops(p, n3, ops);
                    double f[3] = {0.0, 0.0, 0.0};
                    for (int r = 0; r < need; r++) {
                        const double* q = p.fpart + 3 * (base + r);
                        f[0] += __ldcg(q);
                        f[1] += __ldcg(q + 1);
                        f[2] += __ldcg(q + 2);
                    }
                    fs_finish_node(p, du, n3, ops, f);
                }
            }
            cp_async_wait_all();
            bar_sync(kBarHelper, kBlockElems); // the next record is complete and everyone is done with this one
        }
    }
}

// ---- variant C: one CTA per block, no helper -- the CTA's own threads run the node phase after a barrier; several CTAs per SM
// (modes of X and x in shared memory: 125 registers, 4 CTAs) cover each other's latency-bound phases.
constexpr int kBcSmem = kFsRec + ((2 * kFsModes > kFsScratch) ? 2 * kFsModes : kFsScratch);
template <int FORM, int MAT, int MINB, int DBG = 0>
__global__ void __launch_bounds__(kBlockElems, MINB) k_block_fused(const FusedArgs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* srec = reinterpret_cast<uint32_t*>(smem_raw);
    double* se = reinterpret_cast<double*>(smem_raw + kFsRec);
    double* sXm = se;                       // modes live during the integration-point loop, the scratch after it
    double* sxm = se + 21 * kBlockElems;
    const int tid = threadIdx.x;
    const int64_t b = p.b_begin + blockIdx.x;
    {
        const uint4* g = reinterpret_cast<const uint4*>(p.rec + b * (int64_t)kBlockRecWords);
        for (int q = tid; q < kBlockRecWords / 4; q += kBlockElems) cp_async16(reinterpret_cast<uint4*>(srec) + q, g + q);
    }
    int n[8];
    {
        const int32_t* bc = p.bconn + b * (8 * kBlockElems) + tid;
#pragma unroll
        for (int a = 0; a < 8; a++) n[a] = __ldg(bc + a * kBlockElems);
    }
    bool active = n[0] >= 0;
    int64_t e = 0;
    if (p.off || MAT == kExplJ2 || MAT == kJ2Simo) {
        e = __ldg(p.elem + b * kBlockElems + tid);
        if (active && p.off && p.off[e]) active = false;
    }
    Modes A;
    int err = kErrNone;
    if (active) {
        {
            Modes cX, cU;
            load_modes(p.X, n, cX);
            load_modes_coherent(p.d, n, cU);
#pragma unroll
            for (int k2 = 0; k2 < 7; k2++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    sXm[(3 * k2 + i) * kBlockElems + tid] = cX.m[k2][i];
                    sxm[(3 * k2 + i) * kBlockElems + tid] = FORM != kSmallStrain ? cU.m[k2][i] + cX.m[k2][i] : cU.m[k2][i];
                }
        }
        ForceCtx fc;
        fc.mat = p.mat;
        fc.hist = p.hist;
        fc.e = e;
        fc.stride = p.stride;
        fc.iteration = 0;
        Modes dummy;
        err = force_modes<FORM, MAT>(fc, SmemModes(sXm + tid, kBlockElems), SmemModes(sxm + tid, kBlockElems), dummy, A);
    }
    if (err) {
        if (!(p.off || MAT == kExplJ2 || MAT == kJ2Simo)) e = __ldg(p.elem + b * kBlockElems + tid);
        atomicMax(p.status, (unsigned long long)err);
        atomicMin(p.status + 1, (unsigned long long)e);
    }
    int pos[8];
    {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.epos + (b * kBlockElems + tid) * 8));
        pos[0] = q.x & 0xFFFF; pos[1] = q.x >> 16; pos[2] = q.y & 0xFFFF; pos[3] = q.y >> 16;
        pos[4] = q.z & 0xFFFF; pos[5] = q.z >> 16; pos[6] = q.w & 0xFFFF; pos[7] = q.w >> 16;
    }
    cp_async_wait_all();
    __syncthreads(); // every lane is done with the modes: the scratch may overlay them; the record has landed
    if (n[0] >= 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double f[8];
            if (active) modes_to_nodes(A, i, f);
#pragma unroll
            for (int a = 0; a < 8; a++) se[i * (8 * kBlockElems) + pos[a]] = active ? f[a] : 0.0;
        }
    }
    __syncthreads();
    // node phase
    if (DBG == 4) return;
    DofUpdate du;
    du.dt = p.dt;
    du.fext_scale = p.fext_scale;
    du.next_value_scale = p.next_value_scale;
    const int nl = (int)srec[0], nint = (int)srec[1];
    const uint16_t* ioff = reinterpret_cast<const uint16_t*>(srec + kRecIncOff);
    const uint16_t* srank = reinterpret_cast<const uint16_t*>(srec + kRecSRank);
    int old[3] = {-1, -1, -1};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int l = tid + j * kBlockElems;
        if (l >= nl) break;
        const int64_t n3 = 3 * (int64_t)srec[kRecNodes + l];
        NodeOps ops;
        if (l < nint) fs_load_ops(p, n3, ops);
        double f[3] = {0.0, 0.0, 0.0};
        const int o1 = ioff[l + 1];
        for (int o = ioff[l]; o < o1; o++) {
            f[0] += se[o];
            f[1] += se[8 * kBlockElems + o];
            f[2] += se[16 * kBlockElems + o];
        }
        if (l < nint) fs_finish_node(p, du, n3, ops, f);
        else {
            double* out = p.fpart + 3 * ((int64_t)srec[kRecSBase + l] + (srank[l] & 0xFF));
            out[0] = f[0];
            out[1] = f[1];
            out[2] = f[2];
            old[j] = 0;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int l = tid + j * kBlockElems;
        if (old[j] == 0 && DBG != 1 && DBG != 3) asm volatile("atom.add.release.gpu.global.s32 %0, [%1], 1;" : "=r"(old[j]) : "l"(p.cnt + srec[kRecSBase + l]) : "memory");
        if (DBG == 3 && old[j] == 0) old[j] = atomicAdd(p.cnt + srec[kRecSBase + l], 1); // relaxed, no membar (incorrect; timing only)
    }
    if (DBG == 1 || DBG == 2) return;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int l = tid + j * kBlockElems;
        if (old[j] < 0) continue;
        const int need = srank[l] >> 8;
        if (old[j] != need - 1) continue;
        const int64_t base = srec[kRecSBase + l], n3 = 3 * (int64_t)srec[kRecNodes + l];
        p.cnt[base] = 0;
        NodeOps ops;
        fs_load_ops(p, n3, ops);
        double f[3] = {0.0, 0.0, 0.0};
        for (int r = 0; r < need; r++) {
            const double* q = p.fpart + 3 * (base + r);
            f[0] += __ldcg(q);
            f[1] += __ldcg(q + 1);
            f[2] += __ldcg(q + 2);
        }
        fs_finish_node(p, du, n3, ops, f);
    }
}

// Nodes the kernel leaves unfinished -- held nodes (multi-GPU interface) -- are finished from their partials by this pass:
// thread per listed surface node
struct SurfaceArgs {
    int64_t count;
    const int32_t* nodes; // [count] global node ids
    const int32_t* base;  // [count] first partial slot
    const int32_t* nsh;   // [count] number of partials
    const double* fpart;
    int* cnt;
    FusedArgs f;          // nodal arrays and step constants
    double* packed;       // if set: the summed partial goes to packed[3 pslot[s]] instead of updating the node (interface exchange)
    const int32_t* pslot;
};
__global__ void __launch_bounds__(256) k_fused_surface_pass(const SurfaceArgs p)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= p.count) return;
    const int64_t base = p.base[s];
    double f[3] = {0.0, 0.0, 0.0};
    for (int r = 0, n = p.nsh[s]; r < n; r++) {
        const double* q = p.fpart + 3 * (base + r);
        f[0] += q[0];
        f[1] += q[1];
        f[2] += q[2];
    }
    p.cnt[base] = 0;
    if (p.packed) {
        double* out = p.packed + 3 * (int64_t)p.pslot[s];
        out[0] = f[0];
        out[1] = f[1];
        out[2] = f[2];
        return;
    }
    DofUpdate du;
    du.dt = p.f.dt;
    du.fext_scale = p.f.fext_scale;
    du.next_value_scale = p.f.next_value_scale;
    const int64_t n3 = 3 * (int64_t)p.nodes[s];
    NodeOps ops;
    fs_load_ops(p.f, n3, ops);
    fs_finish_node(p.f, du, n3, ops, f);
}

} // namespace tb2
