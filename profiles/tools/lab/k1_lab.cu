// k1_lab.cu -- stand-alone bench of explicit-step kernel variants on a jittered n^3 cube (TL + SimoIso3D), used to choose the
// shipped configuration of tb2_block_step.cuh.  Not part of the product; builds against the product headers:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I tahoe_b200/csrc -I profiles/tools/lab -o build/k1_lab profiles/tools/lab/k1_lab.cu
//   build/k1_lab [n=100] [steps=40]
// Baseline = round 1's scheme (one thread per element gathering from global memory, 192 B/element force scratch, node kernel).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "tb2_fused_step.cuh"

using namespace tb2;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------- baseline kernels (round-1 scheme)
struct BaseArgs {
    int64_t e0 = 0;
    int64_t ne, stride;
    const int* conn;
    const double *X, *u;
    double* fe;
    MatConst mat;
    unsigned long long* status;
};
// SM: 0 = modes in registers, 1 = X modes in shared memory, 2 = X and x modes in shared memory.  NOGATHER: every thread reads
// nodes 0..7 (L1 hits): the integration-point loop alone, i.e. the ceiling of any scheme that hides the gather completely
template <int MINB, int SM, bool NOGATHER, bool PAIRS = false>
__global__ void __launch_bounds__(128, MINB) k_base_force(const BaseArgs p)
{
    __shared__ double sXm[SM >= 1 ? 21 * 128 : 1], sxm[SM >= 2 ? 21 * 128 : 1];
    const int64_t e = p.e0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= p.ne) return;
    int n[8];
#pragma unroll
    for (int a = 0; a < 8; a++) n[a] = __ldg(p.conn + a * p.stride + (NOGATHER ? (int64_t)(e & 255) + 5 * 10101 : e));
    Modes cX, cU, A;
    load_modes(p.X, n, cX);
    load_modes(p.u, n, cU);
#pragma unroll
    for (int k = 0; k < 7; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) cU.m[k][i] += cX.m[k][i];
    ForceCtx fc;
    fc.mat = p.mat;
    fc.e = e;
    fc.stride = p.stride;
    fc.iteration = 0;
    int err;
    if (SM >= 1) {
#pragma unroll
        for (int k = 0; k < 7; k++)
#pragma unroll
            for (int i = 0; i < 3; i++) {
                sXm[(3 * k + i) * 128 + threadIdx.x] = cX.m[k][i];
                if (SM >= 2) sxm[(3 * k + i) * 128 + threadIdx.x] = cU.m[k][i];
            }
    }
    if (PAIRS) {
        if (SM == 2) err = force_modes_neo_pairs<kSimoIso>(p.mat, SmemModes(sXm + threadIdx.x, 128), SmemModes(sxm + threadIdx.x, 128), A);
        else if (SM == 1) err = force_modes_neo_pairs<kSimoIso>(p.mat, SmemModes(sXm + threadIdx.x, 128), RegModes(cU), A);
        else err = force_modes_neo_pairs<kSimoIso>(p.mat, RegModes(cX), RegModes(cU), A);
    } else if (SM == 2) err = force_modes<kTotalLagrangian, kSimoIso>(fc, SmemModes(sXm + threadIdx.x, 128), SmemModes(sxm + threadIdx.x, 128), cU, A);
    else if (SM == 1) err = force_modes<kTotalLagrangian, kSimoIso>(fc, SmemModes(sXm + threadIdx.x, 128), RegModes(cU), cU, A);
    else err = force_modes<kTotalLagrangian, kSimoIso>(fc, RegModes(cX), RegModes(cU), cU, A);
    if (err) atomicMax(p.status, (unsigned long long)err);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double f[8];
        modes_to_nodes(A, i, f);
#pragma unroll
        for (int a = 0; a < 8; a++) p.fe[(int64_t)(3 * a + i) * p.stride + e] = f[a];
    }
}
template <bool DESC, bool NOFEXT, int T = 256, int MINB = 1>
__global__ void __launch_bounds__(T, MINB) k_base_node(int64_t n0, int64_t nn, const int4* __restrict__ inc8, const double* __restrict__ fe, int64_t stride,
                                                  DofUpdate du, const double* __restrict__ fext, const double* __restrict__ minv,
                                                  const unsigned char* __restrict__ code, double* d, double* v)
{
    int64_t n = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nn) return;
    if (DESC) n = nn - 1 - (n - n0); // reverse sweep: the scratch the element sweep wrote last is still in L2
    int ent[8];
    const int4 lo = __ldg(inc8 + 2 * n), hi = __ldg(inc8 + 2 * n + 1);
    ent[0] = lo.x; ent[1] = lo.y; ent[2] = lo.z; ent[3] = lo.w;
    ent[4] = hi.x; ent[5] = hi.y; ent[6] = hi.z; ent[7] = hi.w;
    double f[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (ent[q] >= 0) {
            const int64_t e = ent[q] >> 3;
            const int a3 = 3 * (ent[q] & 7);
            f[0] += __ldg(fe + (int64_t)(a3)*stride + e);
            f[1] += __ldg(fe + (int64_t)(a3 + 1) * stride + e);
            f[2] += __ldg(fe + (int64_t)(a3 + 2) * stride + e);
        }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int64_t q = 3 * n + i;
        double di = d[q], vi = v[q], ai;
        cd_update_dof<true>(du, code[q], f[i], NOFEXT ? 0.0 : fext[q], minv[q], 0.0, di, vi, ai);
        d[q] = di;
        v[q] = vi;
    }
}

// ---------------------------------------------------------------- host side
static double urand(uint64_t& s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

template <class T>
static T* upload(const std::vector<T>& h)
{
    T* d = nullptr;
    CK(cudaMalloc(&d, h.size() * sizeof(T)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

struct Fields {
    double *d, *v, *a, *fint;
};

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 100;
    const int steps = argc > 2 ? atoi(argv[2]) : 40;
    const int64_t ne = (int64_t)n * n * n, p1 = n + 1, nn = p1 * p1 * p1, stride = (ne + 31) / 32 * 32;
    const double h = 1.0 / n;
    std::vector<int32_t> conn(8 * stride, 0);
    std::vector<double> X(3 * nn), u(3 * nn), v(3 * nn), minv(3 * nn), fext(3 * nn, 0.0);
    std::vector<unsigned char> code(3 * nn, 0);
    uint64_t seed = 12345;
    for (int k = 0; k <= n; k++)
        for (int j = 0; j <= n; j++)
            for (int i = 0; i <= n; i++) {
                const int64_t id = ((int64_t)k * p1 + j) * p1 + i;
                const bool interior = i > 0 && i < n && j > 0 && j < n && k > 0 && k < n;
                const double c[3] = {i * h, j * h, k * h};
                for (int q = 0; q < 3; q++) {
                    X[3 * id + q] = c[q] + (interior ? (urand(seed) - 0.5) * 0.2 * h : 0.0);
                    u[3 * id + q] = 0.003 * c[(q + 1) % 3] + 0.002 * c[q] + 1e-4 * h * (urand(seed) - 0.5);
                    v[3 * id + q] = 1e-3 * (urand(seed) - 0.5);
                    minv[3 * id + q] = 1.0 / (h * h * h);
                    if (i == 0) code[3 * id + q] = 1;
                }
            }
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                const int64_t e = ((int64_t)k * n + j) * n + i, n0 = ((int64_t)k * p1 + j) * p1 + i;
                const int64_t nd[8] = {n0, n0 + 1, n0 + 1 + p1, n0 + p1, n0 + p1 * p1, n0 + 1 + p1 * p1, n0 + 1 + p1 + p1 * p1, n0 + p1 + p1 * p1};
                for (int a = 0; a < 8; a++) conn[a * stride + e] = (int32_t)nd[a];
            }
    std::vector<int32_t> inc8(8 * nn, -1), cnt(nn, 0);
    for (int64_t e = 0; e < ne; e++)
        for (int a = 0; a < 8; a++) {
            const int32_t nd = conn[a * stride + e];
            inc8[8 * (int64_t)nd + cnt[nd]++] = (int32_t)(e * 8 + a);
        }
    BlockPlan plan;
    auto t0 = std::chrono::steady_clock::now();
    build_block_plan(ne, nn, stride, conn.data(), X.data(), nullptr, nullptr, plan);
    printf("mesh %d^3: %ld elements, %ld nodes; block plan: %ld blocks, max local nodes %ld, interior %.3f, %ld partial slots, %zu surface nodes, %.2f s\n", n,
           (long)ne, (long)nn, (long)plan.nblocks, (long)plan.max_local, (double)plan.n_interior / nn, (long)plan.npartial, plan.surf_nodes.size(),
           std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    int* d_conn = upload(conn);
    double *d_X = upload(X), *d_minv = upload(minv), *d_fext = upload(fext);
    unsigned char* d_code = upload(code);
    int* d_inc8 = upload(inc8);
    uint32_t* d_rec = upload(plan.rec);
    double *d_fe, *d_fpart, *d_bcval;
    CK(cudaMalloc(&d_fe, 24 * stride * sizeof(double)));
    CK(cudaMalloc(&d_fpart, 3 * (plan.npartial + 1) * sizeof(double)));
    CK(cudaMalloc(&d_bcval, 3 * nn * sizeof(double)));
    CK(cudaMemset(d_bcval, 0, 3 * nn * sizeof(double)));
    unsigned long long* d_status;
    CK(cudaMalloc(&d_status, 16));
    CK(cudaMemset(d_status, 0, 16));
    Fields F;
    CK(cudaMalloc(&F.d, 3 * nn * sizeof(double)));
    CK(cudaMalloc(&F.v, 3 * nn * sizeof(double)));
    CK(cudaMalloc(&F.a, 3 * nn * sizeof(double)));
    CK(cudaMalloc(&F.fint, 3 * nn * sizeof(double)));
    auto reset = [&]() {
        CK(cudaMemcpy(F.d, u.data(), 3 * nn * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(F.v, v.data(), 3 * nn * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemset(F.a, 0, 3 * nn * sizeof(double)));
    };
    MatConst mat{};
    mat.mu = 5.0;
    mat.kappa = 1000.0;
    const double dt = 0.5 * h / std::sqrt((mat.kappa + 4.0 * mat.mu / 3.0) / 1.0);
    DofUpdate du{dt, 1.0, 1.0};
    cudaEvent_t ev0, ev1;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));

    std::vector<double> ref_d, ref_v, got(3 * nn);
    auto compare = [&](const char* name, float ms) {
        CK(cudaMemcpy(got.data(), F.d, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost));
        double dmax = 0, rmax = 0;
        if (ref_d.empty()) ref_d = got;
        for (int64_t q = 0; q < 3 * nn; q++) {
            dmax = std::max(dmax, std::fabs(got[q] - ref_d[q]));
            rmax = std::max(rmax, std::fabs(ref_d[q]));
        }
        CK(cudaMemcpy(got.data(), F.v, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost));
        double vmax = 0, vr = 0;
        if (ref_v.empty()) ref_v = got;
        for (int64_t q = 0; q < 3 * nn; q++) {
            vmax = std::max(vmax, std::fabs(got[q] - ref_v[q]));
            vr = std::max(vr, std::fabs(ref_v[q]));
        }
        unsigned long long st[2];
        CK(cudaMemcpy(st, d_status, 16, cudaMemcpyDeviceToHost));
        printf("%-34s %8.3f us/step  %7.3f G el-upd/s   d relerr %.2e  v relerr %.2e  status %llu\n", name, 1e3 * ms / steps, ne / (1e6 * ms / steps),
               dmax / rmax, vmax / vr, st[0]);
        fflush(stdout);
    };

    // ---- baseline
    auto run_base = [&](auto kern, const char* name, int k5 = 0) {
        BaseArgs b{0, ne, stride, d_conn, d_X, F.d, d_fe, mat, d_status};
        const unsigned ge = (unsigned)((ne + 127) / 128), gn = (unsigned)((nn + 255) / 256);
        for (int rep = 0; rep < 2; rep++) {
            reset();
            CK(cudaEventRecord(ev0));
            for (int s = 0; s < steps; s++) {
                kern<<<ge, 128>>>(b);
                if (k5 == 0) k_base_node<false, false><<<gn, 256>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
                if (k5 == 1) k_base_node<true, false><<<gn, 256>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
                if (k5 == 2) k_base_node<true, true><<<gn, 256>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
                if (k5 == 3) k_base_node<false, true><<<gn, 256>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
            }
            CK(cudaEventRecord(ev1));
            CK(cudaEventSynchronize(ev1));
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        compare(name, ms);
        // the sweep alone
        reset();
        CK(cudaEventRecord(ev0));
        for (int s = 0; s < steps; s++) kern<<<ge, 128>>>(b);
        CK(cudaEventRecord(ev1));
        CK(cudaEventSynchronize(ev1));
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        printf("    sweep alone %8.3f us\n", 1e3 * ms / steps);
    };
    run_base(k_base_force<3, 0, false>, "base regs3 +K5");
    run_base(k_base_force<3, 2, false, true>, "pairs xxsmem3 +K5");
    run_base(k_base_force<3, 2, false, true>, "pairs xxsmem3 +K5 desc", 1);
    run_base(k_base_force<3, 2, false, true>, "pairs xxsmem3 +K5 desc nofext", 2);
    run_base(k_base_force<3, 2, false, true>, "pairs xxsmem3 +K5 asc nofext", 3);
    run_base(k_base_force<3, 1, false, true>, "pairs xsmem3 +K5");
    run_base(k_base_force<3, 0, false, true>, "pairs regs3 +K5");
    run_base(k_base_force<2, 0, false, true>, "pairs regs2 +K5");
    run_base(k_base_force<2, 1, false, true>, "pairs xsmem2 +K5");
    if (getenv("LAB_K5")) { // the node kernel alone (scratch left by one sweep): occupancy / CTA-size variants
        BaseArgs b{0, ne, stride, d_conn, d_X, F.d, d_fe, mat, d_status};
        reset();
        k_base_force<3, 0, false><<<(unsigned)((ne + 127) / 128), 128>>>(b);
        auto t5 = [&](auto kern, int T, const char* name) {
            reset();
            const unsigned gn = (unsigned)((nn + T - 1) / T);
            for (int w = 0; w < 3; w++) kern<<<gn, T>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
            CK(cudaEventRecord(ev0));
            for (int s = 0; s < steps; s++) kern<<<gn, T>>>(0, nn, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
            CK(cudaEventRecord(ev1));
            CK(cudaEventSynchronize(ev1));
            float ms;
            CK(cudaEventElapsedTime(&ms, ev0, ev1));
            printf("K5 %-28s %8.3f us\n", name, 1e3 * ms / steps);
        };
        t5(k_base_node<false, false, 256, 1>, 256, "256 thr, natural regs");
        t5(k_base_node<false, false, 256, 4>, 256, "256 thr, 4 CTAs/SM");
        t5(k_base_node<false, false, 256, 6>, 256, "256 thr, 6 CTAs/SM");
        t5(k_base_node<false, false, 256, 8>, 256, "256 thr, 8 CTAs/SM");
        t5(k_base_node<false, false, 128, 1>, 128, "128 thr, natural regs");
        t5(k_base_node<false, false, 128, 12>, 128, "128 thr, 12 CTAs/SM");
        t5(k_base_node<false, false, 512, 1>, 512, "512 thr, natural regs");
        t5(k_base_node<false, false, 512, 3>, 512, "512 thr, 3 CTAs/SM");
        t5(k_base_node<false, true, 256, 1>, 256, "256 thr, no fext");
        return 0;
    }
    if (getenv("LAB_SWEEPS_ONLY")) return 0;
    // ---- slab pipeline: K1 slabs on stream A, K5 slabs on stream B (co-resident when K1 leaves registers free)
    cudaStream_t sA, sB;
    CK(cudaStreamCreateWithFlags(&sA, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sB, cudaStreamNonBlocking));
    auto run_pipe = [&](auto kern, int S, int t5, const char* name) {
        std::vector<cudaEvent_t> e1(S), e5(S);
        for (int c = 0; c < S; c++) {
            CK(cudaEventCreateWithFlags(&e1[c], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e5[c], cudaEventDisableTiming));
        }
        const int ps = (n + S - 1) / S; // element planes per slab
        auto eb = [&](int c) { return std::min<int64_t>(ne, (int64_t)c * ps * n * n); };
        auto nb = [&](int c) { return c >= S ? nn : std::min<int64_t>(nn, (int64_t)c * ps * p1 * p1); };
        for (int rep = 0; rep < 2; rep++) {
            reset();
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(ev0, sA));
            for (int s = 0; s < steps; s++) {
                for (int c = 0; c < S; c++) {
                    if (s > 0) CK(cudaStreamWaitEvent(sA, e5[std::min(c + 1, S - 1)], 0));
                    BaseArgs b{eb(c), eb(c + 1), stride, d_conn, d_X, F.d, d_fe, mat, d_status};
                    const int64_t cnt = eb(c + 1) - eb(c);
                    if (cnt > 0) kern<<<(unsigned)((cnt + 127) / 128), 128, 0, sA>>>(b);
                    CK(cudaEventRecord(e1[c], sA));
                    if (c > 0) {
                        CK(cudaStreamWaitEvent(sB, e1[c], 0));
                        const int64_t n0 = nb(c - 1), n1 = nb(c);
                        if (n1 > n0) k_base_node<false, false><<<(unsigned)((n1 - n0 + t5 - 1) / t5), t5, 0, sB>>>(n0, n1, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
                        CK(cudaEventRecord(e5[c - 1], sB));
                    }
                }
                CK(cudaStreamWaitEvent(sB, e1[S - 1], 0));
                const int64_t n0 = nb(S - 1), n1 = nn;
                k_base_node<false, false><<<(unsigned)((n1 - n0 + t5 - 1) / t5), t5, 0, sB>>>(n0, n1, (const int4*)d_inc8, d_fe, stride, du, d_fext, d_minv, d_code, F.d, F.v);
                CK(cudaEventRecord(e5[S - 1], sB));
            }
            CK(cudaStreamWaitEvent(sA, e5[S - 1], 0));
            CK(cudaEventRecord(ev1, sA));
            CK(cudaEventSynchronize(ev1));
            CK(cudaGetLastError());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        char label[128];
        snprintf(label, sizeof label, "%s, %d slabs, K5 CTA %d", name, S, t5);
        compare(label, ms);
    };
    for (int S : {8, 16, 25}) {
        run_pipe(k_base_force<3, 0, false>, S, 256, "pipe regs3 (168)");
        run_pipe(k_base_force<3, 2, false>, S, 256, "pipe xxsmem3 (134)");
        run_pipe(k_base_force<3, 2, false>, S, 64, "pipe xxsmem3 (134)");
        run_pipe(k_base_force<4, 2, false>, S, 256, "pipe xxsmem4 (125)");
    }
    // ---- fused, warp-specialised
    int32_t *d_bconn = upload(plan.bconn), *d_elem = upload(plan.elem);
    uint16_t* d_epos = upload(plan.epos);
    int* d_cnt;
    CK(cudaMalloc(&d_cnt, (plan.npartial + 1) * sizeof(int)));
    CK(cudaMemset(d_cnt, 0, (plan.npartial + 1) * sizeof(int)));
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    auto run_fused = [&](auto kern, const char* name, int64_t nblk = -1) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFsSmem));
        FusedArgs p{};
        p.rec = d_rec;
        p.bconn = d_bconn;
        p.elem = d_elem;
        p.epos = d_epos;
        p.b_begin = 0;
        p.b_end = nblk < 0 ? plan.nblocks : nblk;
        p.X = d_X;
        p.d = F.d;
        p.v = F.v;
        p.a = F.a;
        p.fint = F.fint;
        p.minv = d_minv;
        p.fext = d_fext;
        p.bcval = d_bcval;
        p.code = d_code;
        p.fpart = d_fpart;
        p.cnt = d_cnt;
        p.mat = mat;
        p.stride = stride;
        p.dt = dt;
        p.fext_scale = 1.0;
        p.next_value_scale = 1.0;
        p.next_predictor = 1;
        p.status = d_status;
        for (int rep = 0; rep < 2; rep++) {
            reset();
            CK(cudaEventRecord(ev0));
            for (int s = 0; s < steps; s++) kern<<<nsm, kFsThreads, kFsSmem>>>(p);
            CK(cudaEventRecord(ev1));
            CK(cudaEventSynchronize(ev1));
            CK(cudaGetLastError());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        compare(name, ms);
    };
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 3>, "fused DBG3: compute only, no barriers");
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 5>, "fused DBG5: DBG3 with L1-resident gather");
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 3>, "fused DBG3: 0 blocks (launch overhead)", 0);
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 3>, "fused DBG3: 444 blocks (1 round)", 444);
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 3>, "fused DBG3: 888 blocks (2 rounds)", 888);
    run_fused(k_fused_step<kTotalLagrangian, kSimoIso, 3>, "fused DBG3: 7548 blocks (17 rounds)", 7548);
    // ---- variant C: CTA per block, CTA-local node phase
    auto run_c = [&](auto kern, const char* name) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBcSmem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlockElems, kBcSmem));
        FusedArgs p{};
        p.rec = d_rec; p.bconn = d_bconn; p.elem = d_elem; p.epos = d_epos;
        p.b_begin = 0; p.b_end = plan.nblocks;
        p.X = d_X; p.d = F.d; p.v = F.v; p.a = F.a; p.fint = F.fint;
        p.minv = d_minv; p.fext = d_fext; p.bcval = d_bcval; p.code = d_code;
        p.fpart = d_fpart; p.cnt = d_cnt; p.mat = mat; p.stride = stride;
        p.dt = dt; p.fext_scale = 1.0; p.next_value_scale = 1.0; p.next_predictor = 1; p.status = d_status;
        for (int rep = 0; rep < 2; rep++) {
            reset();
            CK(cudaEventRecord(ev0));
            for (int s = 0; s < steps; s++) kern<<<(unsigned)plan.nblocks, kBlockElems, kBcSmem>>>(p);
            CK(cudaEventRecord(ev1));
            CK(cudaEventSynchronize(ev1));
            CK(cudaGetLastError());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        char label[128];
        snprintf(label, sizeof label, "%s [occ %d, smem %d]", name, occ, kBcSmem);
        compare(label, ms);
    };
    run_c(k_block_fused<kTotalLagrangian, kSimoIso, 4>, "C: CTA/block minb4");
    run_c(k_block_fused<kTotalLagrangian, kSimoIso, 4, 4>, "C4: no node phase");
    run_c(k_block_fused<kTotalLagrangian, kSimoIso, 4, 1>, "C1: sums+interior+partials");
    CK(cudaMemset(d_cnt, 0, (plan.npartial + 1) * sizeof(int)));
    run_c(k_block_fused<kTotalLagrangian, kSimoIso, 4, 2>, "C2: + release atomics");
    CK(cudaMemset(d_cnt, 0, (plan.npartial + 1) * sizeof(int)));
    run_c(k_block_fused<kTotalLagrangian, kSimoIso, 4, 3>, "C3: relaxed atomics + completions");
    CK(cudaMemset(d_cnt, 0, (plan.npartial + 1) * sizeof(int)));
    return 0;
}
