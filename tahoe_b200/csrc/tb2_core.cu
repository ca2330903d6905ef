// tb2_core.cu -- library plumbing, device-resident mesh and the node<->element incidence (K9 part 1).
#include <cub/cub.cuh>

#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "tb2_internal.h"

namespace tb2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return TB2_ERR_CUDA;
}

// ---- kernels ------------------------------------------------------------------------------------------
// [ne][8] AoS connectivity (ElementBaseT::fConnectivities, iArray2DT) -> [8][stride] SoA, plus (node, e*8+a) pairs
__global__ void k_conn_to_soa(int64_t ne, int64_t stride, const int* __restrict__ aos, int* __restrict__ soa,
                              int* __restrict__ keys, int* __restrict__ vals)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; // one thread per (e,a), AoS order
    if (t >= ne * 8) return;
    const int64_t e = t >> 3;
    const int a = (int)(t & 7);
    const int n = aos[t];
    soa[a * stride + e] = n;
    keys[t] = n;
    vals[t] = (int)t; // e*8+a
}
__global__ void k_pad_conn(int64_t ne, int64_t stride, int* __restrict__ soa)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t pad = stride - ne;
    if (t >= pad * 8) return;
    soa[(t / pad) * stride + ne + (t % pad)] = 0;
}
// inc_ptr[n] = first position in the sorted key array with key >= n (n = 0..nn)
__global__ void k_lower_bound(int64_t nn, int64_t nkeys, const int* __restrict__ keys, int* __restrict__ ptr)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n > nn) return;
    int64_t lo = 0, hi = nkeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < n) lo = mid + 1;
        else hi = mid;
    }
    ptr[n] = (int)lo;
}

__global__ void k_build_inc8(int64_t nn, const int* __restrict__ inc_ptr, const int* __restrict__ inc, int* __restrict__ inc8)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 8 * nn) return;
    const int64_t n = t >> 3;
    const int q = (int)(t & 7);
    const int k = inc_ptr[n] + q;
    inc8[t] = k < inc_ptr[n + 1] ? inc[k] : -1;
}

// FP64 FMA peak probe: 8 independent dependent-FMA chains per thread (the roofline denominator for K1/K3, which
// MEASURED_PEAKS.json does not carry)
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double x, double y)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

} // namespace tb2

using namespace tb2;

extern "C" {

const char* tb2_version(void) { return "tahoe_b200 0.1 (sm_100a, fp64)"; }
const char* tb2_last_error(void) { return g_err; }

int tb2_device_count(int* count)
{
    TB2_ARG(count);
    TB2_CUDA(cudaGetDeviceCount(count));
    return TB2_OK;
}
int tb2_measure_fp64_peak(int device, double* tflops)
{
    TB2_ARG(tflops);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    TB2_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, T = 256, iters = 1 << 14;
    DevBuf<double> out;
    TB2_CUDA(out.alloc((size_t)blocks * T));
    cudaEvent_t a, b;
    TB2_CUDA(cudaEventCreate(&a));
    TB2_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        TB2_CUDA(cudaEventRecord(a));
        k_fp64_peak<<<blocks, T>>>(out.p, iters, 1.0000001, 1e-9);
        TB2_CUDA(cudaEventRecord(b));
        TB2_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        TB2_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * T / (best * 1e-3) * 1e-12;
    return TB2_OK;
}
int tb2_malloc(int device, size_t bytes, void** d_ptr)
{
    TB2_ARG(d_ptr);
    DeviceGuard g(device);
    TB2_CUDA(cudaMalloc(d_ptr, bytes));
    return TB2_OK;
}
int tb2_free(int device, void* d_ptr)
{
    DeviceGuard g(device);
    TB2_CUDA(cudaFree(d_ptr));
    return TB2_OK;
}
int tb2_memcpy_h2d(int device, void* d_dst, const void* h_src, size_t bytes)
{
    DeviceGuard g(device);
    TB2_CUDA(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
    return TB2_OK;
}
int tb2_memcpy_d2h(int device, void* h_dst, const void* d_src, size_t bytes)
{
    DeviceGuard g(device);
    TB2_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return TB2_OK;
}
int tb2_host_register(void* h_ptr, size_t bytes)
{
    TB2_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
    return TB2_OK;
}
int tb2_host_unregister(void* h_ptr)
{
    TB2_CUDA(cudaHostUnregister(h_ptr));
    return TB2_OK;
}

int tb2_mesh_create(int device, int64_t nn, int64_t ne, const int32_t* h_conn, const double* h_coords, tb2_mesh** out)
{
    TB2_ARG(out && h_conn && h_coords && nn > 0 && ne > 0);
    TB2_ARG(ne < (1LL << 28) && nn < (1LL << 31) / 3); // int32 incidence entries e*8+a and nodal dof indices 3n+i
    int ndev = 0;
    TB2_CUDA(cudaGetDeviceCount(&ndev));
    TB2_ARG(device >= 0 && device < ndev);
    for (int64_t i = 0; i < ne * 8; i++)
        if (h_conn[i] < 0 || h_conn[i] >= nn) {
            set_error("connectivity entry %lld out of range", (long long)i);
            return TB2_ERR_SIZE;
        }
    DeviceGuard g(device);
    tb2_mesh* m = new tb2_mesh;
    m->device = device;
    m->nn = nn;
    m->ne = ne;
    m->stride = (ne + 31) / 32 * 32;
    auto fail = [&](int s) { delete m; return s; };
#define M_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(cuda_fail(_e, #call, __FILE__, __LINE__)); } while (0)
    M_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    M_CUDA(m->conn.alloc(8 * m->stride));
    M_CUDA(m->X.alloc(3 * nn));
    M_CUDA(m->inc_ptr.alloc(nn + 1));
    M_CUDA(m->inc.alloc(8 * ne));
    M_CUDA(m->inc8.alloc(8 * nn));
    M_CUDA(m->fe.alloc(24 * m->stride));
    M_CUDA(cudaMemcpyAsync(m->X.p, h_coords, 3 * nn * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    {
        DevBuf<int> aos, keys, vals, keys_sorted;
        M_CUDA(aos.alloc(8 * ne));
        M_CUDA(keys.alloc(8 * ne));
        M_CUDA(vals.alloc(8 * ne));
        M_CUDA(keys_sorted.alloc(8 * ne));
        M_CUDA(cudaMemcpyAsync(aos.p, h_conn, 8 * ne * sizeof(int), cudaMemcpyHostToDevice, m->stream));
        const int T = 256;
        k_conn_to_soa<<<(unsigned)((8 * ne + T - 1) / T), T, 0, m->stream>>>(ne, m->stride, aos.p, m->conn.p, keys.p, vals.p);
        if (m->stride > ne) k_pad_conn<<<(unsigned)(((m->stride - ne) * 8 + T - 1) / T), T, 0, m->stream>>>(ne, m->stride, m->conn.p);
        // stable radix sort by node id: within a node the entries stay in ascending element order, which is the
        // reference's serial assembly order (SolverT::AssembleRHS, SolverT.cpp:446-477)
        size_t tmp_bytes = 0;
        int end_bit = 1;
        while ((1LL << end_bit) < nn) end_bit++;
        M_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, m->inc.p, (int)(8 * ne), 0, end_bit, m->stream));
        DevBuf<unsigned char> tmp;
        M_CUDA(tmp.alloc(tmp_bytes));
        M_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, vals.p, m->inc.p, (int)(8 * ne), 0, end_bit, m->stream));
        k_lower_bound<<<(unsigned)((nn + 1 + T - 1) / T), T, 0, m->stream>>>(nn, 8 * ne, keys_sorted.p, m->inc_ptr.p);
        k_build_inc8<<<(unsigned)((8 * nn + T - 1) / T), T, 0, m->stream>>>(nn, m->inc_ptr.p, m->inc.p, m->inc8.p);
        M_CUDA(cudaGetLastError());
        M_CUDA(cudaStreamSynchronize(m->stream));
    }
#undef M_CUDA
    *out = m;
    return TB2_OK;
}

int tb2_comm_destroy(tb2_mesh* mesh);

int tb2_mesh_destroy(tb2_mesh* m)
{
    if (!m) return TB2_OK;
    DeviceGuard g(m->device);
    if (m->comm) tb2_comm_destroy(m);
    cudaStreamSynchronize(m->stream);
    for (auto& r : m->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    if (m->ev_k5) cudaEventDestroy(m->ev_k5);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    cudaStreamDestroy(m->stream);
    delete m;
    return TB2_OK;
}
int tb2_mesh_sizes(const tb2_mesh* m, int64_t* nn, int64_t* ne)
{
    TB2_ARG(m);
    if (nn) *nn = m->nn;
    if (ne) *ne = m->ne;
    return TB2_OK;
}
int tb2_mesh_device(const tb2_mesh* m, int* device)
{
    TB2_ARG(m && device);
    *device = m->device;
    return TB2_OK;
}
void* tb2_mesh_stream(const tb2_mesh* m) { return m ? (void*)m->stream : nullptr; }

int tb2_mesh_synchronize(tb2_mesh* m)
{
    TB2_ARG(m);
    DeviceGuard g(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return TB2_OK;
}
int tb2_profile_reserve(tb2_mesh* m, int64_t records)
{
    TB2_ARG(m && records >= 0);
    DeviceGuard g(m->device);
    while ((int64_t)m->prof.size() < records) {
        ProfRec r;
        r.cat = 0;
        TB2_CUDA(cudaEventCreate(&r.a));
        TB2_CUDA(cudaEventCreate(&r.b));
        m->prof.push_back(r);
    }
    return TB2_OK;
}
int tb2_profile_begin(tb2_mesh* m)
{
    TB2_ARG(m);
    DeviceGuard g(m->device);
    m->prof_used = 0;
    m->prof_on = true;
    m->launches = 0;
    return TB2_OK;
}
int tb2_profile_end(tb2_mesh* m, double* h_ms, int64_t* h_count, int64_t* launches)
{
    TB2_ARG(m);
    DeviceGuard g(m->device);
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    m->prof_on = false;
    double ms[kProfNumCat] = {0};
    int64_t cnt[kProfNumCat] = {0};
    for (size_t i = 0; i < m->prof_used; i++) {
        const ProfRec& r = m->prof[i];
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            ms[r.cat] += t;
            cnt[r.cat]++;
        }
    }
    if (const char* path = getenv("TB2_PROF_DUMP")) { // timeline of the last TB2_PROF_DUMP_N (default 200) launches: cat,start_ms,end_ms
        char fname[1024];
        snprintf(fname, sizeof fname, path, m->device); // a %d in the path becomes the device ordinal (one file per rank)
        if (FILE* f = fopen(fname, "w")) {
            const char* nmax = getenv("TB2_PROF_DUMP_N");
            const size_t want = nmax ? (size_t)atol(nmax) : 200;
            const size_t first = m->prof_used > want ? m->prof_used - want : 0;
            fprintf(f, "cat,start_ms,end_ms\n");
            for (size_t i = first; i < m->prof_used; i++) {
                float t0 = 0.f, t1 = 0.f;
                if (cudaEventElapsedTime(&t0, m->prof[first].a, m->prof[i].a) != cudaSuccess) { cudaGetLastError(); continue; }
                if (cudaEventElapsedTime(&t1, m->prof[first].a, m->prof[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
                fprintf(f, "%d,%.4f,%.4f\n", m->prof[i].cat, t0, t1);
            }
            fclose(f);
        }
    }
    m->prof_used = 0;
    for (int c = 0; c < kProfNumCat; c++) {
        if (h_ms) h_ms[c] = ms[c];
        if (h_count) h_count[c] = cnt[c];
    }
    if (launches) *launches = (int64_t)m->launches;
    return TB2_OK;
}

} // extern "C"
