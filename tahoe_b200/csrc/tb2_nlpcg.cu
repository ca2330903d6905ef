// tb2_nlpcg.cu -- the two nonlinear solution drivers of the hot path, device resident: Tahoe's <PCG_solver> (nonlinear PCG) and
// <nonlinear_solver> (Newton) with the device CSR + Jacobi-PCG as its linear solve (tb2_newton_solve, end of file).
//
// Replaces PCGSolver_LS (solvers/PCGSolver_LS.cpp: Iterate :107-118, CGSearch :145-211, Update :213-348, GValue :351-371) running
// inside NLSolver::Solve / ExitIteration (solvers/NLSolver.cpp:57-263, 675-756) with <diagonal_matrix/> as its matrix type, i.e.
// a DiagonalMatrixT in kDiagOnly mode (SolverT.cpp:1097-1102, DiagonalMatrixT.cpp:107-113, 267-323) that is re-formed every
// `restart` iterations.  The solver needs no assembled tangent: one iteration is 2-4 residual sweeps (K1 + ordered node gather)
// and a handful of equation-space vector kernels; displacement, residuals, search directions and the preconditioner never leave
// the device.  The host keeps the control flow of the reference (restart counter, secant line search with its bracketing rules,
// convergence tests) and reads two scalars per residual evaluation -- the same decisions, taken from the same quantities.
//
// Reductions are two-stage with a fixed block order (no float atomics): reruns are bit-identical.  Multi-GPU:
// partial nodal forces / diagonals are summed over the partition interface, dot products count owned equations once and are
// all-reduced (2 scalars).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "tb2_internal.h"
#include "tb2_linesearch.h"

namespace tb2 {

int launch_element_forces(tb2_group* g, const double* d_u, const double* d_ul, int iteration);
int launch_node_gather(tb2_mesh* m, double* d_out, bool per_dof);
bool comm_active(tb2_mesh* m);
const unsigned char* comm_owned_mask(tb2_mesh* m);
int comm_allreduce_scalars(tb2_mesh* m, double* d_vals, int n);

static const int kNlBlocks = 148 * 4;

TB2_DEV double nl_block_sum(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t; // valid in thread 0
}

// FEManagerT::FormRHS on the active equations: R = fext - fint (NodeManagerT::FormRHS adds the nodal forces, the element group
// -fint); partial sums of R.R (SolverT::Residual) and, for the line search, of update.R (PCGSolver_LS::GValue)
__global__ void __launch_bounds__(256) k_nl_residual(int64_t n, const int* __restrict__ eq_node, const double* __restrict__ fext,
                                                    const double* __restrict__ fint, const double* __restrict__ upd,
                                                    const unsigned char* __restrict__ w, double* __restrict__ R, double* __restrict__ partial)
{
    __shared__ double sh[32];
    double rr = 0.0, gr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = __ldg(eq_node + i);
        const double r = fext[k] - fint[k];
        R[i] = r;
        if (!w || w[i]) {
            rr += r * r;
            if (upd) gr += upd[i] * r;
        }
    }
    rr = nl_block_sum(rr, sh);
    gr = nl_block_sum(gr, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = rr;
        partial[2 * blockIdx.x + 1] = gr;
    }
}
// out[0..1] = sums of the interleaved partials, in block order
__global__ void __launch_bounds__(1024) k_nl_sum2(int nparts, const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = nl_block_sum(a, sh);
    b = nl_block_sum(b, sh);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}
// DiagonalMatrixT::Factorize (DiagonalMatrixT.cpp:267-310): reciprocal, pivots with |m| <= kSmall = 1e-12 are left as they are
__global__ void k_nl_factorize(int64_t n, const int* __restrict__ eq_node, const double* __restrict__ diag, double* __restrict__ minv)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = diag[eq_node[i]];
    minv[i] = fabs(d) > 1.0e-12 ? 1.0 / d : d;
}
// CGSearch, restart branch (PCGSolver_LS.cpp:152-165): R_last = R; dir = M^-1 R; dir_last = dir; partial of dir.R
__global__ void __launch_bounds__(256) k_nl_steepest(int64_t n, const double* __restrict__ R, const double* __restrict__ minv,
                                                    const unsigned char* __restrict__ w, double* __restrict__ R_last,
                                                    double* __restrict__ dir, double* __restrict__ partial)
{
    __shared__ double sh[32];
    double g = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double r = R[i], d = r * minv[i];
        R_last[i] = r;
        dir[i] = d;
        if (!w || w[i]) g += d * r;
    }
    g = nl_block_sum(g, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = g; partial[2 * blockIdx.x + 1] = 0.0; }
}
// Bertsekas (6.32) with the scaling matrix (PCGSolver_LS.cpp:166-186): partials of R.M^-1(R - R_last) and R_last.M^-1 R_last
__global__ void __launch_bounds__(256) k_nl_beta_partials(int64_t n, const double* __restrict__ R, const double* __restrict__ R_last,
                                                         const double* __restrict__ minv, const unsigned char* __restrict__ w,
                                                         double* __restrict__ partial)
{
    __shared__ double sh[32];
    double num = 0.0, den = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (w && !w[i]) continue;
        const double r = R[i], rl = R_last[i], mi = minv[i];
        num += r * ((r - rl) * mi);
        den += rl * (rl * mi);
    }
    num = nl_block_sum(num, sh);
    den = nl_block_sum(den, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = num; partial[2 * blockIdx.x + 1] = den; }
}
// (:188-201) beta = num / den unless |den| < 1e-24 (then steepest descent); R_last = R; dir = M^-1 R + beta dir; partial of dir.R
__global__ void __launch_bounds__(256) k_nl_direction(int64_t n, const double* __restrict__ red, const double* __restrict__ R,
                                                     const double* __restrict__ minv, const unsigned char* __restrict__ w,
                                                     double* __restrict__ R_last, double* __restrict__ dir, double* __restrict__ partial)
{
    __shared__ double sh[32];
    const double beta = fabs(red[1]) < 1.0e-24 ? 0.0 : red[0] / red[1];
    double g = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double r = R[i];
        const double d = r * minv[i] + beta * dir[i];
        R_last[i] = r;
        dir[i] = d;
        if (!w || w[i]) g += d * r;
    }
    g = nl_block_sum(g, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = g; partial[2 * blockIdx.x + 1] = 0.0; }
}
// FEManagerT::Update -> FieldT::AssembleUpdate (FieldT.cpp:531-556): u[active] += ds * update
__global__ void k_nl_update(int64_t n, const int* __restrict__ eq_node, double ds, const double* __restrict__ upd, double* __restrict__ u)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) u[eq_node[i]] += ds * upd[i];
}
__global__ void k_nl_eq_owned(int64_t neq, const int* __restrict__ eq_node, const unsigned char* __restrict__ node_owned, unsigned char* __restrict__ w)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < neq) w[i] = node_owned[eq_node[i] / 3];
}
// y = alpha y + x
__global__ void __launch_bounds__(256) k_nl_axpby(int64_t n, double alpha, double* __restrict__ y, const double* __restrict__ x)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = alpha * y[i] + x[i];
}

} // namespace tb2

using namespace tb2;

struct tb2_nlpcg {
    int device = 0; // the solver may outlive the group / mesh it was made for (FEManagerT deletes element groups before solvers):
                    // its destructor touches nothing but its own buffers
    tb2_group* group = nullptr;
    tb2_equations* eqs = nullptr;
    tb2_nlpcg_params prm{};
    DevBuf<double> R, R_last, dir, minv, fint, diag, partial, red;
    DevBuf<double> fma; // [nn][3] inertia force of an implicit step
    DevBuf<unsigned char> eq_owned;
    int64_t residual_sweeps = 0, preconditioner_sweeps = 0;
};

namespace {

struct Ctx {
    tb2_nlpcg* s;
    tb2_mesh* m;
    cudaStream_t st;
    int64_t n;
    unsigned vb, nb1;
    double* u;
    const double* ul;
    const double* fext;
    const unsigned char* w;
    int iteration; // SolverT::fNumIteration
    // implicit dynamics (tb2_newton_solve_dynamic): the unknown is the acceleration increment; null for static solves
    const tb2_dynamics* dyn = nullptr;
    double* v = nullptr;
    double* a = nullptr;
};

int read2(Ctx& c, double out[2])
{
    k_nl_sum2<<<1, 1024, 0, c.st>>>((int)c.vb, c.s->partial.p, c.s->red.p);
    if (comm_active(c.m)) TB2_CHECK(comm_allreduce_scalars(c.m, c.s->red.p, 2));
    TB2_CUDA(cudaMemcpyAsync(out, c.s->red.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.st));
    TB2_CUDA(cudaStreamSynchronize(c.st));
    return TB2_OK;
}
// R(u) on the device; h[0] = R.R, h[1] = update.R (when with_update)
int form_rhs(Ctx& c, bool with_update, double h[2])
{
    tb2_nlpcg* s = c.s;
    TB2_CHECK(launch_element_forces(s->group, c.u, c.ul, c.iteration));
    TB2_CHECK(launch_node_gather(c.m, s->fint.p, true));
    if (c.dyn) { // element residual of an implicit integrator: constKd fint + constMa M a (SolidElementT.cpp:1237-1265)
        if (!s->fma.p) TB2_CUDA(s->fma.alloc(3 * c.m->nn));
        TB2_CHECK(tb2_form_inertial_force(s->group, c.dyn->mass_type, c.dyn->constMa, c.a, s->fma.p));
        const int64_t n3 = 3 * c.m->nn;
        k_nl_axpby<<<(unsigned)((n3 + 255) / 256), 256, 0, c.st>>>(n3, c.dyn->constKd, s->fint.p, s->fma.p);
    }
    if (comm_active(c.m)) TB2_CHECK(tb2_comm_sum_interface(c.m, s->fint.p));
    {
        ProfScope ps(c.m, kProfPcgVec, 2);
        k_nl_residual<<<c.vb, 256, 0, c.st>>>(c.n, s->eqs->eq_node.p, c.fext, s->fint.p, with_update ? s->dir.p : nullptr, c.w, s->R.p, s->partial.p);
        TB2_CHECK(read2(c, h));
    }
    s->residual_sweeps++;
    return tb2_group_status(s->group, nullptr); // kBadJacobianDet etc. surface here, as the exception does in GValue / FormRHS
}
int form_preconditioner(Ctx& c)
{
    tb2_nlpcg* s = c.s;
    TB2_CHECK(tb2_form_stiffness_diagonal(s->group, c.u, c.ul, c.iteration, s->diag.p));
    if (comm_active(c.m)) TB2_CHECK(tb2_comm_sum_interface(c.m, s->diag.p));
    ProfScope ps(c.m, kProfPcgVec);
    k_nl_factorize<<<c.nb1, 256, 0, c.st>>>(c.n, s->eqs->eq_node.p, s->diag.p, s->minv.p);
    s->preconditioner_sweeps++;
    return TB2_OK;
}
// PCGSolver_LS::GValue (:351-371)
int gvalue(Ctx& c, double step, double& s_current, double& G, double& rr)
{
    {
        ProfScope ps(c.m, kProfPcgVec);
        k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, c.s->eqs->eq_node.p, step - s_current, c.s->dir.p, c.u);
    }
    s_current = step;
    double h[2];
    TB2_CHECK(form_rhs(c, true, h));
    rr = h[0];
    G = h[1];
    return TB2_OK;
}

} // namespace

extern "C" {

int tb2_nlpcg_create(tb2_group* g, tb2_equations* eqs, const tb2_nlpcg_params* prm, tb2_nlpcg** out)
{
    TB2_ARG(g && eqs && prm && out && eqs->mesh == g->mesh);
    TB2_ARG(prm->restart >= 0 && prm->line_search_iterations >= 0 && prm->max_iterations >= 0);
    tb2_mesh* m = g->mesh;
    DeviceGuard dg(m->device);
    tb2_nlpcg* s = new tb2_nlpcg;
    s->device = m->device;
    s->group = g;
    s->eqs = eqs;
    s->prm = *prm;
    const int64_t n = eqs->neq;
    cudaError_t e = cudaSuccess;
    for (DevBuf<double>* b : {&s->R, &s->R_last, &s->dir, &s->minv})
        if (e == cudaSuccess) e = b->alloc(n);
    if (e == cudaSuccess) e = s->fint.alloc(3 * m->nn);
    if (e == cudaSuccess) e = s->diag.alloc(3 * m->nn);
    if (e == cudaSuccess) e = s->partial.alloc(2 * kNlBlocks);
    if (e == cudaSuccess) e = s->red.alloc(2);
    if (e != cudaSuccess) {
        delete s;
        return cuda_fail(e, "tb2_nlpcg_create: work vectors", __FILE__, __LINE__);
    }
    *out = s;
    return TB2_OK;
}

int tb2_nlpcg_destroy(tb2_nlpcg* s)
{
    if (!s) return TB2_OK;
    DeviceGuard dg(s->device);
    cudaDeviceSynchronize();
    delete s;
    return TB2_OK;
}

int tb2_nlpcg_counters(const tb2_nlpcg* s, int64_t* residual_sweeps, int64_t* preconditioner_sweeps)
{
    TB2_ARG(s);
    if (residual_sweeps) *residual_sweeps = s->residual_sweeps;
    if (preconditioner_sweeps) *preconditioner_sweeps = s->preconditioner_sweeps;
    return TB2_OK;
}

int tb2_secant_search_host(double (*slope)(double step, void* user), void* user, double slope_at_zero, double max_step, double abs_tolerance,
                           double rel_tolerance, int max_trials, double* final_step, int* evaluations)
{
    TB2_ARG(slope && final_step && max_trials >= 0);
    SecantSearch search(max_step, abs_tolerance, rel_tolerance, max_trials);
    search.begin(slope_at_zero);
    double step = 0.0, current = 0.0;
    int count = 0;
    while (search.propose(&step)) {
        search.observe(slope(step, user));
        current = step;
        count++;
    }
    if (search.stalled() && search.best_step() != current) {
        current = search.best_step();
        slope(current, user); // the caller's state moves to the chosen step, as GValue's last call does
        count++;
    }
    *final_step = current;
    if (evaluations) *evaluations = count;
    return TB2_OK;
}

int tb2_nlpcg_solve(tb2_nlpcg* s, double* d_u, const double* d_u_last, const double* d_fext, int solve_max_iterations, int* status,
                    int* iterations, double* error_out, double* error0_out)
{
    TB2_ARG(s && d_u && d_fext && status);
    tb2_mesh* m = s->group->mesh;
    DeviceGuard dg(m->device);
    const tb2_nlpcg_params& prm = s->prm;
    Ctx c;
    c.s = s;
    c.m = m;
    c.st = m->stream;
    c.n = s->eqs->neq;
    c.nb1 = (unsigned)((c.n + 255) / 256);
    c.vb = c.nb1 < (unsigned)kNlBlocks ? c.nb1 : (unsigned)kNlBlocks;
    c.u = d_u;
    c.ul = d_u_last;
    c.fext = d_fext;
    c.iteration = -1; // SolverT::InitStep
    c.w = nullptr;
    if (comm_active(m)) {
        if (!s->eq_owned.p) {
            TB2_CUDA(s->eq_owned.alloc(c.n));
            k_nl_eq_owned<<<c.nb1, 256, 0, c.st>>>(c.n, s->eqs->eq_node.p, comm_owned_mask(m), s->eq_owned.p);
        }
        c.w = s->eq_owned.p;
    }
    *status = TB2_SOLVER_CONTINUE;
    int rc = TB2_OK;
    double h[2], error = 0.0, error0 = 0.0;
    int num_iterations = 0, tan_iterations = 0, restart_count = -1;
    SecantSearch search(prm.max_step, prm.abs_tolerance, prm.line_search_tolerance, prm.line_search_iterations);

    // NLSolver::ExitIteration (NLSolver.cpp:675-756)
    auto exit_iteration = [&](int iter) {
        if (iter == -1) {
            error0 = error;
            return error0 < prm.abs_tolerance ? TB2_SOLVER_CONVERGED : TB2_SOLVER_CONTINUE;
        }
        const double rel = error / error0;
        if (rel > prm.divergence_tolerance) return TB2_SOLVER_FAILED;
        if (iter < prm.min_iterations - 1) return TB2_SOLVER_CONTINUE;
        if (rel < prm.rel_tolerance || error < prm.abs_tolerance) return TB2_SOLVER_CONVERGED;
        if (iter >= prm.max_iterations) return TB2_SOLVER_FAILED;
        return TB2_SOLVER_CONTINUE;
    };
#define NL_TRY(call)                 \
    do {                             \
        rc = (call);                 \
        if (rc != TB2_OK) goto done; \
    } while (0)

    NL_TRY(form_rhs(c, false, h));
    error = std::sqrt(h[0]);
    *status = exit_iteration(c.iteration);
    while (*status == TB2_SOLVER_CONTINUE && (solve_max_iterations < 0 || num_iterations < solve_max_iterations)) { // NLSolver.cpp:130-131
        num_iterations++;
        tan_iterations++;
        if (num_iterations == 1 || tan_iterations >= prm.restart) { // fReformTangentIterations = fRestart
            tan_iterations = 0;
            NL_TRY(form_preconditioner(c));
        }
        // ---- CGSearch: the new direction (in s->dir); G_a = direction . residual comes with it
        restart_count++;
        {
            ProfScope ps(m, kProfPcgVec, 3);
            if (restart_count == 0 || restart_count == prm.restart) {
                k_nl_steepest<<<c.vb, 256, 0, c.st>>>(c.n, s->R.p, s->minv.p, c.w, s->R_last.p, s->dir.p, s->partial.p);
                restart_count = 0;
            } else {
                k_nl_beta_partials<<<c.vb, 256, 0, c.st>>>(c.n, s->R.p, s->R_last.p, s->minv.p, c.w, s->partial.p);
                k_nl_sum2<<<1, 1024, 0, c.st>>>((int)c.vb, s->partial.p, s->red.p);
                if (comm_active(m)) NL_TRY(comm_allreduce_scalars(m, s->red.p, 2));
                k_nl_direction<<<c.vb, 256, 0, c.st>>>(c.n, s->red.p, s->R.p, s->minv.p, c.w, s->R_last.p, s->dir.p, s->partial.p);
            }
        }
        // ---- Update: secant search for the step s with R(u + s dir) . dir = 0 (PCGSolver_LS.cpp:213-348; tb2_linesearch.h)
        double rr_last = 0.0;
        if (prm.line_search_iterations == 0) {
            k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, s->eqs->eq_node.p, 1.0, s->dir.p, c.u); // NLSolver::Update: the full step
            c.iteration++;
            NL_TRY(form_rhs(c, false, h));
            rr_last = h[0];
        } else {
            NL_TRY(read2(c, h));
            double s_current = 0.0, G_trial = 0.0, rr = 0.0, step = 0.0;
            search.begin(h[0]); // G(0) = direction . residual came with the direction update
            while (search.propose(&step)) {
                NL_TRY(gvalue(c, step, s_current, G_trial, rr));
                search.observe(G_trial);
            }
            if (search.stalled() && search.best_step() != s_current) // settle on the best step tried
                NL_TRY(gvalue(c, search.best_step(), s_current, G_trial, rr));
            // The reference now forms the residual once more at the state GValue left (NLSolver.cpp:174-195): the same sweep on
            // the same displacements, bit for bit.  Only the solver's iteration number differs, which J2Simo3D reads
            // (J2Simo3D.cpp:83-84) -- so that case repeats the sweep and every other material reuses the last one.
            c.iteration++;
            if (s->group->mat.kind == TB2_J2_SIMO) {
                NL_TRY(form_rhs(c, false, h));
                rr = h[0];
            }
            rr_last = rr;
        }
        error = std::sqrt(rr_last);
        *status = exit_iteration(c.iteration);
    }
#undef NL_TRY
done:
    if (iterations) *iterations = c.iteration;
    if (error_out) *error_out = error;
    if (error0_out) *error0_out = error0;
    if (rc == TB2_ERR_BAD_JACOBIAN || rc == TB2_ERR_J2_LOCAL) { // NLSolver::Solve catches the exception and returns kFailed (NLSolver.cpp:247-262)
        *status = TB2_SOLVER_FAILED;
        return rc;
    }
    return rc;
}

int tb2_nlpcg_solve_host(tb2_nlpcg* s, double* h_u, const double* h_u_last, const double* h_fext, int solve_max_iterations, int* status,
                         int* iterations, double* error, double* error0)
{
    TB2_ARG(s && h_u && h_fext && status);
    tb2_mesh* m = s->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    DevBuf<double> u, ul, fext;
    TB2_CUDA(u.alloc(3 * m->nn));
    TB2_CUDA(fext.alloc(3 * m->nn));
    TB2_CUDA(cudaMemcpyAsync(u.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(fext.p, h_fext, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_u_last) {
        TB2_CUDA(ul.alloc(3 * m->nn));
        TB2_CUDA(cudaMemcpyAsync(ul.p, h_u_last, bytes, cudaMemcpyHostToDevice, m->stream));
    }
    const int rc = tb2_nlpcg_solve(s, u.p, h_u_last ? ul.p : nullptr, fext.p, solve_max_iterations, status, iterations, error, error0);
    TB2_CUDA(cudaMemcpyAsync(h_u, u.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return rc;
}

} // extern "C"

// ---- Newton driver ------------------------------------------------------------------------------------------------------
// NLSolver::Solve / Iterate / ExitIteration (solvers/NLSolver.cpp:57-263, 759-766, 675-756) with a CUDA_PCG_matrix as fLHS:
//   residual (K1 + ordered gather) -> |R| -> ExitIteration -> [tangent: GlobalMatrixT::Clear + K3] -> Jacobi-PCG (K6-K8) ->
//   FieldT::AssembleUpdate, all on the device; the host reads one norm per Newton iteration.
namespace {
// update = solve result in equation space -> u[active] += update (full Newton step, NLSolver::Update :635-641)
int newton_update(Ctx& c, const double* d_dx)
{
    ProfScope ps(c.m, kProfPcgVec);
    if (c.dyn) { // nNLHHTalpha::Corrector on the active equations (nNLHHTalpha.cpp:131-160): d += beta dt^2 da, v += gamma dt da, a += da
        k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, c.s->eqs->eq_node.p, c.dyn->dcorr_a, d_dx, c.u);
        k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, c.s->eqs->eq_node.p, c.dyn->vcorr_a, d_dx, c.v);
        k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, c.s->eqs->eq_node.p, 1.0, d_dx, c.a);
        return TB2_OK;
    }
    k_nl_update<<<c.nb1, 256, 0, c.st>>>(c.n, c.s->eqs->eq_node.p, 1.0, d_dx, c.u);
    return TB2_OK;
}
} // namespace

static int newton_core(tb2_nlpcg* s, tb2_matrix* A, const tb2_newton_params* np, const tb2_dynamics* dyn, double* d_u, double* d_v,
                       double* d_a, const double* d_u_last, const double* d_fext, int solve_max_iterations, int* status, int* iterations,
                       double* error_out, double* error0_out, int64_t* linear_iterations)
{
    TB2_ARG(s && A && np && d_u && d_fext && status);
    TB2_ARG(!dyn || (d_v && d_a && (dyn->mass_type == TB2_MASS_CONSISTENT || dyn->mass_type == TB2_MASS_LUMPED)));
    TB2_ARG(A->eqs == s->eqs);
    tb2_mesh* m = s->group->mesh;
    DeviceGuard dg(m->device);
    Ctx c;
    c.s = s;
    c.m = m;
    c.st = m->stream;
    c.n = s->eqs->neq;
    c.nb1 = (unsigned)((c.n + 255) / 256);
    c.vb = c.nb1 < (unsigned)kNlBlocks ? c.nb1 : (unsigned)kNlBlocks;
    c.u = d_u;
    c.ul = d_u_last;
    c.fext = d_fext;
    c.iteration = -1;
    c.w = nullptr;
    c.dyn = dyn;
    c.v = d_v;
    c.a = d_a;
    if (comm_active(m)) {
        if (!s->eq_owned.p) {
            TB2_CUDA(s->eq_owned.alloc(c.n));
            k_nl_eq_owned<<<c.nb1, 256, 0, c.st>>>(c.n, s->eqs->eq_node.p, comm_owned_mask(m), s->eq_owned.p);
        }
        c.w = s->eq_owned.p;
    }
    *status = TB2_SOLVER_CONTINUE;
    int rc = TB2_OK, num_iterations = 0, tan_iterations = 0;
    int64_t lin_total = 0;
    int lin_unconverged = 0;
    double lin_worst = 0.0;
    double h[2], error = 0.0, error0 = 0.0;
    const int reform = np->reform_tangent_iterations > 0 ? np->reform_tangent_iterations : 1;
    auto exit_iteration = [&](int iter) {
        if (iter == -1) {
            error0 = error;
            return error0 < np->abs_tolerance ? TB2_SOLVER_CONVERGED : TB2_SOLVER_CONTINUE;
        }
        const double rel = error / error0;
        if (rel > np->divergence_tolerance) return TB2_SOLVER_FAILED;
        if (iter < np->min_iterations - 1) return TB2_SOLVER_CONTINUE;
        if (rel < np->rel_tolerance || error < np->abs_tolerance) return TB2_SOLVER_CONVERGED;
        if (iter >= np->max_iterations) return TB2_SOLVER_FAILED;
        return TB2_SOLVER_CONTINUE;
    };
    rc = form_rhs(c, false, h);
    if (rc == TB2_OK) {
        error = std::sqrt(h[0]);
        *status = exit_iteration(c.iteration);
    }
    // NLSolver::Solve tests the bound in the loop condition (NLSolver.cpp:130-131): max_iterations == 0 does no iteration at all
    while (rc == TB2_OK && *status == TB2_SOLVER_CONTINUE && (solve_max_iterations < 0 || num_iterations < solve_max_iterations)) {
        num_iterations++;
        tan_iterations++;
        if (num_iterations == 1 || tan_iterations >= reform) { // NLSolver.cpp:143-160
            tan_iterations = 0;
            if ((rc = tb2_matrix_clear(A)) != TB2_OK) break;
            if (!dyn || dyn->constK != 0.0) {
                if ((rc = tb2_form_stiffness(s->group, A, c.u, c.ul, c.iteration)) != TB2_OK) break;
                if (dyn && dyn->constK != 1.0 && (rc = tb2_matrix_scale(A, dyn->constK)) != TB2_OK) break;
            }
            if (dyn && dyn->constM != 0.0 && (rc = tb2_form_mass(s->group, A, dyn->mass_type, dyn->constM)) != TB2_OK) break; // effective mass
            if ((rc = tb2_group_status(s->group, nullptr)) != TB2_OK) break;
        }
        // NLSolver::Iterate: fLHS->Solve(fRHS) -- the update from a zero start guess (the solve overwrites its argument)
        if (cudaMemsetAsync(s->dir.p, 0, c.n * sizeof(double), c.st) != cudaSuccess) { rc = TB2_ERR_CUDA; break; }
        int lin_it = 0;
        double lin_r = 0.0;
        // symmetric tangents: Jacobi-PCG; J2Simo3D's consistent tangent is non-symmetric (J2Simo3D.cpp:18-21): BiCGStab
        if (s->group->mat.kind == TB2_J2_SIMO)
            rc = tb2_matrix_bicgstab(A, s->R.p, s->dir.p, np->pcg_rel_tolerance, np->pcg_abs_tolerance, np->pcg_max_iterations, &lin_it, &lin_r);
        else
            rc = tb2_matrix_pcg(A, s->R.p, s->dir.p, np->pcg_rel_tolerance, np->pcg_abs_tolerance, np->pcg_max_iterations, &lin_it, &lin_r);
        lin_total += lin_it;
        if (rc != TB2_OK) break;
        {   // the reference's direct solve is exact; an iterative solve that ran out of iterations is reported (tb2_last_error)
            int conv = 1;
            double rel = 0.0;
            tb2_matrix_pcg_converged(A, &conv, &rel);
            if (!conv) {
                lin_unconverged++;
                if (rel > lin_worst) lin_worst = rel;
            }
        }
        if ((rc = newton_update(c, s->dir.p)) != TB2_OK) break;
        c.iteration++;
        if ((rc = form_rhs(c, false, h)) != TB2_OK) break;
        error = std::sqrt(h[0]);
        *status = exit_iteration(c.iteration);
    }
    if (iterations) *iterations = c.iteration;
    if (error_out) *error_out = error;
    if (error0_out) *error0_out = error0;
    if (linear_iterations) *linear_iterations = lin_total;
    if (rc == TB2_OK && lin_unconverged > 0)
        set_error("Newton: %d of %d linear solves stopped at pcg_max_iterations = %d without meeting their tolerance (largest |r|/|r0| = %.3e); "
                  "the updates of those iterations are inexact", lin_unconverged, num_iterations, np->pcg_max_iterations, lin_worst);
    if (rc != TB2_OK) *status = TB2_SOLVER_FAILED; // NLSolver::Solve: any exception -> kFailed (NLSolver.cpp:247-262)
    return rc;
}

extern "C" int tb2_newton_solve(tb2_nlpcg* s, tb2_matrix* A, const tb2_newton_params* np, double* d_u, const double* d_u_last,
                                const double* d_fext, int solve_max_iterations, int* status, int* iterations, double* error_out,
                                double* error0_out, int64_t* linear_iterations)
{
    return newton_core(s, A, np, nullptr, d_u, nullptr, nullptr, d_u_last, d_fext, solve_max_iterations, status, iterations, error_out,
                       error0_out, linear_iterations);
}

extern "C" int tb2_newton_solve_dynamic(tb2_nlpcg* s, tb2_matrix* A, const tb2_newton_params* np, const tb2_dynamics* dyn, double* d_u,
                                        double* d_v, double* d_a, const double* d_u_last, const double* d_fext, int solve_max_iterations,
                                        int* status, int* iterations, double* error_out, double* error0_out, int64_t* linear_iterations)
{
    TB2_ARG(dyn);
    return newton_core(s, A, np, dyn, d_u, d_v, d_a, d_u_last, d_fext, solve_max_iterations, status, iterations, error_out, error0_out,
                       linear_iterations);
}

extern "C" int tb2_newton_solve_dynamic_host(tb2_nlpcg* s, tb2_matrix* A, const tb2_newton_params* np, const tb2_dynamics* dyn, double* h_u,
                                             double* h_v, double* h_a, const double* h_u_last, const double* h_fext,
                                             int solve_max_iterations, int* status, int* iterations, double* error, double* error0,
                                             int64_t* linear_iterations)
{
    TB2_ARG(s && A && np && dyn && h_u && h_v && h_a && h_fext && status);
    tb2_mesh* m = s->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    DevBuf<double> u, v, a, ul, fext;
    TB2_CUDA(u.alloc(3 * m->nn));
    TB2_CUDA(v.alloc(3 * m->nn));
    TB2_CUDA(a.alloc(3 * m->nn));
    TB2_CUDA(fext.alloc(3 * m->nn));
    TB2_CUDA(cudaMemcpyAsync(u.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(v.p, h_v, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(a.p, h_a, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(fext.p, h_fext, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_u_last) {
        TB2_CUDA(ul.alloc(3 * m->nn));
        TB2_CUDA(cudaMemcpyAsync(ul.p, h_u_last, bytes, cudaMemcpyHostToDevice, m->stream));
    }
    const int rc = tb2_newton_solve_dynamic(s, A, np, dyn, u.p, v.p, a.p, h_u_last ? ul.p : nullptr, fext.p, solve_max_iterations, status,
                                            iterations, error, error0, linear_iterations);
    TB2_CUDA(cudaMemcpyAsync(h_u, u.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_v, v.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaMemcpyAsync(h_a, a.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return rc;
}

extern "C" int tb2_newton_solve_host(tb2_nlpcg* s, tb2_matrix* A, const tb2_newton_params* np, double* h_u, const double* h_u_last,
                                     const double* h_fext, int solve_max_iterations, int* status, int* iterations, double* error,
                                     double* error0, int64_t* linear_iterations)
{
    TB2_ARG(s && A && np && h_u && h_fext && status);
    tb2_mesh* m = s->group->mesh;
    DeviceGuard dg(m->device);
    const size_t bytes = 3 * m->nn * sizeof(double);
    DevBuf<double> u, ul, fext;
    TB2_CUDA(u.alloc(3 * m->nn));
    TB2_CUDA(fext.alloc(3 * m->nn));
    TB2_CUDA(cudaMemcpyAsync(u.p, h_u, bytes, cudaMemcpyHostToDevice, m->stream));
    TB2_CUDA(cudaMemcpyAsync(fext.p, h_fext, bytes, cudaMemcpyHostToDevice, m->stream));
    if (h_u_last) {
        TB2_CUDA(ul.alloc(3 * m->nn));
        TB2_CUDA(cudaMemcpyAsync(ul.p, h_u_last, bytes, cudaMemcpyHostToDevice, m->stream));
    }
    const int rc = tb2_newton_solve(s, A, np, u.p, h_u_last ? ul.p : nullptr, fext.p, solve_max_iterations, status, iterations, error,
                                    error0, linear_iterations);
    TB2_CUDA(cudaMemcpyAsync(h_u, u.p, bytes, cudaMemcpyDeviceToHost, m->stream));
    TB2_CUDA(cudaStreamSynchronize(m->stream));
    return rc;
}
