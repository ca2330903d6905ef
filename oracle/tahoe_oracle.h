/* oracle/tahoe_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of Tahoe's Hex8 continuum-solid hot path
 * (SURVEY.md section 8a).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call this.
 * The product (tahoe_b200/) never does.
 *
 * Parity status: PINNED against the reference itself -- oracle/_ref/tahoe_dump
 * (the unmodified reference compiled by oracle/build_ref.mk) writes
 * full-precision fixtures into tests/golden/ (tests/golden/make_golden.py);
 * tests/test_oracle_golden.py checks every function below against them.
 *
 * Conventions: node order of HexahedronT.cpp:23-25,
 * 8 integration points at +-1/sqrt(3) in node order with unit weights,
 * symmetric-tensor order 11,22,33,23,13,12 with tensor shear components,
 * F stored column-major F[i + 3*j], nodal arrays [node][dof], element
 * vectors [node a][dof i] -> 3*a+i, element matrices column-major 24x24.
 */
#ifndef TAHOE_ORACLE_H
#define TAHOE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SMALL_STRAIN = 0, ORC_TOTAL_LAGRANGIAN = 1, ORC_UPDATED_LAGRANGIAN = 2,
       ORC_SMALL_STRAIN_BBAR = 3 /* SmallStrainT with strain_displacement="B-bar" (kMeanDilBbar, SmallStrainT.cpp:337-374) */ };
enum { ORC_SSKSTV = 0, ORC_FDKSTV = 1, ORC_SIMO_ISO = 2, ORC_J2_SIMO = 3,
       ORC_EXPL_NEO = 4, ORC_EXPL_J2 = 5 /* <explicit_solid> materials: ExplNeoHookeanT, ExplJ2PlasticityT (hard[0] = sigma_Y, hard[1] = H) */ };
enum { ORC_OK = 0, ORC_BAD_JACOBIAN = 1, ORC_J2_LOCAL_FAIL = 2 };
enum { ORC_J2_NOTINIT = -1, ORC_J2_PLASTIC = 0, ORC_J2_ELASTIC = 1 }; /* J2SimoC0HardeningT.h:33-36 */
enum { ORC_HARD_LINEAR = 0, ORC_HARD_LINEAR_EXP = 1, ORC_HARD_POWER_LAW = 2, ORC_HARD_CUBIC_SPLINE = 3 };
enum { ORC_MAX_KNOTS = 16 };

typedef struct {
    int    kind;       /* ORC_SSKSTV ... */
    double mu, lambda, kappa, density;
    int    hard_kind;  /* J2: hardening function K(alpha) */
    double hard[4];    /* linear: K = hard[0]*alpha + hard[1];
                          linear_exponential: K = hard[0] + hard[1]*alpha + hard[2]*(1-exp(-alpha/hard[3]));
                          power_law: K = hard[0]*(hard[1] + hard[2]*alpha)^hard[3] (C1functions/PowerLawT.cpp:28-37) */
    int    nknots;     /* cubic_spline (C1functions/CubicSplineT.cpp): knots and the nknots+1 coefficient rows set by orc_material_set_spline */
    double knot_x[ORC_MAX_KNOTS];
    double spline[(ORC_MAX_KNOTS + 1) * 4];
} orc_material_t;

/* J2 history of one integration point, field order of
 * J2SimoC0HardeningT::LoadData (J2SimoC0HardeningT.cpp:429-452) */
typedef struct {
    double b_bar[6], unit_norm[6], beta_bar[6], b_bar_trial[6], beta_bar_trial[6];
    double internal[8]; /* alpha, stressnorm, dgamma, ftrial, mu_bar, mu_bar_bar, detF_tot, heat */
    int    flag;        /* ORC_J2_* */
} orc_j2_ip_t;

void orc_material_from_E_nu(orc_material_t* m, int kind, double E, double nu, double density);

/* a3: parent-domain tables. Na[ip][a], DNa[ip][d][a], w[ip] */
void orc_hex8_parent(double Na[8][8], double DNa[8][3][8], double w[8]);

/* a3: dN/dX at the 8 IPs of one element. X[a][d]. returns ORC_BAD_JACOBIAN if any det <= 0 */
int orc_hex8_shape(const double X[8][3], double dNdX[8][3][8], double det[8]);

/* K1: one element's internal force fe[3*a+i] (the +B^T sigma integral; Tahoe's RHS gets -fe).
 * u_last and j2 (8 IPs, may be NULL for non-J2), alloc = pointer to the element's
 * "IsAllocated" flag, iteration = GroupIterationNumber() */
int orc_element_force(int form, const orc_material_t* m, const double X[8][3], const double u[8][3],
                      const double u_last[8][3], orc_j2_ip_t* j2, int* alloc, int iteration, double fe[24]);

/* K3: one element's tangent Ke[r + 24*c] (whole matrix) */
int orc_element_stiffness(int form, const orc_material_t* m, const double X[8][3], const double u[8][3],
                          const double u_last[8][3], orc_j2_ip_t* j2, int* alloc, int iteration, double Ke[576]);

/* K4: one element's lumped mass me[a] (same on the 3 dofs of node a) */
int orc_element_lumped_mass(double density, const double X[8][3], double me[8]);

/* J2 history commit / reset over one element (J2SimoC0HardeningT::Update/Reset) */
void orc_j2_update(const orc_material_t* m, orc_j2_ip_t* j2);
void orc_j2_reset(orc_j2_ip_t* j2);

/* mesh-level sweeps, element order = serial reference order */
int orc_internal_force(int form, const orc_material_t* m, int64_t ne, const int32_t* conn /*[ne][8]*/,
                       const double* X /*[nn][3]*/, const double* u, const double* u_last,
                       orc_j2_ip_t* j2 /*[ne][8] or NULL*/, int* alloc /*[ne] or NULL*/, int iteration,
                       double* f /*[nn][3], accumulated into*/, int64_t* bad_elem);
int orc_lumped_mass(double density, int64_t ne, const int32_t* conn, const double* X, double* mass /*[nn][3] accumulated*/);

/* a24: equation numbers.  bc[nn*3] != 0 marks prescribed dofs.  eqnos 1-based, -1 prescribed. returns n_eq */
int64_t orc_set_equation_numbers(int64_t nn, const uint8_t* bc, int32_t* eqnos);

/* a22: sparsity of the active equations.  CSR in the form of MSRBuilderT::SetSuperLUData
 * (rowptr 0-based, colind 0-based sorted, diagonal in place).  Call with colind==NULL to size. */
int64_t orc_csr_structure(int64_t ne, const int32_t* conn, int64_t nn, const int32_t* eqnos, int64_t neq,
                          int upper_only, int64_t* rowptr /*[neq+1]*/, int32_t* colind);
/* MSR structure data of MSRBuilderT::SetMSRData: bindx[0..neq] row starts, then sorted off-diagonal cols */
int64_t orc_msr_structure(int64_t ne, const int32_t* conn, int64_t nn, const int32_t* eqnos, int64_t neq,
                          int upper_only, int32_t* bindx);

/* greedy element colouring in element order (no reference counterpart: SURVEY section 0.4) */
int orc_greedy_colouring(int64_t ne, const int32_t* conn, int64_t nn, int32_t* colour);

/* assemble K (full CSR) and the residual R = fext - fint on the active equations */
int orc_assemble_stiffness(int form, const orc_material_t* m, int64_t ne, const int32_t* conn,
                           const double* X, const double* u, const double* u_last, orc_j2_ip_t* j2, int* alloc,
                           int iteration, const int32_t* eqnos, int64_t neq, const int64_t* rowptr,
                           const int32_t* colind, double* val);

/* K6: y = A x on CSR (MSRMatrixT::Multx restated on the CSR form) */
void orc_csr_spmv(int64_t n, const int64_t* rowptr, const int32_t* colind, const double* val, const double* x, double* y);

/* K6-K8: linear Jacobi-PCG.  returns iterations; x is start guess and result */
int orc_pcg_jacobi(int64_t n, const int64_t* rowptr, const int32_t* colind, const double* val,
                   const double* b, double* x, double rtol, double atol, int max_iter, double* final_rnorm);

/* a19: central difference.  bc code per dof: 0 free, 1 fixed (kFix), 2 prescribed displacement (kDsp, value in bcval) */
void orc_cd_predictor(int64_t ndof, double dt, double* d, double* v, double* a, const uint8_t* bc, const double* bcval);
void orc_cd_corrector(int64_t ndof, double dt, double* v, double* a, const double* R, const double* mass, const uint8_t* bc);

/* <explicit_solid> on Hex8 (ExplicitElementT::BatchedInternalForce, ExplicitElementT.cpp:649-993; Hex8KernelT;
 * ExplNeoHookeanT / ExplJ2PlasticityT; CFL estimate :404-478; fixed mass scaling :492-571 and LHSDriver :576-618).
 * hist[ne][8][16] (ExplJ2PlasticityT.h:8-11: F_n row-major, sigma_n Voigt, eps_p) is updated on EVERY force evaluation. */
void orc_explicit_solid_init_history(int64_t ne, double* hist);
int orc_explicit_solid_force(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, const double* u, double* hist,
                             double* f /*[nn][3] accumulated*/);
double orc_explicit_solid_stable_dt(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X);
void orc_explicit_solid_mass_scale(const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, double target_dt,
                                   double scale_factor, double* scale /*[ne]*/);
int orc_lumped_mass_scaled(double density, int64_t ne, const int32_t* conn, const double* X, const double* scale, double* mass);
/* inertia (ContinuumElementT::FormMa / FormMass, ContinuumElementT.cpp:678-1002): mass_type 1 consistent, 2 lumped (HRZ).
 * orc_inertial_force: f[nn][3] += scale * M a;  orc_assemble_mass: CSR val += constM * M */
int orc_inertial_force(double density, int mass_type, int64_t ne, const int32_t* conn, const double* X, const double* acc,
                       double scale, double* f);
int orc_assemble_mass(double density, int mass_type, double constM, int64_t ne, const int32_t* conn, const double* X,
                      const int32_t* eqnos, const int64_t* rowptr, const int32_t* colind, double* val);
/* CubicSplineT::SetSpline (CubicSplineT.cpp:254-324): fixity 0 parabolic, 1 free_run; returns nonzero on bad input */
int orc_material_set_spline(orc_material_t* m, int n, const double* x, const double* y, int fixity);
/* the hardening function and its derivative (known-answer checks) */
void orc_hardening(const orc_material_t* m, double alpha, double* K, double* dK);
/* The search that feeds it: Contact3DT::SetActiveStrikers (Contact3DT.cpp:226-334) on the configuration x[nn][3].  For every
 * striker the facet of smallest |h| among the facets -- of surfaces the striker is not a node of -- that Contact3DT::Intersect
 * (:336-391) accepts: |h| <= sqrt(|a x b|)/2 and the projection inside the triangle within |a x b|/50; the first facet in surface /
 * facet order wins a tie.  (The reference looks for candidates through a search grid around the facet midpoint; every striker
 * Intersect accepts lies inside that region, so the grid only prunes.)  hit[ns] = index into facets, -1 for a free striker. */
int orc_contact_search(int64_t nfacets, const int32_t* facets /*[nf][3]*/, const int32_t* facet_surface /*[nf]*/, int64_t nstrikers,
                       const int32_t* strikers /*[ns]*/, int64_t nn, const double* x, int32_t* hit, double* gap /*[ns] h of the hit, or NULL*/);
/* SURVEY 8(f)-4, contact_3D_penalty: PenaltyContact3DT::RHSDriver (PenaltyContact3DT.cpp:262-500) over a given list of active
 * striker-facet pairs (Contact3DT::SetActiveInteractions builds it: three facet nodes, then the striker; 0-based).  Per pair with
 * penetration h = n . (x_s - centroid) < 0 on the configuration X + constKd u: the penalty force dphi dh/du with dphi = -K h area
 * (Contact3DT::Set_dn_du for the normal's variation, Contact3DT.cpp:176-220), velocity-based regularised Coulomb friction
 * (mu > 0 and v given) and normal viscous damping (visc > 0 and v given), each split -1/3 on the facet nodes and +1 on the striker.
 * f[nn][3] receives the pairs' 12-vectors added in pair order (SolverT::AssembleRHS); returns the number of pairs in contact,
 * *h_max the deepest penetration (<= 0). */
int orc_contact_force(int64_t npairs, const int32_t* pairs /*[npairs][4]*/, const double* area /*[npairs]*/, double K, double mu, double eps,
                      double visc, double constKd, int64_t nn, const double* X, const double* u, const double* v /* or NULL */, double* f,
                      double* h_max);
/* natural_bc tractions (ContinuumElementT::ApplyTractionBC, ContinuumElementT.cpp:514-665): f[nn][3] += consistent nodal forces of
 * ncards facet cards (elem, facet 0-based; tract[card][4][3] nodal traction vectors in facet-node order; coord_system 0 global, 1 local) */
int orc_traction_force(int64_t ncards, const int32_t* elem, const int32_t* facet, const int32_t* conn, const double* X,
                       const double* tract, int coord_system, double scale, double* f);
/* one material evaluation (known-answer tests): F row-major [9], h[16] in/out (EXPL_J2), sig[6] */
void orc_explicit_material_stress(const orc_material_t* m, const double* F, double* h, double* sig);

/* nodal Cauchy stress as SolidElementT::ComputeOutput writes it (extrapolated with HexahedronT::SetExtrapolation,
 * averaged over the elements at a node); SSKStV (incl. B-bar), FDKStV, SimoIso3D.  out[nn][6], order 11,22,33,23,13,12 */
int orc_nodal_stress(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X, const double* u,
                     double* out);
/* the same with a history material (J2Simo3D): last converged displacement, element history, iteration number */
int orc_nodal_stress_history(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X,
                             const double* u, const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* out);

/* a21: nonlinear preconditioned CG with secant line search, PCGSolver_LS (solvers/PCGSolver_LS.cpp:107-371) inside
 * NLSolver::Solve / ExitIteration (solvers/NLSolver.cpp:57-263, 675-756), preconditioner = DiagonalMatrixT kDiagOnly
 * (DiagonalMatrixT.cpp:107-113, 267-310).  u[nn][3] in/out (prescribed dofs already hold their values). */
enum { ORC_NLPCG_CONTINUE = 0, ORC_NLPCG_CONVERGED = 1, ORC_NLPCG_FAILED = 2 }; /* SolverT::SolutionStatusT (SolverT.h:59-62) */
typedef struct {
    int    restart;              /* PCGSolver_LS.cpp:53 */
    int    ls_iterations;        /* line_search_iterations */
    double ls_tolerance;         /* line_search_tolerance */
    double max_step;
    double abs_tol, rel_tol, div_tol; /* NLSolver: fZeroTolerance, fTolerance, fDivTolerance */
    int    max_iterations, min_iterations; /* NLSolver: fMaxIterations, fMinIterations */
    int    solve_max_iterations; /* the argument of Solve(int): -1 = no limit */
} orc_nlpcg_params_t;
int orc_stiffness_diagonal(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, const double* X, const double* u,
                           const double* u_last, orc_j2_ip_t* j2, int* alloc, int iteration, double* diag /*[nn][3] accumulated*/);
/* returns ORC_NLPCG_* (or -ORC_BAD_JACOBIAN ...); *iterations = SolverT::IterationNumber() at exit */
int orc_nlpcg_solve(int form, const orc_material_t* m, int64_t ne, const int32_t* conn, int64_t nn, const double* X, double* u,
                    const double* u_last, orc_j2_ip_t* j2, int* alloc, const int32_t* eqnos, int64_t neq, const double* fext,
                    const orc_nlpcg_params_t* prm, int* iterations, double* error, double* error0);

#ifdef __cplusplus
}
#endif
#endif
