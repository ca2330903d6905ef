/* tahoe_b200.h -- C ABI of libtahoe_b200.so: the B200 (sm_100a, FP64) implementation of Tahoe's
 * Hex8 continuum-solid hot path.
 *
 * Tahoe has no FFI today: the boundary is three C++ abstract classes.  Every entry point
 * below names the reference member function (file:line under the Tahoe source tree) whose work it
 * performs, so that the C++ plugin classes in tahoe_b200/host/ (CudaSolidElementT : SolidElementT's
 * subclasses, CudaPCGMatrixT : MSRMatrixT, CudaPCGSolverT : PCGSolver_LS, CudaExplicitSolverT : SolverT with
 * CudaExplicitCDIntegrator : ExplicitCDIntegrator, CudaPenaltyContact3DT : PenaltyContact3DT, FastGeomInputT : TahoeInputT) are thin
 * forwarding shells.  INTEGRATION.md shows the reference-side registration.
 * The host-only entry points (tb2_geom_*, tb2_partition_*, tb2_secant_search_host) make no CUDA call and work without a device.
 *
 * Conventions (identical to the reference):
 *   - nodal arrays are [node][dof] doubles (dArray2DT), ndof = nsd = 3;
 *   - connectivity is [element][8] int32, 0-based, HexahedronT node order (HexahedronT.cpp:23-25);
 *   - equation numbers are 1-based, <= 0 means prescribed (FieldT.cpp:635-659);
 *   - symmetric tensors are ordered 11,22,33,23,13,12 (dSymMatrixT);
 *   - 8 integration points, +-1/sqrt(3), in node order, unit weights (HexahedronT.cpp:1540-1547).
 *   - all arithmetic is FP64.  There is no CPU fallback: every call fails with TB2_ERR_CUDA when no
 *     sm_100-class device is usable.
 *
 * Pointers named d_* are DEVICE pointers on the mesh's device; pointers named h_* are HOST pointers.
 * All calls are synchronous with respect to the host unless stated otherwise; one CUDA stream per mesh.
 * Return value: 0 (TB2_OK) or a tb2_status; the mapping to Tahoe's ExceptionT::CodeT is in tb2_status.
 */
#ifndef TAHOE_B200_H
#define TAHOE_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TB2_OK = 0,
    TB2_ERR_BAD_JACOBIAN = 1, /* ExceptionT::kBadJacobianDet: det J <= 0 (ParentDomainT.cpp:451) or det F <= 0 (TotalLagrangianT.cpp:127) */
    TB2_ERR_J2_LOCAL = 2,     /* ExceptionT::kGeneralFail: J2 local Newton failed (J2SimoC0HardeningT.cpp:203-204) */
    TB2_ERR_CUDA = 3,         /* ExceptionT::kGeneralFail: CUDA runtime error / no device */
    TB2_ERR_ARG = 4,          /* ExceptionT::kBadInputValue */
    TB2_ERR_SIZE = 5,         /* ExceptionT::kSizeMismatch / kOutOfRange */
    TB2_ERR_PCG_BREAKDOWN = 6,/* ExceptionT::kGeneralFail: p.Ap <= 0 */
    TB2_ERR_COMM = 7          /* ExceptionT::kMPIFail */
} tb2_status;

typedef enum { TB2_SMALL_STRAIN = 0, TB2_TOTAL_LAGRANGIAN = 1, TB2_UPDATED_LAGRANGIAN = 2,
               TB2_SMALL_STRAIN_BBAR = 3 /* SmallStrainT with strain_displacement="B-bar": mean-dilatation B-bar of Hughes (4.5.11-23),
                                            SmallStrainT.cpp:337-374,404-421, SolidElementT::Set_B_bar SolidElementT.cpp:956-1044 */
             } tb2_formulation;
/* materials: SSKStV (Hookean/KStV/SSKStV.cpp), FDKStV (FDKStV.cpp), SimoIso3D (Simo/SimoIso3D.cpp), J2Simo3D (plasticity_J2/J2Simo3D.cpp) */
typedef enum { TB2_SSKSTV = 0, TB2_FDKSTV = 1, TB2_SIMO_ISO = 2, TB2_J2_SIMO = 3,
               /* the materials of <explicit_solid> (ExplicitElementT), valid with TB2_UPDATED_LAGRANGIAN:
                  ExplNeoHookeanT (elements/explicit/materials/ExplNeoHookeanT.cpp:79-111) and ExplJ2PlasticityT
                  (ExplJ2PlasticityT.cpp:87-310; hard[0] = sigma_Y, hard[1] = hardening modulus H) */
               TB2_EXPL_NEO_HOOKEAN = 4, TB2_EXPL_J2 = 5 } tb2_material_kind;
typedef enum { TB2_HARD_LINEAR = 0, TB2_HARD_LINEAR_EXP = 1, TB2_HARD_POWER_LAW = 2, TB2_HARD_CUBIC_SPLINE = 3 } tb2_hardening_kind;
enum { TB2_MAX_KNOTS = 16 };
enum { TB2_SPLINE_PARABOLIC = 0, TB2_SPLINE_FREE_RUN = 1 }; /* CubicSplineT::FixityT (CubicSplineT.h:29-30) */
/* kinematic boundary condition codes per dof (KBC_CardT::CodeT subset used by nExplicitCD::ConsistentKBC, nExplicitCD.cpp:20-69) */
typedef enum { TB2_BC_FREE = 0, TB2_BC_FIX = 1, TB2_BC_DSP = 2 } tb2_bc_code;

typedef struct {
    int32_t kind;      /* tb2_material_kind */
    int32_t hard_kind; /* tb2_hardening_kind (J2 only) */
    double  mu, lambda, kappa, density; /* IsotropicT (materials/primitives/IsotropicT.cpp:32-45) */
    double  hard[4];   /* linear: K = hard[0]*alpha + hard[1] (C1functions/LinearT.h:71);
                          linear_exponential: K = hard[0] + hard[1]*alpha + hard[2]*(1 - exp(-alpha/hard[3]));
                          power_law: K = hard[0]*(hard[1] + hard[2]*alpha)^hard[3] (C1functions/PowerLawT.cpp:28-37) */
    /* cubic_spline (C1functions/CubicSplineT.cpp): the <OrderedPair> knots and the fixity; the library forms the spline
       coefficients as CubicSplineT::SetSpline does (:254-324).  3 <= num_knots <= TB2_MAX_KNOTS, knot_x ascending. */
    int32_t num_knots, spline_fixity;
    double  knot_x[TB2_MAX_KNOTS], knot_y[TB2_MAX_KNOTS];
} tb2_material;

typedef struct tb2_mesh      tb2_mesh;      /* device-resident connectivity + reference coordinates (ModelManagerT / ElementBaseT::fConnectivities) */
typedef struct tb2_group     tb2_group;     /* one continuum-solid element group (SolidElementT subclass + its material + history) */
typedef struct tb2_equations tb2_equations; /* equation numbers + sparsity (FieldT::fEqnos, MSRBuilderT) */
typedef struct tb2_matrix    tb2_matrix;    /* device CSR global matrix (GlobalMatrixT subclass; MSRMatrixT semantics) */
typedef struct tb2_geom      tb2_geom;      /* a parsed TahoeII .geom file (host memory) */
typedef struct tb2_traction  tb2_traction;
typedef struct tb2_contact   tb2_contact;   /* contact_3D_penalty: active striker-facet pairs of one PenaltyContact3DT group */  /* natural_bc traction cards of one element group (ContinuumElementT::fTractionList) */
typedef struct tb2_explicit  tb2_explicit;  /* d, v, a, lumped mass, BCs on device: FieldT + nExplicitCD + DiagonalMatrixT */

/* ---- library ---------------------------------------------------------------------------------- */
const char* tb2_version(void);
const char* tb2_last_error(void);          /* text of the last CUDA / argument error on this thread */
int tb2_device_count(int* count);
/* raw device memory helpers for hosts that do not bring their own allocator (the C++ plugin classes) */
int tb2_malloc(int device, size_t bytes, void** d_ptr);
int tb2_free(int device, void* d_ptr);
int tb2_memcpy_h2d(int device, void* d_dst, const void* h_src, size_t bytes);
int tb2_memcpy_d2h(int device, void* h_dst, const void* d_src, size_t bytes);
int tb2_host_register(void* h_ptr, size_t bytes);   /* pin a Tahoe-owned dArrayT for async copies */
int tb2_host_unregister(void* h_ptr);

/* ---- instrumentation ---------------------------------------------------------------------------
 * Per-kernel device times measured with CUDA events on the mesh stream, and the number of kernels this library launched.
 * Categories: 0 element internal force (K1), 1 node gather + central-difference update (K5), 2 predictor, 3 SpMV (K6),
 * 4 PCG vector kernels (K7/K8), 5 stiffness assembly (K3), 6 interface exchange, 7 other.
 * tb2_profile_end synchronises the stream; h_ms[8], h_count[8]. */
int tb2_profile_reserve(tb2_mesh* mesh, int64_t records); /* pre-create the event pairs so that none is created inside a timed region */
int tb2_profile_begin(tb2_mesh* mesh);
int tb2_profile_end(tb2_mesh* mesh, double* h_ms, int64_t* h_count, int64_t* kernel_launches);
int tb2_mesh_synchronize(tb2_mesh* mesh);
/* measured FP64 FMA throughput of the device (dependent-FMA chains, best of 5): the roofline denominator of K1 / K3 */
int tb2_measure_fp64_peak(int device, double* tflops);

/* ---- mesh (ElementBaseT::DefineElements ElementBaseT.cpp:592-658, ModelManagerT coordinates) --- */
/* Uploads connectivity (re-laid-out SoA [8][ne_pad]) and reference coordinates, builds the
 * node->(element,local node) incidence used by the deterministic gather (replaces the serial
 * scatter of SolverT::AssembleRHS, SolverT.cpp:446-477). */
int tb2_mesh_create(int device, int64_t num_nodes, int64_t num_elements, const int32_t* h_conn, const double* h_coords,
                    tb2_mesh** mesh);
int tb2_mesh_destroy(tb2_mesh* mesh);
int tb2_mesh_sizes(const tb2_mesh* mesh, int64_t* num_nodes, int64_t* num_elements);
int tb2_mesh_device(const tb2_mesh* mesh, int* device);
void* tb2_mesh_stream(const tb2_mesh* mesh); /* cudaStream_t all kernels of this mesh are launched on */
/* Greedy element colouring in element order (no reference counterpart; pinned to
 * oracle/tahoe_oracle.c:orc_greedy_colouring).  h_colour[num_elements], returns the colour count. */
int tb2_mesh_colouring(tb2_mesh* mesh, int32_t* h_colour, int32_t* num_colours);

/* ---- TahoeII text geometry at scale: what ModelManagerT gets from TahoeInputT / ModelFileT
 * (toolbox/src/dataio/input/TahoeInputT.cpp, database/ModelFileT.cpp), parsed with all host threads.  No CUDA calls.
 * Ids come back 0-based: connectivity rows in file order, node-set members, side sets as (element in its block, facet).
 * Element and node sections may be inline or in the external files the main file names (cube.1.geom style).  A negative
 * node-set entry (the files' "all model nodes" marker, beam.1.geom) comes back as -1. */
int tb2_geom_open(const char* path, tb2_geom** geom);
int tb2_geom_close(tb2_geom* geom);
int tb2_geom_sizes(const tb2_geom* geom, int64_t* nn, int32_t* nsd, int32_t* nblocks, int32_t* nnodesets, int32_t* nsidesets);
int tb2_geom_coords(const tb2_geom* geom, double* h_X /*[nn][3], zero-padded when nsd < 3*/);
int tb2_geom_block(const tb2_geom* geom, int32_t block, int32_t* id, int64_t* nel, int32_t* nen, int32_t* h_conn /*[nel][nen] or NULL*/);
int tb2_geom_nodeset(const tb2_geom* geom, int32_t set, int32_t* id, int64_t* n, int32_t* h_nodes /*[n] or NULL*/);
int tb2_geom_sideset(const tb2_geom* geom, int32_t set, int32_t* id, int32_t* block_id, int64_t* n, int32_t* h_sides /*[n][2] or NULL*/);

/* ---- element group (SolidElementT / SmallStrainT / TotalLagrangianT / UpdatedLagrangianT) -------- */
int tb2_group_create(tb2_mesh* mesh, int formulation, const tb2_material* material, tb2_group** group);
int tb2_group_destroy(tb2_group* group);
/* SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295) with FormKd (SmallStrainT.cpp:255-282,
 * TotalLagrangianT.cpp:107-144, UpdatedLagrangianT.cpp:145-171): d_fint[nn][3] = sum_e B^T sigma
 * (Tahoe's RHS receives -d_fint).  d_u_last may be NULL unless the material is J2.  iteration is
 * ElementSupportT::IterationNumber() (-1 on the first RHS of a step: J2Simo3D.cpp:83-84).
 * Asynchronous on the mesh stream; errors are reported by tb2_group_status. */
int tb2_form_internal_force(tb2_group* group, const double* d_u, const double* d_u_last, int iteration, double* d_fint);
/* same through host arrays: the call a host-resident RHSDriver() makes (H2D u, kernels, D2H fint) */
int tb2_form_internal_force_host(tb2_group* group, const double* h_u, const double* h_u_last, int iteration, double* h_fint);
/* first failing element (0-based) and code since the last call; resets the flag.  Synchronises. */
int tb2_group_status(tb2_group* group, int64_t* bad_element);
/* ContinuumElementT::FormMass kLumpedMass (ContinuumElementT.cpp:767-842) summed to nodes: d_mass[nn][3] */
int tb2_form_lumped_mass(tb2_group* group, double* d_mass);
int tb2_form_lumped_mass_host(tb2_group* group, double* h_mass);
/* inertia branches of the element loops (implicit dynamics): ContinuumElementT::MassTypeT and
 * FormMa (ContinuumElementT.cpp:868-1002, called from SolidElementT::ElementRHSDriver :1243-1265): d_f[nn][3] = scale * M a with the
 * consistent (sum_ip rho w detJ0 N_a N_b) or lumped (HRZ diagonal, as tb2_form_lumped_mass) element mass on the reference
 * configuration.  Tahoe's RHS receives -constMa * (this).  On a partitioned mesh the result is the rank's partial sum, like the
 * internal force: follow with tb2_comm_sum_interface (the same holds for tb2_traction_form; tb2_form_mass adds to the rank's
 * unassembled sub-domain matrix, which the distributed PCG already sums on the interface rows). */
enum { TB2_MASS_CONSISTENT = 1 /* kConsistentMass */, TB2_MASS_LUMPED = 2 /* kLumpedMass */ };
int tb2_form_inertial_force(tb2_group* group, int mass_type, double scale, const double* d_acc, double* d_f);
int tb2_form_inertial_force_host(tb2_group* group, int mass_type, double scale, const double* h_acc, double* h_f);
/* ElementCardT status flags (ElementBaseT::SetStatus; the element loops skip ElementCardT::kOFF elements: SolidElementT.cpp:1116,
 * 1177).  h_off[ne]: 1 = off, 0 = on; NULL = all on.  Off elements contribute no force, tangent or mass. */
int tb2_group_set_element_status(tb2_group* group, const uint8_t* h_off);

/* <explicit_solid> extras.  ExplicitElementT::ComputeStableTimeStep (ExplicitElementT.cpp:404-478): min over
 * the elements of h / c, h = cbrt of the three-diagonal volume estimate, c = sqrt((kappa + 4 mu / 3) / rho). */
int tb2_group_stable_time_step(tb2_group* group, double* dt);
/* ExplicitElementT::ApplyMassScaling, fixed type (:492-571): elements with h / c < target_dt * scale_factor get the mass factor
 * (target / dt_elem)^2, applied to the density in the lumped-mass assembly (LHSDriver :576-618), i.e. by every later
 * tb2_form_lumped_mass / tb2_explicit_create of this group.  h_scale[ne] (may be NULL) receives the factors. */
int tb2_group_set_mass_scaling(tb2_group* group, double target_dt, double scale_factor, int64_t* num_scaled, double* max_factor,
                               double* h_scale);
/* ExplJ2PlasticityT history in the reference's layout [ip][16][element] (ExplicitElementT.h:75-79): h_hist[8][16][ne] */
int tb2_group_get_explicit_history(tb2_group* group, double* h_hist);

/* Nodal stress output: SolidElementT::ComputeOutput, iNodalStress branch (SolidElementT.cpp:1352-1840) -- Cauchy
 * stress at the integration points, extrapolated with HexahedronT::SetExtrapolation (HexahedronT.cpp:2099-2150) and averaged over
 * the elements at each node (GroupAverageT.cpp:40-49,190-205).  d_stress[nn][6], order 11,22,33,23,13,12.  SSKStV (incl. B-bar),
 * FDKStV and SimoIso3D (J2Simo3D through tb2_group_nodal_stress_at); other materials return TB2_ERR_ARG.  Switched-off elements take
 * no part in the values or in the averaging counts. */
int tb2_group_nodal_stress(tb2_group* group, const double* d_u, double* d_stress);
int tb2_group_nodal_stress_host(tb2_group* group, const double* h_u, double* h_stress);
/* the same for a history material (J2Simo3D): SolidElementT::ComputeOutput calls J2Simo3D::s_ij at every integration point with the
 * element history, the last converged displacement (FiniteStrainT::SetGlobalShape, F_last) and the solver's iteration number, before
 * the step's history update (FEManagerT::CloseStep, FEManagerT.cpp:639-645).  d_u_last = NULL gives tb2_group_nodal_stress. */
int tb2_group_nodal_stress_at(tb2_group* group, const double* d_u, const double* d_u_last, int iteration, double* d_stress);
int tb2_group_nodal_stress_at_host(tb2_group* group, const double* h_u, const double* h_u_last, int iteration, double* h_stress);

/* J2 history: SolidElementT::CloseStep -> J2Simo3D::UpdateHistory (J2SimoC0HardeningT.cpp:341-384), ResetStep -> ResetHistory (:387-407) */
int tb2_group_close_step(tb2_group* group);
int tb2_group_reset_step(tb2_group* group);
/* J2 history download in the reference's ElementCardT layout (J2SimoC0HardeningT.cpp:429-452):
 * h_data[ne][5*48+64] doubles, h_flags[ne][8], h_alloc[ne] (restart hand-off, ContinuumElementT.cpp:217-249) */
int tb2_group_get_history(tb2_group* group, double* h_data, int32_t* h_flags, int32_t* h_alloc);
int tb2_group_set_history(tb2_group* group, const double* h_data, const int32_t* h_flags, const int32_t* h_alloc);

/* ---- natural_bc tractions: ContinuumElementT::ApplyTractionBC (ContinuumElementT.cpp:514-665).
 * One card per loaded element facet, as ContinuumElementT::TakeNaturalBC builds them (:1008-1123): h_elem[ncards] element in the
 * group (0-based), h_facet[ncards] facet 0..5 in HexahedronT::NodesOnFacet numbering (HexahedronT.cpp:1913-1918; the .geom side-set
 * value minus one), h_tract[ncards][4][3] traction vectors at the four facet nodes in facet-node order, coord_system as
 * Traction_CardT::CoordSystemT.  tb2_traction_form integrates the cards on the initial coordinates with the traction scaled by
 * `scale` (the schedule value, Traction_CardT::CurrentValue) and sums them into d_f[nn][3] in card order (accumulate != 0: added
 * to d_f, else d_f is overwritten).  A degenerate facet with a local frame returns TB2_ERR_BAD_JACOBIAN. */
enum { TB2_TRACTION_GLOBAL = 0 /* Traction_CardT::kCartesian */, TB2_TRACTION_LOCAL = 1 /* kLocal: (t1, t2, normal) */ };
int tb2_traction_create(tb2_mesh* mesh, int64_t ncards, const int32_t* h_elem, const int32_t* h_facet, const double* h_tract,
                        int coord_system, tb2_traction** traction);
int tb2_traction_destroy(tb2_traction* traction);
int tb2_traction_form(tb2_traction* traction, double scale, int accumulate, double* d_f);
int tb2_traction_form_host(tb2_traction* traction, double scale, int accumulate, double* h_f);

/* ---- contact_3D_penalty force: PenaltyContact3DT::RHSDriver (PenaltyContact3DT.cpp:262-500).
 * The neighbour search stays host code (Contact3DT::SetActiveInteractions, Contact3DT.cpp:100-175; it runs at relaxation points, not
 * per step); the host hands over its result with tb2_contact_set_pairs: h_pairs[npairs][4] = the three facet nodes and the striker
 * (0-based; the rows of the group's connectivity), h_striker_area[npairs] = ContactT::fStrikerArea of each pair's striker.
 * tb2_contact_form evaluates every pair on X + constKd u (constKd = eIntegratorT::FormKd): for a closed gap h = n . (x_s - centroid) < 0
 * the penalty force -K h area dh/du, velocity-based regularised Coulomb friction (friction_coefficient > 0) and normal viscous damping
 * (viscous_damping > 0), both from d_v[nn][3]; the pairs' 12-vectors are summed into d_f[nn][3] in pair order, the order of
 * ElementSupportT::AssembleRHS in the reference's loop (accumulate != 0: added to d_f, else d_f is overwritten).  The sign is that of
 * the residual the reference assembles.  tb2_contact_tracking returns what ContactT::SetTrackingData receives: the number of pairs in
 * contact and the deepest penetration (<= 0) of the last evaluation.  The slip-based friction of static analyses (per-striker history,
 * PenaltyContact3DT.cpp:427-447) and the contact tangent (LHSDriver) are not on the device. */
int tb2_contact_create(tb2_mesh* mesh, double penalty_stiffness, double friction_coefficient, double friction_epsilon_velocity,
                       double viscous_damping, tb2_contact** contact);
int tb2_contact_destroy(tb2_contact* contact);
int tb2_contact_set_pairs(tb2_contact* contact, int64_t npairs, const int32_t* h_pairs, const double* h_striker_area);
int tb2_contact_form(tb2_contact* contact, double constKd, const double* d_u, const double* d_v /* NULL without friction / damping */,
                     int accumulate, double* d_f);
int tb2_contact_form_host(tb2_contact* contact, double constKd, const double* h_u, const double* h_v, int accumulate, double* h_f);
int tb2_contact_tracking(tb2_contact* contact, int* num_contact, double* h_max);
/* The search on the device, for hosts that keep the whole loop resident: Contact3DT::SetActiveStrikers with Contact3DT::Intersect
 * (Contact3DT.cpp:226-391).  tb2_contact_set_surfaces hands over what the search works on -- the triangulated contact surfaces
 * (h_facets[nfacets][3] node triples, h_facet_surface[nfacets] the surface each belongs to, at most 32 surfaces), the striker nodes
 * and their tributary areas (ContactT::fStrikerTags / fStrikerArea).  tb2_contact_search looks, for every striker, on the configuration
 * X + d_u for the accepted facet of smallest |h| among the surfaces the striker is not a node of (first facet on a tie) and makes the
 * active pairs, in striker order, the group's pair list (the reference's row order follows its search-grid traversal; the set of pairs
 * is the same).  With such a group attached, tb2_explicit_run searches once before its first step and after every step. */
int tb2_contact_set_surfaces(tb2_contact* contact, int64_t nfacets, const int32_t* h_facets, const int32_t* h_facet_surface,
                             int64_t nstrikers, const int32_t* h_strikers, const double* h_striker_area);
int tb2_contact_search(tb2_contact* contact, const double* d_u, int64_t* npairs);
int tb2_contact_get_pairs(tb2_contact* contact, int64_t* npairs, int32_t* h_pairs /*[npairs][4] or NULL*/, double* h_area /*or NULL*/);
int tb2_contact_has_surfaces(const tb2_contact* contact);

/* ---- explicit central difference (nExplicitCD.cpp:72-139, DiagonalMatrixT.cpp:267-323, FieldT.cpp:531-556) */
int tb2_explicit_create(tb2_group* group, tb2_explicit** ex); /* forms and inverts the lumped mass */
int tb2_explicit_destroy(tb2_explicit* ex);
int tb2_explicit_set_state(tb2_explicit* ex, const double* h_d, const double* h_v, const double* h_a);
int tb2_explicit_get_state(tb2_explicit* ex, double* h_d, double* h_v, double* h_a);
/* h_code[nn][3] tb2_bc_code, h_value[nn][3] prescribed displacement, h_fext[nn][3] nodal forces (FieldT::FormRHS, FieldT.cpp:390-411).
 * Any pointer may be NULL to keep the current array. */
int tb2_explicit_set_bc(tb2_explicit* ex, const uint8_t* h_code, const double* h_value, const double* h_fext);
/* sparse refresh of prescribed displacement values (KBC controllers driven by a schedule, nExplicitCD::ConsistentKBC
 * nExplicitCD.cpp:20-69): value[h_dofs[k]] = h_values[k], h_dofs = nodal dof indices 3 n + i */
int tb2_explicit_update_bc_values(tb2_explicit* ex, int64_t count, const int64_t* h_dofs, const double* h_values);
/* FEManagerT::InitialCondition (FEManagerT.cpp:2034): a = M^-1 (fext - fint(d)) on free dofs */
/* A contact_3D_penalty group in the resident step: before every element sweep the attached group's force is re-formed on the predicted
 * d, v (what PenaltyContact3DT::RHSDriver sees inside FEManagerT::FormRHS) and enters the residual beside s fext - fint; also in
 * tb2_explicit_initial_condition.  The pair list may be replaced between runs (tb2_contact_set_pairs, after the host's search at a
 * relaxation point).  NULL detaches.  Single-GPU runs. */
int tb2_explicit_attach_contact(tb2_explicit* ex, tb2_contact* contact);
int tb2_explicit_initial_condition(tb2_explicit* ex);
/* nsteps x { Predictor + ConsistentKBC ; fint ; a = M^-1 R ; Corrector }.  h_fext_scale / h_value_scale
 * (each nsteps long or NULL = 1.0) scale the stored fext / prescribed values at each step (ScheduleT). */
int tb2_explicit_run(tb2_explicit* ex, double dt, int nsteps, const double* h_fext_scale, const double* h_value_scale);
/* one step through host arrays: H2D d,v,a ; step ; D2H d,v,a  (what a drop-in does when Tahoe's FieldT stays authoritative) */
int tb2_explicit_step_host(tb2_explicit* ex, double dt, double* h_d, double* h_v, double* h_a);
/* The resident drop-in's mode of operation (FEManagerT::SolveStep with the fields on the device, FEManagerT.cpp:451-686):
 * nsteps steps as tb2_explicit_run, then the displacement d -> h_d[nn][3] on a copy stream.  Returns WITHOUT waiting: the
 * copy overlaps the kernels of later calls (pinned h_d: tb2_host_register).  *ticket identifies the call; h_d is complete
 * (and element errors of the steps are reported) when tb2_explicit_wait(ex, ticket) returns.  At most two calls may be
 * outstanding per buffer pair: wait for ticket t before reusing the h_d of ticket t. */
int tb2_explicit_run_async(tb2_explicit* ex, double dt, int nsteps, const double* h_fext_scale, const double* h_value_scale,
                           double* h_d, int* ticket);
int tb2_explicit_wait(tb2_explicit* ex, int ticket);
/* device views for cooperating plugins / multi-GPU harness: which = 0 d, 1 v, 2 a, 3 mass, 4 fext, 5 fint */
double* tb2_explicit_device_array(tb2_explicit* ex, int which);

/* ---- equations + sparsity (NodeManagerT::SetEquationNumbers NodeManagerT.cpp:712-767; GraphT::MakeGraph GraphT.cpp:376-488;
 *      MSRBuilderT.cpp:134-181,216-244) ---------------------------------------------------------- */
int tb2_equations_create(tb2_mesh* mesh, const uint8_t* h_bc_code /*[nn][3], nonzero = prescribed*/, tb2_equations** eqs);
int tb2_equations_destroy(tb2_equations* eqs);
int tb2_equations_count(const tb2_equations* eqs, int64_t* num_eq);
int tb2_equations_get(const tb2_equations* eqs, int32_t* h_eqnos /*[nn][3]*/);
const int32_t* tb2_equations_device(const tb2_equations* eqs);

/* ---- global matrix (GlobalMatrixT.h:24-223; storage semantics of MSRMatrixT, solve = AztecMatrixT-style CG+Jacobi) */
int tb2_matrix_create(tb2_equations* eqs, tb2_matrix** A); /* GlobalMatrixT::Initialize + MSRBuilderT: CSR structure built on device */
/* the same matrix type from host CSR arrays (rowptr int64 [neq+1], colind sorted, 0-based): what a host-assembled
 * MSRMatrixT-derived plugin hands over (GenerateRCV / MSRBuilderT::SetSuperLUData form); values by tb2_matrix_set_values */
int tb2_matrix_create_csr(int device, int64_t num_eq, const int64_t* h_rowptr, const int32_t* h_colind, tb2_matrix** A);
int tb2_matrix_set_values(tb2_matrix* A, const double* h_val /*[nnz]*/);
int tb2_matrix_destroy(tb2_matrix* A);
int tb2_matrix_nnz(const tb2_matrix* A, int64_t* nnz);
/* MSRBuilderT::SetSuperLUData form: rowptr[neq+1] (int64), colind[nnz] sorted, diagonal in place */
int tb2_matrix_get_csr(const tb2_matrix* A, int64_t* h_rowptr, int32_t* h_colind, double* h_val /* may be NULL */);
/* MSRBuilderT::SetMSRData form (MSRMatrixT.h:21-23): returns length when h_bindx == NULL */
int tb2_matrix_get_msr(const tb2_matrix* A, int upper_only, int32_t* h_bindx, int64_t* length);
int tb2_matrix_clear(tb2_matrix* A); /* GlobalMatrixT::Clear */
/* SolidElementT::ElementLHSDriver (SolidElementT.cpp:1100-1154) + FormStiffness (SmallStrainT.cpp:285-324,
 * TotalLagrangianT.cpp:40-104, UpdatedLagrangianT.cpp:94-142) + MSRMatrixT::Assemble (MSRMatrixT.cpp:66-216):
 * A += K(u) without float atomics: element matrices of an L2-sized element chunk, then a per-(row node, column node) gather that
 * sums the contributions in ascending element order (the reference's serial assembly order) through the precomputed
 * element->CSR-slot map.  TB2_K3_COLOURED=1 selects the colour-by-colour read-modify-write form instead. */
int tb2_form_stiffness(tb2_group* group, tb2_matrix* A, const double* d_u, const double* d_u_last, int iteration);
int tb2_form_stiffness_host(tb2_group* group, tb2_matrix* A, const double* h_u, const double* h_u_last, int iteration);
/* ElementLHSDriver with formM (SolidElementT.cpp:1100-1154 -> ContinuumElementT::FormMass, ContinuumElementT.cpp:678-866):
 * A += constM * M, mass_type as TB2_MASS_*; with tb2_matrix_scale this forms the effective matrix of an implicit integrator,
 * constM * M + constK * K (eLinearHHTalpha::eComputeParameters: constM = 1, constK = (1 + alpha) beta dt^2) */
int tb2_form_mass(tb2_group* group, tb2_matrix* A, int mass_type, double constM);
int tb2_matrix_scale(tb2_matrix* A, double s); /* val *= s */
/* the same element loop assembled into a DiagonalMatrixT in kDiagOnly mode (DiagonalMatrixT.cpp:107-113: fMatrix[eq] += elMat(i,i)),
 * which is what <PCG_solver><diagonal_matrix/> uses as its preconditioner (SolverT.cpp:1097-1102): d_diag[nn][3] = diag K(u) per nodal dof */
int tb2_form_stiffness_diagonal(tb2_group* group, const double* d_u, const double* d_u_last, int iteration, double* d_diag);
int tb2_form_stiffness_diagonal_host(tb2_group* group, const double* h_u, const double* h_u_last, int iteration, double* h_diag);
/* MSRMatrixT::Multx (MSRMatrixT.cpp:385-420): y = A x on equation-space vectors [neq] */
int tb2_matrix_multx(tb2_matrix* A, const double* d_x, double* d_y);
int tb2_matrix_multx_host(tb2_matrix* A, const double* h_x, double* h_y);
/* GlobalMatrixT::CopyDiagonal */
int tb2_matrix_copy_diagonal(tb2_matrix* A, double* d_diag);
int tb2_matrix_copy_diagonal_host(tb2_matrix* A, double* h_diag); /* GlobalMatrixT::CopyDiagonal on host memory */
/* GlobalMatrixT::Solve -> BackSubstitute (GlobalMatrixT.cpp:77-113): Jacobi-preconditioned CG
 * (preconditioner = DiagonalMatrixT::Factorize semantics, DiagonalMatrixT.cpp:267-310).
 * d_x: start guess in, solution out.  Stops when |r| <= atol or |r| <= rtol*|r0| or max_iter. */
int tb2_matrix_pcg(tb2_matrix* A, const double* d_b, double* d_x, double rtol, double atol, int max_iter, int* iterations,
                   double* final_rnorm);
int tb2_matrix_pcg_host(tb2_matrix* A, const double* h_b, double* h_x, double rtol, double atol, int max_iter,
                        int* iterations, double* final_rnorm);
/* tb2_matrix_pcg returns TB2_OK when it stops at max_iter too (a caller may ask for a fixed number of iterations); whether the
 * last solve met rtol / atol, and its |r|/|r0|, are read here.  GlobalMatrixT::Solve reports failure by returning false
 * (GlobalMatrixT.cpp:77-113): the plugin's BackSubstitute throws when this says 0. */
int tb2_matrix_pcg_converged(const tb2_matrix* A, int* converged, double* relative_residual);
/* GlobalMatrixT::Solve for a NON-symmetric matrix (J2Simo3D::TangentType() = kNonSymmetric, J2Simo3D.cpp:18-21; the reference
 * uses LU there, SolverT.cpp:1108-1109): Jacobi-preconditioned BiCGStab on the device CSR, same arguments and stop test as
 * tb2_matrix_pcg; tb2_matrix_pcg_converged reports on it too.  One GPU. */
int tb2_matrix_bicgstab(tb2_matrix* A, const double* d_b, double* d_x, double rtol, double atol, int max_iter, int* iterations,
                        double* final_rnorm);
int tb2_matrix_bicgstab_host(tb2_matrix* A, const double* h_b, double* h_x, double rtol, double atol, int max_iter,
                             int* iterations, double* final_rnorm);
/* gather / scatter between [nn][3] nodal arrays and [neq] equation vectors (FieldT::AssembleUpdate FieldT.cpp:531-556,
 * SolverT::AssembleRHS SolverT.cpp:446-477): y_eq = x_node[active] ; x_node[active] += s * y_eq */
int tb2_equations_gather(const tb2_equations* eqs, const double* d_nodal, double* d_eqvec);
int tb2_equations_scatter_add(const tb2_equations* eqs, double scale, const double* d_eqvec, double* d_nodal);

/* ---- nonlinear PCG solver (SolverT subclass; the device twin of <PCG_solver><diagonal_matrix/></PCG_solver>) ------------
 * PCGSolver_LS (solvers/PCGSolver_LS.cpp:107-371: CGSearch with Bertsekas' scaled beta, restart every `restart` iterations,
 * secant line search of <= line_search_iterations residual evaluations, step bound max_step) inside NLSolver::Solve /
 * ExitIteration (solvers/NLSolver.cpp:57-263, 675-756); preconditioner = diag K(u) re-formed at every restart
 * (DiagonalMatrixT kDiagOnly, DiagonalMatrixT.cpp:107-113, 267-323).  Parameter names are the XML attributes. */
typedef enum { TB2_SOLVER_CONTINUE = 0, TB2_SOLVER_CONVERGED = 1, TB2_SOLVER_FAILED = 2 } tb2_solver_status; /* SolverT::SolutionStatusT, SolverT.h:59-62 */
typedef struct {
    int32_t restart;                 /* PCGSolver_LS.cpp:53-56 */
    int32_t line_search_iterations;  /* :66-69, default 3; 0 = full steps (NLSolver::Update) */
    double  line_search_tolerance;   /* :71-74, default 0.25 */
    double  max_step;                /* :76-79, default 2.5 */
    double  abs_tolerance;           /* NLSolver fZeroTolerance */
    double  rel_tolerance;           /* NLSolver fTolerance */
    double  divergence_tolerance;    /* NLSolver fDivTolerance */
    int32_t max_iterations;          /* NLSolver fMaxIterations */
    int32_t min_iterations;          /* NLSolver fMinIterations */
} tb2_nlpcg_params;
typedef struct tb2_nlpcg tb2_nlpcg;
int tb2_nlpcg_create(tb2_group* group, tb2_equations* eqs, const tb2_nlpcg_params* params, tb2_nlpcg** solver);
int tb2_nlpcg_destroy(tb2_nlpcg* solver);
/* The solver's line search alone, on a host callback G(s) = R(u + s dir) . dir (PCGSolver_LS::Update's secant search with its
 * bracketing, clamping at max_step, trial budget and best-step fallback, PCGSolver_LS.cpp:213-348): for host programs that own
 * the residual evaluation, and for checking the decision logic without a device.  final_step is the step the callback saw last. */
int tb2_secant_search_host(double (*slope)(double step, void* user), void* user, double slope_at_zero, double max_step,
                           double abs_tolerance, double rel_tolerance, int max_trials, double* final_step, int* evaluations);
/* SolverT::Solve(max_iterations) for one load step (SolverT::InitStep state: iteration number -1).  d_u[nn][3]: displacement
 * with the prescribed dofs already set, updated in place; d_u_last: last converged displacement (J2 only, else NULL);
 * d_fext[nn][3]: nodal forces of this step (FieldT::FormRHS).  solve_max_iterations = -1: no limit beyond params.
 * *status = tb2_solver_status; *iterations = SolverT::IterationNumber() at exit; *error = |R|, *error0 = |R| of the first pass.
 * An element failure (TB2_ERR_BAD_JACOBIAN / TB2_ERR_J2_LOCAL) is returned with *status = TB2_SOLVER_FAILED, as NLSolver::Solve
 * turns the exception into kFailed (NLSolver.cpp:247-262). */
int tb2_nlpcg_solve(tb2_nlpcg* solver, double* d_u, const double* d_u_last, const double* d_fext, int solve_max_iterations,
                    int* status, int* iterations, double* error, double* error0);
int tb2_nlpcg_solve_host(tb2_nlpcg* solver, double* h_u, const double* h_u_last, const double* h_fext, int solve_max_iterations,
                         int* status, int* iterations, double* error, double* error0);
/* residual sweeps (K1) and preconditioner sweeps (K3 diagonal) launched so far */
int tb2_nlpcg_counters(const tb2_nlpcg* solver, int64_t* residual_sweeps, int64_t* preconditioner_sweeps);

/* ---- Newton driver (<nonlinear_solver> with <CUDA_PCG_matrix/>) ----------------------------------------------------------
 * NLSolver::Solve / Iterate / ExitIteration (solvers/NLSolver.cpp:57-263, 759-766, 675-756) for one load step, resident:
 * residual (K1) -> ExitIteration -> every reform_tangent_iterations: GlobalMatrixT::Clear + FormLHS (K3) -> GlobalMatrixT::Solve
 * (Jacobi-PCG, K6-K8, zero start guess) -> FieldT::AssembleUpdate.  Work vectors are those of a tb2_nlpcg object made for the
 * same group and equations (its params are not used); A must have been created from the same tb2_equations. */
typedef struct {
    double  abs_tolerance, rel_tolerance, divergence_tolerance; /* NLSolver fZeroTolerance, fTolerance, fDivTolerance */
    int32_t max_iterations, min_iterations;                     /* fMaxIterations, fMinIterations */
    int32_t reform_tangent_iterations;                          /* fReformTangentIterations (default 1) */
    double  pcg_rel_tolerance, pcg_abs_tolerance;               /* <CUDA_PCG_matrix rel_tolerance abs_tolerance max_iterations/> */
    int32_t pcg_max_iterations;
} tb2_newton_params;
int tb2_newton_solve(tb2_nlpcg* work, tb2_matrix* A, const tb2_newton_params* params, double* d_u, const double* d_u_last,
                     const double* d_fext, int solve_max_iterations, int* status, int* iterations, double* error, double* error0,
                     int64_t* linear_iterations /* PCG iterations summed over the Newton iterations */);
int tb2_newton_solve_host(tb2_nlpcg* work, tb2_matrix* A, const tb2_newton_params* params, double* h_u, const double* h_u_last,
                          const double* h_fext, int solve_max_iterations, int* status, int* iterations, double* error, double* error0,
                          int64_t* linear_iterations);

/* The same driver for an implicit time integrator (FEManagerT::SolveStep with nonlinear_HHT, IntegratorT_factory.cpp:45-47): the
 * unknown is the acceleration increment.  Residual = fext - constKd fint(u) - constMa M a (eNLHHTalpha.cpp:17-36), effective matrix
 * constM M + constK K(u) (eLinearHHTalpha::eComputeParameters: constM = 1, constK = (1 + alpha) beta dt^2), update
 * u += dcorr_a da, v += vcorr_a da, a += da on the active equations (nNLHHTalpha::Corrector, nNLHHTalpha.cpp:131-160,219-229:
 * dcorr_a = beta dt^2, vcorr_a = gamma dt).  The caller applies the predictor and the kinematic BCs before the call.  With dt = 0
 * (constK = dcorr_a = vcorr_a = 0) this is the initial-acceleration solve of FEManagerT::InitialCondition (FEManagerT.cpp:2053-2080). */
typedef struct {
    int32_t mass_type; /* TB2_MASS_CONSISTENT | TB2_MASS_LUMPED */
    double  constM, constK, constMa, constKd, dcorr_a, vcorr_a;
} tb2_dynamics;
int tb2_newton_solve_dynamic(tb2_nlpcg* work, tb2_matrix* A, const tb2_newton_params* params, const tb2_dynamics* dynamics, double* d_u,
                             double* d_v, double* d_a, const double* d_u_last, const double* d_fext, int solve_max_iterations,
                             int* status, int* iterations, double* error, double* error0, int64_t* linear_iterations);
int tb2_newton_solve_dynamic_host(tb2_nlpcg* work, tb2_matrix* A, const tb2_newton_params* params, const tb2_dynamics* dynamics,
                                  double* h_u, double* h_v, double* h_a, const double* h_u_last, const double* h_fext,
                                  int solve_max_iterations, int* status, int* iterations, double* error, double* error0,
                                  int64_t* linear_iterations);

/* ---- multi-GPU (replaces CommManagerT::AllGather / CommunicatorT::Sum) ------------ */
/* ---- element partition for the one-process-per-GPU path (host code, no CUDA calls) -- DecomposeT / GraphBaseT::Partition's role
 * (DecomposeT.cpp:362-592, GraphBaseT.cpp:185-270).  tb2_partition_rcb: owner rank of every element by recursive coordinate
 * bisection of the element centroids (longest extent, weighted median: any rank count; deterministic).  tb2_partition_part: one
 * rank's part from ANY element -> rank map (this one or a graph partitioner's): call it with the output arrays NULL for the sizes,
 * then with arrays of those sizes.  Local nodes are numbered by ascending global id, elements keep their global order;
 * h_if_nodes / h_if_slots / h_node_owned are what tb2_comm_init takes. */
int tb2_partition_rcb(int64_t num_nodes, int64_t num_elements, const int32_t* h_conn /*[ne][8]*/, const double* h_coords /*[nn][3]*/,
                      int nparts, int32_t* h_owner /*[ne]*/);
int tb2_partition_part(int64_t num_nodes, int64_t num_elements, const int32_t* h_conn, const int32_t* h_owner, int nparts, int rank,
                       int64_t* num_local_nodes, int64_t* num_local_elements, int64_t* num_interface_nodes,
                       int64_t* num_global_interface_nodes, int64_t* h_node_gid /*[local nodes]*/, int64_t* h_elem_gid /*[local elements]*/,
                       int32_t* h_local_conn /*[local elements][8]*/, int32_t* h_if_nodes, int32_t* h_if_slots, uint8_t* h_node_owned);

/* One process per GPU.  The harness creates an NCCL unique id on rank 0, distributes it by its own
 * means, and every rank calls tb2_comm_init on its mesh.  h_interface_nodes are this rank's local node
 * ids of nodes shared with other ranks, h_interface_slots their positions in the packed global interface
 * vector (identical on every sharer). */
int tb2_comm_unique_id(char h_id[128]);
int tb2_comm_init(tb2_mesh* mesh, int rank, int nranks, const char h_id[128], int64_t num_interface_nodes,
                  const int32_t* h_interface_nodes, const int32_t* h_interface_slots, int64_t num_global_interface_nodes,
                  const uint8_t* h_node_owned /*[nn]: 1 if this rank owns the node (dot products count it once)*/);
int tb2_comm_destroy(tb2_mesh* mesh);
/* Interface exchange over NVLink peer memory, in place of the reference's point-to-point ghost-node messages and
 * MPI_Allreduce (CommManagerT.cpp:424-436, SolverT.cpp:854-860): every rank exports the 64-byte IPC handle of its exchange
 * window, the host program all-gathers the handles by its own means (MPI_Allgather where CommManagerT lives) and every rank
 * imports the [nranks][64] table.  From then on the explicit step and the distributed PCG publish their partial interface
 * values into their own window and pull the sharers' values inside the consuming kernel -- no collective library on the data
 * path; without the import they run the packed ncclAllReduce.  One process per GPU, all GPUs in one NVLink domain (<= 16).
 * Like a communicator, the windows are torn down collectively: the host program synchronises the ranks (a barrier after the
 * last run has been waited for) before any rank calls tb2_comm_destroy / tb2_mesh_destroy.  A peer that never arrives at an
 * exchange ends the waiting kernel after 300 s with a device trap (every later call on that rank fails). */
int tb2_comm_peer_export(tb2_mesh* mesh, char h_handle[64]);
int tb2_comm_peer_import(tb2_mesh* mesh, const char* h_handles /* [nranks][64], rank order */);
int tb2_comm_peer_enabled(tb2_mesh* mesh); /* 1 after a successful import */
/* back to the packed ncclAllReduce; collective like the import: when the import failed on ANY rank (no IPC between the processes,
 * GPUs outside one NVLink domain) the host program calls this on every rank */
int tb2_comm_peer_disable(tb2_mesh* mesh);
/* d_nodal[nn][3] += contributions of the other sharers on interface nodes (ncclAllReduce on the packed vector) */
int tb2_comm_sum_interface(tb2_mesh* mesh, double* d_nodal);

#ifdef __cplusplus
}
#endif
#endif
