/* CudaSolidElementT.cpp -- see CudaSolidElementT.h */
#include "CudaSolidElementT.h"

#include "CudaPCGMatrixT.h"
#include "CudaPenaltyContact3DT.h"
#include "ElementSupportT.h"
#include "ExceptionT.h"
#include "FDKStV.h"
#include "FEManagerT.h"
#include "FieldT.h"
#include "IsotropicT.h"
#include "J2Simo3D.h"
#include "MaterialListT.h"
#include "OutputSetT.h"
#include "ParameterListT.h"
#include "SSKStV.h"
#include "ScheduleT.h"
#include "SimoIso3D.h"
#include "SolidMaterialT.h"
#include "eIntegratorT.h"
#include "iArray2DT.h"

#include <cstring>
#include <vector>

using namespace Tahoe;

namespace Tahoe {
const char* kCudaSolidElementNames[kNumCudaSolidElementNames] = {"cuda_small_strain", "cuda_total_lagrangian", "cuda_updated_lagrangian",
	"cuda_explicit_solid", "cuda_contact_3D_penalty"};

ElementBaseT* NewCudaSolidElement(const StringT& name, const ElementSupportT& support)
{
	if (name == kCudaSolidElementNames[0]) return new CudaSmallStrainT(support, kCudaSolidElementNames[0], TB2_SMALL_STRAIN);
	if (name == kCudaSolidElementNames[1]) return new CudaTotalLagrangianT(support, kCudaSolidElementNames[1], TB2_TOTAL_LAGRANGIAN);
	if (name == kCudaSolidElementNames[2]) return new CudaUpdatedLagrangianT(support, kCudaSolidElementNames[2], TB2_UPDATED_LAGRANGIAN);
	if (name == kCudaSolidElementNames[3]) return new CudaExplicitSolidT(support, kCudaSolidElementNames[3], TB2_UPDATED_LAGRANGIAN);
	if (name == kCudaSolidElementNames[4]) return new CudaPenaltyContact3DT(support, kCudaSolidElementNames[4]);
	return NULL;
}
} // namespace Tahoe

/* depth-first search of the validated parameter tree for a sub-list by name (J2 hardening function) */
static const ParameterListT* FindList(const ParameterListT& list, const char* name)
{
	if (list.Name() == name) return &list;
	const ArrayT<ParameterListT>& subs = list.Lists();
	for (int i = 0; i < subs.Length(); i++) {
		const ParameterListT* hit = FindList(subs[i], name);
		if (hit) return hit;
	}
	return NULL;
}

template <class BaseT>
CudaSolidElementT<BaseT>::CudaSolidElementT(const ElementSupportT& support, const char* name, int formulation):
	BaseT(support),
	fFormulation(formulation),
	fMesh(NULL),
	fGroup(NULL),
	fEqs(NULL),
	fMatrix(NULL),
	fIsJ2(false),
	fMuted(false),
	fMaterialKind(-1),
	fDevice(0)
{
	this->SetName(name);
}

template <class BaseT>
CudaSolidElementT<BaseT>::~CudaSolidElementT(void)
{
	if (fMatrix) tb2_matrix_destroy(fMatrix);
	if (fEqs) tb2_equations_destroy(fEqs);
	for (size_t i = 0; i < fGroups.size(); i++)
		if (fGroups[i]) tb2_group_destroy(fGroups[i]);
	if (fMesh) tb2_mesh_destroy(fMesh);
}

/* status code -> Tahoe exception (include/tahoe_b200.h: tb2_status) */
template <class BaseT>
void CudaSolidElementT<BaseT>::Check(int status, const char* caller) const
{
	switch (status) {
	case TB2_OK: return;
	case TB2_ERR_BAD_JACOBIAN: ExceptionT::BadJacobianDet(caller, "%s", tb2_last_error());
	case TB2_ERR_ARG: ExceptionT::BadInputValue(caller, "%s", tb2_last_error());
	case TB2_ERR_SIZE: ExceptionT::SizeMismatch(caller, "%s", tb2_last_error());
	case TB2_ERR_COMM: ExceptionT::MPIFail(caller, "%s", tb2_last_error());
	default: ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
	}
}

/* everything the wrapped element class defines, plus the device the group lives on */
template <class BaseT>
void CudaSolidElementT<BaseT>::DefineParameters(ParameterListT& list) const
{
	BaseT::DefineParameters(list);
	ParameterT device(ParameterT::Integer, "device");
	device.SetDefault(0);
	device.AddLimit(0, LimitT::LowerInclusive);
	list.AddParameter(device, ParameterListT::ZeroOrOnce);
}

template <class BaseT>
void CudaSolidElementT<BaseT>::TakeParameterList(const ParameterListT& list)
{
	const char caller[] = "CudaSolidElementT::TakeParameterList";

	/* inherited: connectivity, shape functions, materials, output -- all Tahoe's own */
	BaseT::TakeParameterList(list);
	const ParameterT* device = list.Parameter("device");
	fDevice = device ? int(*device) : 0;

	/* the device path covers exactly the reference's Hex8 / 8-point / standard-B case */
	if (this->GeometryCode() != GeometryT::kHexahedron || this->NumElementNodes() != 8 || this->NumIP() != 8 || this->NumSD() != 3)
		ExceptionT::BadInputValue(caller, "the CUDA element group supports 8-node hexahedra with 8 integration points only");
	const ParameterT* b_opt = list.Parameter("strain_displacement"); /* SmallStrainT only (SmallStrainT.cpp:38-42): 0 = standard, 1 = B-bar */
	if (b_opt && int(*b_opt) != 0) {
		if (fFormulation != TB2_SMALL_STRAIN || int(*b_opt) != 1)
			ExceptionT::BadInputValue(caller, "strain_displacement must be \"standard\" or \"B-bar\"");
		fFormulation = TB2_SMALL_STRAIN_BBAR; /* SmallStrainT::kMeanDilBbar */
	}
	fIsJ2 = false;

	/* one device group per material of the list (= per element block, SmallStrainT::CollectMaterialInfo :161-186); the groups share the
	 * device mesh and each sees the elements of the other materials as switched off */
	std::vector<tb2_material> mats;
	const int num_materials = this->fMaterialList->Length();
	tb2_material mat;
	memset(&mat, 0, sizeof(mat));
	if (this->Name() == kCudaSolidElementNames[3]) {
		if (num_materials != 1) ExceptionT::BadInputValue(caller, "explicit_solid takes one material (ExplicitElementT scans a single mu / kappa)");
		/* <explicit_solid>: mu / kappa / density are scanned from the material sub-tree and <j2_plasticity> selects the law, exactly
		 * as ExplicitElementT::TakeParameterList does (ExplicitElementT.cpp:166-213) */
		struct Finder {
			static void Find(const ParameterListT& p, double& mu, double& kappa, double& density, bool& has_mu, bool& has_kappa) {
				const ParameterT* pm = p.Parameter("mu");
				const ParameterT* pk = p.Parameter("kappa");
				const ParameterT* pd = p.Parameter("density");
				if (pm) { mu = *pm; has_mu = true; }
				if (pk) { kappa = *pk; has_kappa = true; }
				if (pd) density = *pd;
				const ArrayT<ParameterListT>& subs = p.Lists();
				for (int i = 0; i < subs.Length(); i++) Find(subs[i], mu, kappa, density, has_mu, has_kappa);
			}
		};
		double mu = 0.0, kappa = 0.0, density = 1.0;
		bool has_mu = false, has_kappa = false;
		const ParameterListT& block = list.GetList("large_strain_element_block");
		Finder::Find(block.GetListChoice(*this, "large_strain_material_choice"), mu, kappa, density, has_mu, has_kappa);
		if (!has_mu || !has_kappa) ExceptionT::BadInputValue(caller, "explicit_solid needs mu and kappa in its material sub-tree");
		mat.kind = TB2_EXPL_NEO_HOOKEAN;
		mat.mu = mu;
		mat.kappa = kappa;
		mat.lambda = kappa - 2.0 * mu / 3.0;
		mat.density = density;
		if (list.NumLists("j2_plasticity") > 0) {
			const ParameterListT& j2 = list.GetList("j2_plasticity");
			mat.kind = TB2_EXPL_J2;
			mat.hard[0] = j2.GetParameter("sigma_Y");
			mat.hard[1] = j2.GetParameter("hardening");
		}
		/* <mass_scaling type="fixed" | "adaptive">: the scale factors and the scaled lumped mass stay ExplicitElementT's own host code
		 * (ApplyMassScaling / LHSDriver, ExplicitElementT.cpp:492-618).  The adaptive type re-evaluates the factors every
		 * update_interval calls from the INITIAL coordinates (:499-560) and the explicit mass is formed once, so in this reference it
		 * is the fixed type; nothing on the device depends on it. */
		if (list.NumLists("anp_tet4") > 0) ExceptionT::BadInputValue(caller, "anp_tet4 is a Tet4 option; the CUDA group is Hex8 only");
		mats.push_back(mat);
	} else
	for (int im = 0; im < num_materials; im++) {
		memset(&mat, 0, sizeof(mat));
		const char* block_name = (fFormulation == TB2_SMALL_STRAIN || fFormulation == TB2_SMALL_STRAIN_BBAR) ? "small_strain_element_block" : "large_strain_element_block";
		const ParameterListT& scope = (list.NumLists(block_name) == num_materials) ? list.GetList(block_name, im) : list;
		/* material constants */
		ContinuumMaterialT* cmat = (*(this->fMaterialList))[im];
		if (dynamic_cast<SSKStV*>(cmat)) mat.kind = TB2_SSKSTV;
		else if (dynamic_cast<FDKStV*>(cmat)) mat.kind = TB2_FDKSTV;
		else if (dynamic_cast<J2Simo3D*>(cmat)) mat.kind = TB2_J2_SIMO; /* before its base SimoIso3D */
		else if (dynamic_cast<SimoIso3D*>(cmat)) mat.kind = TB2_SIMO_ISO;
		else ExceptionT::BadInputValue(caller, "material \"%s\" has no device implementation", cmat->Name().Pointer());
		const IsotropicT* iso = dynamic_cast<const IsotropicT*>(cmat);
		const SolidMaterialT* smat = dynamic_cast<const SolidMaterialT*>(cmat);
		if (!iso || !smat) ExceptionT::GeneralFail(caller, "material is not isotropic");
		mat.mu = iso->Mu();
		mat.lambda = iso->Lambda();
		mat.kappa = iso->Kappa();
		mat.density = const_cast<SolidMaterialT*>(smat)->Density();
		fIsJ2 = fIsJ2 || (mat.kind == TB2_J2_SIMO);
		if (mat.kind == TB2_J2_SIMO) { /* K(alpha): C1functions/LinearT.h:71, LinearExponentialT.cpp:48-57 */
			const ParameterListT* lin = FindList(scope, "linear_function");
			const ParameterListT* lexp = FindList(scope, "linear_exponential");
			const ParameterListT* plaw = FindList(scope, "power_law");
			const ParameterListT* spline = FindList(scope, "cubic_spline");
			if (lin) {
				mat.hard_kind = TB2_HARD_LINEAR;
				mat.hard[0] = lin->GetParameter("a");
				mat.hard[1] = lin->GetParameter("b");
			} else if (lexp) {
				mat.hard_kind = TB2_HARD_LINEAR_EXP;
				mat.hard[0] = lexp->GetParameter("a");
				mat.hard[1] = lexp->GetParameter("b");
				mat.hard[2] = lexp->GetParameter("c");
				mat.hard[3] = lexp->GetParameter("d");
			} else if (plaw) { /* PowerLawT.cpp:28-37 */
				mat.hard_kind = TB2_HARD_POWER_LAW;
				mat.hard[0] = plaw->GetParameter("a");
				mat.hard[1] = plaw->GetParameter("b");
				mat.hard[2] = plaw->GetParameter("c");
				mat.hard[3] = plaw->GetParameter("n");
			} else if (spline) { /* CubicSplineT::TakeParameterList (CubicSplineT.cpp:352-381): the library forms the coefficients */
				mat.hard_kind = TB2_HARD_CUBIC_SPLINE;
				mat.num_knots = spline->NumLists("OrderedPair");
				if (mat.num_knots > TB2_MAX_KNOTS) ExceptionT::BadInputValue(caller, "cubic_spline hardening: at most %d knots", TB2_MAX_KNOTS);
				int fixity = spline->GetParameter("fixity");
				mat.spline_fixity = fixity == 1 ? TB2_SPLINE_FREE_RUN : TB2_SPLINE_PARABOLIC;
				for (int i = 0; i < mat.num_knots; i++) {
					const ParameterListT* knot = spline->List("OrderedPair", i);
					mat.knot_x[i] = knot->GetParameter("x");
					mat.knot_y[i] = knot->GetParameter("y");
				}
			} else
				ExceptionT::BadInputValue(caller, "Simo_J2 hardening must be linear_function, linear_exponential, power_law or cubic_spline");
		}
		mats.push_back(mat);
	}

	/* connectivity of all blocks, in block order then file order (ElementBaseT.cpp:607-632) */
	std::vector<int32_t> conn;
	for (int b = 0; b < this->fConnectivities.Length(); b++) {
		const iArray2DT& c = *(this->fConnectivities[b]);
		conn.insert(conn.end(), c.Pointer(), c.Pointer() + c.Length());
	}
	const dArray2DT& X = this->ElementSupport().InitialCoordinates();
	Check(tb2_mesh_create(fDevice, X.MajorDim(), (int64_t)(conn.size() / 8), &conn[0], X.Pointer(), &fMesh), caller);
	fGroups.resize(mats.size(), NULL);
	fKinds.resize(mats.size());
	for (size_t im = 0; im < mats.size(); im++) {
		Check(tb2_group_create(fMesh, fFormulation, &mats[im], &fGroups[im]), caller);
		fKinds[im] = mats[im].kind;
	}
	fGroup = fGroups[0];
	fMaterialKind = mats.size() == 1 ? mats[0].kind : -1; /* device stress output: single-material groups */
	fFint.Dimension(X.MajorDim(), 3);
	if (fGroups.size() > 1) {
		fPart.Dimension(X.MajorDim(), 3);
		ArrayT<ElementCardT::StatusT> status;
		this->GetStatus(status);
		SetStatus(status); /* installs the per-material element masks */
	}
}

template <class BaseT>
void CudaSolidElementT<BaseT>::RHSDriver(void)
{
	const char caller[] = "CudaSolidElementT::RHSDriver";

	/* tractions: O(surface) host work, inherited (ContinuumElementT.cpp:505-665) */
	ContinuumElementT::RHSDriver();

	/* components dictated by the integrator (SolidElementT.cpp:1197-1215) */
	double constMa = 0.0, constKd = 0.0;
	int formMa = this->fIntegrator->FormMa(constMa);
	int formKd = this->fIntegrator->FormKd(constKd);
	if (this->fMassType == ContinuumElementT::kNoMass) formMa = 0;
	/* body force (SolidElementT.cpp:1204-1211): formed through the mass operator, also when the integrator itself needs no M a */
	int formBody = 0;
	if (this->fMassType != ContinuumElementT::kNoMass && this->fBodySchedule && this->fBody.Magnitude() > kSmall) {
		formBody = 1;
		if (!formMa) constMa = 1.0;
	}
	if (fMuted) { /* a resident solver forms the internal force and the inertia itself: tractions (above) and the body force remain */
		formKd = 0;
		formMa = 0;
		if (formBody) constMa = 1.0;
	}
	if (!formKd && !formMa && !formBody) return;

	const FieldT& field = this->Field();
	const dArray2DT& disp = field[0];
	fFint = 0.0;
	if (formKd) {
		const double* last = fIsJ2 ? field(-1, 0).Pointer() : NULL;
		int iteration = this->ElementSupport().IterationNumber(this->Group());
		Check(tb2_form_internal_force_host(fGroups[0], disp.Pointer(), fKinds[0] == TB2_J2_SIMO ? last : NULL, iteration, fFint.Pointer()), caller);
		for (size_t i = 1; i < fGroups.size(); i++) { /* further materials: elements of the others are masked off, the parts add up */
			Check(tb2_form_internal_force_host(fGroups[i], disp.Pointer(), fKinds[i] == TB2_J2_SIMO ? last : NULL, iteration, fPart.Pointer()), caller);
			fFint += fPart;
		}
		fFint *= -constKd;
	}

	/* inertia term of an implicit integrator (SolidElementT.cpp:1243-1265): -constMa M a with the group's mass type */
	if (formMa || formBody) {
		if (this->fMassType != ContinuumElementT::kConsistentMass && this->fMassType != ContinuumElementT::kLumpedMass)
			ExceptionT::GeneralFail(caller, "unresolved mass type %d", int(this->fMassType));
		if (fMa.MajorDim() != disp.MajorDim()) fMa.Dimension(disp.MajorDim(), 3);
		const double* acc = field[2].Pointer();
		if (formBody) {
			/* ContinuumElementT::AddBodyForce (ContinuumElementT.cpp:849-865) sets every nodal value to -b * schedule (it does not
			 * add to the acceleration), so the element loop integrates the mass operator against that constant field */
			if (fBodyAcc.MajorDim() != disp.MajorDim()) fBodyAcc.Dimension(disp.MajorDim(), 3);
			const double loadfactor = this->fBodySchedule->Value();
			for (int i = 0; i < 3; i++) fBodyAcc.SetColumn(i, -this->fBody[i] * loadfactor);
			acc = fBodyAcc.Pointer();
		}
		for (size_t i = 0; i < fGroups.size(); i++) { /* each material with its own density */
			Check(tb2_form_inertial_force_host(fGroups[i], int(this->fMassType), -constMa, acc, fMa.Pointer()), caller);
			fFint += fMa;
		}
	}

	/* RHS += -(constKd fint + constMa M a) on the active equations: one call for the whole group (SolverT::AssembleRHS, SolverT.cpp:446-477) */
	this->ElementSupport().AssembleRHS(this->Group(), fFint, field.Equations());
}

template <class BaseT>
void CudaSolidElementT<BaseT>::LHSDriver(GlobalT::SystemTypeT sys_type)
{
	const char caller[] = "CudaSolidElementT::LHSDriver";

	/* mass matrix requested (explicit: once) or a host matrix type: Tahoe's own assembly */
	double constM = 0.0, constK = 0.0;
	int formM = this->fIntegrator->FormM(constM);
	int formK = this->fIntegrator->FormK(constK);
	const GlobalMatrixT& lhs = this->ElementSupport().FEManager().LHS(this->Group());
	CudaPCGMatrixT* cuda_lhs = const_cast<CudaPCGMatrixT*>(dynamic_cast<const CudaPCGMatrixT*>(&lhs));
	if (this->fMassType == ContinuumElementT::kNoMass) formM = 0;
	const bool haveK = formK && fabs(constK) > kSmall, haveM = formM && fabs(constM) > kSmall;
	if (!cuda_lhs || (!haveK && !haveM)) {
		BaseT::LHSDriver(sys_type);
		return;
	}

	/* tractions etc. */
	ContinuumElementT::LHSDriver(sys_type);

	/* device assembly into the cooperating matrix: K3 straight into its CSR, no element matrices cross the bus */
	const FieldT& field = this->Field();
	DeviceEquations();
	if (!fMatrix) Check(tb2_matrix_create(fEqs, &fMatrix), caller);
	const double* last = fIsJ2 ? field(-1, 0).Pointer() : NULL;
	int iteration = this->ElementSupport().IterationNumber(this->Group());
	Check(tb2_matrix_clear(fMatrix), caller);
	if (haveK) {
		for (size_t i = 0; i < fGroups.size(); i++)
			Check(tb2_form_stiffness_host(fGroups[i], fMatrix, field[0].Pointer(), fKinds[i] == TB2_J2_SIMO ? last : NULL, iteration), caller);
		if (fabs(constK - 1.0) > 0.0) Check(tb2_matrix_scale(fMatrix, constK), caller); /* eLinearHHTalpha::FormK: (1 + alpha) beta dt^2 */
	}
	if (haveM) /* effective mass of an implicit integrator: constM M (ContinuumElementT::FormMass) */
		for (size_t i = 0; i < fGroups.size(); i++)
			Check(tb2_form_mass(fGroups[i], fMatrix, int(this->fMassType), constM), caller);
	cuda_lhs->AddDeviceMatrix(fMatrix, 1.0);
}

template <class BaseT>
void CudaSolidElementT<BaseT>::ComputeOutput(const iArrayT& n_codes, dArray2DT& n_values, const iArrayT& e_codes, dArray2DT& e_values)
{
	const char caller[] = "CudaSolidElementT::ComputeOutput";
	const int n_out = n_codes.Sum();
	const bool device_material = fMaterialKind == TB2_SSKSTV || fMaterialKind == TB2_FDKSTV || fMaterialKind == TB2_SIMO_ISO ||
		fMaterialKind == TB2_J2_SIMO;
	const bool device_codes = e_codes.Sum() == 0 && n_out > 0 && n_codes[SolidElementT::iNodalStress] == 6 &&
		n_out == n_codes[SolidElementT::iNodalDisp] + n_codes[SolidElementT::iNodalStress] &&
		(n_codes[SolidElementT::iNodalDisp] == 0 || n_codes[SolidElementT::iNodalDisp] == 3);
	if (!device_material || !device_codes || this->qUseSimo || this->qNoExtrap) {
		/* host output of a Simo_J2 group evaluates J2Simo3D at the integration points: it needs the converged history */
		if (fIsJ2) HistoryToCards();
		BaseT::ComputeOutput(n_codes, n_values, e_codes, e_values);
		return;
	}
	/* SolidElementT::ComputeOutput (SolidElementT.cpp:1352-1840) for [displacements | extrapolated stresses]: IP Cauchy stress,
	 * HexahedronT::SetExtrapolation, GroupAverageT averaging -- one device call for the whole group */
	const dArray2DT& disp = this->Field()[0];
	dArray2DT stress(disp.MajorDim(), 6);
	/* J2Simo3D::s_ij reads the element history, F of the last converged step and the iteration number (J2Simo3D.cpp:64-105) */
	const double* last = fMaterialKind == TB2_J2_SIMO ? this->Field()(-1, 0).Pointer() : NULL;
	Check(tb2_group_nodal_stress_at_host(fGroup, disp.Pointer(), last, this->ElementSupport().IterationNumber(this->Group()), stress.Pointer()), caller);
	static bool announced = false;
	if (!announced) {
		cout << "\n " << caller << ": nodal stresses extrapolated and averaged on the device" << endl;
		announced = true;
	}
	const iArrayT& nodes_used = this->ElementSupport().OutputSet(this->fOutputID).NodesUsed();
	const int ndisp = n_codes[SolidElementT::iNodalDisp];
	n_values.Dimension(nodes_used.Length(), n_out);
	for (int r = 0; r < nodes_used.Length(); r++) {
		double* row = n_values(r);
		const int node = nodes_used[r];
		for (int i = 0; i < ndisp; i++) row[i] = disp(node, i);
		for (int I = 0; I < 6; I++) row[ndisp + I] = stress(node, I);
	}
	e_values.Dimension(this->NumElements(), 0);
}

template <class BaseT>
tb2_equations* CudaSolidElementT<BaseT>::DeviceEquations(void)
{
	const char caller[] = "CudaSolidElementT::DeviceEquations";
	if (!fEqs) { /* equation numbers exist now: prescribed dofs have eqnos <= 0 (FieldT.cpp:635-659) */
		const iArray2DT& eq = this->Field().Equations();
		std::vector<uint8_t> bc(eq.Length());
		for (int i = 0; i < eq.Length(); i++) bc[i] = eq[i] > 0 ? 0 : 1;
		Check(tb2_equations_create(fMesh, &bc[0], &fEqs), caller);
		/* the device numbers equations as NodeManagerT::SetEquationNumbers does (node-major, no renumbering): verify once */
		std::vector<int32_t> dev_eq(eq.Length());
		Check(tb2_equations_get(fEqs, &dev_eq[0]), caller);
		for (int i = 0; i < eq.Length(); i++)
			if ((eq[i] > 0 || dev_eq[i] > 0) && eq[i] != dev_eq[i])
				ExceptionT::GeneralFail(caller, "equation numbering differs from the device's at dof %d (renumbering matrix type?)", i);
	}
	return fEqs;
}

template <class BaseT>
void CudaSolidElementT<BaseT>::SetStatus(const ArrayT<ElementCardT::StatusT>& status)
{
	BaseT::SetStatus(status);
	if (fGroups.empty()) return;
	std::vector<uint8_t> off(status.Length());
	for (size_t g = 0; g < fGroups.size(); g++) {
		bool any = false;
		for (int i = 0; i < status.Length(); i++) {
			/* off for this device group: switched off by the user, or an element of another material */
			off[i] = (status[i] == ElementCardT::kOFF || (fGroups.size() > 1 && this->fElementCards[i].MaterialNumber() != int(g))) ? 1 : 0;
			any = any || off[i];
		}
		Check(tb2_group_set_element_status(fGroups[g], any ? &off[0] : NULL), "CudaSolidElementT::SetStatus");
	}
}

template <class BaseT>
void CudaSolidElementT<BaseT>::CloseStep(void)
{
	BaseT::CloseStep();
	for (size_t i = 0; i < fGroups.size(); i++) Check(tb2_group_close_step(fGroups[i]), "CudaSolidElementT::CloseStep");
}

template <class BaseT>
GlobalT::RelaxCodeT CudaSolidElementT<BaseT>::ResetStep(void)
{
	GlobalT::RelaxCodeT relax = BaseT::ResetStep();
	for (size_t i = 0; i < fGroups.size(); i++) Check(tb2_group_reset_step(fGroups[i]), "CudaSolidElementT::ResetStep");
	return relax;
}

/* J2SimoC0HardeningT keeps 8 flags and 5*48 + 64 doubles per element, allocated at the first plastic step (AllocateElement
 * :312-333); tb2_group_get_history returns exactly that layout, so the cards written here are what the classic element would hold */
static const int kJ2CardInts = 8, kJ2CardDoubles = 5 * 48 + 64;

template <class BaseT>
void CudaSolidElementT<BaseT>::HistoryToCards(void) const
{
	const char caller[] = "CudaSolidElementT::HistoryToCards";
	if (!fIsJ2 || fGroups.empty()) return;
	const int ne = this->fElementCards.Length();
	std::vector<double> data((size_t) ne * kJ2CardDoubles);
	std::vector<int32_t> flags((size_t) ne * kJ2CardInts), alloc(ne);
	for (size_t g = 0; g < fGroups.size(); g++) {
	if (fKinds[g] != TB2_J2_SIMO) continue;
	Check(tb2_group_get_history(fGroups[g], &data[0], &flags[0], &alloc[0]), caller);
	for (int e = 0; e < ne; e++) {
		if (!alloc[e] || (fGroups.size() > 1 && this->fElementCards[e].MaterialNumber() != int(g))) continue;
		ElementCardT& card = const_cast<ElementCardT&>(this->fElementCards[e]);
		card.Dimension(kJ2CardInts, kJ2CardDoubles);
		for (int i = 0; i < kJ2CardInts; i++) card.IntegerData()[i] = flags[(size_t) e * kJ2CardInts + i];
		memcpy(card.DoubleData().Pointer(), &data[(size_t) e * kJ2CardDoubles], sizeof(double) * kJ2CardDoubles);
	}
	}
}

template <class BaseT>
void CudaSolidElementT<BaseT>::HistoryFromCards(void)
{
	const char caller[] = "CudaSolidElementT::HistoryFromCards";
	if (!fIsJ2 || fGroups.empty()) return;
	const int ne = this->fElementCards.Length();
	for (size_t g = 0; g < fGroups.size(); g++) {
	if (fKinds[g] != TB2_J2_SIMO) continue;
	std::vector<double> data((size_t) ne * kJ2CardDoubles, 0.0);
	std::vector<int32_t> flags((size_t) ne * kJ2CardInts, 0), alloc(ne, 0);
	for (int e = 0; e < ne; e++) {
		const ElementCardT& card = this->fElementCards[e];
		if (!card.IsAllocated() || (fGroups.size() > 1 && card.MaterialNumber() != int(g))) continue;
		if (card.IntegerData().Length() != kJ2CardInts || card.DoubleData().Length() != kJ2CardDoubles)
			ExceptionT::SizeMismatch(caller, "element %d: restart data is not a Simo_J2 Hex8 record", e + 1);
		alloc[e] = 1;
		for (int i = 0; i < kJ2CardInts; i++) flags[(size_t) e * kJ2CardInts + i] = card.IntegerData()[i];
		memcpy(&data[(size_t) e * kJ2CardDoubles], card.DoubleData().Pointer(), sizeof(double) * kJ2CardDoubles);
	}
	Check(tb2_group_set_history(fGroups[g], &data[0], &flags[0], &alloc[0]), caller);
	}
}

template <class BaseT>
void CudaSolidElementT<BaseT>::ReadRestart(istream& in)
{
	BaseT::ReadRestart(in); /* status flags + element cards */
	HistoryFromCards();
	/* status flags may have changed */
	ArrayT<ElementCardT::StatusT> status(this->fElementCards.Length());
	for (int i = 0; i < status.Length(); i++) status[i] = this->fElementCards[i].Flag();
	SetStatus(status);
}

template <class BaseT>
void CudaSolidElementT<BaseT>::WriteRestart(ostream& out) const
{
	HistoryToCards();
	BaseT::WriteRestart(out);
}

/* explicit instantiation */
namespace Tahoe {
template class CudaSolidElementT<SmallStrainT>;
template class CudaSolidElementT<TotalLagrangianT>;
template class CudaSolidElementT<UpdatedLagrangianT>;
template class CudaSolidElementT<ExplicitElementT>;
}
