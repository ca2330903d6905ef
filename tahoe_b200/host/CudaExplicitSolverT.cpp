/* CudaExplicitSolverT.cpp -- see CudaExplicitSolverT.h */
#include "CudaExplicitSolverT.h"

#include "CudaPenaltyContact3DT.h"
#include "CudaSolidElementT.h"
#include "ElementBaseT.h"
#include "ExceptionT.h"
#include "FEManagerT.h"
#include "FieldT.h"
#include "NodeManagerT.h"
#include "ParameterListT.h"
#include "TimeManagerT.h"
#include "iArray2DT.h"

#include <cmath>
#include <cstring>

using namespace Tahoe;

namespace Tahoe {
const char* kCudaExplicitCDName = "CUDA_central_difference";
const char* kCudaExplicitSolverName = "CUDA_explicit_solver";

IntegratorT* NewCudaIntegrator(int type)
{
	if (type == CudaExplicitCDIntegrator::kCode) return new CudaExplicitCDIntegrator;
	return NULL;
}
} // namespace Tahoe

void CudaExplicitCDIntegrator::Predictor(BasicFieldT& field, int fieldstart, int fieldend)
{
	if (fResident) return; /* tb2_explicit_run applies it to the device-resident fields */
	ExplicitCDIntegrator::Predictor(field, fieldstart, fieldend);
}

CudaExplicitSolverT::CudaExplicitSolverT(FEManagerT& fe_manager, int group):
	SolverT(fe_manager, group),
	fEx(NULL),
	fDev(NULL),
	fContact(NULL),
	fHasLoads(false),
	fLoadsChecked(false),
	fSteps(0),
	fDownloads(0),
	fRestartInc(0)
{
	SetName(kCudaExplicitSolverName);
}

CudaExplicitSolverT::~CudaExplicitSolverT(void)
{
	if (fEx) tb2_explicit_destroy(fEx);
}

void CudaExplicitSolverT::DefineParameters(ParameterListT& list) const
{
	SolverT::DefineParameters(list);
	ParameterT restart_inc(ParameterT::Integer, "restart_output_inc");
	restart_inc.SetDefault(0);
	restart_inc.AddLimit(0, LimitT::LowerInclusive);
	list.AddParameter(restart_inc);
}

void CudaExplicitSolverT::TakeParameterList(const ParameterListT& list)
{
	SolverT::TakeParameterList(list);
	fRestartInc = list.GetParameter("restart_output_inc");
}

void CudaExplicitSolverT::Check(int status, const char* caller) const
{
	if (status == TB2_OK) return;
	if (status == TB2_ERR_BAD_JACOBIAN) throw ExceptionT::kBadJacobianDet;
	ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
}

CudaStiffnessSourceT* CudaExplicitSolverT::FindDeviceGroup(void) const
{
	const char caller[] = "CudaExplicitSolverT::FindDeviceGroup";
	CudaStiffnessSourceT* found = NULL;
	CudaPenaltyContact3DT* contact = NULL;
	for (int i = 0; i < fFEManager.NumElementGroups(); i++) {
		ElementBaseT* group = fFEManager.ElementGroup(i);
		if (!group->InGroup(Group())) continue;
		CudaPenaltyContact3DT* c = dynamic_cast<CudaPenaltyContact3DT*>(group);
		if (c && !contact) { contact = c; continue; } /* one contact group rides along: its force is formed inside the device step */
		CudaStiffnessSourceT* dev = dynamic_cast<CudaStiffnessSourceT*>(group);
		if (!dev || found)
			ExceptionT::BadInputValue(caller, "CUDA_explicit_solver needs exactly one cuda_* continuum element group (and at most one "
				"cuda_contact_3D_penalty group) in its solver group (element group %d is %s)", i + 1, dev ? "a second one" : "a host group");
		found = dev;
	}
	if (!found) ExceptionT::BadInputValue(caller, "no cuda_* element group in solver group %d", Group() + 1);
	const_cast<CudaExplicitSolverT*>(this)->fContact = contact;
	return found;
}

/* device state from Tahoe's FieldT: d, v, a as they stand, a kinematic condition on every dof without an equation */
void CudaExplicitSolverT::Setup(CudaStiffnessSourceT* dev)
{
	const char caller[] = "CudaExplicitSolverT::Setup";
	if (!dev->DeviceGroup())
		ExceptionT::BadInputValue(caller, "CUDA_explicit_solver drives a single-material CUDA element group (this one has several materials)");
	const FieldT& field = dev->DeviceField();
	CudaExplicitCDIntegrator* integrator = dynamic_cast<CudaExplicitCDIntegrator*>(const_cast<nIntegratorT*>(&field.nIntegrator()));
	if (!integrator)
		ExceptionT::BadInputValue(caller, "field \"%s\" must use integrator=\"%s\": with the host integrator the predictor would run twice",
			field.FieldName().Pointer(), kCudaExplicitCDName);
	if (field.Order() != 2) ExceptionT::BadInputValue(caller, "expecting a second-order field");
	Check(tb2_explicit_create(dev->DeviceGroup(), &fEx), caller);
	const iArray2DT& eqnos = field.Equations();
	const int ndof = eqnos.Length();
	std::vector<uint8_t> code(ndof, TB2_BC_FREE);
	std::vector<double> value(ndof, 0.0);
	fPrescribed.clear();
	for (int k = 0; k < ndof; k++)
		if (eqnos[k] < 1) {
			code[k] = TB2_BC_DSP;
			value[k] = field[0][k];
			fPrescribed.push_back(k);
		}
	fPrescribedValue.assign(fPrescribed.size(), 0.0);
	for (size_t q = 0; q < fPrescribed.size(); q++) fPrescribedValue[q] = value[fPrescribed[q]];
	fFext.assign(ndof, 0.0);
	Check(tb2_explicit_set_bc(fEx, &code[0], &value[0], &fFext[0]), caller);
	Check(tb2_explicit_set_state(fEx, field[0].Pointer(), field[1].Pointer(), field[2].Pointer()), caller);
	fHasLoads = const_cast<FieldT&>(field).ForceBC().Length() > 0;
	integrator->SetResident(true);
	fDev = dev;
	if (fContact) {
		if (fContact->DeviceMesh() != dev->DeviceMesh())
			ExceptionT::GeneralFail(caller, "the contact group and the continuum group do not share a device mesh");
		/* the device object searches for itself from now on (tb2_contact_search after every step, as LinearSolver::Solve relaxes the
		 * system after its update, LinearSolver.cpp:76-90): nothing of the contact crosses the bus per step */
		fContact->SendSurfaces();
		Check(tb2_explicit_attach_contact(fEx, fContact->DeviceContact()), caller);
	}
}

SolverT::SolutionStatusT CudaExplicitSolverT::Solve(int)
{
	const char caller[] = "CudaExplicitSolverT::Solve";
	try {
		if (!fEx) Setup(FindDeviceGroup());
		const FieldT& field = fDev->DeviceField();
		const iArray2DT& eqnos = field.Equations();
		const int ndof = eqnos.Length();

		/* prescribed values of this step: Tahoe's KBC controllers wrote them into the host field (FieldT::InitStep ->
		 * nExplicitCD::ConsistentKBC, nExplicitCD.cpp:20-69); they go up when they changed */
		if (!fPrescribed.empty()) {
			const double* d = field[0].Pointer();
			bool changed = false;
			fScratch.resize(fPrescribed.size());
			for (size_t q = 0; q < fPrescribed.size(); q++) {
				fScratch[q] = d[fPrescribed[q]];
				changed = changed || fScratch[q] != fPrescribedValue[q];
			}
			if (changed) {
				fPrescribedValue = fScratch;
				Check(tb2_explicit_update_bc_values(fEx, (int64_t)fPrescribed.size(), &fPrescribed[0], &fPrescribedValue[0]), caller);
			}
		}

		/* external load on the active equations: Tahoe's own FormRHS (nodal forces of FieldT::FormRHS, tractions and body forces
		 * of the element group) with the group's internal force left out -- formed while there is something to form */
		/* loads follow time through the schedules only (FBC cards, tractions, body forces all scale with a ScheduleT value): they are
		 * formed again when a schedule value moved, not every step */
		bool schedules_moved = !fLoadsChecked;
		{
			const TimeManagerT* tm = fFEManager.TimeManager();
			const int ns = tm->NumSchedule();
			if ((int)fScheduleValue.size() != ns) { fScheduleValue.assign(ns, 0.0); schedules_moved = true; }
			for (int k = 0; k < ns; k++) {
				const double v = tm->ScheduleValue(k);
				if (v != fScheduleValue[k]) { fScheduleValue[k] = v; schedules_moved = true; }
			}
		}
		if ((fHasLoads && schedules_moved) || !fLoadsChecked) {
			fRHS_lock = kOpen;
			fLHS_lock = kIgnore;
			fRHS = 0.0;
			fDev->MuteInternalForce(true);
			if (fContact) fContact->MuteForce(true);
			try { fFEManager.FormRHS(Group()); }
			catch (ExceptionT::CodeT code) { fDev->MuteInternalForce(false); if (fContact) fContact->MuteForce(false); throw code; }
			fDev->MuteInternalForce(false);
			if (fContact) fContact->MuteForce(false);
			fRHS_lock = kLocked;
			double biggest = 0.0;
			for (int k = 0; k < ndof; k++) {
				fFext[k] = eqnos[k] > 0 ? fRHS[eqnos[k] - 1] : 0.0;
				biggest = fabs(fFext[k]) > biggest ? fabs(fFext[k]) : biggest;
			}
			if (!fLoadsChecked) {
				fLoadsChecked = true;
				fHasLoads = fHasLoads || biggest > 0.0 || fDev->HasSurfaceOrBodyLoads();
			}
			if (fHasLoads || biggest > 0.0) Check(tb2_explicit_set_bc(fEx, NULL, NULL, &fFext[0]), caller);
		}

		/* predictor + KBC values, internal force (+ contact force on the predicted state), a = M^-1 R, corrector; with a contact group
		 * attached the run also searches for the striker-facet pairs of the corrected configuration, on the device */
		Check(tb2_explicit_run(fEx, fFEManager.TimeStep(), 1, NULL, NULL), caller);
		fSteps++;
		return kConverged;
	}
	catch (ExceptionT::CodeT code) {
		cout << "\n " << caller << ": exception at step number " << fFEManager.StepNumber() << " with step " << fFEManager.TimeStep()
		     << "\n     " << code << ": " << ExceptionT::ToString(code) << endl;
		return kFailed;
	}
}

/* FEManagerT::CloseStep calls the solvers first, then WriteOutput and WriteRestart (FEManagerT.cpp:625-660): the host fields
 * are brought up to date exactly when one of the two will read them */
void CudaExplicitSolverT::CloseStep(void)
{
	SolverT::CloseStep();
	if (!fEx) return;
	const int step = fFEManager.StepNumber();
	const bool last = step == fFEManager.NumberOfSteps();
	const bool output = fFEManager.TimeManager()->WriteOutput();
	const bool restart = fRestartInc > 0 && step % fRestartInc == 0;
	if (!(output || restart || last || fabs(fFEManager.TimeStep()) < kSmall)) return;
	FieldT& field = const_cast<FieldT&>(fDev->DeviceField());
	Check(tb2_explicit_get_state(fEx, field[0].Pointer(), field[1].Pointer(), field[2].Pointer()), "CudaExplicitSolverT::CloseStep");
	fDownloads++;
	if (fContact) { /* Tahoe's own contact object catches up for its log and output (ContactT::RelaxSystem writes the contact info) */
		fFEManager.NodeManager()->UpdateCurrentCoordinates();
		fFEManager.RelaxSystem(Group());
	}
}

void CudaExplicitSolverT::ResetStep(void)
{
	ExceptionT::GeneralFail("CudaExplicitSolverT::ResetStep", "an explicit device step cannot be taken back");
}
