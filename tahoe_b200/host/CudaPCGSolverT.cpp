/* CudaPCGSolverT.cpp -- see CudaPCGSolverT.h */
#include "CudaPCGSolverT.h"

#include "CudaExplicitSolverT.h"
#include "CudaSolidElementT.h"
#include "ElementBaseT.h"
#include "ExceptionT.h"
#include "FEManagerT.h"
#include "FieldT.h"
#include "ParameterListT.h"
#include "iArray2DT.h"

#include <cstring>
#include <vector>

using namespace Tahoe;

namespace Tahoe {
const char* kCudaPCGSolverName = "CUDA_PCG_solver";

SolverT* NewCudaSolver(FEManagerT& fe_manager, const char* name, int group)
{
	if (strcmp(name, kCudaPCGSolverName) == 0) return new CudaPCGSolverT(fe_manager, group);
	if (strcmp(name, kCudaExplicitSolverName) == 0) return new CudaExplicitSolverT(fe_manager, group);
	return NULL;
}
} // namespace Tahoe

CudaPCGSolverT::CudaPCGSolverT(FEManagerT& fe_manager, int group):
	PCGSolver_LS(fe_manager, group),
	fSolver(NULL),
	fLastSweeps(0)
{
	SetName(kCudaPCGSolverName);
	memset(&fParams, 0, sizeof(fParams));
}

CudaPCGSolverT::~CudaPCGSolverT(void)
{
	if (fSolver) tb2_nlpcg_destroy(fSolver);
}

void CudaPCGSolverT::TakeParameterList(const ParameterListT& list)
{
	/* inherited: NLSolver tolerances, the (host) diagonal_matrix, PCGSolver_LS's own copies of the attributes below */
	PCGSolver_LS::TakeParameterList(list);

	fParams.restart = list.GetParameter("restart");
	fParams.line_search_iterations = list.GetParameter("line_search_iterations");
	fParams.line_search_tolerance = list.GetParameter("line_search_tolerance");
	fParams.max_step = list.GetParameter("max_step");
	fParams.abs_tolerance = fZeroTolerance;
	fParams.rel_tolerance = fTolerance;
	fParams.divergence_tolerance = fDivTolerance;
	fParams.max_iterations = fMaxIterations;
	fParams.min_iterations = fMinIterations;
}

/* the one CudaSolidElementT group of this solver group; any other element group in it has no device residual */
CudaStiffnessSourceT* CudaPCGSolverT::FindDeviceGroup(void) const
{
	const char caller[] = "CudaPCGSolverT::FindDeviceGroup";
	CudaStiffnessSourceT* found = NULL;
	for (int i = 0; i < fFEManager.NumElementGroups(); i++) {
		ElementBaseT* group = fFEManager.ElementGroup(i);
		if (!group->InGroup(Group())) continue;
		CudaStiffnessSourceT* dev = dynamic_cast<CudaStiffnessSourceT*>(group);
		if (!dev || found)
			ExceptionT::BadInputValue(caller, "CUDA_PCG_solver needs exactly one cuda_* continuum element group in its solver group "
				"(element group %d is %s)", i + 1, dev ? "a second one" : "a host group");
		found = dev;
	}
	if (!found) ExceptionT::BadInputValue(caller, "no cuda_* element group in solver group %d", Group() + 1);
	return found;
}

SolverT::SolutionStatusT CudaPCGSolverT::Solve(int max_iterations)
{
	const char caller[] = "CudaPCGSolverT::Solve";
	try {
		CudaStiffnessSourceT* dev = FindDeviceGroup();
		const FieldT& field = dev->DeviceField();
		const iArray2DT& eqnos = field.Equations();
		const dArray2DT& disp = field[0];
		const int ndof = eqnos.Length();
		if (!fSolver) {
			if (!dev->DeviceGroup())
				ExceptionT::BadInputValue(caller, "CUDA_PCG_solver drives a single-material CUDA element group (this one has several materials)");
			int status = tb2_nlpcg_create(dev->DeviceGroup(), dev->DeviceEquations(), &fParams, &fSolver);
			if (status != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
		}

		/* external load of this step on the active equations: Tahoe's own FormRHS (FieldT::FormRHS nodal forces, tractions of
		 * ContinuumElementT::RHSDriver) with the group's internal force left out */
		fRHS_lock = kOpen;
		fLHS_lock = kIgnore;
		fRHS = 0.0;
		dev->MuteInternalForce(true);
		try { fFEManager.FormRHS(Group()); }
		catch (ExceptionT::CodeT code) { dev->MuteInternalForce(false); throw code; }
		dev->MuteInternalForce(false);
		fRHS_lock = kLocked;
		std::vector<double> fext(ndof, 0.0), u(disp.Pointer(), disp.Pointer() + ndof);
		for (int k = 0; k < ndof; k++)
			if (eqnos[k] > 0) fext[k] = fRHS[eqnos[k] - 1];
		const double* u_last = dev->NeedsLastDisplacement() ? field(-1, 0).Pointer() : NULL;

		/* the whole PCGSolver_LS::Solve on the device */
		int status = kContinue, iterations = -1;
		double error = 0.0, error0 = 0.0;
		int64_t sweeps0 = 0, sweeps1 = 0;
		tb2_nlpcg_counters(fSolver, &sweeps0, NULL);
		int rc = tb2_nlpcg_solve_host(fSolver, &u[0], u_last, &fext[0], max_iterations, &status, &iterations, &error, &error0);
		tb2_nlpcg_counters(fSolver, &sweeps1, NULL);
		fLastSweeps = int(sweeps1 - sweeps0);
		if (rc == TB2_ERR_BAD_JACOBIAN) throw ExceptionT::kBadJacobianDet;
		if (rc != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());

		/* Tahoe's FieldT stays authoritative: one update with the total increment (FEManagerT::Update -> FieldT::AssembleUpdate) */
		dArrayT update(fRHS.Length());
		update = 0.0;
		for (int k = 0; k < ndof; k++)
			if (eqnos[k] > 0) update[eqnos[k] - 1] = u[k] - disp[k];
		fFEManager.Update(Group(), update);
		fNumIteration = iterations;
		fError0 = error0;
		cout << "\n Group : " << fGroup + 1 << "\n Absolute error = " << error0 << "\n"
		     << setw(kIntWidth) << iterations << ": Relative error = " << (error0 > 0.0 ? error / error0 : 0.0)
		     << (status == TB2_SOLVER_CONVERGED ? " (converged, device PCG, " : " (device PCG, ") << fLastSweeps << " residual sweeps)\n";

		/* the residual Tahoe sees at the converged state (output, reaction forces) */
		fRHS_lock = kOpen;
		fRHS = 0.0;
		fFEManager.FormRHS(Group());
		fRHS_lock = kLocked;

		if (status == TB2_SOLVER_CONVERGED) return DoConverged();
		return status == TB2_SOLVER_FAILED ? kFailed : kContinue;
	}
	catch (ExceptionT::CodeT code) {
		cout << "\n " << caller << ": exception at step number " << fFEManager.StepNumber() << " with step " << fFEManager.TimeStep()
		     << "\n     " << code << ": " << ExceptionT::ToString(code) << endl;
		return kFailed;
	}
}
