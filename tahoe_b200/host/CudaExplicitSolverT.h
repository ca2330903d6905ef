/* CudaExplicitSolverT.h -- Tahoe plugin pair that keeps an explicit central-difference run RESIDENT on a B200.
 *
 * Reference path: FEManagerT::InitStep -> FieldT::InitStep -> nExplicitCD::Predictor (nExplicitCD.cpp:72-96);
 * FEManagerT::SolveStep -> LinearSolver::Solve (LinearSolver.cpp:37-101): FormRHS (element loop), DiagonalMatrixT::Solve
 * (DiagonalMatrixT.cpp:267-323), FEManagerT::Update -> FieldT::AssembleUpdate (FieldT.cpp:531-556) -> nExplicitCD::Corrector
 * (:98-139).  With <cuda_*> element groups alone, u goes to the device and the force comes back every step, and Tahoe's host
 * loops do the O(N_n) updates.  The two classes below take the whole step:
 *
 *   integrator="CUDA_central_difference"   CudaExplicitCDIntegrator : ExplicitCDIntegrator -- the host predictor / corrector
 *                                          are no-ops while a resident solver owns the field (the device runs them);
 *   <CUDA_explicit_solver>                 CudaExplicitSolverT : SolverT -- Solve() = tb2_explicit_run(dt, 1) on the cooperating
 *                                          cuda_* element group: predictor + ConsistentKBC values, internal force, M^-1 R,
 *                                          corrector on device-resident d, v, a.  Prescribed values and external loads are
 *                                          Tahoe's own (KBC controllers / FormRHS with the internal force muted) and go up only
 *                                          when they exist; d, v, a come back to FieldT when a step writes output or a restart
 *                                          file (SolverT::CloseStep is called before FEManagerT::WriteOutput).
 */
#ifndef _CUDA_EXPLICIT_SOLVER_T_H_
#define _CUDA_EXPLICIT_SOLVER_T_H_

#include "ExplicitCDIntegrator.h"
#include "SolverT.h"

#include "tahoe_b200.h"

#include <vector>

namespace Tahoe {

class CudaStiffnessSourceT;
class CudaPenaltyContact3DT;
class FieldT;

class CudaExplicitCDIntegrator: public ExplicitCDIntegrator
{
public:

	/** value of FieldT's "integrator" enumeration (FieldT.cpp:1040-1052) for "CUDA_central_difference" */
	enum { kCode = 105 };

	CudaExplicitCDIntegrator(void): fResident(false) {}

	/** nExplicitCD::Predictor unless a resident solver owns the field */
	virtual void Predictor(BasicFieldT& field, int fieldstart = 0, int fieldend = -1);

	void SetResident(bool resident) { fResident = resident; }
	bool Resident(void) const { return fResident; }

private:

	bool fResident;
};

/** factory used by the registration in IntegratorT::New; NULL for other codes */
IntegratorT* NewCudaIntegrator(int type);
extern const char* kCudaExplicitCDName;

class CudaExplicitSolverT: public SolverT
{
public:

	CudaExplicitSolverT(FEManagerT& fe_manager, int group);
	virtual ~CudaExplicitSolverT(void);

	/** one explicit step on the device */
	virtual SolutionStatusT Solve(int max_iterations);

	/** brings d, v, a back to Tahoe's FieldT when this step writes output or a restart file */
	virtual void CloseStep(void);

	virtual void ResetStep(void);

	/** SolverT's parameters + restart_output_inc: FEManagerT keeps its own copy private, and the fields must be current when
	 * FEManagerT::WriteRestart (FEManagerT.cpp:2150-2200) writes them -- give the run's value here */
	virtual void DefineParameters(ParameterListT& list) const;
	virtual void TakeParameterList(const ParameterListT& list);

	/** steps taken on the device / state downloads so far */
	int DeviceSteps(void) const { return fSteps; }
	int Downloads(void) const { return fDownloads; }

private:

	CudaStiffnessSourceT* FindDeviceGroup(void) const;
	void Setup(CudaStiffnessSourceT* dev);
	void Check(int status, const char* caller) const;

	tb2_explicit* fEx;
	CudaStiffnessSourceT* fDev;
	CudaPenaltyContact3DT* fContact;   /**< a cuda_contact_3D_penalty group of the solver group: its force is formed in the device step */
	std::vector<int64_t> fPrescribed;  /**< nodal dof indices 3 n + i with a kinematic boundary condition */
	std::vector<double> fPrescribedValue, fScratch;
	std::vector<double> fFext;         /**< [nn][3] external load of the step */
	std::vector<double> fScheduleValue; /**< schedule values the loads on the device were formed with */
	bool fHasLoads;                    /**< the field has nodal forces, or the first muted FormRHS was non-zero */
	bool fLoadsChecked;
	int fSteps, fDownloads, fRestartInc;
};

extern const char* kCudaExplicitSolverName;

} // namespace Tahoe
#endif
