/* CudaPCGSolverT.h -- Tahoe solver plugin: the nonlinear PCG solver <PCG_solver> running resident on a B200.
 *
 * Drop-in through <CUDA_PCG_solver ...><diagonal_matrix/></CUDA_PCG_solver> in the `solvers` choice: the attributes are those of
 * <PCG_solver> (the class derives from PCGSolver_LS, so DefineParameters / the generated tahoe.xsd entry are inherited).
 * Solve() hands the load step to tb2_nlpcg_solve (include/tahoe_b200.h): residual sweeps, preconditioner, search directions
 * and the line search of PCGSolver_LS (solvers/PCGSolver_LS.cpp:107-371) run on the device against the cooperating
 * CudaSolidElementT group; Tahoe's FieldT receives the converged update through FEManagerT::Update, and external nodal forces /
 * tractions are still formed by Tahoe's own FormRHS (with the group's internal force muted).
 */
#ifndef _CUDA_PCG_SOLVER_T_H_
#define _CUDA_PCG_SOLVER_T_H_

#include "PCGSolver_LS.h"

#include "tahoe_b200.h"

namespace Tahoe {

class CudaStiffnessSourceT;

class CudaPCGSolverT: public PCGSolver_LS
{
public:

	CudaPCGSolverT(FEManagerT& fe_manager, int group);
	virtual ~CudaPCGSolverT(void);

	/** PCGSolver_LS::Solve on the device */
	virtual SolutionStatusT Solve(int max_iterations);

	/** reads restart / line-search attributes (private in PCGSolver_LS) for the device solver */
	virtual void TakeParameterList(const ParameterListT& list);

	/** residual sweeps of the last Solve */
	int LastResidualSweeps(void) const { return fLastSweeps; }

private:

	CudaStiffnessSourceT* FindDeviceGroup(void) const;

	tb2_nlpcg_params fParams;
	tb2_nlpcg* fSolver;
	int fLastSweeps;
};

/** factory used by the one-line registration in SolverT::New; NULL for other names */
SolverT* NewCudaSolver(FEManagerT& fe_manager, const char* name, int group);
extern const char* kCudaPCGSolverName;

} // namespace Tahoe
#endif
