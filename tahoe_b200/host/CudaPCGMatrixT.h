/* CudaPCGMatrixT.h -- Tahoe global-matrix plugin: device CSR + Jacobi-preconditioned CG on a B200.
 *
 * Drop-in through <CUDA_PCG_matrix rel_tolerance= abs_tolerance= max_iterations=/> in the solver's matrix_type_choice.
 * Precedent in the reference: AztecMatrixT (primitives/globalmatrix/aztec/AztecMatrixT.h:18-75), an iterative CG/Jacobi
 * matrix type behind the same GlobalMatrixT::Solve template method (GlobalMatrixT.cpp:77-113).
 *
 * Storage semantics are MSRMatrixT's (this class derives from it), so every Tahoe element group can assemble into it on
 * the host exactly as into SPOOLES_matrix.  Two value sources:
 *   - host assembly (MSRMatrixT::Assemble): converted MSR -> CSR and uploaded at solve time (tb2_matrix_create_csr);
 *   - device assembly by a cooperating CudaSolidElementT (AddDeviceMatrix): the tangent never leaves the GPU.
 * The host MSR structure (MSRBuilderT graph, fbindx, fval: 12 B per non-zero of host memory and a serial graph build) is set up
 * LAZILY, by the first host-side Assemble / Multx / CopyDiagonal: an analysis whose only contributor is a device-assembling
 * group never builds it -- the sparsity lives in the library's tb2_equations / tb2_matrix alone.
 * BackSubstitute() runs tb2_matrix_pcg (K6-K8).
 *
 * Clone(): as for every MSRMatrixT-derived type of the reference (SPOOLESMatrixT::Clone copy-constructs through
 * MSRMatrixT(const MSRMatrixT&), which is "not implemented", MSRMatrixT.cpp:23-27), cloning is not available.
 */
#ifndef _CUDA_PCG_MATRIX_T_H_
#define _CUDA_PCG_MATRIX_T_H_

#include "MSRMatrixT.h"

#include "tahoe_b200.h"

namespace Tahoe {

class CudaPCGMatrixT: public MSRMatrixT
{
public:

	CudaPCGMatrixT(ostream& out, int check_code, bool symmetric, const CommunicatorT& comm, double rel_tol, double abs_tol,
		int max_iterations);
	virtual ~CudaPCGMatrixT(void);

	virtual void Initialize(int tot_num_eq, int loc_num_eq, int start_eq);
	virtual void Clear(void);

	/** \name structure and host assembly: forwarded to MSRMatrixT, the structure itself is built on first use */
	/*@{*/
	virtual void AddEquationSet(const iArray2DT& eqnos);
	virtual void AddEquationSet(const RaggedArray2DT<int>& eqnos);
	virtual void Assemble(const ElementMatrixT& elMat, const ArrayT<int>& eqnos);
	virtual void Assemble(const ElementMatrixT& elMat, const ArrayT<int>& row_eqnos, const ArrayT<int>& col_eqnos);
	virtual void Assemble(const nArrayT<double>& diagonal_elMat, const ArrayT<int>& eqnos);
	virtual bool CopyDiagonal(dArrayT& diags) const;
	virtual void Multx(const dArrayT& x, dArrayT& b) const;
	/*@}*/

	/** true once the host MSR structure exists (tests: a device-assembled analysis leaves it unbuilt) */
	bool HasHostStructure(void) const { return fHostStructure; }
	virtual bool SolvePreservesData(void) const { return true; };
	virtual GlobalT::SystemTypeT MatrixType(void) const { return fSymmetric ? GlobalT::kSymmetric : GlobalT::kNonSymmetric; };
	virtual GlobalMatrixT* Clone(void) const;

	/** a cooperating element group hands over its device-assembled tangent (values stay in HBM) */
	void AddDeviceMatrix(tb2_matrix* A, double scale);

	/** iterations and residual norm of the last solve */
	int LastIterations(void) const { return fLastIterations; }
	double LastResidualNorm(void) const { return fLastResidual; }

protected:

	virtual void Factorize(void) {}; /* Jacobi: the diagonal is extracted on the device inside the solve */
	virtual void BackSubstitute(dArrayT& result);

private:

	void UploadHostMatrix(void);
	void EnsureHostStructure(void);

	double fRelTol, fAbsTol;
	int fMaxIterations;
	tb2_matrix* fHostCSR;     /**< device copy of the host-assembled matrix (owned) */
	tb2_matrix* fDeviceMatrix; /**< tangent assembled on the device by an element group (not owned) */
	int fLastIterations;
	double fLastResidual;
	bool fHostStructure;  /**< MSRMatrixT::Initialize has run for the current equation system */
	bool fHostAssembled;  /**< a host Assemble arrived since the last Clear */
	bool fGroupsStale;    /**< the builder still holds the equation sets of a system that has been initialised */
};

} // namespace Tahoe
#endif
