/* CudaPCGMatrixT.cpp -- see CudaPCGMatrixT.h */
#include "CudaPCGMatrixT.h"

#include "ExceptionT.h"
#include "MSRBuilderT.h"
#include "dArrayT.h"

#include <vector>

using namespace Tahoe;

CudaPCGMatrixT::CudaPCGMatrixT(ostream& out, int check_code, bool symmetric, const CommunicatorT& comm, double rel_tol,
	double abs_tol, int max_iterations):
	MSRMatrixT(out, check_code, symmetric, comm),
	fRelTol(rel_tol),
	fAbsTol(abs_tol),
	fMaxIterations(max_iterations),
	fHostCSR(NULL),
	fDeviceMatrix(NULL),
	fLastIterations(0),
	fLastResidual(0.0),
	fHostStructure(false),
	fHostAssembled(false),
	fGroupsStale(false)
{
}

CudaPCGMatrixT::~CudaPCGMatrixT(void)
{
	if (fHostCSR) tb2_matrix_destroy(fHostCSR);
}

GlobalMatrixT* CudaPCGMatrixT::Clone(void) const
{
	/* the reference's MSR family cannot be cloned either: MSRMatrixT's copy constructor fails (MSRMatrixT.cpp:23-27) */
	ExceptionT::GeneralFail("CudaPCGMatrixT::Clone", "MSRMatrixT-derived matrices cannot be copied (as SPOOLES_matrix)");
	return NULL;
}

/* the equation sets go to the MSR builder as always (MSRMatrixT.cpp:50-60); sets left over from a system that was initialised
 * without ever building its host structure are dropped first (MSRMatrixT::SetMSRData clears them only when it runs) */
void CudaPCGMatrixT::AddEquationSet(const iArray2DT& eqnos)
{
	if (fGroupsStale) { fMSRBuilder->ClearGroups(); fGroupsStale = false; }
	MSRMatrixT::AddEquationSet(eqnos);
}

void CudaPCGMatrixT::AddEquationSet(const RaggedArray2DT<int>& eqnos)
{
	if (fGroupsStale) { fMSRBuilder->ClearGroups(); fGroupsStale = false; }
	MSRMatrixT::AddEquationSet(eqnos);
}

void CudaPCGMatrixT::Initialize(int tot_num_eq, int loc_num_eq, int start_eq)
{
	/* the dimensions only (GlobalMatrixT.cpp); MSRMatrixT::Initialize -- graph, fbindx, fval -- waits for the first host access */
	GlobalMatrixT::Initialize(tot_num_eq, loc_num_eq, start_eq);
	if (fTotNumEQ != fLocNumEQ || fStartEQ != 1)
		ExceptionT::GeneralFail("CudaPCGMatrixT::Initialize", "one process owns all equations (multi-GPU runs are element-partitioned inside the library)");
	if (fHostCSR) tb2_matrix_destroy(fHostCSR);
	fHostCSR = NULL;
	fDeviceMatrix = NULL;
	fHostStructure = false;
	fHostAssembled = false;
	fGroupsStale = true;
}

void CudaPCGMatrixT::EnsureHostStructure(void)
{
	if (fHostStructure) return;
	MSRMatrixT::Initialize(fTotNumEQ, fLocNumEQ, fStartEQ); /* consumes and clears the builder's equation sets */
	MSRMatrixT::Clear();
	fHostStructure = true;
	fGroupsStale = false;
}

void CudaPCGMatrixT::Clear(void)
{
	if (fHostStructure) MSRMatrixT::Clear();
	fHostAssembled = false;
	fDeviceMatrix = NULL;
}

void CudaPCGMatrixT::Assemble(const ElementMatrixT& elMat, const ArrayT<int>& eqnos)
{
	EnsureHostStructure();
	fHostAssembled = true;
	MSRMatrixT::Assemble(elMat, eqnos);
}

void CudaPCGMatrixT::Assemble(const ElementMatrixT& elMat, const ArrayT<int>& row_eqnos, const ArrayT<int>& col_eqnos)
{
	EnsureHostStructure();
	fHostAssembled = true;
	MSRMatrixT::Assemble(elMat, row_eqnos, col_eqnos);
}

void CudaPCGMatrixT::Assemble(const nArrayT<double>& diagonal_elMat, const ArrayT<int>& eqnos)
{
	EnsureHostStructure();
	fHostAssembled = true;
	MSRMatrixT::Assemble(diagonal_elMat, eqnos);
}

/* host-side reads of a device-assembled matrix go to the device copy */
bool CudaPCGMatrixT::CopyDiagonal(dArrayT& diags) const
{
	if (fDeviceMatrix && !fHostAssembled) {
		diags.Dimension(fLocNumEQ);
		if (tb2_matrix_copy_diagonal_host(fDeviceMatrix, diags.Pointer()) != TB2_OK)
			ExceptionT::GeneralFail("CudaPCGMatrixT::CopyDiagonal", "%s", tb2_last_error());
		return true;
	}
	const_cast<CudaPCGMatrixT*>(this)->EnsureHostStructure();
	return MSRMatrixT::CopyDiagonal(diags);
}

void CudaPCGMatrixT::Multx(const dArrayT& x, dArrayT& b) const
{
	if (fDeviceMatrix && !fHostAssembled) {
		if (tb2_matrix_multx_host(fDeviceMatrix, x.Pointer(), b.Pointer()) != TB2_OK)
			ExceptionT::GeneralFail("CudaPCGMatrixT::Multx", "%s", tb2_last_error());
		return;
	}
	const_cast<CudaPCGMatrixT*>(this)->EnsureHostStructure();
	MSRMatrixT::Multx(x, b);
}

void CudaPCGMatrixT::AddDeviceMatrix(tb2_matrix* A, double scale)
{
	const char caller[] = "CudaPCGMatrixT::AddDeviceMatrix";
	if (fDeviceMatrix && fDeviceMatrix != A) ExceptionT::GeneralFail(caller, "one device-assembling element group per solver group");
	if (fabs(scale - 1.0) > 1.0e-14) ExceptionT::GeneralFail(caller, "tangent scale %g != 1 (static analyses only)", scale);
	fDeviceMatrix = A;
}

/* MSR (MSRMatrixT.h:21-23: fval[0..n-1] diagonal, fbindx[0..n] row starts, then off-diagonal columns; upper triangle only
 * when symmetric) -> full CSR with the diagonal in place, uploaded to the device */
void CudaPCGMatrixT::UploadHostMatrix(void)
{
	const char caller[] = "CudaPCGMatrixT::UploadHostMatrix";
	const int n = fLocNumEQ;
	const int* bindx = fbindx.Pointer();
	const double* val = fval.Pointer();
	std::vector<int64_t> rowptr(n + 1, 0);
	for (int r = 0; r < n; r++) {
		rowptr[r + 1] += 1;
		for (int k = bindx[r]; k < bindx[r + 1]; k++) {
			rowptr[r + 1] += 1;
			if (fSymmetric) rowptr[bindx[k] + 1] += 1;
		}
	}
	for (int r = 0; r < n; r++) rowptr[r + 1] += rowptr[r];
	std::vector<int32_t> colind(rowptr[n]);
	std::vector<double> v(rowptr[n]);
	std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
	/* rows are visited in ascending order, so transposed (lower) entries land before the diagonal and stay sorted */
	for (int r = 0; r < n; r++) {
		colind[fill[r]] = r;
		v[fill[r]++] = val[r];
		for (int k = bindx[r]; k < bindx[r + 1]; k++) {
			const int c = bindx[k];
			if (fSymmetric) { /* (r,c) with c > r, and its mirror (c,r) */
				colind[fill[c]] = r;
				v[fill[c]++] = val[k];
			}
		}
		if (!fSymmetric) { /* full rows: merge the diagonal into the sorted off-diagonals */
			fill[r] = rowptr[r];
			bool diag_done = false;
			for (int k = bindx[r]; k < bindx[r + 1]; k++) {
				if (!diag_done && bindx[k] > r) { colind[fill[r]] = r; v[fill[r]++] = val[r]; diag_done = true; }
				colind[fill[r]] = bindx[k];
				v[fill[r]++] = val[k];
			}
			if (!diag_done) { colind[fill[r]] = r; v[fill[r]++] = val[r]; }
		}
	}
	if (fSymmetric) /* second pass: the upper entries after the diagonal */
		for (int r = 0; r < n; r++)
			for (int k = bindx[r]; k < bindx[r + 1]; k++) {
				colind[fill[r]] = bindx[k];
				v[fill[r]++] = val[k];
			}
	if (!fHostCSR && tb2_matrix_create_csr(0, n, &rowptr[0], &colind[0], &fHostCSR) != TB2_OK)
		ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
	if (tb2_matrix_set_values(fHostCSR, &v[0]) != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
}

void CudaPCGMatrixT::BackSubstitute(dArrayT& result)
{
	const char caller[] = "CudaPCGMatrixT::BackSubstitute";
	tb2_matrix* A = NULL;
	if (fDeviceMatrix) {
		if (fHostAssembled)
			ExceptionT::GeneralFail(caller, "mixed host- and device-assembled contributions are not supported yet");
		A = fDeviceMatrix;
	} else {
		EnsureHostStructure(); /* nothing assembled at all: the zero matrix, as MSRMatrixT would hold it */
		UploadHostMatrix();
		A = fHostCSR;
	}
	fOut << " CudaPCGMatrixT: " << (fDeviceMatrix ? "device-assembled tangent" : "host-assembled MSR values")
	     << ", host MSR structure " << (fHostStructure ? "built" : "not built") << '\n';
	dArrayT x(result.Length());
	x = 0.0;
	/* a non-symmetric system (J2Simo3D's consistent tangent: GlobalT::kNonSymmetric, which the reference hands to an LU, SolverT.cpp:1108-1109)
	 * goes to the Jacobi-preconditioned BiCGStab, a symmetric one to the PCG */
	int status = fSymmetric
		? tb2_matrix_pcg_host(A, result.Pointer(), x.Pointer(), fRelTol, fAbsTol, fMaxIterations, &fLastIterations, &fLastResidual)
		: tb2_matrix_bicgstab_host(A, result.Pointer(), x.Pointer(), fRelTol, fAbsTol, fMaxIterations, &fLastIterations, &fLastResidual);
	if (status != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
	fOut << " CudaPCGMatrixT: " << fLastIterations << (fSymmetric ? " PCG" : " BiCGStab") << " iterations, |r| = " << fLastResidual << '\n';
	/* an iterative solve that ran out of iterations is a failed solve: GlobalMatrixT::Solve catches the exception and returns
	 * false (GlobalMatrixT.cpp:77-113), as it does for a zero pivot of a direct solver */
	int converged = 1;
	double rel = 0.0;
	tb2_matrix_pcg_converged(A, &converged, &rel);
	if (!converged)
		ExceptionT::GeneralFail(caller, "no convergence in %d iterations: |r|/|r0| = %g (rel_tolerance %g, abs_tolerance %g)",
			fLastIterations, rel, fRelTol, fAbsTol);
	result = x;
}
