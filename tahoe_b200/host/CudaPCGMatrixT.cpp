/* CudaPCGMatrixT.cpp -- see CudaPCGMatrixT.h */
#include "CudaPCGMatrixT.h"

#include "ExceptionT.h"
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
	fLastResidual(0.0)
{
}

CudaPCGMatrixT::~CudaPCGMatrixT(void)
{
	if (fHostCSR) tb2_matrix_destroy(fHostCSR);
}

GlobalMatrixT* CudaPCGMatrixT::Clone(void) const
{
	ExceptionT::GeneralFail("CudaPCGMatrixT::Clone", "not implemented");
	return NULL;
}

void CudaPCGMatrixT::Initialize(int tot_num_eq, int loc_num_eq, int start_eq)
{
	/* inherited: builds the MSR structure from the equation sets (MSRMatrixT.cpp:35-45) */
	MSRMatrixT::Initialize(tot_num_eq, loc_num_eq, start_eq);
	if (fTotNumEQ != fLocNumEQ || fStartEQ != 1)
		ExceptionT::GeneralFail("CudaPCGMatrixT::Initialize", "one process owns all equations (multi-GPU runs are element-partitioned inside the library)");
	if (fHostCSR) tb2_matrix_destroy(fHostCSR);
	fHostCSR = NULL;
	fDeviceMatrix = NULL;
}

void CudaPCGMatrixT::Clear(void)
{
	MSRMatrixT::Clear();
	fDeviceMatrix = NULL;
}

void CudaPCGMatrixT::AddDeviceMatrix(tb2_matrix* A, double scale)
{
	const char caller[] = "CudaPCGMatrixT::AddDeviceMatrix";
	if (fDeviceMatrix && fDeviceMatrix != A) ExceptionT::GeneralFail(caller, "one device-assembling element group per solver group");
	if (fabs(scale - 1.0) > 1.0e-14) ExceptionT::GeneralFail(caller, "tangent scale %g != 1 (static analyses only)", scale);
	fDeviceMatrix = A;
}

bool CudaPCGMatrixT::HostValuesAreZero(void) const
{
	const double* v = fval.Pointer();
	for (int i = 0; i < fval.Length(); i++)
		if (v[i] != 0.0) return false;
	return true;
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
		if (!HostValuesAreZero())
			ExceptionT::GeneralFail(caller, "mixed host- and device-assembled contributions are not supported yet");
		A = fDeviceMatrix;
	} else {
		UploadHostMatrix();
		A = fHostCSR;
	}
	dArrayT x(result.Length());
	x = 0.0;
	int status = tb2_matrix_pcg_host(A, result.Pointer(), x.Pointer(), fRelTol, fAbsTol, fMaxIterations, &fLastIterations, &fLastResidual);
	if (status != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
	fOut << " CudaPCGMatrixT: " << fLastIterations << " PCG iterations, |r| = " << fLastResidual << '\n';
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
