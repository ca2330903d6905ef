/* CudaSolidElementT.h -- Tahoe element-group plugin that runs the Hex8 continuum-solid element loop on a B200 through
 * the C ABI of libtahoe_b200.so (include/tahoe_b200.h).
 *
 * The class template derives from Tahoe's own SmallStrainT / TotalLagrangianT / UpdatedLagrangianT, so the XML
 * parameters, material lists, output, mass matrix and restart format are inherited unchanged and an input file differs
 * from a classic one by the element tag only (<total_lagrangian> -> <cuda_total_lagrangian>, <explicit_solid> ->
 * <cuda_explicit_solid>).  What is replaced is the
 * per-element virtual-call loop:
 *   RHSDriver()  : SolidElementT::ElementRHSDriver (SolidElementT.cpp:1166-1295)  -> tb2_form_internal_force_host
 *                  (+ tb2_form_inertial_force_host when the integrator asks for M a: implicit dynamics)
 *   LHSDriver()  : SolidElementT::ElementLHSDriver (SolidElementT.cpp:1100-1154)  -> tb2_form_stiffness into the
 *                  device CSR of a cooperating CudaPCGMatrixT; any other matrix type keeps Tahoe's host assembly
 *   CloseStep()/ResetStep() : J2 history commit / reset on the device.
 *   ReadRestart()/WriteRestart() and host-side output of a Simo_J2 group: the device-resident history is moved through the
 *                  group's ElementCardT storage, so restart files are interchangeable with the classic element's.
 * Results leave through ElementSupportT::AssembleRHS (ElementSupportT.h:43) with the field's equation array, i.e. one
 * call for the whole group instead of one per element.
 */
#ifndef _CUDA_SOLID_ELEMENT_T_H_
#define _CUDA_SOLID_ELEMENT_T_H_

#include "SmallStrainT.h"
#include "TotalLagrangianT.h"
#include "UpdatedLagrangianT.h"
#include "ExplicitElementT.h"
#include "dArray2DT.h"

#include "tahoe_b200.h"

#include <vector>

namespace Tahoe {

class CudaPCGMatrixT;

class FieldT;

/** interface the cooperating matrix / solver plugins use to reach an element group's device objects */
class CudaStiffnessSourceT
{
public:
	virtual ~CudaStiffnessSourceT(void) {}
	virtual tb2_mesh* DeviceMesh(void) = 0;
	virtual tb2_group* DeviceGroup(void) = 0;
	/** equation numbers of the group's field on the device (built on first use, after Tahoe has numbered the equations) */
	virtual tb2_equations* DeviceEquations(void) = 0;
	virtual const FieldT& DeviceField(void) const = 0;
	virtual int SolverGroup(void) const = 0;
	virtual bool NeedsLastDisplacement(void) const = 0;
	/** while set, RHSDriver() adds the tractions only: a device-resident solver (CudaPCGSolverT) forms -fint itself */
	virtual void MuteInternalForce(bool mute) = 0;
	/** natural_bc tractions or a body force: loads the group itself adds to the residual (they may follow a schedule) */
	virtual bool HasSurfaceOrBodyLoads(void) const = 0;
};

template <class BaseT>
class CudaSolidElementT: public BaseT, public CudaStiffnessSourceT
{
public:

	/** \param name XML tag, \param formulation tb2_formulation */
	CudaSolidElementT(const ElementSupportT& support, const char* name, int formulation);
	virtual ~CudaSolidElementT(void);

	/** build the device mesh / element group after Tahoe has read connectivity and materials */
	virtual void DefineParameters(ParameterListT& list) const;
	virtual void TakeParameterList(const ParameterListT& list);

	/** element status flags (ElementCardT::kOFF elements are skipped by the element loops, SolidElementT.cpp:1116,1177): mirrored
	 * on the device */
	virtual void SetStatus(const ArrayT<ElementCardT::StatusT>& status);

	/** \name history */
	/*@{*/
	virtual void CloseStep(void);
	virtual GlobalT::RelaxCodeT ResetStep(void);
	/** restart files in ContinuumElementT's format (ContinuumElementT.cpp:217-249): the device J2 history passes through the element cards */
	virtual void ReadRestart(istream& in);
	virtual void WriteRestart(ostream& out) const;
	/*@}*/

	virtual tb2_mesh* DeviceMesh(void) { return fMesh; }
	/** the device group of a single-material element group (what the resident solvers drive); NULL when the group has several materials */
	virtual tb2_group* DeviceGroup(void) { return fGroups.size() == 1 ? fGroup : NULL; }
	virtual tb2_equations* DeviceEquations(void);
	virtual const FieldT& DeviceField(void) const { return this->Field(); }
	virtual int SolverGroup(void) const { return this->Group(); }
	virtual bool NeedsLastDisplacement(void) const { return fIsJ2; }
	virtual void MuteInternalForce(bool mute) { fMuted = mute; }
	virtual bool HasSurfaceOrBodyLoads(void) const { return this->fTractionList.Length() > 0 || this->fBodySchedule != NULL; }

protected:

	/** residual: tractions on the host (ContinuumElementT::RHSDriver), element forces on the device */
	virtual void RHSDriver(void);

	/** tangent: device assembly when the solver's matrix is a CudaPCGMatrixT, else inherited */
	virtual void LHSDriver(GlobalT::SystemTypeT sys_type);

	/** nodal output: displacements and extrapolated, nodally averaged Cauchy stresses from the device (tb2_group_nodal_stress_host)
	 * when nothing else is requested; any other combination of output codes runs SolidElementT::ComputeOutput on the host */
	virtual void ComputeOutput(const iArrayT& n_codes, dArray2DT& n_values, const iArrayT& e_codes, dArray2DT& e_values);

private:

	void Check(int status, const char* caller) const;

	/** device J2 history -> ElementCardT storage in J2SimoC0HardeningT's layout (AllocateElement :312-333, LoadData :429-452) */
	void HistoryToCards(void) const;
	/** ElementCardT storage -> device */
	void HistoryFromCards(void);

	int fFormulation;
	tb2_mesh* fMesh;
	tb2_group* fGroup;               /**< fGroups[0] */
	std::vector<tb2_group*> fGroups; /**< one device group per material of the list; they share fMesh and mask each other's elements off */
	std::vector<int> fKinds;         /**< tb2_material_kind of each */
	dArray2DT fPart;                 /**< [nn][3] one material's share of a nodal force */
	tb2_equations* fEqs;   /**< built lazily: equation numbers are set after TakeParameterList */
	tb2_matrix* fMatrix;   /**< device tangent of this group (structure from the mesh) */
	bool fIsJ2;
	bool fMuted;           /**< see MuteInternalForce */
	dArray2DT fFint;       /**< [nn][3] internal force of the whole group */
	dArray2DT fMa;         /**< [nn][3] inertia force of the whole group (implicit integrators) */
	dArray2DT fBodyAcc;    /**< [nn][3] the constant nodal field -b * schedule of a body force */
	int fMaterialKind;     /**< tb2_material_kind */
	int fDevice;           /**< CUDA device ordinal of this group (attribute "device", default 0; one process per GPU sets its own) */
};

typedef CudaSolidElementT<SmallStrainT> CudaSmallStrainT;
typedef CudaSolidElementT<TotalLagrangianT> CudaTotalLagrangianT;
typedef CudaSolidElementT<UpdatedLagrangianT> CudaUpdatedLagrangianT;
/** <cuda_explicit_solid>: ExplicitElementT's XML (incl. <j2_plasticity>, <mass_scaling>) and host-side lumped mass, the batched internal force on the device */
typedef CudaSolidElementT<ExplicitElementT> CudaExplicitSolidT;

/** factory used by the one-line registration in ElementListT::NewElement; returns NULL for other names */
ElementBaseT* NewCudaSolidElement(const StringT& name, const ElementSupportT& support);
/** the XML tags handled by NewCudaSolidElement */
static const int kNumCudaSolidElementNames = 5; /* four continuum groups + cuda_contact_3D_penalty (CudaPenaltyContact3DT.h) */
extern const char* kCudaSolidElementNames[kNumCudaSolidElementNames];

} // namespace Tahoe
#endif
