/* CudaPenaltyContact3DT.h -- Tahoe element-group plugin: contact_3D_penalty with its force loop on a B200.
 *
 * Drop-in through <cuda_contact_3D_penalty ...> in place of <contact_3D_penalty ...> (same attributes and sub-lists).  Everything
 * but the force loop is inherited from the reference's PenaltyContact3DT / Contact3DT / ContactT: the surface set-up, the
 * neighbour search that maintains the active striker-facet pairs (Contact3DT::SetActiveInteractions, Contact3DT.cpp:100-175),
 * the tangent (PenaltyContact3DT::LHSDriver) and the output of the contact state.  RHSDriver (PenaltyContact3DT.cpp:262-500: the
 * loop the reference runs under OpenMP with a critical section around its assembly) becomes tb2_contact_form_host -- pair kernel +
 * ordered per-node sums -- followed by ONE ElementSupportT::AssembleRHS for the whole group.
 *
 * The device mesh is shared with a cuda_* continuum group of the same analysis when there is one (CudaStiffnessSourceT), else the
 * group keeps a coordinates-only mesh of its own.  The slip-based friction of static analyses (per-striker history,
 * PenaltyContact3DT.cpp:427-447) has no device form: that combination is rejected at input, not run on the host.
 */
#ifndef _CUDA_PENALTY_CONTACT_3D_T_H_
#define _CUDA_PENALTY_CONTACT_3D_T_H_

#include "PenaltyContact3DT.h"
#include "dArray2DT.h"

#include <vector>

#include "tahoe_b200.h"

namespace Tahoe {

class CudaPenaltyContact3DT: public PenaltyContact3DT
{
public:

	CudaPenaltyContact3DT(const ElementSupportT& support, const char* name);
	virtual ~CudaPenaltyContact3DT(void);

	virtual void TakeParameterList(const ParameterListT& list);

	/** the group's device object with the current pair list (a resident explicit step attaches it with tb2_explicit_attach_contact) */
	tb2_contact* DeviceContact(void);

	/** while set, RHSDriver() adds nothing: a resident solver that attached the group forms the contact force on the device itself */
	void MuteForce(bool mute) { fMuted = mute; }

	/** hands the triangulated surfaces, the strikers and their areas to the device object, whose own search (tb2_contact_search:
	 * Contact3DT::SetActiveStrikers on the device) then maintains the pair list of a resident run */
	void SendSurfaces(void);

	/** the mesh the device object lives on */
	tb2_mesh* DeviceMesh(void) { EnsureDevice(); return fMesh; }

protected:

	virtual void RHSDriver(void);

private:

	void EnsureDevice(void);
	void SyncPairs(void);
	void Check(int status, const char* caller) const;

	tb2_mesh* fMesh;        /**< shared with a cuda_* continuum group, or own (coordinates only) */
	bool fOwnMesh;
	bool fMuted;
	tb2_contact* fContact;
	dArray2DT fForce;       /**< nodal contact forces of the last evaluation */
	std::vector<int> fPairsSent;   /**< the pair list the device holds */
	std::vector<double> fAreaSent;
};

} // namespace Tahoe
#endif
