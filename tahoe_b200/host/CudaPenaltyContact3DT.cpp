/* CudaPenaltyContact3DT.cpp -- see CudaPenaltyContact3DT.h */
#include "CudaPenaltyContact3DT.h"

#include "CudaSolidElementT.h"
#include "ElementSupportT.h"
#include "ExceptionT.h"
#include "FEManagerT.h"
#include "FieldT.h"
#include "ParameterListT.h"
#include "eIntegratorT.h"
#include "iArray2DT.h"

#include <cstring>

using namespace Tahoe;

CudaPenaltyContact3DT::CudaPenaltyContact3DT(const ElementSupportT& support, const char* name):
	PenaltyContact3DT(support),
	fMesh(NULL),
	fOwnMesh(false),
	fMuted(false),
	fContact(NULL)
{
	SetName(name);
}

CudaPenaltyContact3DT::~CudaPenaltyContact3DT(void)
{
	if (fContact) tb2_contact_destroy(fContact);
	if (fOwnMesh && fMesh) tb2_mesh_destroy(fMesh);
}

void CudaPenaltyContact3DT::Check(int status, const char* caller) const
{
	if (status != TB2_OK) ExceptionT::GeneralFail(caller, "%s", tb2_last_error());
}

void CudaPenaltyContact3DT::TakeParameterList(const ParameterListT& list)
{
	/* inherited: surfaces, strikers, penalty stiffness, friction, damping -- all Tahoe's own */
	PenaltyContact3DT::TakeParameterList(list);
	if (fMu > 0.0 && fImplicitFriction)
		ExceptionT::BadInputValue("CudaPenaltyContact3DT::TakeParameterList",
			"slip-based friction of a static analysis has no device form: use contact_3D_penalty for this group");
}

/* the device objects: made on first use, when the other element groups exist */
void CudaPenaltyContact3DT::EnsureDevice(void)
{
	const char caller[] = "CudaPenaltyContact3DT::EnsureDevice";
	if (fContact) return;
	const FEManagerT& fe = ElementSupport().FEManager();
	for (int i = 0; i < fe.NumElementGroups() && !fMesh; i++) {
		CudaStiffnessSourceT* dev = dynamic_cast<CudaStiffnessSourceT*>(fe.ElementGroup(i));
		if (dev && dev->DeviceMesh()) fMesh = dev->DeviceMesh();
	}
	if (!fMesh) {
		/* coordinates only: the contact force reads X, u, v by node; one placeholder element carries the array to the device */
		const dArray2DT& X = ElementSupport().InitialCoordinates();
		if (X.MajorDim() < 8) ExceptionT::GeneralFail(caller, "fewer than 8 nodes");
		int32_t conn[8] = {0, 1, 2, 3, 4, 5, 6, 7};
		Check(tb2_mesh_create(0, X.MajorDim(), 1, conn, X.Pointer(), &fMesh), caller);
		fOwnMesh = true;
	}
	Check(tb2_contact_create(fMesh, fK, fMu, fFrictionEps, fViscousDamping, &fContact), caller);
	fForce.Dimension(ElementSupport().InitialCoordinates().MajorDim(), NumDOF());
}

/* the active pairs as the search left them: rows of fConnectivities[0] (three facet nodes, then the striker) with the striker's
 * area (ContactT::fStrikerArea); sent again only when they changed */
void CudaPenaltyContact3DT::SyncPairs(void)
{
	const iArray2DT& pairs = *fConnectivities[0];
	const int np = pairs.MajorDim();
	if (np > 0 && pairs.MinorDim() != 4) ExceptionT::SizeMismatch("CudaPenaltyContact3DT::SyncPairs", "expecting 4 nodes per pair");
	std::vector<double> area((size_t)np);
	for (int i = 0; i < np; i++) area[i] = fStrikerArea[fStrikerTags_map.Map(pairs(i, 3))];
	const size_t n4 = (size_t)np * 4;
	const bool same = fPairsSent.size() == n4 && fAreaSent == area && (n4 == 0 || memcmp(&fPairsSent[0], pairs.Pointer(), n4 * sizeof(int)) == 0);
	if (same) return;
	fPairsSent.assign(pairs.Pointer(), pairs.Pointer() + n4);
	fAreaSent = area;
	Check(tb2_contact_set_pairs(fContact, np, np ? &fPairsSent[0] : NULL, np ? &fAreaSent[0] : NULL), "CudaPenaltyContact3DT::SyncPairs");
}

void CudaPenaltyContact3DT::SendSurfaces(void)
{
	const char caller[] = "CudaPenaltyContact3DT::SendSurfaces";
	EnsureDevice();
	std::vector<int32_t> facets, surface;
	for (int s = 0; s < fSurfaces.Length(); s++) {
		if (fSurfaces[s].MajorDim() > 0 && fSurfaces[s].MinorDim() != 3) ExceptionT::SizeMismatch(caller, "expecting triangular facets");
		for (int f = 0; f < fSurfaces[s].MajorDim(); f++) {
			for (int a = 0; a < 3; a++) facets.push_back(fSurfaces[s](f, a));
			surface.push_back(s);
		}
	}
	if (fStrikerTags.Length() == 0) ExceptionT::GeneralFail(caller, "the group has no striker list (all-nodes strikers are not supported on the device)");
	Check(tb2_contact_set_surfaces(fContact, (int64_t)surface.size(), facets.empty() ? NULL : &facets[0], surface.empty() ? NULL : &surface[0],
		fStrikerTags.Length(), fStrikerTags.Pointer(), fStrikerArea.Pointer()), caller);
}

tb2_contact* CudaPenaltyContact3DT::DeviceContact(void)
{
	EnsureDevice();
	SyncPairs();
	return fContact;
}

void CudaPenaltyContact3DT::RHSDriver(void)
{
	const char caller[] = "CudaPenaltyContact3DT::RHSDriver";
	double constKd = 0.0;
	int formKd = fIntegrator->FormKd(constKd);
	if (!formKd || fMuted) return;
	EnsureDevice();
	SyncPairs();

	/* the whole pair loop on the device; the per-striker normal forces of the log (fStrikerForce2D) are not formed */
	const FieldT& field = Field();
	fStrikerForce2D = 0.0;
	const double* vel = field.Order() >= 1 ? field[1].Pointer() : NULL;
	Check(tb2_contact_form_host(fContact, constKd, field[0].Pointer(), vel, 0, fForce.Pointer()), caller);
	ElementSupport().AssembleRHS(Group(), fForce, field.Equations());

	int num_contact = 0;
	double h_max = 0.0;
	Check(tb2_contact_tracking(fContact, &num_contact, &h_max), caller);
	SetTrackingData(num_contact, h_max);
}
