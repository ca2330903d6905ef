/* FastGeomInputT.cpp -- see FastGeomInputT.h */
#include "FastGeomInputT.h"

#include "ExceptionT.h"
#include "dArray2DT.h"
#include "iArray2DT.h"
#include "iArrayT.h"

#include <cstdlib>
#include <vector>

using namespace Tahoe;

FastGeomInputT::FastGeomInputT(ostream& out):
	TahoeInputT(out),
	fGeom(NULL),
	fLog(out)
{
}

FastGeomInputT::~FastGeomInputT(void)
{
	if (fGeom) tb2_geom_close(fGeom);
}

bool FastGeomInputT::Open(const StringT& filename)
{
	/* header, names and dimensions: the reference's own ModelFileT */
	if (!TahoeInputT::Open(filename)) return false;
	if (fGeom) tb2_geom_close(fGeom);
	fGeom = NULL;
	fBlock.clear();
	fNodeSet.clear();
	fSideSet.clear();
	/* the bulk sections, once, with all host threads */
	if (tb2_geom_open(filename.Pointer(), &fGeom) != TB2_OK) {
		fLog << "\n FastGeomInputT::Open: " << tb2_last_error() << ": reading through TahoeInputT\n";
		fGeom = NULL;
		return true; /* the inherited reader still serves everything */
	}
	int64_t nn = 0;
	int32_t nsd = 0, nb = 0, nns = 0, nss = 0;
	tb2_geom_sizes(fGeom, &nn, &nsd, &nb, &nns, &nss);
	for (int b = 0; b < nb; b++) {
		int32_t id = 0, nen = 0;
		int64_t nel = 0;
		tb2_geom_block(fGeom, b, &id, &nel, &nen, NULL);
		fBlock[id] = b;
	}
	for (int s = 0; s < nns; s++) {
		int32_t id = 0;
		int64_t n = 0;
		tb2_geom_nodeset(fGeom, s, &id, &n, NULL);
		fNodeSet[id] = s;
	}
	for (int s = 0; s < nss; s++) {
		int32_t id = 0, block = 0;
		int64_t n = 0;
		tb2_geom_sideset(fGeom, s, &id, &block, &n, NULL);
		fSideSet[id] = s;
	}
	if (nn != NumNodes() || nsd != NumDimensions() || nb != NumElementGroups() || nns != NumNodeSets() || nss != NumSideSets())
		ExceptionT::DatabaseFail("FastGeomInputT::Open", "the two readers disagree on the dimensions of %s", filename.Pointer());
	fLog << " FastGeomInputT: " << nn << " nodes, " << nb << " element sets, " << nns << " node sets, " << nss
	     << " side sets parsed by tb2_geom_open\n";
	return true;
}

void FastGeomInputT::Close(void)
{
	if (fGeom) tb2_geom_close(fGeom);
	fGeom = NULL;
	TahoeInputT::Close();
}

int FastGeomInputT::Find(const std::map<int, int>& index, const StringT& name, const char* what) const
{
	std::map<int, int>::const_iterator it = index.find(atoi(name.Pointer()));
	if (it == index.end()) ExceptionT::DatabaseFail("FastGeomInputT", "%s \"%s\" not found", what, name.Pointer());
	return it->second;
}

void FastGeomInputT::ReadCoordinates(dArray2DT& coords)
{
	const char caller[] = "FastGeomInputT::ReadCoordinates";
	if (!fGeom) { TahoeInputT::ReadCoordinates(coords); return; }
	const int nn = NumNodes(), nsd = NumDimensions();
	if (coords.MajorDim() != nn || coords.MinorDim() != nsd) ExceptionT::SizeMismatch(caller);
	if (nsd == 3) {
		if (tb2_geom_coords(fGeom, coords.Pointer()) != TB2_OK) ExceptionT::DatabaseFail(caller, "%s", tb2_last_error());
	} else { /* the library pads to three columns */
		std::vector<double> X((size_t)nn * 3);
		if (tb2_geom_coords(fGeom, &X[0]) != TB2_OK) ExceptionT::DatabaseFail(caller, "%s", tb2_last_error());
		for (int n = 0; n < nn; n++)
			for (int i = 0; i < nsd; i++) coords(n, i) = X[(size_t)n * 3 + i];
	}
}

void FastGeomInputT::ReadCoordinates(dArray2DT& coords, iArrayT& node_id)
{
	ReadCoordinates(coords);
	ReadNodeID(node_id);
}

void FastGeomInputT::ReadConnectivity(const StringT& name, iArray2DT& connects)
{
	const char caller[] = "FastGeomInputT::ReadConnectivity";
	if (!fGeom) { TahoeInputT::ReadConnectivity(name, connects); return; }
	const int b = Find(fBlock, name, "element set");
	int32_t id = 0, nen = 0;
	int64_t nel = 0;
	tb2_geom_block(fGeom, b, &id, &nel, &nen, NULL);
	if (connects.MajorDim() != nel || connects.MinorDim() != nen) ExceptionT::SizeMismatch(caller);
	/* 0-based already (TahoeInputT: connects += -1) */
	if (tb2_geom_block(fGeom, b, &id, &nel, &nen, connects.Pointer()) != TB2_OK) ExceptionT::DatabaseFail(caller, "%s", tb2_last_error());
}

void FastGeomInputT::ReadNodeSet(const StringT& name, iArrayT& nodes)
{
	const char caller[] = "FastGeomInputT::ReadNodeSet";
	if (!fGeom) { TahoeInputT::ReadNodeSet(name, nodes); return; }
	const int s = Find(fNodeSet, name, "node set");
	int32_t id = 0;
	int64_t n = 0;
	tb2_geom_nodeset(fGeom, s, &id, &n, NULL);
	if (nodes.Length() != n) ExceptionT::SizeMismatch(caller);
	if (n == 0) return;
	if (tb2_geom_nodeset(fGeom, s, &id, &n, nodes.Pointer()) != TB2_OK) ExceptionT::DatabaseFail(caller, "%s", tb2_last_error());
	/* a negative entry is the files' "all model nodes" marker: leave its interpretation to the reference's reader */
	for (int k = 0; k < nodes.Length(); k++)
		if (nodes[k] < 0) { TahoeInputT::ReadNodeSet(name, nodes); return; }
}

StringT FastGeomInputT::SideSetGroupName(const StringT& name) const
{
	if (!fGeom) return TahoeInputT::SideSetGroupName(name);
	const int s = Find(fSideSet, name, "side set");
	int32_t id = 0, block = 0;
	int64_t n = 0;
	tb2_geom_sideset(fGeom, s, &id, &block, &n, NULL);
	StringT elname;
	elname.Append(block);
	return elname;
}

void FastGeomInputT::ReadSideSetLocal(const StringT& name, iArray2DT& sides) const
{
	const char caller[] = "FastGeomInputT::ReadSideSetLocal";
	if (!fGeom) { TahoeInputT::ReadSideSetLocal(name, sides); return; }
	const int s = Find(fSideSet, name, "side set");
	int32_t id = 0, block = 0;
	int64_t n = 0;
	tb2_geom_sideset(fGeom, s, &id, &block, &n, NULL);
	if (sides.MajorDim() != n || (n > 0 && sides.MinorDim() != 2)) ExceptionT::SizeMismatch(caller);
	if (n == 0) return;
	/* (element in its block, facet), 0-based already (TahoeInputT: sides += -1) */
	if (tb2_geom_sideset(fGeom, s, &id, &block, &n, sides.Pointer()) != TB2_OK) ExceptionT::DatabaseFail(caller, "%s", tb2_last_error());
}
