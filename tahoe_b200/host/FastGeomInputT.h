/* FastGeomInputT.h -- TahoeII .geom input behind ModelManagerT through the library's threaded reader.
 *
 * ModelManagerT reads geometry through an InputBaseT (IOBaseT::NewInput, IOBaseT.cpp:170-215); for the TahoeII text format that is
 * TahoeInputT over ModelFileT, which parses the bulk sections token by token with formatted stream reads (ModelFileT.cpp) -- minutes
 * of host time on a mesh of tens of millions of elements.  This class keeps TahoeInputT for the header and every query about names
 * and dimensions, and serves the bulk arrays -- coordinates, connectivities, node sets, side sets -- from tb2_geom_open, which reads
 * the file in one piece and converts it with all host threads (tb2_geom.cu).  Registration: one line in IOBaseT::NewInput
 * (registration.patch).  The arrays are the ones TahoeInputT returns (tests/test_geom_reader.py compares them on the reference's own
 * geometry files; tests/test_plugin_binary.py runs host-only analyses through the plugin executable against the reference).
 */
#ifndef _FAST_GEOM_INPUT_T_H_
#define _FAST_GEOM_INPUT_T_H_

#include "TahoeInputT.h"

#include <map>

#include "tahoe_b200.h"

namespace Tahoe {

class FastGeomInputT: public TahoeInputT
{
public:

	FastGeomInputT(ostream& out);
	virtual ~FastGeomInputT(void);

	virtual bool Open(const StringT& filename);
	virtual void Close(void);

	virtual void ReadCoordinates(dArray2DT& coords);
	virtual void ReadCoordinates(dArray2DT& coords, iArrayT& node_id);
	virtual void ReadConnectivity(const StringT& name, iArray2DT& connects);
	virtual void ReadNodeSet(const StringT& name, iArrayT& nodes);
	virtual StringT SideSetGroupName(const StringT& name) const;
	virtual void ReadSideSetLocal(const StringT& name, iArray2DT& sides) const;

private:

	int Find(const std::map<int, int>& index, const StringT& name, const char* what) const;

	tb2_geom* fGeom;
	std::map<int, int> fBlock, fNodeSet, fSideSet; /**< id in the file -> position */
	ostream& fLog;
};

} // namespace Tahoe
#endif
