"""Test-side reader/writer for the reference's input formats (test infrastructure).

* TahoeII ``.geom`` text meshes (reference: toolbox/src/dataio/input/TahoeInputT.cpp,
  database/ModelFileT.cpp; layout as benchmark_XML/level.0/geometry/cube.1.geom and the
  generator benchmark_XML/level.5/explicit_benchmark/generate_3d_mesh.py:17-127).
* the subset of the XML parameter tree the Hex8 hot path uses (tahoe.xsd: <time>, <nodes>/<field>
  with kinematic_BC / force_BC, one solid element group, one solver).

Used by tests/golden/make_golden.py to drive the real reference and by the parity tests to
feed the same case description to the oracle and to the CUDA path.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

# ----------------------------------------------------------------------------
# structured jittered cube
# ----------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """vectorised splitmix64 finaliser on uint64 arrays (same in tahoe_b200/host/mesh.hpp)"""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def jitter_u01(node_ids, dof, seed=12345):
    """uniform [0,1) from (seed, global node id, dof) -- partition independent"""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * np.uint64(0x100000001B3) + node_ids.astype(np.uint64) * np.uint64(3) + np.uint64(dof)
    return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def structured_cube(nx, ny=None, nz=None, jitter=0.1, seed=12345, lengths=(1.0, 1.0, 1.0)):
    """nx*ny*nz Hex8 on a box; node id = k(ny+1)(nx+1)+j(nx+1)+i, element order k-major, i-minor,
    connectivity in HexahedronT order (HexahedronT.cpp:23-25).  Interior nodes are moved by
    +-jitter*h.  Returns coords [nn,3], conn [ne,8] (0-based int32), nodesets {id: 0-based ids}:
    1: x=0, 2: x=L, 3: y=0, 4: y=L, 5: z=0, 6: z=L."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    px, py, pz = nx + 1, ny + 1, nz + 1
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    h = np.array([lengths[0] / nx, lengths[1] / ny, lengths[2] / nz])
    coords = np.stack([i * h[0], j * h[1], k * h[2]], axis=1).astype(np.float64)
    ids = np.arange(px * py * pz, dtype=np.int64)
    interior = (i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz)
    if jitter:
        for d in range(3):
            coords[:, d] += np.where(interior, (jitter_u01(ids, d, seed) - 0.5) * 2.0 * jitter * h[d], 0.0)
    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (ek * py * px + ej * px + ei).ravel()
    conn = np.stack([n0, n0 + 1, n0 + 1 + px, n0 + px,
                     n0 + px * py, n0 + 1 + px * py, n0 + 1 + px + px * py, n0 + px + px * py], axis=1).astype(np.int32)
    nodesets = {1: ids[i == 0], 2: ids[i == nx], 3: ids[j == 0], 4: ids[j == ny], 5: ids[k == 0], 6: ids[k == nz]}
    return coords, conn, {s: v.astype(np.int32) for s, v in nodesets.items()}


def cube_side_sets(nx, ny=None, nz=None):
    """side sets of structured_cube as {id: [[element, facet], ...]} (0-based; facets in HexahedronT::NodesOnFacet numbering,
    HexahedronT.cpp:1913-1918), ids as the node sets: 1: x=0, 2: x=L, 3: y=0, 4: y=L, 5: z=0, 6: z=L"""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    ek, ej, ei = [a.ravel() for a in np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")]
    e = np.arange(nx * ny * nz)
    pick = {1: (ei == 0, 5), 2: (ei == nx - 1, 3), 3: (ej == 0, 2), 4: (ej == ny - 1, 4), 5: (ek == 0, 0), 6: (ek == nz - 1, 1)}
    return {sid: np.stack([e[m], np.full(m.sum(), f)], axis=1).astype(np.int32) for sid, (m, f) in pick.items()}


def warp(coords):
    """smooth map that keeps the x = 0 face fixed and makes every other face of the unit cube curved (non-planar facets,
    non-constant surface Jacobians for the traction tests)"""
    x, y, z = coords[:, 0].copy(), coords[:, 1].copy(), coords[:, 2].copy()
    out = coords.copy()
    out[:, 0] += 0.08 * x * (y - 0.5) * (z + 0.3)
    out[:, 1] += 0.12 * x * z * (1.0 - 0.5 * y)
    out[:, 2] += 0.10 * x * x * y
    return out


# ----------------------------------------------------------------------------
# TahoeII .geom
# ----------------------------------------------------------------------------
def write_geom(path, coords, conn, nodesets, title="structured hex block", sidesets=None, block_sizes=None, sideset_blocks=None):
    """block_sizes: element counts of consecutive element blocks (ids 1, 2, ...); default one block.  sideset_blocks: {side set id:
    element block id} (default block 1); a side set's element numbers count within its block"""
    sidesets = sidesets or {}
    sideset_blocks = sideset_blocks or {}
    block_sizes = block_sizes or [conn.shape[0]]
    assert sum(block_sizes) == conn.shape[0]
    nn, ne = coords.shape[0], conn.shape[0]
    with open(path, "w") as f:
        f.write("*version\n1.0\n*title\n%s\n*dimensions\n" % title)
        f.write("%d  # number of nodes\n3  # number of spatial dimensions\n%d  # number of element sets\n" % (nn, len(block_sizes)))
        f.write("# [ID] [nel] [nen]\n" + "".join("%d %d 8\n" % (b + 1, nb) for b, nb in enumerate(block_sizes)))
        f.write("%d  # number of node sets\n# [ID] [nnd]\n" % len(nodesets))
        for sid in sorted(nodesets):
            f.write("%d %d\n" % (sid, len(nodesets[sid])))
        f.write("%d  # number of side sets\n" % len(sidesets))
        if sidesets:
            f.write("# [ID] [element set ID] [ns]\n")
        for sid in sorted(sidesets):
            f.write("%d %d %d\n" % (sid, sideset_blocks.get(sid, 1), len(sidesets[sid])))
        f.write("# end dimensions\n*nodesets\n")
        for sid in sorted(nodesets):
            ids = np.asarray(nodesets[sid]) + 1
            f.write("*set\n%d  # number of nodes\n" % len(ids))
            for s in range(0, len(ids), 10):
                f.write(" ".join(str(int(v)) for v in ids[s:s + 10]) + "\n")
        f.write("# end node sets\n*sidesets\n")
        for sid in sorted(sidesets):  # [element in its block] [facet], both 1-based
            f.write("*set\n%d  # number of sides\n" % len(sidesets[sid]))
            for e, fc in sidesets[sid]:
                f.write("%d %d\n" % (e + 1, fc + 1))
        f.write("*elements\n")
        e0 = 0
        for nb in block_sizes:
            f.write("*set\n%d  # number of elements\n8  # number of element nodes\n" % nb)
            for e in range(e0, e0 + nb):  # element ids restart in every block
                f.write("%d %s\n" % (e - e0 + 1, " ".join(str(int(v) + 1) for v in conn[e])))
            e0 += nb
        f.write("# end elements\n*nodes\n%d  # number of nodes\n3  # number of spatial dimensions\n" % nn)
        for n in range(nn):
            f.write("%d %.17e %.17e %.17e\n" % (n + 1, coords[n, 0], coords[n, 1], coords[n, 2]))


def _tokens(path):
    out = []
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if line:
                out.extend(line.split())
    return out


def read_geom(path):
    """returns coords, conn (0-based, element blocks concatenated in file order), nodesets {id: ids}"""
    base = os.path.dirname(path)
    tok = _tokens(path)
    p = tok.index("*dimensions") + 1
    nn, nsd, nblocks = int(tok[p]), int(tok[p + 1]), int(tok[p + 2])
    p += 3
    blocks = []
    for _ in range(nblocks):
        blocks.append((int(tok[p]), int(tok[p + 1]), int(tok[p + 2])))
        p += 3
    nns = int(tok[p]); p += 1
    ns_dims = []
    for _ in range(nns):
        ns_dims.append((int(tok[p]), int(tok[p + 1]))); p += 2
    p = tok.index("*nodesets") + 1
    nodesets = {}
    for sid, _cnt in ns_dims:
        assert tok[p] == "*set"; p += 1
        cnt = int(tok[p]); p += 1
        nodesets[sid] = np.array([int(v) - 1 for v in tok[p:p + cnt]], dtype=np.int32)
        p += cnt
    p = tok.index("*elements") + 1
    conns = []
    for _bid, nel, nen in blocks:
        assert tok[p] == "*set"; p += 1
        try:
            int(tok[p]); t, q = tok, p
        except ValueError:  # external file name
            t, q = _tokens(os.path.join(base, tok[p])), 0
            p += 1
        nel2, nen2 = int(t[q]), int(t[q + 1]); q += 2
        rows = np.array([int(v) for v in t[q:q + nel2 * (nen2 + 1)]], dtype=np.int64).reshape(nel2, nen2 + 1)
        conns.append(rows[:, 1:] - 1)
        if t is tok:
            p = q + nel2 * (nen2 + 1)
    p = tok.index("*nodes") + 1
    try:
        int(tok[p]); t, q = tok, p
    except ValueError:
        t, q = _tokens(os.path.join(base, tok[p])), 0
    nn2, nsd2 = int(t[q]), int(t[q + 1]); q += 2
    rows = np.array([float(v) for v in t[q:q + nn2 * (nsd2 + 1)]]).reshape(nn2, nsd2 + 1)
    coords = np.zeros((nn2, 3))
    coords[(rows[:, 0] - 1).astype(int)] = rows[:, 1:]
    return coords, np.concatenate(conns).astype(np.int32), nodesets


def read_sidesets(path):
    """side sets of a single-block .geom as {id: [[element, facet], ...]} 0-based (ModelManagerT::SideSet)"""
    tok = _tokens(path)
    p = tok.index("*dimensions") + 1
    nblocks = int(tok[p + 2])
    p += 3 + 3 * nblocks
    nns = int(tok[p]); p += 1 + 2 * nns
    nss = int(tok[p]); p += 1
    dims = []
    for _ in range(nss):
        dims.append((int(tok[p]), int(tok[p + 1]), int(tok[p + 2]))); p += 3
    p = tok.index("*sidesets") + 1
    out = {}
    for sid, _blk, _cnt in dims:
        assert tok[p] == "*set"; p += 1
        cnt = int(tok[p]); p += 1
        out[sid] = np.array([int(v) - 1 for v in tok[p:p + 2 * cnt]], dtype=np.int32).reshape(cnt, 2)
        p += 2 * cnt
    return out


# ----------------------------------------------------------------------------
# XML parameter tree (hot-path subset)
# ----------------------------------------------------------------------------
ELEMENT_TAGS = ("small_strain", "total_lagrangian", "updated_lagrangian", "explicit_solid")
MATERIAL_TAGS = ("small_strain_StVenant", "large_strain_StVenant", "Simo_isotropic", "Simo_J2", "RG_split_general")
SOLVER_TAGS = ("nonlinear_solver", "linear_solver", "PCG_solver")


def parse_xml(path):
    root = ET.parse(path).getroot()
    d = {"geometry_file": root.get("geometry_file")}
    t = root.find("time")
    sched = []
    for sf in t.findall("schedule_function"):
        pw = sf.find("piecewise_linear")
        sched.append([(float(o.get("x")), float(o.get("y"))) for o in pw.findall("OrderedPair")])
    d["time"] = {"num_steps": int(t.get("num_steps")), "time_step": float(t.get("time_step")), "schedules": sched}
    fld = root.find("nodes").find("field")
    d["integrator"] = fld.get("integrator", "static")
    d["kbc"] = [{"nodeset": int(k.get("node_ID")), "dof": int(k.get("dof")), "type": k.get("type", "fixed"),
                 "schedule": int(k.get("schedule", 0)), "value": float(k.get("value", 0.0))}
                for k in fld.findall("kinematic_BC")]
    d["fbc"] = [{"nodeset": int(k.get("node_ID")), "dof": int(k.get("dof")), "schedule": int(k.get("schedule", 0)),
                 "value": float(k.get("value", 0.0))} for k in fld.findall("force_BC")]
    el = None
    for tag in ELEMENT_TAGS:
        el = root.find("element_list").find(tag)
        if el is not None:
            break
    d["element"] = {"type": el.tag, "mass_type": el.get("mass_type", "automatic"),
                    "strain_displacement": el.get("strain_displacement", "standard"),
                    # ContinuumElementT::TakeNaturalBC (ContinuumElementT.cpp:1008-1095): 1 vector for the whole facet or one per facet node
                    "natural_bc": [{"side_set": int(nb.get("side_set_ID")), "schedule": int(nb.get("schedule")),
                                    "coordinate_system": nb.get("coordinate_system", "global"),
                                    "values": [[float(x.get("value")) for x in dl.findall("Double")] for dl in nb.findall("DoubleList")]}
                                   for nb in el.findall("natural_bc")]}
    mat = None
    for m in el.iter():
        if m.tag in MATERIAL_TAGS:
            mat = m
    md = {"type": mat.tag, "density": float(mat.get("density", 1.0))}
    en = mat.find("E_and_nu")
    if en is not None:
        md["E"], md["nu"] = float(en.get("Young_modulus")), float(en.get("Poisson_ratio"))
    bs = mat.find("bulk_and_shear")
    if bs is not None:
        md["kappa"], md["mu"] = float(bs.get("bulk_modulus")), float(bs.get("shear_modulus"))
    lf = mat.find("linear_function")
    if lf is not None:
        md["hardening"] = {"type": "linear_function", "a": float(lf.get("a")), "b": float(lf.get("b"))}
    le = mat.find("linear_exponential")
    if le is not None:
        md["hardening"] = {"type": "linear_exponential", **{k: float(le.get(k)) for k in "abcd"}}
    pl = mat.find("power_law")
    if pl is not None:
        md["hardening"] = {"type": "power_law", **{k: float(pl.get(k)) for k in "abcn"}}
    cs = mat.find("cubic_spline")
    if cs is not None:
        md["hardening"] = {"type": "cubic_spline", "fixity": cs.get("fixity", "parabolic"),
                           "points": [[float(o.get("x")), float(o.get("y"))] for o in cs.findall("OrderedPair")]}
    if mat.tag == "RG_split_general":  # explicit_solid: ExplicitElementT scans the sub-tree for mu / kappa (ExplicitElementT.cpp:176-200)
        nh = mat.find("rg_eq_potential").find("neo-hookean")
        md = {"type": "explicit_neo_hookean", "density": float(mat.get("density", 1.0)), "kappa": float(nh.get("kappa")), "mu": float(nh.get("mu"))}
        j2 = el.find("j2_plasticity")
        if j2 is not None:
            md["type"] = "explicit_J2"
            md["sigma_Y"], md["hardening_modulus"] = float(j2.get("sigma_Y")), float(j2.get("hardening"))
        ms = el.find("mass_scaling")
        if ms is not None:
            d["element"]["mass_scaling"] = {k: ms.get(k) for k in ("type", "target_dt", "scale_factor", "update_interval") if ms.get(k)}
    d["material"] = md
    for tag in SOLVER_TAGS:
        s = root.find(tag)
        if s is not None:
            d["solver"] = {"type": tag, **{k: v for k, v in s.attrib.items()}, "matrix": list(s)[0].tag if len(s) else None}
    return d


def write_xml(path, d):
    t = d["time"]
    top = "".join(' %s="%s"' % kv for kv in d.get("tahoe_attrs", {}).items())  # restart_file / restart_output_inc (FEManagerT.cpp:1412-1419)
    L = ['<?xml version="1.0"?>', '<tahoe geometry_file="%s"%s>' % (d["geometry_file"], top),
         '  <time num_steps="%d" output_inc="%d" time_step="%.17g">' % (t["num_steps"], d.get("output_inc", 0), t["time_step"])]
    for s in t["schedules"]:
        L.append("    <schedule_function><piecewise_linear>")
        L += ['      <OrderedPair x="%.17g" y="%.17g"/>' % p for p in s]
        L.append("    </piecewise_linear></schedule_function>")
    L.append("  </time>\n  <nodes>")
    integ = ' integrator="%s"' % d["integrator"] if d["integrator"] != "static" else ""
    L.append('    <field field_name="displacement"%s>' % integ)
    L.append('      <dof_labels><String value="D_X"/><String value="D_Y"/><String value="D_Z"/></dof_labels>')
    for k in d["kbc"]:
        if k["type"] == "fixed":
            L.append('      <kinematic_BC dof="%d" node_ID="%d"/>' % (k["dof"], k["nodeset"]))
        else:
            L.append('      <kinematic_BC dof="%d" node_ID="%d" schedule="%d" type="%s" value="%.17g"/>'
                     % (k["dof"], k["nodeset"], k["schedule"], k["type"], k["value"]))
    for k in d["fbc"]:
        L.append('      <force_BC dof="%d" node_ID="%d" schedule="%d" value="%.17g"/>'
                 % (k["dof"], k["nodeset"], k["schedule"], k["value"]))
    L.append("    </field>\n  </nodes>\n  <element_list>")
    e, m = d["element"], d["material"]
    mass = ' mass_type="%s"' % e["mass_type"] if e.get("mass_type", "automatic") != "automatic" else ""
    if e.get("strain_displacement", "standard") != "standard":  # SmallStrainT only (SmallStrainT.cpp:38-42)
        mass += ' strain_displacement="%s"' % e["strain_displacement"]
    L.append('    <%s field_name="displacement"%s>\n      <hexahedron/>' % (e.get("tag", e["type"]), mass))
    if e.get("body_force"):  # ContinuumElementT::NewSub("body_force"): schedule + one <Double> per direction
        bf = e["body_force"]
        L.append('      <body_force schedule="%d">%s</body_force>' % (bf["schedule"], "".join('<Double value="%.17g"/>' % v for v in bf["vector"])))
    for nb in e.get("natural_bc") or []:
        L.append('      <natural_bc schedule="%d" side_set_ID="%d" coordinate_system="%s">' % (nb["schedule"], nb["side_set"], nb["coordinate_system"]))
        for vec in nb["values"]:
            L.append("        <DoubleList>" + "".join('<Double value="%.17g"/>' % v for v in vec) + "</DoubleList>")
        L.append("      </natural_bc>")
    if e.get("nodal_output"):
        L.append('      <solid_element_nodal_output displacements="1"%s/>' % (' stress="1"' if e["nodal_output"] == "stress" else ""))
    small = e["type"] == "small_strain"
    blk = "small_strain_element_block" if small else "large_strain_element_block"
    mlist = "small_strain_material_3D" if small else "large_strain_material_3D"
    explicit = m["type"] in ("explicit_neo_hookean", "explicit_J2")
    if explicit:
        L.append('      <%s><block_ID_list><String value="1"/></block_ID_list>\n        <%s>' % (blk, mlist))
        L.append('          <RG_split_general density="%.17g"><rg_eq_potential><neo-hookean kappa="%.17g" mu="%.17g"/></rg_eq_potential>'
                 % (m["density"], m["kappa"], m["mu"]))
        L.append("          </RG_split_general>\n        </%s>\n      </%s>" % (mlist, blk))
        if m["type"] == "explicit_J2":
            L.append('      <j2_plasticity sigma_Y="%.17g" hardening="%.17g"/>' % (m["sigma_Y"], m["hardening_modulus"]))
        if e.get("mass_scaling"):
            L.append("      <mass_scaling %s/>" % " ".join('%s="%s"' % kv for kv in e["mass_scaling"].items()))
    # one element block per material: block i+1 of the .geom takes materials[i] (default: one block, d["material"])
    for ib, mi in enumerate([] if explicit else (d.get("materials") or [m])):
        L.append('      <%s><block_ID_list><String value="%d"/></block_ID_list>\n        <%s>' % (blk, ib + 1, mlist))
        L.append('          <%s density="%.17g">' % (mi["type"], mi["density"]))
        if "E" in mi:
            L.append('            <E_and_nu Poisson_ratio="%.17g" Young_modulus="%.17g"/>' % (mi["nu"], mi["E"]))
        else:
            L.append('            <bulk_and_shear bulk_modulus="%.17g" shear_modulus="%.17g"/>' % (mi["kappa"], mi["mu"]))
        h = mi.get("hardening")
        if h and h["type"] == "cubic_spline":
            L.append('            <cubic_spline fixity="%s">' % h["fixity"])
            L += ['              <OrderedPair x="%.17g" y="%.17g"/>' % (x, y) for x, y in h["points"]]
            L.append("            </cubic_spline>")
        elif h:
            L.append("            <%s %s/>" % (h["type"], " ".join('%s="%.17g"' % (k, v) for k, v in h.items() if k != "type")))
        L.append("          </%s>\n        </%s>\n      </%s>" % (mi["type"], mlist, blk))
    L.append("    </%s>\n  </element_list>" % e.get("tag", e["type"]))
    s = d["solver"]
    attrs = " ".join('%s="%s"' % (k, v) for k, v in s.items() if k not in ("type", "matrix", "matrix_attrs"))
    L.append("  <%s %s><%s %s/></%s>\n</tahoe>" % (s["type"], attrs, s["matrix"], s.get("matrix_attrs", ""), s["type"]))
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")


def schedule_value(sched, t):
    """piecewise linear schedule (toolbox C1functions PiecewiseLinearT): constant outside the range"""
    xs = [p[0] for p in sched]
    ys = [p[1] for p in sched]
    return float(np.interp(t, xs, ys))


def traction_cards(desc, sidesets):
    """natural_bc lists -> [(elem, facet, tract[ncards,4,3], coordinate_system, schedule index)] as ContinuumElementT::TakeNaturalBC
    builds the cards (one per side, the same nodal vectors for every side of a set)"""
    out = []
    for nb in desc["element"].get("natural_bc") or []:
        sides = sidesets[nb["side_set"]]
        vals = np.asarray(nb["values"], np.float64)
        nodal = np.broadcast_to(vals, (4, 3)) if vals.shape[0] == 1 else vals
        out.append((sides[:, 0].copy(), sides[:, 1].copy(), np.broadcast_to(nodal, (len(sides), 4, 3)).copy(),
                    nb["coordinate_system"], nb["schedule"] - 1))
    return out


def bc_arrays(desc, nodesets, nn, t):
    """(code[nn,3] uint8: 0 free / 1 fixed / 2 prescribed-u, value[nn,3], fext[nn,3]) at time t (nodal forces only; tractions are
    added by cases.Case.bc through the implementation under test)"""
    code = np.zeros((nn, 3), np.uint8)
    val = np.zeros((nn, 3))
    fext = np.zeros((nn, 3))
    sch = desc["time"]["schedules"]
    for k in desc["kbc"]:
        ids = nodesets[k["nodeset"]]
        if k["type"] == "fixed":
            code[ids, k["dof"] - 1] = 1
        elif k["type"] == "u":
            code[ids, k["dof"] - 1] = 2
            val[ids, k["dof"] - 1] = k["value"] * schedule_value(sch[k["schedule"] - 1], t)
        else:
            raise NotImplementedError(k["type"])
    for k in desc["fbc"]:
        ids = nodesets[k["nodeset"]]
        fext[ids, k["dof"] - 1] += k["value"] * schedule_value(sch[k["schedule"] - 1], t)
    return code, val, fext


def read_nodal_table(run_file):
    """last 'Nodal data' table of a Tahoe text .run file (TextOutputT, 12 digits after the point) -> (labels, array [nn][nvalues])"""
    import glob
    import re
    parts = sorted(glob.glob(run_file + ".ps*"))  # one file per print step next to the table of contents
    text = open(parts[-1] if parts else run_file).read()
    block = text[text.rindex("Nodal data:"):]
    rows, labels = [], []
    for line in block.splitlines():
        f = line.split()
        if len(f) >= 3 and f[0] == "index" and f[1] == "node":
            labels = f[2:]
        elif len(f) >= 5 and re.match(r"^\d+$", f[0]) and re.match(r"^\d+$", f[1]):
            rows.append([float(x) for x in f[2:]])
        elif rows and not f:
            break
    return labels, np.array(rows)
