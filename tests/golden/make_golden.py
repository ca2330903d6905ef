#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and oracle/_ref/tahoe_dump, built by
`make -f oracle/build_ref.mk`).  Two kinds of cases:

* the reference's own regression inputs for the hot path (benchmark_XML level.0 / level.1,
  SURVEY.md section 8c), run as shipped;
* synthetic jittered cubes written as TahoeII .geom + XML, so that Jacobians
  are non-trivial and every formulation x material pair of the hot path is pinned, including the
  assembled tangent and MSR structure (`SPOOLES_matrix` is MSRMatrixT-derived and does not renumber).

Each fixture holds the case description (JSON), the mesh, node sets, and the reference's in-memory
arrays at full precision (oracle/ref_dump.cpp).  The fixtures travel; the reference does not.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "tests"))
import tahoe_input as ti  # noqa: E402

DUMP = os.path.join(REPO, "oracle", "_ref", "tahoe_dump")
REF = "/root/reference/benchmark_XML"


def load_dump(d):
    out = {}
    for line in open(os.path.join(d, "manifest.txt")):
        name, typ, d0, d1 = line.split()
        a = np.fromfile(os.path.join(d, name + ".bin"), dtype=np.float64 if typ == "f8" else np.int32)
        out[name] = a.reshape(int(d0), int(d1)) if int(d1) else a
    return out


def run_case(name, xml_path, flags, desc, coords, conn, nodesets, sidesets=None):
    out = tempfile.mkdtemp(prefix="dump_")
    r = subprocess.run([DUMP, os.path.basename(xml_path), out] + flags, cwd=os.path.dirname(xml_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout[-3000:])
        raise RuntimeError("tahoe_dump failed for " + name)
    dump = load_dump(out)
    shutil.rmtree(out)
    assert np.array_equal(dump["conn"], conn), "connectivity mismatch vs geom reader"
    assert np.allclose(dump["coords"], coords, rtol=0, atol=0), "coords mismatch vs geom reader"
    payload = {"desc": np.array(json.dumps(desc))}
    for sid, ids in nodesets.items():
        payload["ns_%d" % sid] = np.asarray(ids, np.int32)
    used = {nb["side_set"] for nb in desc["element"].get("natural_bc") or []}
    for sid, sides in (sidesets or {}).items():
        if sid in used:
            payload["ss_%d" % sid] = np.asarray(sides, np.int32)
    for k, v in dump.items():
        payload["ref_" + k] = v
    if desc["element"].get("nodal_output") == "stress":
        # SolidElementT::ComputeOutput (SolidElementT.cpp:1352-1840): Cauchy stress extrapolated to the nodes and averaged, as the
        # reference's own TextOutputT wrote it during the run (13 significant digits)
        labels, table = ti.read_nodal_table(os.path.splitext(xml_path)[0] + ".io0.run")
        assert labels[-6:] == ["s11", "s22", "s33", "s23", "s13", "s12"], labels
        payload["ref_nodal_stress"] = table[:, -6:]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print("%-28s %7.1f kB  %s" % (name, os.path.getsize(path) / 1024, " ".join(sorted(dump))))


def reference_case(name, rel_xml, flags):
    """one of the reference's own inputs, run as shipped from a scratch copy of its directory tree"""
    src_dir = os.path.dirname(os.path.join(REF, rel_xml))
    work = tempfile.mkdtemp(prefix="ref_")
    level_dir = os.path.dirname(src_dir)
    shutil.copytree(level_dir, os.path.join(work, "lvl"), ignore=shutil.ignore_patterns("benchmark", "*.run", "*.out"))
    xml = os.path.join(work, "lvl", os.path.basename(src_dir), os.path.basename(rel_xml))
    desc = ti.parse_xml(xml)
    desc["source"] = "benchmark_XML/" + rel_xml
    geom = os.path.normpath(os.path.join(os.path.dirname(xml), desc["geometry_file"]))
    coords, conn, nodesets = ti.read_geom(geom)
    run_case(name, xml, flags, desc, coords, conn, nodesets, ti.read_sidesets(geom))
    shutil.rmtree(work)


def pcg_history(name, rel_xml):
    """convergence history of a PCG_solver run: the reference executable on a scratch copy of its own input with
    output_flag="all_iterations" added (PCGSolver_LS.cpp:59-64), `Relative error` lines of NLSolver::ExitIteration"""
    src_dir = os.path.dirname(os.path.join(REF, rel_xml))
    work = tempfile.mkdtemp(prefix="hist_")
    shutil.copytree(os.path.dirname(src_dir), os.path.join(work, "lvl"), ignore=shutil.ignore_patterns("benchmark", "*.run", "*.out"))
    xml = os.path.join(work, "lvl", os.path.basename(src_dir), os.path.basename(rel_xml))
    text = open(xml).read().replace("<PCG_solver ", '<PCG_solver output_flag="all_iterations" ')
    open(xml, "w").write(text)
    r = subprocess.run([os.path.join(REPO, "oracle", "_ref", "tahoe"), "-f", os.path.basename(xml)], cwd=os.path.dirname(xml),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    hist = [ln.split("=")[-1].split()[0] for ln in r.stdout.splitlines() if "Relative error =" in ln]
    with open(os.path.join(HERE, name + "_history.txt"), "w") as f:
        f.write("# relative error |R|/|R0| after iteration 0, 1, ... of the reference's run of benchmark_XML/%s\n" % rel_xml)
        f.write("\n".join(hist) + "\n")
    shutil.rmtree(work)
    print("%-28s %d iterations" % (name + "_history", len(hist)))


def synthetic_case(name, n, desc, flags, jitter=0.1):
    work = tempfile.mkdtemp(prefix="syn_")
    dims = n if isinstance(n, tuple) else (n, n, n)
    coords, conn, nodesets = ti.structured_cube(*dims, jitter=jitter, seed=12345)
    sidesets = None
    if desc["element"].get("natural_bc"):  # traction cases: curved faces, side sets in the .geom
        coords = ti.warp(coords)
        sidesets = ti.cube_side_sets(*dims)
    ti.write_geom(os.path.join(work, "mesh.geom"), coords, conn, nodesets, sidesets=sidesets)
    coords, conn2, nodesets2 = ti.read_geom(os.path.join(work, "mesh.geom"))  # what the reference will read
    assert np.array_equal(conn, conn2)
    desc = dict(desc, geometry_file="mesh.geom", source="synthetic %dx%dx%d jitter %g seed 12345" % (*dims, jitter))
    xml = os.path.join(work, name + ".xml")
    ti.write_xml(xml, desc)
    if sidesets:
        back = ti.read_sidesets(os.path.join(work, "mesh.geom"))
        assert all(np.array_equal(back[k], sidesets[k]) for k in sidesets)
        desc["source"] += " warped"
    run_case(name, xml, flags, desc, coords, conn, nodesets2, sidesets)
    shutil.rmtree(work)


def contact_case(name, rel_xml, flags, edits=()):
    """SURVEY 8(f)-4, contact_3D_penalty (PenaltyContact3DT): one of the reference's own contact inputs, optionally with text edits
    (attributes added to the contact tag, fewer steps), run through tahoe_dump --contact.  The fixture holds what the force needs --
    coordinates, equation numbers, d (and v) at the dumped steps, the active striker-facet pairs with the striker areas as the
    reference's own search left them, the parameters -- and the contact group's own FormRHS at those states."""
    src_dir = os.path.dirname(os.path.join(REF, rel_xml))
    work = tempfile.mkdtemp(prefix="contact_")
    shutil.copytree(os.path.dirname(src_dir), os.path.join(work, "lvl"), ignore=shutil.ignore_patterns("benchmark", "*.run", "*.out", "*.exo"))
    xml = os.path.join(work, "lvl", os.path.basename(src_dir), os.path.basename(rel_xml))
    text = open(xml).read()
    for old, new in edits:
        assert old in text, old
        text = text.replace(old, new)
    open(xml, "w").write(text)
    out = tempfile.mkdtemp(prefix="dump_")
    r = subprocess.run([DUMP, os.path.basename(xml), out] + flags + ["--contact"], cwd=os.path.dirname(xml), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout[-3000:])
        raise RuntimeError("tahoe_dump failed for " + name)
    dump = load_dump(out)
    shutil.rmtree(out)
    shutil.rmtree(work)
    steps = sorted(int(k.split("_")[1]) for k in dump if k.startswith("cpairs_"))
    payload = {"source": np.array("benchmark_XML/" + rel_xml + ("; edits: " + json.dumps(edits) if edits else "")),
               "steps": np.asarray(steps, np.int32)}
    keep = ("coords", "eqnos", "cparams", "cfacets", "cfacet_surface", "cstrikers", "cstriker_area") + tuple("%s_%d" % (a, k) for k in steps for a in ("d", "v", "cpairs", "carea", "crhs"))
    for k in keep:
        if k in dump:
            payload["ref_" + k] = dump[k]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    npairs = [int(dump["cpairs_%d" % k].shape[0]) for k in steps]
    print("%-28s %7.1f kB  steps %s pairs %s" % (name, os.path.getsize(path) / 1024, steps, npairs))


RAMP = [(0.0, 0.0), (1.0, 1.0)]
RAMP_FAST = [(0.0, 0.0), (0.4, 1.0), (10.0, 1.0)]
CLAMP_X0 = [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)]
NEWTON = {"type": "nonlinear_solver", "abs_tolerance": "1.0e-12", "rel_tolerance": "1.0e-12",
          "divergence_tolerance": "1.0e+03", "max_iterations": "25", "matrix": "SPOOLES_matrix"}
EXPLICIT = {"type": "linear_solver", "matrix": "diagonal_matrix"}
# PCGSolver_LS with its DiagonalMatrixT preconditioner: the attributes of level.0/3D.elastostatic/beam.PCG.xml
PCG = {"type": "PCG_solver", "abs_tolerance": "1.0e-12", "divergence_tolerance": "10.0", "line_search_iterations": "10",
       "line_search_tolerance": "0.1", "max_iterations": "2000", "max_step": "2.5", "quick_solve_iter": "100",
       "rel_tolerance": "1.0e-10", "restart": "30", "matrix": "diagonal_matrix"}


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    # ---- the reference's own regression inputs ----
    ref_cases = [
        ("ref_traction_a", "level.0/3D.elastostatic/traction.a.xml", ["--fint"]),
        ("ref_beam_newton_totlag", "level.0/3D.elastostatic/beam.Newton.TotLag.xml", ["--fint"]),
        ("ref_beam_newton", "level.0/3D.elastostatic/beam.Newton.xml", ["--fint"]),
        ("ref_explicit_1", "level.0/3D.elastodynamic/explicit.1.xml", ["--every", "25", "--fint"]),
        ("ref_explicit_2", "level.0/3D.elastodynamic/explicit.2.xml", ["--every", "25", "--fint"]),
        ("ref_implicit_1", "level.0/3D.elastodynamic/implicit.1.xml", ["--every", "1", "--fint"]),  # nonlinear_HHT + consistent_mass
        ("ref_beam_pcg", "level.0/3D.elastostatic/beam.PCG.xml", ["--fint"]),
        ("ref_beam_bbar", "level.0/3D.elastostatic/beam.B_bar.xml", ["--fint"]),
        ("ref_mat_1_a", "level.1/material.solid/3D/material.01/mat.1.a.xml", ["--fint"]),
        ("ref_mat_2_a", "level.1/material.solid/3D/material.02/mat.2.a.xml", ["--fint"]),
        ("ref_mat_5_a", "level.1/material.solid/3D/material.05/mat.5.a.xml", ["--fint"]),
        ("ref_mat_09_a", "level.1/material.solid/3D/material.09/mat.09.a.xml", ["--every", "1", "--fint"]),
        ("ref_mat_09_b", "level.1/material.solid/3D/material.09/mat.09.b.xml", ["--every", "1", "--fint"]),  # linear_exponential K(alpha): local Newton
        ("ref_mat_09_c", "level.1/material.solid/3D/material.09/mat.09.c.xml", ["--every", "1", "--fint"]),  # cubic_spline K(alpha)
        ("ref_mat_09_d", "level.1/material.solid/3D/material.09/mat.09.d.xml", ["--every", "1", "--fint"]),  # power_law K(alpha)
    ]
    for name, rel, flags in ref_cases:
        if want(name):
            reference_case(name, rel, flags)
    if want("ref_beam_pcg"):
        pcg_history("ref_beam_pcg", "level.0/3D.elastostatic/beam.PCG.xml")
    # contact_3D_penalty: the reference's static two-cube case as shipped, and its explicit sliding-friction case cut to 1200 steps with
    # viscous damping switched on as well (all three force terms of PenaltyContact3DT::RHSDriver active, velocity-based friction)
    if want("ref_contact_cubes_1"):
        contact_case("ref_contact_cubes_1", "level.2/contact_simple/cubes.1.xml", ["--every", "1"])
    # more of the reference's own contact inputs, for the oracle's force and search (CPU tests): a second static two-cube case, the
    # quasi-static sliding case (30 steps, pairs change from step to step), the damped impact and the Hertz sphere cut short
    if want("ref_contact_cubes_2"):
        contact_case("ref_contact_cubes_2", "level.2/contact_simple/cubes.2.xml", ["--every", "1"])
    if want("ref_contact_sliding_3d"):
        contact_case("ref_contact_sliding_3d", "level.2/contact_simple/sliding.3D.xml", ["--every", "5"])
    if want("ref_contact_impact_damped"):
        contact_case("ref_contact_impact_damped", "level.5/explicit_benchmark/vectorized_cubes_impact_damped.xml", ["--every", "250"],
                     edits=(('output_format="ExodusII"', ""), ('num_steps="5000"', 'num_steps="1500"')))
    if want("ref_contact_hertz_explicit"):
        contact_case("ref_contact_hertz_explicit", "level.5/hertz/hertz_explicit.xml", ["--every", "400"],
                     edits=(('output_format="ExodusII"', ""), ('num_steps="10000"', 'num_steps="1200"')))
    if want("ref_contact_sliding_friction"):
        contact_case("ref_contact_sliding_friction", "level.5/explicit_benchmark/vectorized_cubes_friction.xml", ["--every", "300"],
                     edits=(('output_format="ExodusII"', ""), ('num_steps="5000"', 'num_steps="1200"'),
                            ('friction_epsilon_velocity="0.001">', 'friction_epsilon_velocity="0.001" viscous_damping="20.0">')))

    # ---- synthetic jittered cubes ----
    kstv = {"type": "small_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25}
    fdkstv = {"type": "large_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25}
    simo = {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 5.0}
    simo_soft = {"type": "Simo_isotropic", "density": 1.0, "E": 100.0, "nu": 0.25}
    j2 = {"type": "Simo_J2", "density": 1.0, "E": 100.0, "nu": 0.25,
          "hardening": {"type": "linear_function", "a": 0.05, "b": 0.25}}
    pull_f = [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}, {"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.005}]

    def pull_u(v):
        return CLAMP_X0 + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": v},
                           {"nodeset": 2, "dof": 3, "type": "u", "schedule": 1, "value": 0.3 * v}]

    def static(nsteps):
        return {"num_steps": nsteps, "time_step": 1.0 / nsteps, "schedules": [RAMP]}

    syn = [
        ("syn_ss_kstv_static", 4, {"time": static(1), "integrator": "static", "kbc": CLAMP_X0, "fbc": pull_f,
                                   "element": {"type": "small_strain"}, "material": kstv, "solver": NEWTON},
         ["--fint", "--lhs"]),
        ("syn_tl_simo_static", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": [],
                                   "element": {"type": "total_lagrangian"}, "material": simo_soft, "solver": NEWTON},
         ["--every", "1", "--fint", "--lhs"]),
        ("syn_ul_simo_static", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": [],
                                   "element": {"type": "updated_lagrangian"}, "material": simo_soft, "solver": NEWTON},
         ["--every", "1", "--fint", "--lhs"]),
        ("syn_tl_fdkstv_static", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": [],
                                     "element": {"type": "total_lagrangian"}, "material": fdkstv, "solver": NEWTON},
         ["--every", "1", "--fint", "--lhs"]),
        ("syn_ul_fdkstv_static", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": [],
                                     "element": {"type": "updated_lagrangian"}, "material": fdkstv, "solver": NEWTON},
         ["--every", "1", "--fint", "--lhs"]),
        ("syn_ul_j2_static", 3, {"time": static(4), "integrator": "static", "kbc": pull_u(0.06), "fbc": [],
                                 "element": {"type": "updated_lagrangian"}, "material": j2, "solver": NEWTON},
         ["--every", "1", "--fint", "--lhs"]),
        # Simo_J2 with a parabolic-run-out cubic_spline K(alpha) on non-uniform knots (the local Newton crosses several spline intervals)
        ("syn_ul_j2_spline_static", 3, {"time": static(4), "integrator": "static", "kbc": pull_u(0.08), "fbc": [],
                                        "element": {"type": "updated_lagrangian"},
                                        "material": dict(j2, hardening={"type": "cubic_spline", "fixity": "parabolic",
                                                                        "points": [[0.0, 0.25], [0.005, 0.255], [0.02, 0.26], [0.04, 0.30], [0.1, 0.31]]}),
                                        "solver": NEWTON}, ["--every", "1", "--fint", "--lhs"]),
        # a5: mean-dilatation B-bar (SmallStrainT strain_displacement="B-bar"), nearly incompressible so that it matters
        ("syn_ss_kstv_bbar_static", 4, {"time": static(1), "integrator": "static", "kbc": CLAMP_X0, "fbc": pull_f,
                                        "element": {"type": "small_strain", "strain_displacement": "B-bar"},
                                        "material": dict(kstv, nu=0.49), "solver": NEWTON}, ["--fint", "--lhs"]),
        # nodal stress output (extrapolation + averaging) after a static solve
        ("syn_tl_simo_stress", 3, {"time": static(1), "integrator": "static", "kbc": pull_u(0.15), "fbc": [], "output_inc": 1,
                                   "element": {"type": "total_lagrangian", "nodal_output": "stress"}, "material": simo_soft, "solver": NEWTON},
         ["--fint"]),
        ("syn_ss_kstv_stress", 4, {"time": static(1), "integrator": "static", "kbc": CLAMP_X0, "fbc": pull_f, "output_inc": 1,
                                   "element": {"type": "small_strain", "nodal_output": "stress"}, "material": kstv, "solver": NEWTON},
         ["--fint"]),
        ("syn_ul_fdkstv_stress", 3, {"time": static(1), "integrator": "static", "kbc": pull_u(0.15), "fbc": [], "output_inc": 1,
                                     "element": {"type": "updated_lagrangian", "nodal_output": "stress"}, "material": fdkstv, "solver": NEWTON},
         ["--fint"]),
        # natural_bc tractions on curved faces: a follower-free pressure-like load in the facet frame with a different
        # vector at each facet node, plus a constant global shear on another face
        ("syn_tl_simo_traction", 3, {"time": static(2), "integrator": "static", "kbc": CLAMP_X0, "fbc": [],
                                     "element": {"type": "total_lagrangian", "natural_bc": [
                                         {"side_set": 2, "schedule": 1, "coordinate_system": "local",
                                          "values": [[0.3, 0.1, -2.0], [0.2, 0.0, -2.5], [0.1, -0.1, -3.0], [0.0, 0.2, -1.5]]},
                                         {"side_set": 6, "schedule": 1, "coordinate_system": "global", "values": [[0.0, 0.5, 0.25]]}]},
                                     "material": simo_soft, "solver": NEWTON}, ["--every", "1", "--fint"]),
        ("syn_ss_kstv_traction", 4, {"time": static(1), "integrator": "static", "kbc": CLAMP_X0, "fbc": pull_f,
                                     "element": {"type": "small_strain", "natural_bc": [
                                         {"side_set": 4, "schedule": 1, "coordinate_system": "local", "values": [[0.0, 0.0, -1.0]]}]},
                                     "material": kstv, "solver": NEWTON}, ["--fint", "--lhs"]),
        ("syn_ul_j2_stress", 3, {"time": static(3), "integrator": "static", "kbc": pull_u(0.06), "fbc": [], "output_inc": 3,
                                 "element": {"type": "updated_lagrangian", "nodal_output": "stress"}, "material": j2, "solver": NEWTON},
         ["--every", "1", "--fint"]),
        # a21: nonlinear PCG (PCGSolver_LS) -- linear, finite-strain and J2 cases
        ("syn_ss_kstv_pcg", 3, {"time": static(1), "integrator": "static", "kbc": CLAMP_X0, "fbc": pull_f,
                                "element": {"type": "small_strain"}, "material": kstv, "solver": PCG}, ["--fint"]),
        ("syn_ul_fdkstv_pcg", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": [],
                                  "element": {"type": "updated_lagrangian"}, "material": fdkstv, "solver": PCG},
         ["--every", "1", "--fint"]),
        ("syn_tl_simo_pcg", 3, {"time": static(2), "integrator": "static", "kbc": pull_u(0.15), "fbc": pull_f,
                                "element": {"type": "total_lagrangian"}, "material": simo_soft, "solver": PCG},
         ["--every", "1", "--fint"]),
        ("syn_ul_j2_pcg", 3, {"time": static(4), "integrator": "static", "kbc": pull_u(0.06), "fbc": [],
                              "element": {"type": "updated_lagrangian"}, "material": j2, "solver": PCG},
         ["--every", "1", "--fint"]),
        ("syn_tl_simo_explicit", 4, {"time": {"num_steps": 40, "time_step": 0.5 * 0.25 / np.sqrt(1000.0 + 4.0 * 5.0 / 3.0),
                                              "schedules": [[(0.0, 1.0)]]},
                                     "integrator": "central_difference", "kbc": CLAMP_X0,
                                     "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02},
                                             {"nodeset": 2, "dof": 3, "schedule": 1, "value": -0.01}],
                                     "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"},
                                     "material": simo, "solver": EXPLICIT},
         ["--every", "10", "--fint"]),
        ("syn_ul_fdkstv_explicit", 4, {"time": {"num_steps": 40, "time_step": 0.01, "schedules": [[(0.0, 1.0)]]},
                                       "integrator": "central_difference",
                                       "kbc": CLAMP_X0 + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.0}],
                                       "fbc": [{"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.05}],
                                       "element": {"type": "updated_lagrangian", "mass_type": "lumped_mass"},
                                       "material": fdkstv, "solver": EXPLICIT},
         ["--every", "10", "--fint"]),
        # a2/a16 inertia branches: implicit dynamics (nonlinear_HHT = Newmark beta 1/4, gamma 1/2) with consistent and lumped mass
        ("syn_ul_fdkstv_implicit", 3, {"time": {"num_steps": 4, "time_step": 0.05, "schedules": [[(0.0, 1.0)]]},
                                       "integrator": "nonlinear_HHT", "kbc": CLAMP_X0,
                                       "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.05},
                                               {"nodeset": 2, "dof": 3, "schedule": 1, "value": -0.02}],
                                       "element": {"type": "updated_lagrangian", "mass_type": "consistent_mass"},
                                       "material": fdkstv, "solver": NEWTON}, ["--every", "1", "--fint"]),
        ("syn_tl_simo_implicit", 3, {"time": {"num_steps": 4, "time_step": 0.05, "schedules": [RAMP_FAST]},
                                     "integrator": "nonlinear_HHT",
                                     "kbc": CLAMP_X0 + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.1}],
                                     "fbc": [{"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.05}],
                                     "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"},
                                     "material": simo_soft, "solver": NEWTON}, ["--every", "1", "--fint"]),
        # explicit_solid (ExplicitElementT: batched UL force, ExplNeoHookeanT / ExplJ2PlasticityT, mass scaling)
        ("syn_xs_neo_explicit", 4, {"time": {"num_steps": 40, "time_step": 0.5 * 0.25 / np.sqrt(1000.0 + 4.0 * 5.0 / 3.0),
                                             "schedules": [[(0.0, 1.0)]]},
                                    "integrator": "central_difference", "kbc": CLAMP_X0,
                                    "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02},
                                            {"nodeset": 2, "dof": 3, "schedule": 1, "value": -0.01}],
                                    "element": {"type": "explicit_solid", "mass_type": "lumped_mass"},
                                    "material": {"type": "explicit_neo_hookean", "density": 1.0, "kappa": 1000.0, "mu": 5.0},
                                    "solver": EXPLICIT}, ["--every", "10", "--fint"]),
        ("syn_xs_j2_explicit", 4, {"time": {"num_steps": 400, "time_step": 0.4 * 0.25 / np.sqrt(1000.0 + 4.0 * 50.0 / 3.0),
                                            "schedules": [RAMP_FAST]},
                                   "integrator": "central_difference",
                                   "kbc": CLAMP_X0 + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.08}],
                                   "fbc": [],
                                   "element": {"type": "explicit_solid", "mass_type": "lumped_mass"},
                                   "material": {"type": "explicit_J2", "density": 1.0, "kappa": 1000.0, "mu": 50.0, "sigma_Y": 2.0,
                                                "hardening_modulus": 100.0},
                                   "solver": EXPLICIT}, ["--every", "100", "--fint"]),
        ("syn_xs_neo_massscaled_explicit", 4, {"time": {"num_steps": 40, "time_step": 0.5 * 0.25 / np.sqrt(1000.0 + 4.0 * 5.0 / 3.0),
                                                        "schedules": [[(0.0, 1.0)]]},
                                               "integrator": "central_difference", "kbc": CLAMP_X0,
                                               "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}],
                                               "element": {"type": "explicit_solid", "mass_type": "lumped_mass",
                                                           "mass_scaling": {"type": "fixed", "target_dt": "0.0077", "scale_factor": "0.9"}},
                                               "material": {"type": "explicit_neo_hookean", "density": 1.0, "kappa": 1000.0, "mu": 5.0},
                                               "solver": EXPLICIT}, ["--every", "10", "--fint"]),
    ]
    for name, n, desc, flags in syn:
        if want(name):
            synthetic_case(name, n, desc, flags)


if __name__ == "__main__":
    main()
