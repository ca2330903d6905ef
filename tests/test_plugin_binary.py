"""The drop-in itself: a Tahoe executable built from the UNMODIFIED reference libraries + tahoe_b200/host (plugin classes and
the registration patch) must, for inputs that differ from classic ones by the element / matrix tag only, reproduce the
reference executable's nodal output.  The reference's own regression criterion is rel 1e-8 on the 12-digit .run files
(benchmark_XML/comparator/src/ComparatorT.cpp:27-28); here 1e-9.

CPU part (no GPU): the plugin binary validates the new tags against its parameter tree and then fails loudly because no
CUDA device exists -- it must not fall back to the host element loop."""
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import tahoe_input as ti

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(REPO, "oracle", "_ref", "tahoe")
PLUGIN_BIN = os.path.join(REPO, "tahoe_b200", "host", "_build", "tahoe_b200")
COMPARE_BIN = os.path.join(REPO, "oracle", "_ref", "compare")  # the reference's own regression comparator (benchmark_XML/comparator)
needs_bins = pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(PLUGIN_BIN)),
                                reason="reference / plugin executables are built in the authoring container (make -C tahoe_b200/host)")

CLAMP = [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)]
RAMP = [(0.0, 0.0), (1.0, 1.0)]


def _cases():
    kstv = {"type": "small_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25}
    simo = {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 5.0}
    simo_soft = {"type": "Simo_isotropic", "density": 1.0, "E": 100.0, "nu": 0.25}
    j2 = {"type": "Simo_J2", "density": 1.0, "E": 100.0, "nu": 0.25, "hardening": {"type": "linear_function", "a": 0.05, "b": 0.25}}
    newton = {"type": "nonlinear_solver", "abs_tolerance": "1.0e-12", "rel_tolerance": "1.0e-12", "divergence_tolerance": "1.0e+03",
              "max_iterations": "25", "matrix": "SPOOLES_matrix"}
    pcg = dict(newton, matrix="CUDA_PCG_matrix", matrix_attrs='rel_tolerance="1.0e-13" max_iterations="20000"')
    # a21: the reference's nonlinear PCG (PCGSolver_LS + diagonal_matrix) against the device-resident CUDA_PCG_solver
    nlpcg = {"type": "PCG_solver", "abs_tolerance": "1.0e-13", "divergence_tolerance": "10.0", "line_search_iterations": "10",
             "line_search_tolerance": "0.1", "max_iterations": "3000", "max_step": "2.5", "quick_solve_iter": "100",
             "rel_tolerance": "1.0e-12", "restart": "40", "matrix": "diagonal_matrix"}
    cuda_nlpcg = dict(nlpcg, type="CUDA_PCG_solver")
    pull = CLAMP + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.06}, {"nodeset": 2, "dof": 3, "type": "u", "schedule": 1, "value": 0.02}]
    n = 5
    dt = 0.25 * (1.0 / n) / np.sqrt(1000.0 + 4.0 * 5.0 / 3.0)
    resident = {"type": "CUDA_explicit_solver", "matrix": "diagonal_matrix", "integrator": "CUDA_central_difference"}
    return {
        # explicit dynamics: device K1, Tahoe's own lumped mass / DiagonalMatrixT / nExplicitCD on the host
        "explicit_tl_simo": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                              "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}],
                              "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"}, "material": simo,
                              "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        # the same run RESIDENT on the device: integrator="CUDA_central_difference" + <CUDA_explicit_solver> (d, v, a stay on the GPU,
        # FieldT receives them when a step writes output): nodal force, body force, and a prescribed displacement following a schedule
        "explicit_tl_simo_resident": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                       "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}],
                                       "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"}, "material": simo,
                                       "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, resident),
        "explicit_tl_simo_gravity_resident": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                               "kbc": CLAMP, "fbc": [],
                                               "element": {"type": "total_lagrangian", "mass_type": "lumped_mass",
                                                           "body_force": {"schedule": 1, "vector": [0.0, 0.3, -9.81]}},
                                               "material": simo, "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, resident),
        "explicit_ul_simo_pull_resident": ({"time": {"num_steps": 40, "time_step": dt, "schedules": [[(0.0, 0.0), (40 * dt, 1.0)]]},
                                            "integrator": "central_difference",
                                            "kbc": CLAMP + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.01}], "fbc": [],
                                            "element": {"type": "updated_lagrangian", "mass_type": "lumped_mass"}, "material": simo,
                                            "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, resident),
        # the reference's batched <explicit_solid> against <cuda_explicit_solid>: Neo-Hookean with fixed mass scaling
        # (host-side lumped mass incl. the scaling, device force) and Hughes-Winget J2 pulled past yield (device-resident history)
        "explicit_solid_neo": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}],
                                "element": {"type": "explicit_solid", "mass_type": "lumped_mass",
                                            "mass_scaling": {"type": "fixed", "target_dt": "%.6g" % (1.25 * 0.874 * 0.2 / np.sqrt(1000.0 + 20.0 / 3.0)),
                                                             "scale_factor": "0.9"}},
                                "material": {"type": "explicit_neo_hookean", "density": 1.0, "kappa": 1000.0, "mu": 5.0},
                                "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        "explicit_solid_neo_adaptive": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                         "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}],
                                         "element": {"type": "explicit_solid", "mass_type": "lumped_mass",
                                                     "mass_scaling": {"type": "adaptive", "target_dt": "%.6g" % (1.25 * 0.874 * 0.2 / np.sqrt(1000.0 + 20.0 / 3.0)),
                                                                      "scale_factor": "0.9", "update_interval": "7"}},
                                         "material": {"type": "explicit_neo_hookean", "density": 1.0, "kappa": 1000.0, "mu": 5.0},
                                         "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        "explicit_solid_j2": ({"time": {"num_steps": 300, "time_step": 0.4 * 0.2 / np.sqrt(1000.0 + 200.0 / 3.0), "schedules": [[(0.0, 0.0), (0.3, 1.0), (10.0, 1.0)]]},
                               "integrator": "central_difference",
                               "kbc": CLAMP + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.08}], "fbc": [],
                               "element": {"type": "explicit_solid", "mass_type": "lumped_mass"},
                               "material": {"type": "explicit_J2", "density": 1.0, "kappa": 1000.0, "mu": 50.0, "sigma_Y": 2.0, "hardening_modulus": 100.0},
                               "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        # static Newton: device K1 + device K3 + device PCG against the reference's SPOOLES LU
        "static_ss_kstv_pcg": ({"time": {"num_steps": 1, "time_step": 1.0, "schedules": [RAMP]}, "integrator": "static", "kbc": CLAMP,
                                "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}, {"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.005}],
                                "element": {"type": "small_strain"}, "material": kstv, "solver": newton}, pcg),
        # a5: mean-dilatation B-bar at nu = 0.49, device K1 + K3 (B-bar variants) + device PCG
        "static_ss_bbar_pcg": ({"time": {"num_steps": 1, "time_step": 1.0, "schedules": [RAMP]}, "integrator": "static", "kbc": CLAMP,
                                "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02}, {"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.005}],
                                "element": {"type": "small_strain", "strain_displacement": "B-bar"}, "material": dict(kstv, nu=0.49),
                                "solver": newton}, pcg),
        # nodal stress output through the plugin's ComputeOutput (device extrapolation + averaging)
        "static_tl_simo_stress_out": ({"time": {"num_steps": 1, "time_step": 1.0, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                       "element": {"type": "total_lagrangian", "nodal_output": "stress"}, "material": simo_soft, "solver": newton}, pcg),
        "static_tl_simo_pcg": ({"time": {"num_steps": 2, "time_step": 0.5, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                "element": {"type": "total_lagrangian"}, "material": simo_soft, "solver": newton}, pcg),
        "static_tl_simo_nlpcg": ({"time": {"num_steps": 2, "time_step": 0.5, "schedules": [RAMP]}, "integrator": "static", "kbc": pull,
                                  "fbc": [{"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.004}],
                                  "element": {"type": "total_lagrangian"}, "material": simo_soft, "solver": nlpcg}, cuda_nlpcg),
        "static_ul_j2_nlpcg": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                "element": {"type": "updated_lagrangian"}, "material": j2, "solver": nlpcg}, cuda_nlpcg),
        # J2 stress output on the device (J2Simo3D::s_ij with the device-resident history, before the step's history update)
        "static_ul_j2_stress_out": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                          "element": {"type": "updated_lagrangian", "nodal_output": "stress"}, "material": j2, "solver": newton}, None),
        # body force in an explicit run: -M b formed on the device through the mass operator (SolidElementT.cpp:1204-1265)
        "explicit_tl_simo_gravity": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                      "kbc": CLAMP, "fbc": [],
                                      "element": {"type": "total_lagrangian", "mass_type": "lumped_mass",
                                                  "body_force": {"schedule": 1, "vector": [0.0, 0.3, -9.81]}},
                                      "material": simo, "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        # (with an implicit integrator the reference's AddBodyForce overwrites the nodal accelerations, ContinuumElementT.cpp:858-863,
        #  and its own Newton loop cannot converge -- not a usable combination, so it is not a test case)
        # implicit dynamics (nonlinear_HHT): device fint + M a, device K and M assembled into the CUDA_PCG_matrix, against SPOOLES
        "implicit_ul_kstv_consistent_pcg": ({"time": {"num_steps": 4, "time_step": 0.05, "schedules": [[(0.0, 1.0)]]}, "integrator": "nonlinear_HHT",
                                             "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.05},
                                                                   {"nodeset": 2, "dof": 3, "schedule": 1, "value": -0.02}],
                                             "element": {"type": "updated_lagrangian", "mass_type": "consistent_mass"},
                                             "material": {"type": "large_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25},
                                             "solver": newton}, pcg),
        "implicit_tl_simo_lumped_pcg": ({"time": {"num_steps": 4, "time_step": 0.05, "schedules": [[(0.0, 0.0), (0.1, 1.0), (10.0, 1.0)]]},
                                         "integrator": "nonlinear_HHT",
                                         "kbc": CLAMP + [{"nodeset": 2, "dof": 1, "type": "u", "schedule": 1, "value": 0.05}],
                                         "fbc": [{"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.02}],
                                         "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"}, "material": simo_soft,
                                         "solver": newton}, pcg),
        # Simo_J2 with tabulated / power-law hardening: the plugin hands the knots (or a, b, c, n) to the library
        "static_ul_j2_spline_lu": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                    "element": {"type": "updated_lagrangian"},
                                    "material": dict(j2, hardening={"type": "cubic_spline", "fixity": "free_run",
                                                                    "points": [[0.0, 0.25], [0.01, 0.255], [0.05, 0.26], [0.10, 0.30]]}),
                                    "solver": newton}, None),
        "static_ul_j2_powerlaw_lu": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                      "element": {"type": "updated_lagrangian"},
                                      "material": dict(j2, hardening={"type": "power_law", "a": 0.25, "b": 1.0, "c": 400.0, "n": 0.8}),
                                      "solver": newton}, None),
        # two element blocks with different materials in one element group (a1: ElementCardT material ids): one device group per material
        "static_tl_two_materials_pcg": ({"time": {"num_steps": 2, "time_step": 0.5, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                         "element": {"type": "total_lagrangian", "nodal_output": "stress"}, "material": simo_soft,
                                         "materials": [simo_soft, {"type": "large_strain_StVenant", "density": 2.0, "E": 300.0, "nu": 0.3}],
                                         "solver": newton}, pcg),
        "static_ul_j2_and_elastic_lu": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                         "element": {"type": "updated_lagrangian"}, "material": j2,
                                         "materials": [dict(j2, hardening={"type": "power_law", "a": 0.25, "b": 1.0, "c": 400.0, "n": 0.8}), simo_soft, ],
                                         "solver": newton}, None),
        "explicit_tl_two_materials_gravity": ({"time": {"num_steps": 30, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                                               "kbc": CLAMP, "fbc": [],
                                               "element": {"type": "total_lagrangian", "mass_type": "lumped_mass",
                                                           "body_force": {"schedule": 1, "vector": [0.0, 0.3, -9.81]}},
                                               "material": simo, "materials": [simo, dict(simo, density=3.0, mu=8.0)],
                                               "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}, None),
        # configs[3] through the executable: UL + Simo_J2 under Tahoe's Newton with the device-assembled non-symmetric tangent solved by the
        # Jacobi-BiCGStab of <CUDA_PCG_matrix> (the reference: LU)
        "static_ul_j2_bicgstab": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                                   "element": {"type": "updated_lagrangian"}, "material": j2, "solver": newton},
                                  dict(newton, matrix="CUDA_PCG_matrix", matrix_attrs='rel_tolerance="1.0e-12" max_iterations="20000"')),
        # J2: device K1 with history, Tahoe's host tangent + SPOOLES (non-symmetric tangent)
        "static_ul_j2_lu": ({"time": {"num_steps": 3, "time_step": 1.0 / 3, "schedules": [RAMP]}, "integrator": "static", "kbc": pull, "fbc": [],
                             "element": {"type": "updated_lagrangian"}, "material": j2, "solver": newton}, None),
    }


def _write(work, name, desc, cuda, solver_override, n=5, tahoe_attrs=None, suffix=""):
    X, conn, ns = ti.structured_cube(n, jitter=0.15)
    if not os.path.exists(os.path.join(work, "mesh.geom")):
        nmat = len(desc.get("materials") or [0])  # several materials: the bottom two element layers are block 1, the rest block 2
        ti.write_geom(os.path.join(work, "mesh.geom"), X, conn, ns, block_sizes=None if nmat == 1 else [2 * n * n, (n - 2) * n * n])
    d = dict(desc, geometry_file="mesh.geom", output_inc=desc["time"]["num_steps"])
    d["element"] = dict(desc["element"], nodal_output=desc["element"].get("nodal_output", True))
    if cuda:
        d["element"]["tag"] = "cuda_" + desc["element"]["type"]
        if solver_override:
            d["solver"] = {k: v for k, v in solver_override.items() if k != "integrator"}
            if "integrator" in solver_override:  # the resident explicit pair: solver tag + the field's integrator
                d["integrator"] = solver_override["integrator"]
    if tahoe_attrs:
        d["tahoe_attrs"] = tahoe_attrs
    path = os.path.join(work, name + (".cuda" if cuda else ".ref") + suffix + ".xml")
    ti.write_xml(path, d)
    return path


def _nodal_output(run_file):
    """last 'Nodal data' table of a Tahoe text .run file -> array [nn, nvalues]"""
    import glob
    parts = sorted(glob.glob(run_file + ".ps*"))  # one file per print step next to the table of contents
    text = open(parts[-1] if parts else run_file).read()
    block = text[text.rindex("Nodal data:"):]
    rows = []
    for line in block.splitlines():
        f = line.split()
        if len(f) >= 5 and re.match(r"^\d+$", f[0]) and re.match(r"^\d+$", f[1]):
            rows.append([float(x) for x in f[2:]])
        elif rows and not f:
            break
    return np.array(rows)


def _run(binary, xml):
    return subprocess.run([binary, "-f", os.path.basename(xml)], cwd=os.path.dirname(xml), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)


@needs_bins
def test_plugin_binary_accepts_new_tags_and_fails_loudly_without_gpu():
    from tahoe_b200 import capi
    try:
        if capi.device_count() > 0:
            pytest.skip("a CUDA device is present")
    except capi.Tb2Error:
        pass
    work = tempfile.mkdtemp(prefix="tb2_plugin_")
    try:
        desc, override = _cases()["static_ss_kstv_pcg"]
        xml = _write(work, "case", desc, True, override, n=2)
        r = _run(PLUGIN_BIN, xml)
        assert "cuda_small_strain" not in r.stdout or "unrecognized" not in r.stdout.lower()
        assert "CUDA error" in r.stdout or "no CUDA" in r.stdout or "cuda" in r.stdout.lower(), r.stdout[-2000:]
        assert not os.path.exists(os.path.join(work, "case.cuda.io0.run"))  # nothing was computed
    finally:
        shutil.rmtree(work, ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(PLUGIN_BIN), reason="the plugin executable is built in the authoring container (make -C tahoe_b200/host)")
def test_plugin_schema_lists_the_cuda_tags():
    """`tahoe -xsd` (FEExecutionManagerT.cpp:219-221) generates tahoe.xsd from DefineParameters / DefineSubs / NewSub: the plugin's
    element groups, matrix, solvers and integrator must appear in the schema the patched executable writes (no GPU needed)"""
    work = tempfile.mkdtemp(prefix="tb2_xsd_")
    try:
        r = subprocess.run([PLUGIN_BIN, "-xsd"], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        xsd = os.path.join(work, "tahoe.xsd")
        assert os.path.exists(xsd), r.stdout[-2000:]
        text = open(xsd).read()
        for tag in ("cuda_small_strain", "cuda_total_lagrangian", "cuda_updated_lagrangian", "cuda_explicit_solid", "cuda_contact_3D_penalty", "CUDA_PCG_matrix",
                    "CUDA_PCG_solver", "CUDA_explicit_solver", "CUDA_central_difference"):
            assert tag in text, tag
        # the matrix's attributes and the explicit solver's restart attribute are declared, not just the names
        assert re.search(r"element name='CUDA_PCG_matrix'>.*?attribute name='rel_tolerance'", text, re.S)
        assert re.search(r"element name='CUDA_explicit_solver'>.*?attribute name='restart_output_inc'", text, re.S)
        assert re.search(r"enumeration value='CUDA_central_difference'", text)
    finally:
        shutil.rmtree(work, ignore_errors=True)


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_cases()))
def test_plugin_reproduces_reference_output(name):
    desc, override = _cases()[name]
    work = tempfile.mkdtemp(prefix="tb2_plugin_")
    try:
        ref_xml = _write(work, name, desc, False, None)
        cuda_xml = _write(work, name, desc, True, override)
        r0 = _run(REF_BIN, ref_xml)
        assert r0.returncode == 0, r0.stdout[-2000:]
        r1 = _run(PLUGIN_BIN, cuda_xml)
        assert r1.returncode == 0 and "End Execution" in r1.stdout and "ExceptionT::Throw" not in r1.stdout, r1.stdout[-3000:]
        a = _nodal_output(os.path.join(work, name + ".ref.io0.run"))
        b = _nodal_output(os.path.join(work, name + ".cuda.io0.run"))
        assert a.shape == b.shape and a.shape[0] == 6 ** 3
        assert np.abs(a).max() > 1e-6
        # two nonlinear-CG runs stopped at |R| < 1e-12 |R0| agree to the conditioning of the problem, not to the LU cases' 1e-9
        tol = 1e-7 if name.endswith("nlpcg") else 1e-9
        assert np.abs(a - b).max() < tol * np.abs(a).max()
        if name.endswith("nlpcg"):
            assert "device PCG" in r1.stdout
        if name.endswith("resident"):
            assert "CUDA_explicit_solver" in open(cuda_xml).read() and "CUDA_central_difference" in open(cuda_xml).read()
        if override is None and os.path.exists(COMPARE_BIN):
            # the reference's own acceptance test (run_benchmarks.sh) for the cases that keep the reference's solver: its comparator,
            # with its default tolerances (a value fails when it is off by more than 1e-8 relative AND 1e-10 absolute), checks the CUDA
            # run's output files against the classic run's as the `benchmark/` reference
            import glob
            os.makedirs(os.path.join(work, "benchmark"), exist_ok=True)
            for f in glob.glob(os.path.join(work, name + ".ref.io0.*")):
                dst = os.path.join(work, "benchmark", os.path.basename(f).replace(".ref.", ".cuda."))
                with open(f) as src, open(dst, "w") as out:  # the .run table of contents names its .ps files
                    out.write(src.read().replace(name + ".ref.", name + ".cuda."))
            rc = subprocess.run([COMPARE_BIN, "-f", name + ".cuda.xml"], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
            assert (name + ".cuda.xml: PASS") in rc.stdout, rc.stdout[-2500:]
        if name.endswith("stress_out") or name == "static_tl_two_materials_pcg":
            assert a.shape[1] == 9
            # single-material groups take the device path; the two-material group goes through Tahoe's host ComputeOutput
            assert ("nodal stresses extrapolated and averaged on the device" in r1.stdout) == name.endswith("stress_out")
            for col in range(9):  # every column against its own scale: D_X D_Y D_Z s11 s22 s33 s23 s13 s12
                assert np.abs(a[:, col] - b[:, col]).max() < 1e-9 * np.abs(a[:, col]).max()
    finally:
        shutil.rmtree(work, ignore_errors=True)


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("writer", ["ref", "cuda"])
def test_plugin_restart_files_are_interchangeable_with_the_reference(writer):
    """SURVEY 8(f)-2 restart hand-off: a Simo_J2 run is stopped after step 2 of 3 by one executable and finished from the restart
    file by the other (FEManagerT::WriteRestart/ReadRestart, FEManagerT.cpp:2086-2240; ContinuumElementT.cpp:217-249 element
    records).  The device-resident plastic history has to survive the hand-off in both directions."""
    desc, _ = _cases()["static_ul_j2_lu"]
    work = tempfile.mkdtemp(prefix="tb2_restart_")
    try:
        full = _write(work, "full", desc, False, None)
        r = _run(REF_BIN, full)
        assert r.returncode == 0, r.stdout[-2000:]
        want = _nodal_output(os.path.join(work, "full.ref.io0.run"))
        bins = {"ref": REF_BIN, "cuda": PLUGIN_BIN}
        reader = "cuda" if writer == "ref" else "ref"
        first = _write(work, "first", desc, writer == "cuda", None, tahoe_attrs={"restart_output_inc": "2"})
        r = _run(bins[writer], first)
        assert r.returncode == 0, r.stdout[-2000:]
        rs = "first.%s.rs2of3" % writer
        assert os.path.exists(os.path.join(work, rs)) and os.path.exists(os.path.join(work, rs + ".elem0"))
        # the element record holds allocated (yielded) elements: "1" lines followed by "<flag> 8 304"
        elem = open(os.path.join(work, rs + ".elem0")).read()
        assert " 8 304" in elem
        second = _write(work, "second", desc, reader == "cuda", None, tahoe_attrs={"restart_file": rs})
        r = _run(bins[reader], second)
        assert r.returncode == 0 and "Restart file" in r.stdout, r.stdout[-3000:]
        got = _nodal_output(os.path.join(work, "second.%s.io0.run" % reader))
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    finally:
        shutil.rmtree(work, ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(PLUGIN_BIN), reason="the plugin executable is built in the authoring container (make -C tahoe_b200/host)")
@pytest.mark.gpu
def test_plugin_at_scale_50_cubed():
    """The plugin executable on 50^3 = 125 k elements (384 k equations, 30 M non-zeros), where the reference's own SPOOLES LU is no
    longer a practical comparison: (i) static small strain through <cuda_small_strain> + <CUDA_PCG_matrix> against the C oracle's
    assembled K solved by SciPy's CG -- and the host MSR structure (MSRBuilderT graph, fbindx, fval) must never have been built;
    (ii) 20 steps of the resident explicit pair (<CUDA_explicit_solver> + CUDA_central_difference) against the oracle's
    predictor / force sweep / corrector."""
    import sys
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as oracle
    oracle.build()
    n = 50
    X, conn, ns = ti.structured_cube(n, jitter=0.15)
    kstv = {"type": "small_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25}
    simo = {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 5.0}
    newton = {"type": "nonlinear_solver", "abs_tolerance": "1.0e-12", "rel_tolerance": "1.0e-9", "divergence_tolerance": "1.0e+03",
              "max_iterations": "5", "matrix": "CUDA_PCG_matrix", "matrix_attrs": 'rel_tolerance="1.0e-12" max_iterations="20000"'}
    work = tempfile.mkdtemp(prefix="tb2_plugin50_")
    try:
        # ---- (i) static
        desc = {"time": {"num_steps": 1, "time_step": 1.0, "schedules": [RAMP]}, "integrator": "static", "kbc": CLAMP,
                "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02 / n ** 2}, {"nodeset": 2, "dof": 2, "schedule": 1, "value": 0.005 / n ** 2}],
                "element": {"type": "small_strain"}, "material": kstv, "solver": newton}
        xml = _write(work, "static50", desc, True, newton, n=n)
        r = _run(PLUGIN_BIN, xml)
        assert r.returncode == 0 and "End Execution" in r.stdout and "ExceptionT::Throw" not in r.stdout, r.stdout[-3000:]
        log = open(os.path.join(work, "static50.cuda.out")).read()
        assert "device-assembled tangent, host MSR structure not built" in log
        assert "host MSR structure built" not in log
        got = _nodal_output(os.path.join(work, "static50.cuda.io0.run"))
        assert got.shape[0] == (n + 1) ** 3
        code = np.zeros(X.shape, np.uint8)
        code[ns[1]] = 1
        eq, neq = oracle.equation_numbers(code)
        rp, ci = oracle.csr_structure(conn, eq, neq)
        err, kv = oracle.assemble_stiffness(oracle.SMALL_STRAIN, oracle.material(kstv), conn, X, np.zeros_like(X), eq, neq, rp, ci)
        assert err == 0
        K = sp.csr_matrix((kv, ci, rp), shape=(neq, neq))
        f = np.zeros_like(X)
        f[ns[2], 0] = 0.02 / n ** 2
        f[ns[2], 1] = 0.005 / n ** 2
        act = eq.reshape(-1) > 0
        dinv = 1.0 / K.diagonal()
        x, info = spla.cg(K, f.reshape(-1)[act], rtol=1e-12, maxiter=20000, M=spla.LinearOperator((neq, neq), matvec=lambda v: dinv * v))
        assert info == 0
        want = np.zeros(3 * X.shape[0])
        want[act] = x
        want = want.reshape(-1, 3)
        assert np.abs(got[:, :3] - want).max() < 1e-7 * np.abs(want).max()
        # ---- (ii) resident explicit
        dt = 0.25 * (1.0 / n) / np.sqrt(1000.0 + 4.0 * 5.0 / 3.0)
        nsteps = 20
        resident = {"type": "CUDA_explicit_solver", "matrix": "diagonal_matrix", "integrator": "CUDA_central_difference"}
        desc = {"time": {"num_steps": nsteps, "time_step": dt, "schedules": [[(0.0, 1.0)]]}, "integrator": "central_difference",
                "kbc": CLAMP, "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02 / n ** 2}],
                "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"}, "material": simo,
                "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}
        xml = _write(work, "explicit50", desc, True, resident, n=n)
        r = _run(PLUGIN_BIN, xml)
        assert r.returncode == 0 and "End Execution" in r.stdout and "ExceptionT::Throw" not in r.stdout, r.stdout[-3000:]
        got = _nodal_output(os.path.join(work, "explicit50.cuda.io0.run"))
        omat = oracle.material(simo)
        mass = oracle.lumped_mass(1.0, conn, X)
        fext = np.zeros_like(X)
        fext[ns[2], 0] = 0.02 / n ** 2
        d, v = np.zeros_like(X), np.zeros_like(X)
        # FEManagerT::InitialCondition: a0 = M^-1 (fext - fint(0)) on the free dofs
        a = np.where(code > 0, 0.0, fext / mass)
        for _ in range(nsteps):
            oracle.cd_predictor(dt, d, v, a, code, np.zeros_like(X))
            e, fi = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, omat, conn, X, d)
            assert e == 0
            oracle.cd_corrector(dt, v, a, fext - fi, mass, code)
        assert np.abs(got[:, :3] - d).max() < 1e-9 * np.abs(d).max()
    finally:
        shutil.rmtree(work, ignore_errors=True)


# ---- <cuda_contact_3D_penalty> (CudaPenaltyContact3DT) against the reference's <contact_3D_penalty> ---------------

CONTACT_XML = """<?xml version="1.0"?>
<tahoe geometry_file="mesh.geom" title="two stacked cubes, penalty contact">
    <time num_steps="%(num_steps)d" output_inc="%(num_steps)d" time_step="%(dt).6e">
        <schedule_function><piecewise_linear><OrderedPair x="0.0" y="0.0"/><OrderedPair x="1.0" y="1.0"/></piecewise_linear></schedule_function>
    </time>
    <nodes>
        <field field_name="displacement"%(integrator)s>
            <dof_labels><String value="D_X"/><String value="D_Y"/><String value="D_Z"/></dof_labels>
%(conditions)s
        </field>
    </nodes>
    <element_list>
        <%(solid_tag)s field_name="displacement"%(mass)s>
            <hexahedron/>
            <solid_element_nodal_output displacements="1"/>
            <large_strain_element_block>
                <block_ID_list><String value="1"/><String value="2"/></block_ID_list>
                <large_strain_material_3D>
                    <Simo_isotropic density="1.0"><E_and_nu Poisson_ratio="0.25" Young_modulus="100.0"/></Simo_isotropic>
                </large_strain_material_3D>
            </large_strain_element_block>
        </%(solid_tag)s>
        <%(tag)s field_name="displacement" %(contact_attrs)s>
            <contact_surface><surface_side_set side_set_ID="1"/></contact_surface>
            <contact_surface><surface_side_set side_set_ID="2"/></contact_surface>
            <contact_nodes><node_ID_list><String value="3"/><String value="4"/></node_ID_list></contact_nodes>
        </%(tag)s>
    </element_list>
    %(solver)s
</tahoe>
"""


def _two_cubes(work, n=3):
    """two stacked n^3 cubes as element blocks 1 (lower) and 2 (upper, its own nodes, zero gap): node sets 1 bottom of the lower cube,
    2 top of the upper, 3 / 4 the two faces that meet; side sets 1 (block 1, z = L facets) and 2 (block 2, z = 0 facets)"""
    Xl, cl, nsl = ti.structured_cube(n, jitter=0.0)
    X = np.vstack([Xl, Xl + np.array([0.0, 0.0, 1.0])])
    conn = np.vstack([cl, cl + Xl.shape[0]]).astype(np.int32)
    ns = {1: nsl[5], 2: nsl[6] + Xl.shape[0], 3: nsl[6], 4: nsl[5] + Xl.shape[0]}
    ss = ti.cube_side_sets(n)
    ti.write_geom(os.path.join(work, "mesh.geom"), X, conn, ns, sidesets={1: ss[6], 2: ss[5]}, block_sizes=[n ** 3, n ** 3],
                  sideset_blocks={1: 1, 2: 2})
    return X


def _contact_cases():
    clamp = "\n".join('            <kinematic_BC dof="%d" node_ID="1"/>' % d for d in (1, 2, 3))
    return {
        # the shape of the reference's level.5 impact benchmark: the upper cube comes down sliding; friction and damping on
        "impact_friction_damping": dict(
            num_steps=250, dt=2.0e-4, integrator=' integrator="central_difference"', mass=' mass_type="lumped_mass"',
            conditions='            <initial_condition dof="3" node_ID="4" type="D_u" value="-5.0"/>\n'
                       '            <initial_condition dof="3" node_ID="2" type="D_u" value="-5.0"/>\n'
                       '            <initial_condition dof="1" node_ID="4" type="D_u" value="0.5"/>\n'
                       '            <initial_condition dof="1" node_ID="2" type="D_u" value="0.5"/>\n' + clamp,
            contact_attrs='penalty_stiffness="500.0" friction_coefficient="0.3" friction_epsilon_velocity="0.001" viscous_damping="20.0"',
            solver="<linear_solver><diagonal_matrix/></linear_solver>"),
        # the shape of level.2/contact_simple/cubes.1.xml: the top face pushed down quasi-statically, Newton with the inherited contact tangent
        "static_push": dict(
            num_steps=4, dt=0.25, integrator="", mass="",
            conditions=clamp + '\n            <kinematic_BC dof="1" node_ID="2"/>\n            <kinematic_BC dof="2" node_ID="2"/>\n'
                               '            <kinematic_BC dof="3" node_ID="2" schedule="1" type="u" value="-0.1"/>',
            contact_attrs='penalty_stiffness="50.0"',
            solver='<nonlinear_solver abs_tolerance="1.0e-10" divergence_tolerance="1.0e+03" max_iterations="10" rel_tolerance="1.0e-12">'
                   '<profile_matrix/></nonlinear_solver>'),
    }


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_contact_cases()))
def test_plugin_contact_reproduces_reference_output(name):
    work = tempfile.mkdtemp(prefix="tb2_contact_")
    try:
        X = _two_cubes(work)
        case = _contact_cases()[name]
        for tag, suffix in (("contact_3D_penalty", "ref"), ("cuda_contact_3D_penalty", "cuda")):
            open(os.path.join(work, "%s.%s.xml" % (name, suffix)), "w").write(CONTACT_XML % dict(case, tag=tag, solid_tag="updated_lagrangian"))
        r0 = _run(REF_BIN, os.path.join(work, name + ".ref.xml"))
        assert r0.returncode == 0 and "End Execution" in r0.stdout, r0.stdout[-3000:]
        r1 = _run(PLUGIN_BIN, os.path.join(work, name + ".cuda.xml"))
        assert r1.returncode == 0 and "End Execution" in r1.stdout and "ExceptionT::Throw" not in r1.stdout, r1.stdout[-3000:]
        a = _nodal_output(os.path.join(work, name + ".ref.io0.run"))
        b = _nodal_output(os.path.join(work, name + ".cuda.io0.run"))
        assert a.shape == b.shape and a.shape[0] == X.shape[0]
        assert np.abs(a).max() > 1e-3
        assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()
        # contact did happen: the lower cube moved although nothing else loads it
        assert np.abs(a[: X.shape[0] // 2, :3]).max() > 1e-4
    finally:
        shutil.rmtree(work, ignore_errors=True)


# ---- the threaded .geom reader behind ModelManagerT (FastGeomInputT) -- host-only analyses, no GPU needed ----------

@needs_bins
@pytest.mark.parametrize("case", ["static_one_block", "contact_two_blocks"])
def test_plugin_reads_geometry_through_the_fast_reader(case):
    """The plugin executable registers FastGeomInputT for TahoeII files (IOBaseT::NewInput): coordinates, connectivities, node sets and
    side sets of every run come from tb2_geom_open.  Classic inputs (no cuda_* tag) therefore run entirely on the host through the new
    reader and must reproduce the reference executable's output: a one-block static analysis with nodal loads on node sets, and the
    two-block contact case whose surfaces are side sets of different element blocks."""
    work = tempfile.mkdtemp(prefix="tb2_fastgeom_")
    try:
        if case == "static_one_block":
            desc, _ = _cases()["static_ss_kstv_pcg"]
            ref_xml = _write(work, case, desc, False, None)
            cuda_xml = os.path.join(work, case + ".plugin.xml")
            shutil.copy(ref_xml, cuda_xml)
            nn = 6 ** 3
        else:
            X = _two_cubes(work)
            nn = X.shape[0]
            ref_xml, cuda_xml = os.path.join(work, case + ".ref.xml"), os.path.join(work, case + ".plugin.xml")
            for path in (ref_xml, cuda_xml):
                open(path, "w").write(CONTACT_XML % dict(_contact_cases()["static_push"], tag="contact_3D_penalty", solid_tag="updated_lagrangian"))
        r0 = _run(REF_BIN, ref_xml)
        assert r0.returncode == 0 and "End Execution" in r0.stdout, r0.stdout[-2000:]
        r1 = _run(PLUGIN_BIN, cuda_xml)
        assert r1.returncode == 0 and "End Execution" in r1.stdout and "ExceptionT::Throw" not in r1.stdout, r1.stdout[-3000:]
        log = open(os.path.splitext(cuda_xml)[0] + ".out").read()
        assert "FastGeomInputT: %d nodes" % nn in log and "parsed by tb2_geom_open" in log
        assert "reading through TahoeInputT" not in log
        a = _nodal_output(os.path.splitext(ref_xml)[0] + ".io0.run")
        b = _nodal_output(os.path.splitext(cuda_xml)[0] + ".io0.run")
        assert a.shape == b.shape and a.shape[0] == nn and np.abs(a).max() > 1e-6
        assert np.array_equal(a, b)  # the same host code on the same arrays: identical output
    finally:
        shutil.rmtree(work, ignore_errors=True)


@needs_bins
@pytest.mark.gpu
def test_plugin_resident_explicit_run_with_contact():
    """The impact case with everything on the device: <cuda_updated_lagrangian> + <cuda_contact_3D_penalty> under
    integrator="CUDA_central_difference" + <CUDA_explicit_solver>.  The step (predictor, element sweep, contact force on the predicted
    state, update) runs resident; after every step the displacements come down for the reference's own contact search
    (ContactT::RelaxSystem) and the pair list goes up when the search changed it.  Against the classic executable: 1e-9."""
    work = tempfile.mkdtemp(prefix="tb2_contact_res_")
    try:
        X = _two_cubes(work)
        case = _contact_cases()["impact_friction_damping"]
        name = "impact_resident"
        open(os.path.join(work, name + ".ref.xml"), "w").write(CONTACT_XML % dict(case, tag="contact_3D_penalty", solid_tag="updated_lagrangian"))
        resident = dict(case, integrator=' integrator="CUDA_central_difference"',
                        solver='<CUDA_explicit_solver restart_output_inc="0"><diagonal_matrix/></CUDA_explicit_solver>')
        open(os.path.join(work, name + ".cuda.xml"), "w").write(CONTACT_XML % dict(resident, tag="cuda_contact_3D_penalty",
                                                                                 solid_tag="cuda_updated_lagrangian"))
        r0 = _run(REF_BIN, os.path.join(work, name + ".ref.xml"))
        assert r0.returncode == 0 and "End Execution" in r0.stdout, r0.stdout[-3000:]
        r1 = _run(PLUGIN_BIN, os.path.join(work, name + ".cuda.xml"))
        assert r1.returncode == 0 and "End Execution" in r1.stdout and "ExceptionT::Throw" not in r1.stdout, r1.stdout[-3000:]
        a = _nodal_output(os.path.join(work, name + ".ref.io0.run"))
        b = _nodal_output(os.path.join(work, name + ".cuda.io0.run"))
        assert a.shape == b.shape and a.shape[0] == X.shape[0] and np.abs(a).max() > 1e-3
        assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()
        assert np.abs(a[: X.shape[0] // 2, :3]).max() > 1e-4  # the lower cube moved: contact happened
    finally:
        shutil.rmtree(work, ignore_errors=True)
